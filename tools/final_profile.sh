#!/bin/bash
# End-of-round measurements on one B200 (run through gpurun); everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; tail -2 gpurun_out/final_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python tools/eager_steps.py 3 > gpurun_out/final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_mlp_v3_kernel -c 2 -o gpurun_out/final_v3 python tools/v3_timeline.py > gpurun_out/final_ncu.log 2>&1
ncu -i gpurun_out/final_v3.ncu-rep --page raw --csv > gpurun_out/final_v3_raw.csv 2>/dev/null
tail -c 600 gpurun_out/final_bench.json; tail -c 300 gpurun_out/final_reference.json
