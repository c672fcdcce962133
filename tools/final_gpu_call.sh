#!/bin/bash
# One short gpurun call on the shipped build (budget: < 6 minutes of box time): smoke gate, the full GPU test suite, the driver's bench
# command, the ncu launch list and one `ncu --set full` capture of the edge kernel, the cfg2 / cfg5 workloads, one real 1000-step chain.
#   gpurun --timeout 560 -- 'bash tools/final_gpu_call.sh r02b'
tag=${1:-r02b}
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
if ! timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; then
  echo "SMOKE FAILED / TIMED OUT"; tail -5 gpurun_out/${tag}_smoke.log; exit 1
fi
el "$(tail -1 gpurun_out/${tag}_smoke.log)"
timeout 400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_golden.py::test_hybrid_cutoff_vs_oracle --deselect tests/test_gpu_reference_golden.py::test_hybrid_cutoff_rejects_what_does_not_fit > gpurun_out/${tag}_pytest_gpu.log 2>&1
el "pytest (without hybrid): $(tail -1 gpurun_out/${tag}_pytest_gpu.log)"
timeout 200 python -m pytest tests/test_gpu_reference_golden.py -m gpu -q -k hybrid > gpurun_out/${tag}_pytest_hybrid.log 2>&1
el "pytest hybrid: $(tail -1 gpurun_out/${tag}_pytest_hybrid.log)"
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
el "bench: $(cut -c1-260 gpurun_out/${tag}_bench.json)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/eager_steps.py 3 > gpurun_out/${tag}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_shares.csv 2>/dev/null; el "launch list"; head -8 gpurun_out/${tag}_launch_shares.csv
TDIFF_FREE_DEPTH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_mlp_v4_kernel -s 52 -c 2 -o gpurun_out/${tag}_v4 python tools/eager_steps.py 2 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_v4.ncu-rep --page raw --csv > gpurun_out/${tag}_v4_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_v4.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_v4_source.csv 2>/dev/null
el "ncu full capture: $(ls -la gpurun_out/${tag}_v4.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
el "cfg2: $(cut -c1-200 gpurun_out/${tag}_bench_cfg2.json)"
timeout 200 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err
el "cfg5: $(cut -c1-200 gpurun_out/${tag}_bench_cfg5.json)"
timeout 300 python bench.py --full-chain --no-cpu-baseline > gpurun_out/${tag}_bench_full_chain.json 2> gpurun_out/${tag}_bench_full_chain.err
el "full chain: $(cut -c1-200 gpurun_out/${tag}_bench_full_chain.json)"
