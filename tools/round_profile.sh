#!/bin/bash
# Measurements of one build on one B200 (run through gpurun); everything lands in gpurun_out/<tag>_*.
#   tools/round_profile.sh <tag> [tests|bench|full|ncu|audit ...]
tag=${1:-r02}; shift
what=${*:-tests bench ncu}
mkdir -p gpurun_out
# gate: a tiny forward + chain against the oracle under a short timeout -- a hung or wrong kernel stops the whole call here
if ! timeout 240 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; then
  echo "SMOKE FAILED / TIMED OUT -- aborting"; tail -5 gpurun_out/${tag}_smoke.log; exit 1
fi
tail -1 gpurun_out/${tag}_smoke.log
for w in $what; do
  case $w in
    tests) timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log ;;
    bench) timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err; cut -c1-330 gpurun_out/${tag}_bench.json ;;
    reference) timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err; cut -c1-300 gpurun_out/${tag}_reference.json ;;
    full) timeout 900 python bench.py --full-chain --no-cpu-baseline > gpurun_out/${tag}_bench_full_chain.json 2> gpurun_out/${tag}_bench_full_chain.err; cut -c1-330 gpurun_out/${tag}_bench_full_chain.json ;;
    cfg2) timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; cut -c1-330 gpurun_out/${tag}_bench_cfg2.json ;;
    cfg5) timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err; cut -c1-330 gpurun_out/${tag}_bench_cfg5.json ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/eager_steps.py 3 > gpurun_out/${tag}_launches.log 2>&1
      python tools/launch_shares.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_shares.csv; head -14 gpurun_out/${tag}_launch_shares.csv
      # the two big launches of a middle layer of the second step: x2h key (softmax epilogue) and x2h value (+ fused aggregation)
      TDIFF_FREE_DEPTH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_mlp_v4_kernel -s 52 -c 2 -o gpurun_out/${tag}_v4 python tools/eager_steps.py 2 > gpurun_out/${tag}_ncu.log 2>&1
      ncu -i gpurun_out/${tag}_v4.ncu-rep --page raw --csv > gpurun_out/${tag}_v4_raw.csv 2>/dev/null
      ncu -i gpurun_out/${tag}_v4.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_v4_source.csv 2>/dev/null
      ncu -i gpurun_out/${tag}_v4.ncu-rep --page source --csv --print-source cuda > gpurun_out/${tag}_v4_source_cuda.csv 2>/dev/null ;;
    aggh)
      timeout 300 ncu --set full --clock-control none -k regex:aggregate_h_kernel -s 3 -c 1 -o gpurun_out/${tag}_agg_h python tools/agg_h_standalone.py > gpurun_out/${tag}_agg_h.log 2>&1
      ncu -i gpurun_out/${tag}_agg_h.ncu-rep --page raw --csv > gpurun_out/${tag}_agg_h_raw.csv 2>/dev/null
      timeout 120 python tools/agg_h_standalone.py | tail -1 ;;
    scale2)
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${tag}_scale_2gpu.json 2> gpurun_out/${tag}_scale_2gpu.err; cut -c1-330 gpurun_out/${tag}_scale_2gpu.json; tail -3 gpurun_out/${tag}_scale_2gpu.err
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/${tag}_scale_2gpu_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_scale_2gpu_reference.json ;;
    audit) timeout 900 python tools/precision_audit.py --out gpurun_out/${tag}_precision_audit.json > gpurun_out/${tag}_audit.log 2>&1; tail -8 gpurun_out/${tag}_audit.log ;;
  esac
done
