#!/bin/bash
# Last-minute A/B of one variant build against the shipped one; the full GPU test suite runs on the variant only if it is faster.
#   gpurun --timeout 170 -- 'bash tools/ab_then_test.sh r02f packlog'
tag=$1; v=$2
mkdir -p gpurun_out
lib=$PWD/targetdiff_b200/libtdiff_$v.so
t0=$(date +%s)
TDIFF_LIB=$lib timeout 60 python __graft_entry__.py smoke > gpurun_out/${tag}_${v}_smoke.log 2>&1 || { echo "SMOKE FAILED: $(tail -2 gpurun_out/${tag}_${v}_smoke.log)"; exit 1; }
TDIFF_LIB=$lib timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_${v}_bench.json 2> gpurun_out/${tag}_${v}_bench.err
timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_default_bench.json 2> gpurun_out/${tag}_default_bench.err
python - <<PY
import json
a=json.load(open('gpurun_out/${tag}_${v}_bench.json')); b=json.load(open('gpurun_out/${tag}_default_bench.json'))
print('[%d s] %s %.3f ms/step (edge MLP %.3f ms/layer) vs shipped %.3f (%.3f)' % ($(date +%s) - $t0, '$v', a['ms_per_step'], a['roofline']['ms_per_layer'], b['ms_per_step'], b['roofline']['ms_per_layer']))
open('gpurun_out/${tag}_faster','w').write('1' if a['ms_per_step'] < 0.997*b['ms_per_step'] else '0')
PY
if [ "$(cat gpurun_out/${tag}_faster)" = "1" ]; then
  TDIFF_LIB=$lib timeout 100 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_${v}_pytest_gpu.log 2>&1
  echo "[$(( $(date +%s) - t0 )) s] pytest on $v: $(tail -1 gpurun_out/${tag}_${v}_pytest_gpu.log)"
fi
