#!/bin/bash
# A/B of kernel builds in one gpurun call: for every libtdiff_<name>.so given, the smoke parity gate and the driver's bench command
# (device-timed part only); then the in-situ wait accounting of the instrumented builds.
#   gpurun --timeout 420 -- 'bash tools/variant_ab.sh r02c ref "" tile64 lean prefirst leanpre t64lean'
tag=$1; shift
mkdir -p gpurun_out
t0=$(date +%s)
for v in "$@"; do
  lib=$PWD/targetdiff_b200/libtdiff${v:+_$v}.so
  name=${v:-default}
  if ! TDIFF_LIB=$lib timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_${name}_smoke.log 2>&1; then
    echo "[$(( $(date +%s) - t0 )) s] $name: SMOKE FAILED: $(tail -2 gpurun_out/${tag}_${name}_smoke.log | tr '\n' ' ')"; continue
  fi
  TDIFF_LIB=$lib timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_${name}_bench.json 2> gpurun_out/${tag}_${name}_bench.err
  echo "[$(( $(date +%s) - t0 )) s] $name: $(python -c "import json,sys; d=json.load(open('gpurun_out/${tag}_${name}_bench.json')); print('%.3f ms/step  %.2f mol/s  edge-MLP %.3f ms/layer  clocks %s' % (d['ms_per_step'], d['value'], d['roofline']['ms_per_layer'], d['clocks']['sm_mhz']))" 2>&1 | tail -1)"
done
for v in waitstats wspre; do
  [ -f targetdiff_b200/libtdiff_$v.so ] || continue
  TDIFF_LIB=$PWD/targetdiff_b200/libtdiff_$v.so timeout 150 python tools/wait_stats.py 5 > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  echo "[$(( $(date +%s) - t0 )) s] $v: $(python -c "import json; d=json.load(open('gpurun_out/${tag}_${v}.json')); d.pop('raw'); print(json.dumps(d))" 2>&1 | tail -1)"
done
