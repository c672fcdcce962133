#!/bin/bash
# BASELINE configs[3] on one 8-GPU box (run through `gpurun --gpus 8`): the 100-pocket sweep through the product CLI, both schedules.
#   tools/cfg4_run.sh <tag> [pockets] [samples] [steps]
tag=${1:-r02}; pockets=${2:-100}; samples=${3:-100}; steps=${4:-1000}
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
if ! timeout 240 python __graft_entry__.py smoke > gpurun_out/${tag}_cfg4_smoke.log 2>&1; then echo "smoke failed"; tail -3 gpurun_out/${tag}_cfg4_smoke.log; exit 1; fi
for sched in round_robin longest_first; do
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 tools/cfg4_sweep.py \
    --pockets $pockets --samples $samples --steps $steps --schedule $sched --out gpurun_out/${tag}_cfg4_${sched}.json > gpurun_out/${tag}_cfg4_${sched}.log 2>&1
  tail -2 gpurun_out/${tag}_cfg4_${sched}.log | cut -c1-600
done
