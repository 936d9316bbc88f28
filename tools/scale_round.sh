#!/bin/bash
# usage (GPU box with >= max(N) GPUs): tools/scale_round.sh <tag> "2 4 8" [steps]
# The driver's scaling run by hand: the default bench line at every N (4K W4 frame in bands across the ranks, p2p band
# stores; `also` = 8K frame tiled + 1080p weak), the NCCL exchange at the largest N, the c5 pose sweep at the largest N.
tag=${1:-r02}; ns=${2:-"2 4 8"}; steps=${3:-100}; mkdir -p gpurun_out
run() {  # N name args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; fi
  echo "$name exit $?"; grep -h "^{" gpurun_out/${tag}_$name.json | python tools/bench_line.py || tail -5 gpurun_out/${tag}_$name.err
}
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
last=1
for n in $ns; do run $n default_n$n --steps $steps --warmup 3; last=$n; done
if [ "$last" != 1 ]; then
  run $last nccl_n$last --steps $steps --warmup 3 --gather nccl
  run $last c5_n$last --steps $steps --warmup 3 --workload c5
fi
