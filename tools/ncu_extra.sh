#!/bin/bash
# usage (GPU box): tools/ncu_extra.sh <tag>  — `ncu --set full` of the primary kernel on the workloads gpu_round.sh does not capture (c4, c1);
# their DRAM bytes go into profiles/traffic.json (bench.py's roofline.traffic)
tag=${1:-r02}; mkdir -p gpurun_out
for w in c4 c1; do
  timeout 600 ncu --set full --clock-control none -k regex:"primary_kernel" -c 4 -f -o gpurun_out/${tag}_${w}_full python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/${tag}_ncu_$w.log 2>&1; echo "ncu $w exit $?"
done
