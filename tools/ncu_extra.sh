mkdir -p gpurun_out
for w in c4 c1; do
timeout 600 ncu --set full --clock-control none -k regex:"primary_kernel" -c 4 -f -o gpurun_out/r02g_${w}_full python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r02g_ncu_$w.log 2>&1; echo "ncu $w exit $?"
done
