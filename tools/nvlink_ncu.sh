#!/bin/bash
# usage (box with N GPUs): tools/nvlink_ncu.sh <tag> [N=2]
# Counter evidence for the peer-to-peer band stores: ONE process (uvt_group, 2 members) renders the 4K frame; ncu profiles
# only device 1 (--devices 1: the member that does not present) and reads the NVLink / fabric counters of its shade_kernel,
# whose stores go straight into member 0's frame.
tag=${1:-r02}; n=${2:-2}; devs=$(seq -s, 1 $((n-1))); mkdir -p gpurun_out
ncu --devices 1 --query-metrics 2>/dev/null | grep -i -E "nvl|fabric|pcie|peer" > gpurun_out/${tag}_nvlink_metrics_available.txt
wc -l gpurun_out/${tag}_nvlink_metrics_available.txt
M=$(grep -o -E "^(nvltx__bytes|nvlrx__bytes|nvltx__bytes_data_user|nvlrx__bytes_data_user|lts__t_sectors_srcunit_ltcfabric|lts__t_sectors_srcunit_ltcfabric_op_write|lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_st|lts__t_sectors_aperture_peer|lts__t_sectors_aperture_peer_op_write|lts__t_sectors_aperture_peer_op_read|l1tex__m_l1tex2xbar_write_bytes|pcie__write_bytes|pcie__read_bytes)\b" gpurun_out/${tag}_nvlink_metrics_available.txt | sort -u | sed 's/$/.sum/' | paste -sd, -)
echo "metrics: $M"
timeout 900 ncu --devices $devs --clock-control none -k regex:"shade_kernel" -c $((2*(n-1))) --metrics gpu__time_duration.sum,lts__t_sectors_op_write.sum,$M \
  --csv --log-file gpurun_out/${tag}_nvlink_ncu.csv python tools/group_bench.py --gpus $n --workload c3 --steps 1 --warmup 1 > gpurun_out/${tag}_nvlink_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${tag}_nvlink_ncu.log
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${tag}_nvlink_ncu.csv")) if len(r) > 14 and r[0].isdigit()]
for r in rows:
    if r[12].startswith(("nvl", "gpu__time")): print(r[0], "device", r[9], r[4][:24], r[12], r[13], r[14])
PY
