#!/bin/bash
# usage (on the GPU box, through gpurun): tools/gpu_round.sh <tag> [tests]
# One measurement round: GPU tests, the default bench line (+ the reference arm), the ncu launch list of the same
# command and `ncu --set full` captures of the traversal kernels on c3 (4K / W4) and c2 (1080p / W1).
tag=${1:-r02}; mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/${tag}_gputests.log
fi
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_default_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|secondary_kernel|shade_kernel" -c 14 -f -o gpurun_out/${tag}_c3_full \
  python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/${tag}_ncu_c3.log 2>&1; echo "ncu c3 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"primary_kernel" -c 6 -f -o gpurun_out/${tag}_c2_full \
  python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/${tag}_ncu_c2.log 2>&1; echo "ncu c2 exit $?"
ls -la gpurun_out | tail -20
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_default.json", "gpurun_out/${tag}_bench_default_reference.json"):
    for ln in open(f):
        if ln.startswith("{"):
            j = json.loads(ln)
            print(f, "ms", round(j["ms_per_step"], 4), "Grays/s", round(j["value"], 4), "e2e", round(j["e2e"]["value"], 4), j.get("pass_ms"), j.get("clocks"))
            r = j.get("roofline")
            if r: print("  roofline", r["bound"], round(r["achieved"]), round(r["peak"]), round(r["frac"], 3))
            for k, v in j.get("also", {}).items():
                print("  also", k, "ms", round(v["ms_per_step"], 4), "Grays/s", round(v["value"], 3), "e2e", round(v["e2e"]["value"], 3), v.get("pass_ms"), "frac", round(v["roofline"]["frac"], 3))
PY
