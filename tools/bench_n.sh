#!/bin/bash
# usage: bench_n.sh N tag [bench args...]   — one torchrun bench line, summarised
N=$1; tag=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err || tail -20 gpurun_out/$tag.err
python - <<PY
import json
for ln in open("gpurun_out/$tag.json"):
    if not ln.startswith("{"): continue
    j = json.loads(ln)
    print("$tag", "N=$N", j["config"]["workload"][:3], "ms", round(j["ms_per_step"], 4), "Grays/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3),
          "e2e_ms", round(j["e2e"]["ms_per_step"], 4), {k: v for k, v in j["config"].items() if k in ("frame_equals_torch_gather",)}, j["e2e"].get("host_frame_equals_device_frame"), j.get("pass_ms"))
    for k, v in j.get("also", {}).items():
        print("   also", k, "ms", round(v["ms_per_step"], 4), "Grays/s", round(v["value"], 3), "e2e", round(v["e2e"]["value"], 3), v.get("frame_equals_torch_gather"))
PY
