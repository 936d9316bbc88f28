"""Summarise bench.py JSON lines from stdin (one per line)."""
import json
import sys

for ln in sys.stdin:
    if not ln.startswith("{"):
        continue
    j = json.loads(ln)
    print(f"  N={j['n_gpus']} {j['config']['workload'][:2]} ms {j['ms_per_step']:.4f} Grays/s {j['value']:.3f} e2e {j['e2e']['value']:.3f} ({j['e2e'].get('ms_per_step', 0):.3f} ms) "
          f"{j['scaling']} pass_ms {j.get('pass_ms')} clocks {j.get('clocks', {}).get('sm_mhz')} {j.get('clocks', {}).get('reasons')}")
    if "nvlink" in j:
        print("    nvlink", {k: v for k, v in j["nvlink"].items() if k not in ("source", "note")})
    for k in ("frame_equals_torch_gather",):
        if k in j["config"]:
            print("    ", k, j["config"][k], "host frame ok:", j["e2e"].get("host_frame_equals_device_frame"), "d2h GB/s per rank", j["e2e"].get("d2h_gb_per_s_per_rank"))
    for k, v in j.get("also", {}).items():
        print(f"    also {k}: ms {v['ms_per_step']:.4f} Grays/s {v['value']:.3f} e2e {v['e2e']['value']:.3f} {v.get('scaling')} eq {v.get('frame_equals_torch_gather')} "
              f"nvlink {({kk: vv for kk, vv in v['nvlink'].items() if kk in ('rank0_rx_over_expected', 'unavailable')}) if 'nvlink' in v else None}")
