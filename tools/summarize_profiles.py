"""Turn gpurun_out/ ncu artefacts into the committed summaries under profiles/.

  python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep> [workload]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
workload = sys.argv[4] if len(sys.argv) > 4 else "c2"
out_dir = os.path.join(ROOT, "profiles")

# ---- launch list: per-kernel device time and share of the captured command
rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
per = {}
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    if "<" in r[4]:
        name = r[4].split("(uvt::WorldArgs")[0].replace("void ", "")
    unit, val = r[13], float(r[14].replace(",", ""))
    us = val / 1000.0 if unit in ("nsecond", "ns") else val * (1.0 if unit in ("usecond", "us") else 1000.0)
    per.setdefault(name, []).append(us)
total = sum(sum(v) for v in per.values())
with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 4 --warmup 3 --no-cpu-baseline` (workload {workload})\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: compare SHARES.\n\n")
    f.write("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {100*sum(v)/total:.1f}% |\n")
    f.write("\nNotes: `primary_kernel<World, 0, ...>` is the timed step; `<World, 1, ...>` / `<World, 2, ...>` are the one-off counting variants run "
            "outside the timed region; `FillFunctor` is the untimed 256 MiB L2 flush between timed steps; `l2_read_kernel` is the in-run L2 peak probe; "
            "the `field_*`, `count_virtual`, `build_chunks2`, `repack_bricks`, `brick_rowmask`, `column_tops`, `quad_clear`, `clearance`, `dense_fill` "
            "kernels build the B200 layout once at `uvt_world_commit`.\n")

# ---- full capture of the dominant kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active"]
r = data[-1]
vals = {w: (r[hdr.index(w)], units[hdr.index(w)]) for w in want if w in hdr}


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


dram = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
with open(os.path.join(out_dir, f"{tag}_primary_ncu_full.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none` of the dominant kernel (workload {workload})\n\n| metric | value | unit |\n|---|---|---|\n")
    for k, (v, u) in vals.items():
        f.write(f"| {k} | {v} | {u} |\n")
    stalls = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    f.write("\nWarp stall reasons (warps stalled per issue-active cycle, top 8):\n\n")
    for h, v in sorted(stalls, key=lambda kv: -kv[1])[:8]:
        f.write(f"* {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}\n")
    f.write(f"\nDRAM traffic per launch: {dram/1e6:.2f} MB (read + write).\n")
tj = os.path.join(out_dir, "traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t[workload] = dram
json.dump(t, open(tj, "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launches.md")).read())
print(open(os.path.join(out_dir, f"{tag}_primary_ncu_full.md")).read())
