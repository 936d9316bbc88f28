"""Turn the artefacts of tools/gpu_round.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/.

  python tools/summarize_round.py <tag>

Writes profiles/<tag>_launches.{csv,md}, <tag>_{c3,c2}_ncu_full.md (last launch of every traversal kernel of the
`ncu --set full` captures), <tag>_instr_mix_*.md (tools/instr_mix.py), copies the bench JSON lines and refreshes
profiles/traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py for roofline.traffic).
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        # per-pipe instruction issue (share of each pipe's peak): where the issue slots go
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active"]


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def short(name):
    return name.replace("void ", "").split("(")[0]


def launches():
    src = os.path.join(G, f"{tag}_launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(P, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    per = {}
    for r in rows:
        unit, val = r[13], float(r[14].replace(",", ""))
        us = val / 1000.0 if unit in ("nsecond", "ns") else val * (1.0 if unit in ("usecond", "us") else 1000.0)
        per.setdefault(short(r[4]), []).append(us)
    total = sum(sum(v) for v in per.values())
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 4 --warmup 3 --no-cpu-baseline` (default workload: c3 = 4K / W4, then the `also` records c4 and c2)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: compare SHARES.\n\n")
        f.write("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {100*sum(v)/total:.1f}% |\n")
        f.write("\nNotes: `primary_kernel<World, 0, ...>` / `secondary_kernel<World, 0>` / `shade_kernel` are the timed step; the `<World, 1, ...>` / `<World, 2, ...>` "
                "variants are the one-off exact-counter runs outside the timed region; `FillFunctor` is the untimed 256 MiB L2 flush between timed steps; "
                "`l2_read_kernel` is the in-run L2 peak probe; the `pg_*` kernels are the device procgen; every other kernel builds the B200 layout once at "
                "`uvt_world_commit`.\n")
    print(open(os.path.join(P, f"{tag}_launches.md")).read())


def full(wl):
    rep = os.path.join(G, f"{tag}_{wl}_full.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]
    last = {}
    for i, r in enumerate(data):
        k = short(r[hdr.index("Kernel Name")])
        if k.startswith("primary_kernel") or k.startswith("secondary_kernel"):
            if k.split("<")[1].split(",")[1].strip().rstrip(">") != "0":
                continue  # counting variants
        last[k] = (i, r)
    names = sorted(last, key=lambda k: last[k][0])
    with open(os.path.join(P, f"{tag}_{wl}_ncu_full.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the frame kernels, workload {wl}\n\n")
        f.write("Last captured launch of each kernel (launch index in the report in brackets).\n\n")
        f.write("| metric | " + " | ".join(f"`{k}` [{last[k][0]}]" for k in names) + " | unit |\n|---|" + "---|" * (len(names) + 1) + "\n")
        for w in WANT:
            if w not in hdr:
                continue
            j = hdr.index(w)
            f.write(f"| {w} | " + " | ".join(last[k][1][j] for k in names) + f" | {units[j]} |\n")
        for k in names:
            r = last[k][1]
            stalls = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
            f.write(f"\nWarp stall reasons of `{k}` (warps stalled per issue-active cycle, top 6): " +
                    ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for h, v in sorted(stalls, key=lambda kv: -kv[1])[:6]) + "\n")
        dom = [k for k in names if k.startswith("primary_kernel")][0]
        r = last[dom][1]
        dram = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")]) + \
            to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        f.write(f"\nDRAM traffic per launch of the dominant kernel (`{dom}`): {dram/1e6:.2f} MB (read + write).\n")
    tj = os.path.join(P, "traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    t[wl] = dram
    json.dump(t, open(tj, "w"), indent=1)
    print(open(os.path.join(P, f"{tag}_{wl}_ncu_full.md")).read())
    # instruction mix per source region
    for k in names:
        if k.startswith("shade"):
            continue
        kind = "primary" if k.startswith("primary") else "secondary"
        out = os.path.join(P, f"{tag}_instr_mix_{wl}_{kind}.md")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "instr_mix.py"), rep, str(last[k][0]), "--md", out, "--title",
                        f"{tag}: instruction mix of `{k}`, workload {wl}"], stdout=subprocess.DEVNULL)
        print(open(out).read())


launches()
full("c3")
full("c2")
for n in ("bench_default.json", "bench_default_reference.json"):
    src = os.path.join(G, f"{tag}_{n}")
    if os.path.exists(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(P, f"{tag}_{n}"))
