"""Per-source-region and per-opcode instruction mix of one kernel launch from an ncu report
(`ncu --set full --import-source on`, code compiled with -lineinfo).

  python tools/instr_mix.py <report.ncu-rep> <launch-index> [--lines] [--md out.md --title "..."]

Warp-level "Instructions Executed" of every SASS instruction is summed (a) by the source line ncu
attributes it to, grouped into the regions of the traversal named in REGIONS (matched by file and
by marker text, so the table survives line shifts), and (b) by SASS opcode.
"""
import csv
import subprocess
import sys
from collections import defaultdict


def split_row(line, n):
    """ncu does not escape quotes inside the Source column: split on '","' and fold the surplus back into column 1."""
    f = line.strip()
    if f.startswith('"'):
        f = f[1:]
    if f.endswith('"'):
        f = f[:-1]
    parts = f.split('","')
    if n and len(parts) > n:
        extra = len(parts) - n
        parts = [parts[0], '","'.join(parts[1:2 + extra])] + parts[2 + extra:]
    return parts


def load(rep, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch),
                          "--launch-count", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    files, cur, hdr, kernel, seen = {}, None, None, None, set()
    for line in out.splitlines():
        r = split_row(line, len(hdr) if hdr else 0)
        if not r or r == [""]:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            files.setdefault(cur, {"lines": {}, "sass": []})
            continue
        if r[0] == "Function Name":
            kernel = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if cur is None or hdr is None or len(r) < len(hdr):
            continue
        ix = hdr.index("Instructions Executed")
        if r[0] != "":  # a source line with the totals of its SASS
            try:
                files[cur]["lines"][int(r[0])] = (r[1], int(r[ix]), int(r[hdr.index("# Samples")] or 0))
                files[cur]["last"] = int(r[0])
            except ValueError:
                pass
        elif r[2].startswith("0x") and (cur, files[cur].get("last", 0), r[2]) not in seen:  # the listing may repeat: count an address once per line
            seen.add((cur, files[cur].get("last", 0), r[2]))
            try:
                files[cur]["sass"].append((files[cur].get("last", 0), r[2], r[3].strip(), int(r[ix])))
            except ValueError:
                pass
    return kernel, files


# (region name, file, first-line marker, last-line marker) — markers are substrings of source lines
REGIONS = [
    ("ray set-up (primary_ray, AABB clip, zero patch)", "kernels.cuh", "__device__ __forceinline__ void intersect_aabb", "const float t_far_unused_marker"),
]


def region_of(fname, line, text, bounds):
    for name, f, lo, hi in bounds:
        if f == fname and lo <= line <= hi:
            return name
    return f"other ({fname})"


def find_line(path, marker, start=1):
    with open(path) as fh:
        for i, ln in enumerate(fh, 1):
            if i >= start and marker in ln:
                return i
    raise SystemExit(f"marker not found in {path}: {marker}")


def main():
    import os
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep, launch = args[0], int(args[1])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tr = os.path.join(root, "unnamed-voxel-tracer_b200", "csrc", "trace.cuh")
    ke = os.path.join(root, "unnamed-voxel-tracer_b200", "csrc", "kernels.cuh")
    L = lambda m, s=1: find_line(tr, m, s)
    fast = L("trace_map_fast(const WorldCompact &w")
    loop = L("for (;;) {", fast)
    short_lo = L("DENSE && COUNT != 1 && slow) {", loop)
    gen_lo = L("if (slow) {", short_lo)
    walk = L("column-tops walk: many free trips at once", gen_lo)
    vote = L("how many trips can the whole warp run", walk)
    steps = L("k DDA steps, branch-free", vote)
    seg = L("if (trip >= max_steps) {", steps)
    end = L("if (COUNT == 1 && fast) tc.t_in = out.trips;", seg)
    bounds = [
        ("generic verbatim loop (trace_map; NaN / huge-origin lanes only)", "trace.cuh", L("__device__ __forceinline__ void trace_map("), L("conservative \"nothing ahead\" test") - 1),
        ("sky_sealed (pooled scheduler only)", "trace.cuh", L("conservative \"nothing ahead\" test"), L("free trips from the column tops") - 1),
        ("column-tops walk (line_free_trips: free runs + sealing)", "trace.cuh", L("free trips from the column tops"), L("// One DDA step (map.glsl:157-162)") - 1),
        ("DDA step arithmetic (dda_step / dda_step_last, inline PTX)", "trace.cuh", L("// One DDA step (map.glsl:157-162)"), L("// ---- traceMap, B200 fast path") - 1),
        ("traversal set-up (reciprocals, grid/within, phase constants)", "trace.cuh", fast, loop),
        ("lookup: short common path (dense grid, empty block)", "trace.cuh", short_lo, gen_lo - 1),
        ("lookup: general path (materials, sub-voxels, carries, faces, seals)", "trace.cuh", gen_lo, walk - 1),
        ("column-tops walk: call site", "trace.cuh", walk, vote - 1),
        ("round bookkeeping (warp min vote, run length)", "trace.cuh", vote, steps - 1),
        ("DDA run loop control + minIdx report", "trace.cuh", steps, seg - 1),
        ("iteration cap", "trace.cuh", seg, end),
    ]
    kernel, files = load(rep, launch)
    per = defaultdict(int)
    total = 0
    for fname, d in files.items():
        for line, (text, n, _) in d["lines"].items():
            name = region_of(fname, line, text, bounds)
            if fname == "kernels.cuh":
                name = "kernel body: ray generation, AABB clip, sky colour, G-buffer stores (kernels.cuh)"
            per[name] += n
            total += n
    ops = defaultdict(int)
    addr_seen = set()
    for fname, d in files.items():
        for _, addr, sass, n in d["sass"]:
            if addr in addr_seen:
                continue
            addr_seen.add(addr)
            t = sass.split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
            ops[op.split(".")[0]] += n
    lines = []
    lines.append(f"kernel: `{kernel}`\n")
    lines.append(f"warp-level instructions executed (sum over SASS): **{total/1e6:.1f} M**\n")
    lines.append("| source region | M warp-instr | share |\n|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
        lines.append(f"| {k} | {v/1e6:.1f} | {100*v/total:.1f}% |")
    lines.append("\n| SASS opcode | M warp-instr | share |\n|---|---|---|")
    tot_ops = sum(ops.values())
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:28]:
        lines.append(f"| {k} | {v/1e6:.1f} | {100*v/tot_ops:.1f}% |")
    text = "\n".join(lines)
    print(text)
    if "--lines" in sys.argv:
        for fname, d in files.items():
            for line, (src, n, smp) in sorted(d["lines"].items()):
                if n > total * 0.002:
                    print(f"{fname}:{line:4d} {n/1e6:8.2f} M  samples {smp:6d}  {src.strip()[:110]}")
    if "--md" in sys.argv:
        out = sys.argv[sys.argv.index("--md") + 1]
        title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else rep
        with open(out, "w") as f:
            f.write(f"# {title}\n\n`python tools/instr_mix.py` over the `ncu --set full --import-source on` report: `Instructions Executed` of every SASS "
                    "instruction, summed by the source line ncu attributes it to (grouped into the regions of `trace_map_fast`) and by opcode.\n\n" + text + "\n")


if __name__ == "__main__":
    main()
