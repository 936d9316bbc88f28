#!/usr/bin/env python
"""Latency of publishing world edits: full uvt_world_commit against the incremental
uvt_world_commit_region that VoxelBrickmap.bind() issues for a dirty box.  Prints one JSON line per world.

    python tools/edit_latency.py [--dims 512 2048]
"""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
uvt = importlib.import_module("unnamed-voxel-tracer_b200")


def timed(fn, n):
    ts = []
    for i in range(n):
        t0 = time.perf_counter()
        fn(i)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs="+", default=[512])
    args = ap.parse_args()
    V = uvt.voxel.Voxel
    for dim in args.dims:
        with uvt.Context(0, map_dim=dim) as ctx:
            bm = uvt.voxel.VoxelBrickmap.init(dim, 8, ctx)
            uvt.procgen.procgen(dim, bm)
            bm.bind(9)
            mid = dim // 2
            top = max(y for y in range(dim) if bm.get(mid, y, mid)) + 1

            def full(i):
                bm.mark_dirty()
                bm.bind(9)

            def in_brick(i):  # toggles a surface block: the chunk table does not change
                bm.set(mid + i % 5, top - 1, mid, V(11, True) if i % 2 else 0)
                bm.bind(9)

            def new_brick(i):  # a block in empty air: new brick, new virtual bricks, new chunk distances
                bm.set(mid + 8 * (i % 8), min(top + 40 + 8 * (i // 8), dim - 1), mid, V(11, True))
                bm.bind(9)

            def clean(i):
                bm.bind(9)

            out = {"world": f"procgen({dim})", "n_bricks": bm.n_bricks,
                   "full_commit_ms": round(timed(full, 3), 3),
                   "edit_in_existing_brick_ms": round(timed(in_brick, 20), 3),
                   "edit_new_brick_ms": round(timed(new_brick, 16), 3),
                   "bind_clean_ms": round(timed(clean, 20), 4)}
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
