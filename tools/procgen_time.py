#!/usr/bin/env python
"""Wall-clock of world generation: serial host procgen vs the device procgen (plan + fill incl. the copy back to the host staging)
and the commit that follows.  python tools/procgen_time.py [--dims 512 2048]"""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
uvt = importlib.import_module("unnamed-voxel-tracer_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs="+", default=[512, 2048])
args = ap.parse_args()
out = {}
for dim in args.dims:
    rec = {}
    for mode in ("device", "host", "device"):
        with uvt.Context(0, map_dim=dim) as ctx:
            bm = uvt.voxel.VoxelBrickmap.init(dim, 8, ctx)
            t0 = time.perf_counter()
            uvt.procgen.procgen(dim, bm, device=mode)
            t1 = time.perf_counter()
            bm.bind(9)
            t2 = time.perf_counter()
            rec[mode] = {"procgen_ms": round((t1 - t0) * 1e3, 2), "commit_ms": round((t2 - t1) * 1e3, 2), "n_bricks": bm.n_bricks}
    out[str(dim)] = rec
print(json.dumps(out))
