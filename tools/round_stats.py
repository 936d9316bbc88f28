#!/usr/bin/env python
"""Lockstep-round statistics of the traversal kernels (experiment build -DUVT_ROUND_STATS, loaded through UVT_LIB_PATH):
rounds per warp and how many rounds advance a single trip because one lane needs a lookup every trip.

    UVT_LIB_PATH=variants/libuvt_stats.so python tools/round_stats.py [--workload c2]
"""
import argparse
import ctypes
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

uvt = importlib.import_module("unnamed-voxel-tracer_b200")
N = uvt._native


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    args = ap.parse_args()
    dim, W, H, shadows, desc = bench.WORKLOADS[args.workload]
    ctx = uvt.Context(0, map_dim=dim)
    uvt.scenes.build_world(ctx, dim, bench.load_models())
    ctx.resize(W, H)
    ctx.set_camera(uvt.scenes.camera_k0(dim) if args.workload in ("c1", "c2") else uvt.scenes.camera_k1(dim))
    ctx.dispatch_primary()
    ctx.dispatch_secondary()
    ctx.sync()
    out = {}
    for name, which in (("primary", 2), ("secondary", 3)):
        c = N.Counters()
        ctx.check(ctx.L.uvt_count_pass(ctx.handle, which, ctypes.byref(c)))
        d = c.as_dict()
        out[name] = {"rays": d["rays"], "rounds": d["t_in"], "single_trip_rounds_sub_voxel_lane": d["t_chunk"],
                     "single_trip_rounds_block_lane": d["t_block"], "hits": d["hits"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
