#!/usr/bin/env python
"""Single-process multi-GPU frame time (uvt_group): the C4 8K frame (and any other workload size) rendered by all
visible GPUs of one process, device time from CUDA events on member 0 bracketed by group syncs.

    python tools/group_bench.py [--gpus N] [--workload c4] [--steps 30]
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

uvt = importlib.import_module("unnamed-voxel-tracer_b200")


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    n = args.gpus or torch.cuda.device_count()
    dim, W, H, shadows, desc = bench.WORKLOADS[args.workload]
    models = bench.load_models()
    with uvt.Group(list(range(n)), map_dim=dim) as grp:
        bm = uvt.voxel.VoxelBrickmap.init(dim, 8, grp)
        uvt.procgen.procgen(dim, bm)
        atlas = uvt.voxel.VoxelModelAtlas.init(grp)
        for m in models:
            atlas.append_model(m)
        bm.bind(9)
        grp.resize(W, H)
        grp.set_camera(uvt.scenes.camera_k0(dim) if args.workload in ("c1", "c2") else uvt.scenes.camera_k1(dim))
        for _ in range(args.warmup):
            grp.dispatch_frame()
        grp.sync()
        ts = []
        for _ in range(args.steps):
            grp.sync()
            t0 = time.perf_counter()
            grp.dispatch_frame()
            grp.sync()
            ts.append((time.perf_counter() - t0) * 1e3)
        frame = grp.readback_frame()
        ts.sort()
        rays = W * H * (2 if shadows else 1)
        med = ts[len(ts) // 2]
        print(json.dumps({"workload": desc, "n_gpus": n, "mode": "single process, uvt_group, peer-to-peer band stores",
                          "host_ms_per_frame_median": round(med, 4), "host_ms_per_frame_min": round(ts[0], 4),
                          "grays_per_s": round(rays / med / 1e6, 3), "frame_checksum": int(frame.astype(np.uint64).sum())}))


if __name__ == "__main__":
    main()
