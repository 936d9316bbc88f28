#!/usr/bin/env python
"""Frame time of ONE GPU's share of a tiled frame (uvt_set_partition(16, P, 0)) against the number of row chunks
(uvt_set_frame_chunks): does running the chunks on two streams hide the pass tails?

    python tools/chunk_bench.py [--workload c3] [--parts 1 2 4 8] [--chunks 1 2 3 4 6]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

uvt = importlib.import_module("unnamed-voxel-tracer_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--parts", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--chunks", type=int, nargs="+", default=[1, 2, 3, 4, 6])
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    dim, W, H, shadows, desc = bench.WORKLOADS[args.workload]
    with uvt.Context(0, map_dim=dim) as ctx:
        uvt.scenes.build_world(ctx, dim, bench.load_models())
        ctx.enable_timing(True)
        cam = uvt.scenes.camera_k0(dim) if args.workload in ("c1", "c2") else uvt.scenes.camera_k1(dim)
        out = {}
        for parts in args.parts:
            ctx.set_partition(16, parts, 0)
            ctx.resize(W, H)
            ctx.set_camera(cam)
            ref = None
            for n in args.chunks:
                ctx.set_frame_chunks(n)
                ts = []
                for i in range(args.steps + 5):
                    ctx.dispatch_frame()
                    ctx.sync()
                    if i >= 5:
                        ts.append(ctx.last_pass_ms("frame"))
                frame = ctx.readback("frame")
                if ref is None:
                    ref = frame
                assert np.array_equal(frame, ref), (parts, n)
                out[f"P{parts}_chunks{n}"] = round(float(np.median(ts)), 4)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
