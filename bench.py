#!/usr/bin/env python
"""bench.py — the headline measurement of the voxel ray-traversal pass on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl ours|reference]

A "step" is one pass of the hot path over one frame of synthetic input (procgen world, fixed
camera).  Default workload (N=1) is BASELINE.json configs[1]: the default procgen world (W1) at
1920x1080, primary rays only, camera K0; metric = primary-ray throughput in Grays/s.

  value      device-timed throughput, inputs resident in HBM, CUDA events on the launch stream,
             L2 flushed (256 MiB memset) between timed steps; max over ranks.
  e2e        the same metric through the C ABI with HOST buffers: camera uniform block uploaded
             and the RGBA8 result read back into pinned host memory inside the timed region.
  roofline   algorithmic bytes (4*T_in + 4*T_chunk + 4*T_block + 24 B/px, exact
             counters from the counting variant of the same kernel) / measured kernel time.
  cpu_baseline  the CPU oracle (port of the reference GLSL) on this box's host cores, rank 0, N=1.

N > 1 (torchrun): the world is replicated; the default workload shards independent frames across
ranks (weak scaling, no data-path collective).  --workload c4 is the 8K frame tiled across ranks
with the NCCL band gather to rank 0 (strong scaling); --workload c5 the 256-pose sweep.
--impl reference times the reference's CPU implementation (the oracle port — the reference GLSL
cannot run in this image) on the same config.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dim, W, H, shadows, description)
    "c1": (512, 1280, 720, True, "c1: W1 procgen(512) world, camera K0, 1280x720, primary + shadow rays + shade"),
    "c2": (512, 1920, 1080, False, "c2: W1 procgen(512) world, camera K0, 1920x1080, primary rays only"),
    "c3": (2048, 3840, 2160, True, "c3: W4 procgen(2048) world, camera K1, 3840x2160, primary + shadow rays + shade"),
    "c4": (512, 7680, 4320, True, "c4: W1 world, camera K1, 7680x4320 tiled in interleaved 32-row bands across ranks, NCCL gather to rank 0"),
    "c5": (2048, 1920, 1080, False, "c5: W4 world, 256 random poses at 1920x1080 sharded across ranks, primary rays"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--layout", default="compact", choices=["compact", "reference"])
    ap.add_argument("--scheduler", default="tile", choices=["pool", "tile"],
                    help="tile = one pixel per thread (default); pool = per-CTA ray pool compacted between trip phases")
    ap.add_argument("--no-dense", action="store_true", help="traverse through chunk table + bricks instead of the dense block grid")
    ap.add_argument("--fused-frame", action="store_true", help="frame = ONE fused kernel instead of the default primary + secondary + shade launches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between timed steps (reported in config)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="c4 only: p2p = kernels store finished bands straight into rank 0's frame over NVLink (CUDA IPC peer mapping); "
                         "nccl = band buffers gathered with NCCL send/recv + reassembly kernel")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_models():
    return np.load(os.path.join(ROOT, "tests", "golden", "atlas_models.npy"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(workload)
    return None


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path: the oracle port (OpenMP over rows, all host
    threads).  The real reference is GLSL on OpenGL and cannot run in this image."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import oracle
    oracle.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread
    uvt = importlib.import_module("unnamed-voxel-tracer_b200")
    dim, W, H, shadows, desc = WORKLOADS[args.workload]
    if args.workload in ("c3", "c4", "c5"):
        # bounded sample: these frames are minutes of CPU work; time a 1/16-area frame and scale per ray
        scale = 4
    else:
        scale = 1
    Ws, Hs = W // scale, H // scale
    bm = uvt.voxel.VoxelBrickmap.init(dim)
    uvt.procgen.procgen(dim, bm)
    ow = oracle.World(dim, bm.chunks(), bm.bricks(), oracle.atlas_from_models(load_models()))
    cam = uvt.scenes.camera_k0(dim) if args.workload in ("c1", "c2") else uvt.scenes.camera_k1(dim)
    prm = oracle.params(dim)

    def step():
        t0 = time.perf_counter()
        if shadows:
            r = oracle.render(ow, cam, Ws, Hs, prm, want_hits=False)
            rays = Ws * Hs + r["secondary_counters"]["rays"]
        else:
            oracle.primary(ow, cam, Ws, Hs, prm, want_hits=False)
            rays = Ws * Hs
        return time.perf_counter() - t0, rays

    for _ in range(max(1, min(args.warmup, 1))):
        step()
    times, rays = [], 0
    for _ in range(args.steps):
        dt, rays = step()
        times.append(dt)
    total = float(np.sum(times))
    value = rays * args.steps / total / 1e9
    sample = f"{Ws}x{Hs} frame ({'full' if scale == 1 else '1/%d-area sample of' % (scale * scale)} {W}x{H}), {args.steps} steps"
    line = {"impl": "reference", "metric": "rays_per_second", "value": value, "unit": "Grays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3 * (scale * scale), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic (procgen world, reference seeds)",
            "config": {"workload": desc, "rays_per_step": rays * scale * scale, "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": "Grays/s", "cores": oracle.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Grays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU restatement of the reference GLSL (oracle/oracle.c); Mesa llvmpipe / Zig / GL are not available in this image"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    use_dist = world > 1
    torch.cuda.set_device(local_rank)
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    uvt = importlib.import_module("unnamed-voxel-tracer_b200")
    dim, W, H, shadows, desc = WORKLOADS[args.workload]
    models = load_models()

    ctx = uvt.Context(local_rank, map_dim=dim, layout=args.layout, dense=not args.no_dense, fused_frame=args.fused_frame)
    ctx.set_scheduler(args.scheduler)
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)  # launch on a torch stream so torch.cuda events / NCCL ordering see the work
    t0 = time.time()
    uvt.scenes.build_world(ctx, dim, models)
    build_s = time.time() - t0

    tiled = args.workload == "c4"
    sweep = args.workload == "c5"
    band = 32
    if tiled:
        ctx.set_partition(band, world, rank)
    ctx.resize(W, H)
    if sweep:
        poses = uvt.scenes.sweep_poses(dim, 256)
        lo, hi = uvt.tiles.shard_poses(256, world, rank)
        my_poses = poses[lo:hi]
        cam = my_poses[0]
    else:
        cam = uvt.scenes.camera_k0(dim) if args.workload in ("c1", "c2") else uvt.scenes.camera_k1(dim)
    ctx.set_camera(cam)
    ctx.enable_timing(True)

    rpp = uvt.tiles.rows_per_part(H, band, world) if tiled else H
    gather_buf = None
    p2p = tiled and args.gather == "p2p"
    shared_ptr = None
    if tiled:
        gather_buf = torch.zeros((rpp, W), dtype=torch.int32, device=f"cuda:{local_rank}")
        full_frame = torch.empty((H, W), dtype=torch.int32, device=f"cuda:{local_rank}") if rank == 0 else None
        if p2p:
            # the presenting rank owns the frame; every rank's kernels store their bands straight into it
            box = [None]
            if rank == 0:
                shared_ptr, handle = ctx.shared_frame_create()
                box[0] = handle
            if use_dist:
                dist.broadcast_object_list(box, src=0)
            if rank != 0:
                shared_ptr = ctx.shared_frame_open(box[0])
            ctx.bind_frame_target(shared_ptr, global_rows=True)
        else:
            ctx.bind_frame_target(gather_buf.data_ptr(), global_rows=False)

    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    pinned = ctx.pinned_empty(W * ctx.local_rows() * 4, np.uint32)

    def device_step(i):
        """One pass of the hot path, inputs resident.  Returns device ms measured with CUDA events on the launch stream."""
        if sweep:
            ctx.set_camera(my_poses[i % len(my_poses)])
        if shadows:
            ctx.dispatch_frame()
            ms = ctx.last_pass_ms("frame")
        else:
            ctx.dispatch_primary()
            ms = ctx.last_pass_ms("primary")
        return ms

    def gather_step():
        with torch.cuda.stream(stream):
            g = uvt.tiles.gather_bands(gather_buf, 0)
            if rank == 0:
                ctx.deinterleave(g.data_ptr(), full_frame.data_ptr(), rpp)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # exact algorithmic bytes from the counting variant of the same kernels (outside the timed region)
    cp = ctx.count_pass("primary")
    pixels_local = W * (ctx.local_rows() if tiled else H)
    if tiled:
        pixels_local = cp["rays"]
    alg_primary = 4 * (cp["t_in"] + cp["t_chunk"] + cp["t_block"]) + 24 * cp["rays"]
    rays_step = cp["rays"]
    alg_step = alg_primary
    fetch = ctx.fetch_stats("primary") if ctx.effective_layout() == "compact" else None
    if shadows:
        ctx.dispatch_primary()
        cs = ctx.count_pass("secondary")
        alg_step += 4 * (cs["t_in"] + cs["t_chunk"] + cs["t_block"]) + 24 * cs["rays"] + 20 * cs["early_out"] + 32 * cp["rays"]
        rays_step += cs["rays"]

    for i in range(args.warmup):
        device_step(i)
        if tiled and not p2p:
            gather_step()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    kernel_ms = []
    gather_ms = []
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(i & 0xFF)  # evict L2 between timed iterations (untimed)
        ms = device_step(i)
        kernel_ms.append(ms)
        if tiled and not p2p:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gather_step()
            e1.record(stream)
            e1.synchronize()
            gather_ms.append(e0.elapsed_time(e1))
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()

    pass_ms = None
    if shadows:  # per-pass device time of one frame (outside the timed region)
        ctx.dispatch_primary(); ctx.dispatch_secondary(); ctx.shade(); ctx.sync()
        pass_ms = {k: ctx.last_pass_ms(k) for k in ("primary", "secondary", "shade")}

    verified = None
    if p2p:
        # the peer-stored frame must equal the NCCL-gathered one (checked outside the timed region)
        host_p2p = np.empty((H, W), np.uint32)
        if rank == 0:
            ctx.read_device(shared_ptr, host_p2p)
        ctx.bind_frame_target(gather_buf.data_ptr(), global_rows=False)
        device_step(0)
        gather_step()
        barrier()
        if rank == 0:
            verified = bool(np.array_equal(full_frame.cpu().numpy().view(np.uint32), host_p2p))
        ctx.bind_frame_target(shared_ptr, global_rows=True)

    step_ms_local = float(np.sum(kernel_ms) + np.sum(gather_ms))
    t = torch.tensor([step_ms_local], dtype=torch.float64, device=f"cuda:{local_rank}")
    totals = torch.tensor([float(rays_step), float(alg_step)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    rays_all, alg_all = float(totals[0].item()), float(totals[1].item())
    value = rays_all * args.steps / (total_ms * 1e-3) / 1e9

    # ---- e2e: the same metric through the C ABI with host buffers (camera H2D + result D2H inside the timed region)
    cam_host = np.array(cam)
    out_kind = "frame" if shadows else "albedo"
    full_pinned = ctx.pinned_empty(W * H * 4, np.uint32) if (tiled and rank == 0) else None
    pinned2 = [pinned, ctx.pinned_empty(pinned.nbytes, np.uint32)]  # double-buffered host target of the pipelined readback

    def e2e_step(i):
        ctx.set_camera(my_poses[i % len(my_poses)] if sweep else cam_host)
        (ctx.dispatch_frame if shadows else ctx.dispatch_primary)()
        if not tiled:
            # pipelined D2H into pinned host memory: frame i lands while frame i+1 renders (uvt_readback_async)
            ctx.readback_async(out_kind, pinned2[i & 1])
            return
        # tiled frame: the step's result is the assembled frame on the presenting rank
        if p2p:
            ctx.sync()
            if use_dist:
                dist.barrier()
            if rank == 0:
                ctx.read_device(shared_ptr, full_pinned)
        else:
            gather_step()
            if rank == 0:
                ctx.read_device(full_frame.data_ptr(), full_pinned)

    for i in range(2):
        e2e_step(i)
    ctx.readback_wait()
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    ctx.readback_wait()  # every frame of the timed region has landed in host memory
    barrier()
    e2e_s = time.perf_counter() - e0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if use_dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays_all * args.steps / float(te.item()) / 1e9
    d2h_bytes = W * H * 4 if tiled else int(pinned.nbytes)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        kms = float(np.mean(kernel_ms))
        achieved = alg_step / (kms * 1e-3) / 1e9
        l2_gbps = ctx.measure_l2_read_gbps(32 << 20, 50)
        line = {
            "metric": "rays_per_second", "value": value, "unit": "Grays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if tiled else "weak", "vs_baseline": None,
            "dtype": "f32+i32", "data": "synthetic (procgen world, reference seeds; 29 block models from the reference .vox set)",
            "config": {"workload": desc, "rays_per_step_all_ranks": rays_all, "layout": ctx.effective_layout(), "scheduler": args.scheduler, "dense_grid": not args.no_dense,
                       "parallelism": ("interleaved %d-row bands over %d ranks + NCCL gather" % (band, world)) if tiled else
                                      ("poses sharded over %d ranks" % world if sweep else "one independent frame per rank per step, world replicated, no collective"),
                       "l2": "not flushed" if args.no_flush else "flushed between timed steps by a 256 MiB fill (untimed)",
                       "timing": "CUDA events on the launch stream per step, summed; max over ranks", "world_build_s": round(build_s, 2)},
            "e2e": {"value": e2e_value, "unit": "Grays/s", "h2d_bytes_per_step": 96, "d2h_bytes_per_step": d2h_bytes,
                    "what": ("uvt_set_camera + dispatch + (band exchange) + D2H of the assembled RGBA8 frame on rank 0 into pinned host memory, wall clock" if tiled
                             else "per step: uvt_set_camera + dispatch + uvt_readback_async of the RGBA8 %s into pinned host memory (frame i is copied out while frame i+1 renders; all frames landed before the clock stops), wall clock" % out_kind)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": ncu_traffic(args.workload), "peak_source": peak_src, "kernel_ms": kms,
                         "algorithmic_bytes_per_launch": alg_step,
                         "primary_trips": cp["t_in"], "primary_trips_with_fetch": fetch["lookups"] if fetch else cp["t_in"],
                         "note": "algorithmic bytes = reference access pattern (4*T_in+4*T_chunk+4*T_block+G-buffer), exact counters; "
                                 "the traversal data is cache resident, so the binding roofline is L2 (roofline_l2)"},
            "roofline_l2": {"bound": "l2", "achieved": achieved, "peak": l2_gbps, "unit": "GB/s", "frac": achieved / l2_gbps,
                            "peak_source": "measured in this run: 16-B ld.global.cg reads of a 32 MiB L2-resident buffer"},
            "wall_ms_per_step": wall / args.steps * 1e3,
        }
        if pass_ms:
            line["pass_ms"] = pass_ms
        if tiled:
            line["gather_ms"] = float(np.mean(gather_ms)) if gather_ms else 0.0
            line["config"]["gather"] = ("p2p: kernels store bands into rank 0's frame over NVLink (CUDA IPC peer mapping), no collective"
                                        if p2p else "nccl: torch.distributed.gather of band buffers + reassembly kernel")
            if verified is not None:
                line["config"]["p2p_frame_equals_nccl_gather"] = verified
        if args.gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(uvt, args, dim, W, H, shadows, cam)
        print(json.dumps(line))
    if p2p and rank != 0:
        ctx.bind_frame_target(0, global_rows=False)
        ctx.shared_frame_close(shared_ptr)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(uvt, args, dim, W, H, shadows, cam):
    """The oracle (CPU port of the reference GLSL) on this box's host cores: bounded sample of the same workload."""
    import oracle
    oracle.set_num_threads(os.cpu_count() or 1)
    scale = 1 if W * H <= 1920 * 1080 and dim <= 512 else 4
    Ws, Hs = W // scale, H // scale
    bm = uvt.voxel.VoxelBrickmap.init(dim)
    uvt.procgen.procgen(dim, bm)
    ow = oracle.World(dim, bm.chunks(), bm.bricks(), oracle.atlas_from_models(load_models()))
    prm = oracle.params(dim)
    best, rays = None, 0
    for _ in range(3):
        t0 = time.perf_counter()
        if shadows:
            r = oracle.render(ow, cam, Ws, Hs, prm, want_hits=False)
            rays = Ws * Hs + r["secondary_counters"]["rays"]
        else:
            oracle.primary(ow, cam, Ws, Hs, prm, want_hits=False)
            rays = Ws * Hs
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": rays / best / 1e9, "unit": "Grays/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{Ws}x{Hs} frame of the same workload, best of 3, OpenMP over rows", "ms_per_frame_sample": best * 1e3}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
