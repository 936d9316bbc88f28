#!/usr/bin/env python
"""bench.py — the headline measurement of the voxel ray-traversal pass on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl ours|reference]

A "step" is one pass of the hot path over one frame of synthetic input (procgen world, fixed
camera).  Default workload is the configuration BASELINE.json's metric ("Grays/s and 4K frame ms") is
quoted on, configs[2]: the 4x-scaled procgen world (W4) at 3840x2160, primary + shadow rays + shade,
camera K1; metric = rays per second (primary + shadow), ms_per_step = the 4K frame time.  The same
JSON line carries `also`: the 8K tiled frame (configs[3]) and the 1080p primary-only frame (configs[1]).

  value      device-timed throughput, inputs resident in HBM, CUDA events on the launch stream,
             L2 flushed (256 MiB memset) between timed steps; max over ranks.
  e2e        the same metric through the C ABI with HOST buffers: camera uniform block uploaded
             and the RGBA8 result read back into pinned host memory inside the timed region.
  roofline   algorithmic bytes (4*T_in + 4*T_chunk + 4*T_block + 24 B/px, exact
             counters from the counting variant of the same kernel) / measured kernel time.
  cpu_baseline  the CPU oracle (port of the reference GLSL) on this box's host cores, rank 0, N=1.

N > 1 (torchrun): the world is replicated and the SAME frame is cut into interleaved 16-row bands
across the ranks (strong scaling); the band exchange is inside the timed region: every rank's shade
kernel stores its finished pixels straight into the presenting rank's frame over NVLink (--gather
p2p, default) or the bands are gathered with NCCL (--gather nccl).  c2 at N > 1 runs one independent
frame per rank (weak, no collective) and c5 shards the 256 poses.
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
    "c3": (2048, 3840, 2160, True, "c3: W4 procgen(2048) world, camera K1, 3840x2160, primary + shadow rays + shade (the 4K frame)"),
    "c4": (512, 7680, 4320, True, "c4: W1 world, camera K1, 7680x4320 tiled in interleaved row bands across ranks, bands assembled on rank 0"),
    "c5": (2048, 1920, 1080, False, "c5: W4 world, 256 random poses at 1920x1080 sharded across ranks, primary rays"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-also", action="store_true", help="skip the `also` sub-records (8K tiled frame, 1080p primary-only frame)")
    ap.add_argument("--layout", default="compact", choices=["compact", "reference"])
    ap.add_argument("--scheduler", default="tile", choices=["pool", "tile"],
                    help="tile = one pixel per thread (default); pool = per-CTA ray pool compacted between trip phases")
    ap.add_argument("--no-dense", action="store_true", help="traverse through chunk table + bricks instead of the dense block grid")
    ap.add_argument("--fused-frame", action="store_true", help="frame = ONE fused kernel instead of the default primary + secondary + shade launches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between timed steps (reported in config)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="tiled frames: p2p = kernels store finished bands straight into rank 0's frame over NVLink (CUDA IPC peer mapping); "
                         "nccl = uvt_dispatch_frame_nccl: grouped ncclSend/ncclRecv per band group on a second stream, overlapped with the traversal of the next group")
    ap.add_argument("--band-rows", type=int, default=0, help="rows per band of a tiled frame (a multiple of 16); default 16 for the p2p band stores (keeps the ranks' row counts "
                    "within 1 %% of each other at 4K / 8 ranks), 32 for the NCCL exchange (half as many send/recv pairs)")
    ap.add_argument("--frame-chunks", type=int, default=0, help="uvt_set_frame_chunks for the timed frames (0 = 2 for tiled p2p frames, 1 otherwise)")
    ap.add_argument("--nccl-groups", type=int, default=4, help="band groups per rank of the NCCL exchange (1 = exchange after the whole frame)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  Sampled through NVML
    from a thread of this process every 2 ms — a timed region of a few tens of ms still gets samples, which an
    `nvidia-smi -lms` child (slow to start) does not guarantee; nvidia-smi is the fallback when NVML is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []   # (t, sm_mhz, reason mask)
        self.sm_max = None
        self.t0 = self.t1 = None
        self.stop_flag = False
        self.thread = None
        self.source = None
        self.proc = None
        self.lines = []

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            reasons_fn(h)

            def loop():
                while not self.stop_flag:
                    try:
                        self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons_fn(h))))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi"
        except OSError:
            self.proc = None

    def _read(self):
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                mask = 0
                for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        mask |= names[n]
                self.samples.append((time.perf_counter(), float(f[1]), mask))
                self.sm_max = max(self.sm_max or 0.0, float(f[2]))
            except ValueError:
                continue

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        time.sleep(0.01)
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e300)]
        window = "timed region"
        if not inside:  # a region shorter than the sampling period: the samples of the whole run under load (warm-up included)
            inside, window = list(self.samples), "whole run (timed region shorter than the sampling period)"
        mask = 0
        for x in inside:
            mask |= x[2]
        sm = [x[1] for x in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(n for b, n in self.REASONS.items() if mask & b),
                "samples": len(sm), "window": window, "source": self.source}


class NvlinkCounters:
    """NVLink data bytes received / sent by every GPU of the box (NVML field values
    NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX / _TX, KiB, summed over the links): read before and after the timed region
    on rank 0, this is the counter evidence for what the band exchange moves."""

    def __init__(self, n):
        self.n, self.ok, self.err = n, False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(n)]
            self.read()
            self.ok = True
        except Exception as e:  # no NVML / no NVLink counters in this container: reported, not fatal
            self.err = f"{type(e).__name__}: {e}"

    def read(self):
        nv = self.nv
        out = []
        for h in self.handles:
            vals = nv.nvmlDeviceGetFieldValues(h, [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF), (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF)])
            row = []
            for v in vals:
                if v.nvmlReturn != 0:
                    raise RuntimeError(f"NVML field {v.fieldId}: return {v.nvmlReturn}")
                row.append(int(v.value.ullVal) * 1024)
            out.append(row)
        return out

    @staticmethod
    def delta(a, b):
        return [[y - x for x, y in zip(ra, rb)] for ra, rb in zip(a, b)]


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
def reference_cpu():
    """The reference's CPU implementation of the path: its own shader text compiled for the host (oracle/_ref/libglslref.so,
    kind "reference") when that library is here, else the oracle port (kind "port").  Returns (kind, render(world, cam, W, H, shadows)
    -> rays traced, threads, note)."""
    import oracle
    n = os.cpu_count() or 1
    try:
        from oracle import glslref
        ok = glslref.available()
    except Exception:
        ok = False
    if ok:
        glslref.lib()
        glslref.set_num_threads(n)  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread

        def render(world, cam, W, H, shadows):
            if shadows:
                r = glslref.render(world, cam, W, H)
                return W * H + glslref.shadow_rays(r["position"])
            glslref.primary(world, cam, W, H)
            return W * H
        return "reference", render, glslref.num_threads(), ("the reference's own GLSL (assets/shaders/primary.comp, secondary.comp, blit.fragment + includes) compiled for the "
                                                            "CPU against a GLSL-in-C++ shim (oracle/glsl_ref), OpenMP over work groups; the GL program itself cannot run in this image")
    oracle.set_num_threads(n)
    prm_cache = {}

    def render(world, cam, W, H, shadows):
        prm = prm_cache.setdefault(world.dim, oracle.params(world.dim))
        if shadows:
            r = oracle.render(world, cam, W, H, prm, want_hits=False)
            return W * H + r["secondary_counters"]["rays"]
        oracle.primary(world, cam, W, H, prm, want_hits=False)
        return W * H
    return "port", render, oracle.num_threads(), "CPU restatement of the reference GLSL (oracle/oracle.c); oracle/_ref is not built here"


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores, all threads, on a bounded sample of the
    same workload.  The real reference is GLSL on OpenGL and cannot run in this image; its shader text can (oracle/_ref)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import oracle
    kind, render, threads, note = reference_cpu()
    uvt = importlib.import_module("unnamed-voxel-tracer_b200")
    dim, W, H, shadows, desc = WORKLOADS[args.workload]
    # bounded sample: the 4K / 8K / sweep frames are seconds of CPU work each; time a 1/16-area frame and scale per ray
    scale = 4 if args.workload in ("c3", "c4", "c5") else 1
    Ws, Hs = W // scale, H // scale
    bm = uvt.voxel.VoxelBrickmap.init(dim)
    uvt.procgen.procgen(dim, bm)
    ow = oracle.World(dim, bm.chunks(), bm.bricks(), oracle.atlas_from_models(load_models()))
    cam = uvt.scenes.camera_k0(dim) if args.workload in ("c1", "c2") else uvt.scenes.camera_k1(dim)

    def step():
        t0 = time.perf_counter()
        rays = render(ow, cam, Ws, Hs, shadows)
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
            "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3 * (scale * scale), "higher_is_better": True,
            "scaling": "strong" if (args.gpus > 1 and args.workload in ("c1", "c3", "c4")) else "weak",
            "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic (procgen world, reference seeds)",
            "config": {"workload": desc, "rays_per_step": rays * scale * scale, "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": "Grays/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Grays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "note": note}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing (NCCL): barriers, max / sum over ranks, object broadcast."""

    def __init__(self, torch, rank, local_rank, world):
        self.torch, self.rank, self.local_rank, self.world = torch, rank, local_rank, world
        self.on = world > 1
        self.dev = f"cuda:{local_rank}"
        if self.on:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.on:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.on:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def bcast(self, obj):
        box = [obj]
        if self.on:
            self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather_obj(self, obj):
        if not self.on:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.on:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure(uvt, torch, D, args, workload, steps, warmup, models, world_cache, want_cpu_baseline):
    """Build the workload on this rank, time `steps` passes of the hot path on the device, then the same through the
    C ABI with host buffers (e2e).  Returns the record on rank 0 (None elsewhere)."""
    rank, local_rank, world = D.rank, D.local_rank, D.world
    dim, W, H, shadows, desc = WORKLOADS[workload]
    sweep = workload == "c5"
    tiled = world > 1 and workload in ("c1", "c3", "c4")   # one frame cut into bands across the ranks (strong scaling)
    band = args.band_rows or (16 if args.gather == "p2p" else 32)

    ctx = uvt.Context(local_rank, map_dim=dim, layout=args.layout, dense=not args.no_dense, fused_frame=args.fused_frame)
    ctx.set_scheduler(args.scheduler)
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)  # launch on a torch stream so torch.cuda events / NCCL ordering see the work
    t0 = time.time()
    bm, _ = uvt.scenes.build_world(ctx, dim, models)
    build_s = time.time() - t0
    if tiled:
        ctx.set_partition(band, world, rank)
    ctx.resize(W, H)
    if sweep:
        poses = uvt.scenes.sweep_poses(dim, 256)
        lo, hi = uvt.tiles.shard_poses(256, world, rank)
        my_poses = poses[lo:hi]
        cam = my_poses[0]
    else:
        cam = uvt.scenes.camera_k0(dim) if workload in ("c1", "c2") else uvt.scenes.camera_k1(dim)
    ctx.set_camera(cam)
    ctx.enable_timing(True)

    rpp = uvt.tiles.rows_per_part(H, band, world) if tiled else H
    p2p = tiled and args.gather == "p2p"
    gather_buf = full_frame = shared_ptr = None
    if tiled:
        gather_buf = torch.zeros((rpp, W), dtype=torch.int32, device=D.dev)
        full_frame = torch.empty((H, W), dtype=torch.int32, device=D.dev) if rank == 0 else None
        if p2p:
            # the presenting rank owns the frame; every rank's shade kernel stores its bands straight into it over NVLink
            handle = None
            if rank == 0:
                shared_ptr, handle = ctx.shared_frame_create()
            handle = D.bcast(handle)
            if rank != 0:
                shared_ptr = ctx.shared_frame_open(handle)
            ctx.bind_frame_target(shared_ptr, global_rows=True)
        else:
            ctx.bind_frame_target(gather_buf.data_ptr(), global_rows=False)
            ctx.nccl_init(D.bcast(uvt.Context.nccl_unique_id() if rank == 0 else None), world, rank)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=D.dev)

    # tiled frames: the rank's rows in two chunks on two streams, so that the tail of one chunk's pass runs under the other's work
    # (uvt_set_frame_chunks; -12 % at 1/8 of a 4K frame per GPU).  N = 1 keeps whole-frame launches: every pass has its own time.
    chunks = args.frame_chunks if args.frame_chunks else (2 if (tiled and p2p) else 1)
    ctx.set_frame_chunks(chunks)

    def device_step(i):
        """One pass of the hot path, inputs resident.  Device ms from CUDA events on the launch stream."""
        if sweep:
            ctx.set_camera(my_poses[i % len(my_poses)])
        if tiled and not p2p:
            # render in band groups; each group's ncclSend/ncclRecv runs on a second stream under the next group's traversal
            ctx.dispatch_frame_nccl(full_frame.data_ptr() if rank == 0 else None, args.nccl_groups)
            return ctx.last_pass_ms("frame"), None
        if shadows:
            ctx.dispatch_frame()
            return ctx.last_pass_ms("frame"), (None if chunks > 1 else ctx.last_pass_ms("primary"))
        ctx.dispatch_primary()
        ms = ctx.last_pass_ms("primary")
        return ms, ms

    def gather_step():
        with torch.cuda.stream(stream):
            g = uvt.tiles.gather_bands(gather_buf, 0)
            if rank == 0:
                ctx.deinterleave(g.data_ptr(), full_frame.data_ptr(), rpp)

    # exact algorithmic bytes from the counting variant of the same kernels (outside the timed region)
    cp = ctx.count_pass("primary")
    alg_primary = 4 * (cp["t_in"] + cp["t_chunk"] + cp["t_block"]) + 24 * cp["rays"]
    rays_step, alg_step = cp["rays"], alg_primary
    fetch = ctx.fetch_stats("primary") if ctx.effective_layout() == "compact" else None
    if shadows:
        ctx.dispatch_primary()
        cs = ctx.count_pass("secondary")
        alg_step += 4 * (cs["t_in"] + cs["t_chunk"] + cs["t_block"]) + 24 * cs["rays"] + 20 * cs["early_out"] + 32 * cp["rays"]
        rays_step += cs["rays"]

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(warmup):
        device_step(i)
    D.barrier()

    nvl = NvlinkCounters(world) if (tiled and rank == 0) else None
    launches0 = ctx.launch_count()
    step_ms, primary_ms = [], []
    D.barrier()
    nvl0 = nvl.read() if (nvl and nvl.ok) else None
    sampler.begin()
    wall0 = time.perf_counter()
    for i in range(steps):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(i & 0xFF)  # evict L2 between timed iterations (untimed)
        ms, pm = device_step(i)    # p2p: the band stores to rank 0 happen inside the shade kernel; nccl: inside the bracket of the call
        step_ms.append(ms)
        primary_ms.append(pm)
    D.barrier()
    wall = time.perf_counter() - wall0
    sampler.end()
    nvl1 = nvl.read() if (nvl and nvl.ok) else None
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()

    pass_ms = None
    chunked_ms = None
    if shadows:  # per-pass device time, averaged over the timed steps is not kept per pass: take one more (whole-frame) frame
        ctx.set_frame_chunks(1)
        ctx.dispatch_frame(); ctx.sync()
        pass_ms = {k: ctx.last_pass_ms(k) for k in ("primary", "secondary", "shade")}
        if chunks == 1 and not (tiled and not p2p):  # what uvt_set_frame_chunks(2) would give (outside the timed region, L2 warm on both sides)
            t1, t2 = [], []
            for n_, acc in ((1, t1), (2, t2)):
                ctx.set_frame_chunks(n_)
                for _ in range(12):
                    ctx.dispatch_frame(); ctx.sync()
                    acc.append(ctx.last_pass_ms("frame"))
            chunked_ms = {"whole_frame_launches": float(np.median(t1[2:])), "two_row_chunks_on_two_streams": float(np.median(t2[2:]))}
        ctx.set_frame_chunks(chunks)

    verified = None
    if tiled:
        # the frame assembled inside the timed region (peer stores / NCCL send-recv) must equal the one gathered with
        # torch.distributed.gather + the reassembly kernel (checked outside the timed region)
        host_p2p = np.empty((H, W), np.uint32)
        D.barrier()
        if rank == 0:
            ctx.read_device(shared_ptr if p2p else full_frame.data_ptr(), host_p2p)
        ctx.bind_frame_target(gather_buf.data_ptr(), global_rows=False)
        ctx.dispatch_frame()
        gather_step()
        D.barrier()
        if rank == 0:
            verified = bool(np.array_equal(full_frame.cpu().numpy().view(np.uint32), host_p2p))
        if p2p:
            ctx.bind_frame_target(shared_ptr, global_rows=True)

    # a step ends when its slowest rank ends: per-step max over ranks, then the sum over the timed steps
    per_step_max = D.reduce(step_ms, "max")
    total_ms = float(np.sum(per_step_max))
    rays_all, alg_all = D.reduce([float(rays_step), float(alg_step)], "sum")
    value = rays_all * steps / (total_ms * 1e-3) / 1e9

    # ---- e2e: the same metric through the C ABI with HOST buffers: camera H2D + the step's result D2H inside the timed region
    cam_host = np.array(cam)
    out_kind = "frame" if shadows else "albedo"
    host_frame = shm = None
    if tiled:
        # every rank copies ITS bands straight into one shared, page-locked host frame: N PCIe links in parallel
        name = D.bcast("uvt_frame_%d_%s" % (os.getpid(), workload) if rank == 0 else None)
        if rank == 0:
            shm = uvt.tiles.SharedHostFrame(name, W, H, create=True)
        D.barrier()
        if rank != 0:
            shm = uvt.tiles.SharedHostFrame(name, W, H, create=False)
        host_frame = shm.array
        ctx.host_register(host_frame)
        ctx.bind_frame_target(0, global_rows=False)   # bands stay in this rank's own frame buffer
    else:
        pinned2 = [ctx.pinned_empty(W * H * 4, np.uint32), ctx.pinned_empty(W * H * 4, np.uint32)]

    def e2e_step(i):
        ctx.set_camera(my_poses[i % len(my_poses)] if sweep else cam_host)
        (ctx.dispatch_frame if shadows else ctx.dispatch_primary)()
        if tiled:
            ctx.readback_bands_async(host_frame)
        else:
            ctx.readback_async(out_kind, pinned2[i & 1])  # frame i lands while frame i+1 renders

    for i in range(2):
        e2e_step(i)
    ctx.readback_wait()
    D.barrier()
    e0 = time.perf_counter()
    for i in range(steps):
        e2e_step(i)
    ctx.readback_wait()  # every frame of the timed region has landed in host memory
    D.barrier()
    e2e_s = time.perf_counter() - e0
    e2e_s = D.reduce([e2e_s], "max")[0]
    e2e_value = rays_all * steps / e2e_s / 1e9
    d2h_local = W * ctx.local_rows() * 4
    e2e_ok = None
    if tiled:
        # the frame assembled in host memory by the N copies equals the frame assembled on rank 0's GPU
        D.barrier()
        if rank == 0:
            e2e_ok = bool(np.array_equal(host_frame, host_p2p))
    d2h_rates = D.gather_obj(round(d2h_local * steps / e2e_s / 1e9, 2))
    all_clocks = D.gather_obj(clocks)

    rec = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        kms = float(np.mean(primary_ms)) if primary_ms[0] is not None else float(pass_ms["primary"])
        achieved = alg_primary / (kms * 1e-3) / 1e9
        l2_gbps = ctx.measure_l2_read_gbps(32 << 20, 50)
        sm = [c["sm_mhz"] for c in all_clocks if c.get("sm_mhz")]
        reasons = sorted({r for c in all_clocks for r in c.get("reasons", [])})
        rec = {
            "metric": "rays_per_second", "value": value, "unit": "Grays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong" if tiled else "weak", "vs_baseline": None,
            "dtype": "f32+i32", "data": "synthetic (procgen world, reference seeds; 29 block models from the reference .vox set)",
            "config": {"workload": desc, "rays_per_step_all_ranks": rays_all, "layout": ctx.effective_layout(), "scheduler": args.scheduler, "dense_grid": not args.no_dense,
                       "parallelism": ("one %dx%d frame in interleaved %d-row bands over %d ranks, world replicated" % (W, H, band, world)) if tiled else
                                      ("poses sharded over %d ranks" % world if sweep else
                                       ("one independent frame per rank per step, world replicated, no collective" if world > 1 else "single GPU")),
                       "l2": "not flushed" if args.no_flush else "flushed between timed steps by a 256 MiB fill (untimed)",
                       "timing": "CUDA events on the launch stream per step; per step the max over ranks, summed over the steps", "world_build_s": round(build_s, 2)},
            "e2e": {"value": e2e_value, "unit": "Grays/s", "h2d_bytes_per_step": 96 * world, "d2h_bytes_per_step": W * H * 4 if tiled else d2h_local * world,
                    "ms_per_step": e2e_s / steps * 1e3, "d2h_gb_per_s_per_rank": d2h_rates,
                    "what": ("per step and rank: uvt_set_camera + dispatch + uvt_readback_bands_async of the rank's RGBA8 bands into ONE shared page-locked host frame "
                             "(N device-to-host copies in parallel, no GPU-to-GPU hop); all frames landed before the clock stops; wall clock, max over ranks" if tiled else
                             "per step: uvt_set_camera + dispatch + uvt_readback_async of the RGBA8 %s (4 B/px; the 24 B/px G-buffer and the 28 B/px hit buffer stay on the device) into pinned host "
                             "memory, frame i copied out while frame i+1 renders; all frames landed before the clock stops; wall clock" % out_kind)},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max([c.get("sm_max_mhz") or 0 for c in all_clocks]) or None,
                       "reasons": reasons, "samples": int(sum(c.get("samples", 0) for c in all_clocks)),
                       "per_rank_sm_mhz": [c.get("sm_mhz") for c in all_clocks], "window": all_clocks[0].get("window"), "source": all_clocks[0].get("source")},
            "roofline": {"bound": "l2", "kernel": "primary_kernel (the dominant launch of the step)", "achieved": achieved, "peak": l2_gbps, "unit": "GB/s",
                         "frac": achieved / l2_gbps, "traffic": ncu_traffic(workload),
                         "peak_source": "L2 read bandwidth measured in this run (16-B ld.global.cg over a 32 MiB resident buffer); MEASURED_PEAKS.json has no L2 figure",
                         "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_primary,
                         "hbm_frac": achieved / peaks["hbm_gbs"], "hbm_peak": peaks["hbm_gbs"], "hbm_peak_source": peak_src,
                         "primary_trips": cp["t_in"], "primary_trips_with_fetch": fetch["lookups"] if fetch else cp["t_in"],
                         "note": "algorithmic bytes = reference access pattern of this rank's primary pass (4*T_in + 4*T_chunk + 4*T_block + 24 B/px), exact counters; "
                                 "the kernel itself is bound by instruction issue (profiles/), the traversal data is cache resident"},
            "wall_ms_per_step": wall / steps * 1e3,
        }
        if pass_ms:
            rec["pass_ms"] = pass_ms
        rec["config"]["frame_chunks"] = chunks
        if chunked_ms:
            rec["frame_chunks_ms"] = chunked_ms
        if tiled:
            rec["config"]["exchange"] = ("p2p: every rank's shade kernel stores its finished bands into rank 0's frame over NVLink (CUDA IPC peer mapping), inside the timed region"
                                         if p2p else "nccl: uvt_dispatch_frame_nccl, %d band groups per rank; each group's bands go to their rows of rank 0's frame by grouped "
                                                     "ncclSend/ncclRecv on a second stream while the next group is traversed; inside the timed region" % args.nccl_groups)
            if verified is not None:
                rec["config"]["frame_equals_torch_gather"] = verified
            if nvl is not None:
                if nvl0 is not None:
                    d = NvlinkCounters.delta(nvl0, nvl1)
                    expect = W * H * 4 * (world - 1) / world
                    rec["nvlink"] = {"source": "NVML NVLINK_THROUGHPUT_DATA_RX/TX field values (KiB granularity), all links of each GPU, read on rank 0 around the timed steps",
                                     "rx_bytes_per_step": [round(r[0] / steps) for r in d], "tx_bytes_per_step": [round(r[1] / steps) for r in d],
                                     "expected_rank0_rx_bytes_per_step": round(expect),
                                     "rank0_rx_over_expected": round(d[0][0] / steps / expect, 4) if expect else None,
                                     "note": "expected = the (N-1)/N of the RGBA8 frame that the other ranks' bands hold (payload); the counters also see whatever else crosses NVLink in the timed region (the per-step barrier-free loop has no other traffic; NCCL adds its protocol)"}
                else:
                    rec["nvlink"] = {"unavailable": nvl.err, "evidence": "profiles/r02_nvlink_p2p.md: ncu nvltx__bytes_data_user of every sender's shade_kernel == its band bytes"}
            if e2e_ok is not None:
                rec["e2e"]["host_frame_equals_device_frame"] = e2e_ok
        if want_cpu_baseline:
            rec["cpu_baseline"] = cpu_baseline(uvt, bm, dim, W, H, shadows, cam, models)

    if tiled:
        ctx.readback_wait()
        ctx.host_unregister(host_frame)
        host_frame = None
        D.barrier()
        shm.close()
    if p2p and rank != 0:
        ctx.bind_frame_target(0, global_rows=False)
        ctx.shared_frame_close(shared_ptr)
    D.barrier()
    ctx.close()
    del flush, gather_buf, full_frame
    torch.cuda.empty_cache()
    if rec is not None and workload == "c2" and world == 1:
        rec["e2e"]["hit_buffer"] = e2e_hit_buffer(uvt, local_rank, dim, W, H, cam, models, min(steps, 20))
    return rec


def e2e_hit_buffer(uvt, device, dim, W, H, cam, models, steps):
    """The same end-to-end loop when the step's result is the explicit 28 B/px HIT BUFFER (the north-star parity artefact:
    hit voxel, face, material, distance, trips) instead of the 4 B/px image: the primary pass writes it (HITBUF variant)
    and every frame's records are copied to pinned host memory.  PCIe-bound by construction."""
    with uvt.Context(device, map_dim=dim, hit_buffer=True) as ctx:
        uvt.scenes.build_world(ctx, dim, models)
        ctx.resize(W, H)
        ctx.set_camera(cam)
        nbytes = W * H * 28
        pinned = [ctx.pinned_empty(nbytes, np.uint8), ctx.pinned_empty(nbytes, np.uint8)]

        def step(i):
            ctx.set_camera(cam)
            ctx.dispatch_primary()
            ctx.readback_async("hit", pinned[i & 1])
        for i in range(2):
            step(i)
        ctx.readback_wait()
        t0 = time.perf_counter()
        for i in range(steps):
            step(i)
        ctx.readback_wait()
        dt = time.perf_counter() - t0
    return {"value": W * H * steps / dt / 1e9, "unit": "Grays/s", "d2h_bytes_per_step": nbytes, "ms_per_step": dt / steps * 1e3, "steps": steps,
            "d2h_gb_per_s": round(nbytes * steps / dt / 1e9, 2),
            "what": "uvt_set_camera + primary pass with the hit buffer on + uvt_readback_async of the 28 B/px hit records into pinned host memory, pipelined like e2e.value"}


def run_ours(args):
    import torch
    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    D = Dist(torch, rank, local_rank, world)
    uvt = importlib.import_module("unnamed-voxel-tracer_b200")
    models = load_models()
    line = measure(uvt, torch, D, args, args.workload, args.steps, args.warmup, models, None,
                   want_cpu_baseline=(args.gpus == 1 and not args.no_cpu_baseline))
    if not args.no_also and args.workload == "c3":
        also = {}
        for wl, st in (("c4", max(30, min(args.steps, 60))), ("c2", args.steps)):
            r = measure(uvt, torch, D, args, wl, st, args.warmup, models, None, want_cpu_baseline=False)
            if r:
                also[wl] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "e2e", "roofline", "gpu_launches", "nvlink") if k in r}
                also[wl]["workload"] = r["config"]["workload"]
                also[wl]["parallelism"] = r["config"]["parallelism"]
                if "pass_ms" in r:
                    also[wl]["pass_ms"] = r["pass_ms"]
                for k in ("exchange", "frame_equals_torch_gather"):
                    if k in r["config"]:
                        also[wl][k] = r["config"][k]
        if line:
            line["also"] = also
    if line:
        print(json.dumps(line))
    D.close()


def cpu_baseline(uvt, bm, dim, W, H, shadows, cam, models):
    """The reference's CPU implementation (its own shader text compiled for the host when oracle/_ref is here, else the
    oracle port) on this box's host cores: a bounded sample of the same workload."""
    import oracle
    kind, render, threads, note = reference_cpu()
    scale = 1 if W * H <= 1920 * 1080 and dim <= 512 else 4
    Ws, Hs = W // scale, H // scale
    ow = oracle.World(dim, bm.chunks(), bm.bricks(), oracle.atlas_from_models(models))
    best, rays = None, 0
    for _ in range(3):
        t0 = time.perf_counter()
        rays = render(ow, cam, Ws, Hs, shadows)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": rays / best / 1e9, "unit": "Grays/s", "cores": threads, "kind": kind,
            "sample": f"{Ws}x{Hs} frame of the same workload ({'full size' if scale == 1 else '1/%d of the pixels' % (scale * scale)}), best of 3, OpenMP",
            "ms_per_frame_sample": best * 1e3, "what": note}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
