import sys, os, time, ctypes, importlib
sys.path.insert(0, os.getcwd())
uvt = importlib.import_module("unnamed-voxel-tracer_b200")
N = uvt._native; L = N.load()
for dim in (512, 2048):
    with uvt.Context(0, map_dim=dim) as ctx:
        bm = uvt.voxel.VoxelBrickmap.init(dim, 8, ctx)
        for rep in range(2):
            n = ctypes.c_size_t()
            t0 = time.perf_counter()
            rc = L.uvt_world_procgen_plan(ctx.handle, 0.0, 0.0, ctypes.byref(n)); assert rc == 0
            t1 = time.perf_counter()
            print(dim, "plan ms", round((t1 - t0) * 1e3, 2), "n", n.value)
        cap = bm.capacity
        t1 = time.perf_counter()
        while cap < n.value:
            cap *= 2
            p = ctypes.c_void_p()
            rc = L.uvt_world_grow(ctx.handle, cap, ctypes.byref(p)); assert rc == 0
        t2 = time.perf_counter()
        print(dim, "grow ms", round((t2 - t1) * 1e3, 2), "cap", cap)
        rc = L.uvt_world_procgen_fill(ctx.handle); assert rc == 0, ctx.L.uvt_last_error(ctx.handle)
        t3 = time.perf_counter()
        print(dim, "fill ms", round((t3 - t2) * 1e3, 2))
