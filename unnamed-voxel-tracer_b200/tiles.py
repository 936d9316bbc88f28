"""Multi-GPU partition of the path: rays are independent, the world is replicated,
so a frame is cut into interleaved row bands (balances sky vs ground rays) and camera batches are
sharded by pose.  The only exchange is the gather of finished bands to the presenting rank.

Host-side logic only; it works with any torch.distributed backend (NCCL on the GPU box, gloo in
the CPU tests).
"""
import numpy as np


def n_bands(H, band_rows):
    return (H + band_rows - 1) // band_rows


def storage_rows(H, band_rows, n_parts, part):
    """Rows a part stores (band-granular, like uvt_local_rows)."""
    if n_parts == 1:
        return H
    return len(range(part, n_bands(H, band_rows), n_parts)) * band_rows


def rows_per_part(H, band_rows, n_parts):
    """Uniform per-part row count used for the gather (the largest part)."""
    return storage_rows(H, band_rows, n_parts, 0)


def local_to_global_rows(H, band_rows, n_parts, part, rows=None):
    """global image row of every local storage row, -1 for padding."""
    rows = storage_rows(H, band_rows, n_parts, part) if rows is None else rows
    ly = np.arange(rows)
    if n_parts == 1:
        return np.where(ly < H, ly, -1)
    lb = ly // band_rows
    y = (lb * n_parts + part) * band_rows + ly % band_rows
    return np.where(y < H, y, -1)


def assemble(gathered, H, band_rows):
    """gathered: [n_parts, rows_per_part, W] array-like (numpy or torch) -> [H, W] frame."""
    n_parts, rpp = gathered.shape[0], gathered.shape[1]
    src_part = np.empty(H, np.int64)
    src_row = np.empty(H, np.int64)
    for part in range(n_parts):
        g = local_to_global_rows(H, band_rows, n_parts, part, rpp)
        ok = g >= 0
        src_part[g[ok]] = part
        src_row[g[ok]] = np.nonzero(ok)[0]
    if isinstance(gathered, np.ndarray):
        return gathered[src_part, src_row]
    import torch
    return gathered[torch.as_tensor(src_part, device=gathered.device), torch.as_tensor(src_row, device=gathered.device)]


def gather_bands(local, dst=0, group=None):
    """Gather every rank's [rows_per_part, W] band buffer to `dst`; returns [n_parts, rows_per_part, W] there, None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local.unsqueeze(0)  # single process: nothing to gather
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return local.unsqueeze(0)
    if rank == dst:
        out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        dist.gather(local, list(out.unbind(0)), dst=dst, group=group)
        return out
    dist.gather(local, None, dst=dst, group=group)
    return None


def shard_poses(n_poses, world, rank):
    """Contiguous pose shards, sizes differing by at most one."""
    base, extra = divmod(n_poses, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class SharedHostFrame:
    """One W x H RGBA8 frame in POSIX shared memory, mapped by every rank of a tiled frame.  Each rank page-locks its
    mapping (Context.host_register) and copies its own bands into it (Context.readback_bands_async): the frame is
    assembled in host memory by N parallel device-to-host copies."""

    def __init__(self, name, W, H, create):
        from multiprocessing import shared_memory
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=W * H * 4) if create else shared_memory.SharedMemory(name=name)
        if not create:  # an attachment must not be unlinked by this process's resource tracker (Python < 3.13 registers it)
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray((H, W), dtype=np.uint32, buffer=self.shm.buf)
        self.owner = create

    def close(self):
        self.array = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except (FileNotFoundError, BufferError):
            pass
