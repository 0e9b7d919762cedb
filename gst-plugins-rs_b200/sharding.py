"""Frame-parallel sharding across the GPUs of one box.

Frames are independent, so frame i goes to rank i mod G and no data-path collective is
needed; torch.distributed is used only to line ranks up (barrier) and to take the maximum of
the per-rank device timings.  Works on any backend (nccl on GPUs, gloo in the CPU tests)."""
import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def frames_for_rank(n_frames, rank, world_size):
    """Round-robin: indices of the frames rank `rank` processes, in stream order."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_frames, world_size))


def owner_of(frame_index, world_size):
    return frame_index % world_size


def row_bands(height, world_size):
    """If ONE frame has to be split instead: contiguous row bands, one per rank,
    sizes differing by at most one row; returns [(first_row, n_rows)] — still no halo, no
    exchange (every output pixel depends on one input pixel)."""
    base, extra = divmod(height, world_size)
    bands, row = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        bands.append((row, n))
        row += n
    return bands


def merge_in_order(per_rank_results, n_frames):
    """Inverse of frames_for_rank: per_rank_results[r][k] is the result of frame r + k*G."""
    g = len(per_rank_results)
    out = [None] * n_frames
    for r, res in enumerate(per_rank_results):
        idx = frames_for_rank(n_frames, r, g)
        if len(res) != len(idx):
            raise ValueError(f"rank {r}: expected {len(idx)} results, got {len(res)}")
        for i, v in zip(idx, res):
            out[i] = v
    return out


def init_process_group(backend, device=None):
    """init_process_group on 127.0.0.1 defaults; returns True when a group was created."""
    import torch.distributed as dist
    rank, _, size = world()
    if size <= 1:
        return False
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    kwargs = {"device_id": device} if (device is not None and backend == "nccl") else {}
    dist.init_process_group(backend, rank=rank, world_size=size, **kwargs)
    return True


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """max of a float over all ranks (the multi-GPU time of a step is its slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
