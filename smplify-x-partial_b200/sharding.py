"""Multi-GPU plumbing: frames are independent, so a batch is sharded in contiguous blocks over
the ranks (one process per GPU, ``torch.distributed``) with no communication during the fit;
the only exchange is one all-gather of the fitted per-frame parameters at the end
(SURVEY.md section 8e).  Works with NCCL on GPUs and with gloo on CPU (tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises the default process group from RANK / WORLD_SIZE / LOCAL_RANK (torchrun).
    Returns (rank, world_size); a single process needs no group."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend)
    return rank, world


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()


def shard_range(n_frames, rank, world):
    """Contiguous block of frames of ``rank``: sizes differ by at most one, earlier ranks take
    the larger blocks."""
    base, extra = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def gather_frames(local, n_frames=None):
    """All-gather of a per-frame tensor [b_rank, ...] -> [sum b_rank, ...] in frame order on
    every rank.  Equal shards use one ``all_gather_into_tensor``; ragged shards are padded to
    the largest shard and trimmed."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    if n_frames is None:
        sizes = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes)
        counts = [int(s.item()) for s in all_sizes]
    else:
        counts = [len(shard_range(n_frames, r, world)) for r in range(world)]
    m = max(counts)
    padded = local
    if local.shape[0] < m:
        pad = torch.zeros((m - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    if all(c == m for c in counts):
        return out
    return torch.cat([out[r * m:r * m + counts[r]] for r in range(world)], dim=0)
