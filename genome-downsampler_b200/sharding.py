"""Multi-GPU host logic: the path shards only across independent units (samples / contigs /
zero-coverage components), so each rank solves its own block of samples with no data-path
collective, and the per-sample kept bitmaps are gathered at the end (NCCL all-gather over
NVLink on GPUs, gloo in the CPU tests).  SURVEY.md §8(e).

`solve_fn(sample_ids) -> (bitmap_words_tensor, words_per_sample)` is supplied by the caller: the
CUDA solver in bench.py, a stand-in in the gloo tests.  This module never imports the oracle.
"""
import torch
import torch.distributed as dist


def shard_samples(n_samples, world_size, rank):
    """Static block partition: rank r owns samples [r*n/W, (r+1)*n/W) (equal-size samples need no
    balancing; unequal ones should be LPT-ordered by reads x length before calling this)."""
    lo = (n_samples * rank) // world_size
    hi = (n_samples * (rank + 1)) // world_size
    return list(range(lo, hi))


def bitmap_words(n_reads):
    """Per-sample bitmap size in 32-bit words, padded to 16 bytes so gathered slices stay aligned."""
    return ((n_reads + 31) // 32 + 3) // 4 * 4


def gather_bitmaps(local_bitmaps, group=None):
    """local_bitmaps: int32 tensor [n_local_samples, words] on this rank's device.  Returns the
    [n_total_samples, words] tensor of every rank's bitmaps in rank order (all ranks get it)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_bitmaps
    world = dist.get_world_size(group)
    out = torch.empty((world * local_bitmaps.shape[0], local_bitmaps.shape[1]),
                      dtype=local_bitmaps.dtype, device=local_bitmaps.device)
    dist.all_gather_into_tensor(out, local_bitmaps.contiguous(), group=group)
    return out


def gather_scalars(values, device, group=None):
    """values: list of python ints for this rank's samples (F*, n_kept, ...).  -> [world*len] int64."""
    t = torch.tensor(values, dtype=torch.int64, device=device)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out
