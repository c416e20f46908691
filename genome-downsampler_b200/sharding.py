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


def gather_bitmaps_to_root(local_bitmaps, out=None, dst=0, group=None, async_op=False):
    """Only rank `dst` needs the whole batch's bitmaps (it writes the output): a gather — grouped
    ncclSend/ncclRecv under NCCL — instead of an all-gather, so the other ranks receive nothing
    (round 1: every rank took in 896 MB per step at 8 GPUs, bandwidth its own kernels wanted).

    local_bitmaps: int32 [n_local, words] (the same n_local on every rank: pad the last block).
    out: on `dst`, an int32 [world * n_local, words] tensor to receive into (allocated if None).
    Returns (out or None, work handle or None)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_bitmaps, None
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    local_bitmaps = local_bitmaps.contiguous()
    gather_list = None
    if rank == dst:
        if out is None:
            out = torch.empty((world * local_bitmaps.shape[0], local_bitmaps.shape[1]),
                              dtype=local_bitmaps.dtype, device=local_bitmaps.device)
        gather_list = list(out.view(world, *local_bitmaps.shape).unbind(0))
    work = dist.gather(local_bitmaps, gather_list, dst=dst, group=group, async_op=async_op)
    return (out if rank == dst else None), (work if async_op else None)


def solve_sharded(solver, start, end, n_per_sample, ref_len, max_coverage, n_samples_total,
                  sample_ids, device, params=None, len_hint=None, dst=0, group=None):
    """The multi-GPU entry point of the path: this rank solves the samples it owns —
    `start`/`end` are device int32 tensors holding exactly those samples, n_per_sample reads each,
    in `sample_ids` order — with ONE gds_solve, and the kept bitmaps of the whole batch arrive on
    rank `dst` in global sample order.  No data-path collective besides that gather.

    Returns (bitmaps on dst: int32 [n_samples_total, words] else None, this rank's gds_result)."""
    import numpy as np
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_local = len(sample_ids)
    per_rank = -(-n_samples_total // world)  # the collective needs equal blocks: pad the last one
    words = bitmap_words(n_per_sample)
    if (n_per_sample // 32) * 32 != n_per_sample or n_per_sample // 32 != words:
        raise ValueError("reads per sample must be a multiple of 128 (word-aligned, 16-byte slices)")
    local = torch.zeros((per_rank, words), dtype=torch.int32, device=device)
    read_off = np.arange(n_local + 1, dtype=np.uint64) * np.uint64(n_per_sample)
    res = solver.solve_device(start.data_ptr(), end.data_ptr(), n_local * n_per_sample,
                              np.full(n_local, ref_len, np.uint32), max_coverage, local.data_ptr(),
                              read_off=read_off, params=params, len_hint=len_hint)
    torch.cuda.current_stream(device).synchronize() if local.is_cuda else None
    out, _ = gather_bitmaps_to_root(local, dst=dst, group=group)
    if out is not None and world > 1:
        out = out[:n_samples_total] if per_rank * world == n_samples_total else torch.cat(
            [out[r * per_rank:r * per_rank + len(shard_samples(n_samples_total, world, r))]
             for r in range(world)])
    elif out is not None:
        out = out[:n_samples_total]
    return out, res
