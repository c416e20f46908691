"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sample sharding + bitmap gather
(SURVEY.md §8e).  The per-sample solve is a deterministic stand-in here (the CUDA solver needs a
GPU); what is tested is that every sample is solved exactly once, on the rank that owns it, and
that the gathered bitmaps come back in global sample order on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_SAMPLES = 7          # odd on purpose: ranks own 3 and 4 samples
READS_PER_SAMPLE = 1000


def _fake_bitmap(sample_id, words):
    rng = np.random.default_rng(1000 + sample_id)
    return torch.from_numpy(rng.integers(0, 2 ** 31, size=words).astype(np.int32))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load_package
    load_package()
    from genome_downsampler_b200 import sharding
    mine = sharding.shard_samples(N_SAMPLES, world, rank)
    words = sharding.bitmap_words(READS_PER_SAMPLE)
    assert words % 4 == 0 and words * 32 >= READS_PER_SAMPLE
    # equal-size collective: pad every rank's block to the largest block
    per_rank = max(len(sharding.shard_samples(N_SAMPLES, world, r)) for r in range(world))
    local = torch.zeros((per_rank, words), dtype=torch.int32)
    for j, k in enumerate(mine):
        local[j] = _fake_bitmap(k, words)
    gathered = sharding.gather_bitmaps(local)
    assert gathered.shape == (world * per_rank, words)
    scal = sharding.gather_scalars([k * 10 for k in mine] + [-1] * (per_rank - len(mine)), "cpu")
    # every rank reconstructs the global order
    got = {}
    for r in range(world):
        for j, k in enumerate(sharding.shard_samples(N_SAMPLES, world, r)):
            got[k] = gathered[r * per_rank + j]
            assert int(scal[r * per_rank + j]) == k * 10
    assert sorted(got) == list(range(N_SAMPLES))
    for k in range(N_SAMPLES):
        assert torch.equal(got[k], _fake_bitmap(k, words))
    # the gather the product uses: only the root receives, blocks arrive in rank order
    root_out, work = sharding.gather_bitmaps_to_root(local, dst=0, async_op=True)
    if work is not None:
        work.wait()
    if rank == 0:
        assert root_out.shape == (world * per_rank, words) and torch.equal(root_out, gathered)
    else:
        assert root_out is None
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_shard_partition_is_exact():
    from __graft_entry__ import load_package
    load_package()
    from genome_downsampler_b200 import sharding
    for n in (1, 7, 64, 512, 513):
        for w in (1, 2, 3, 8):
            parts = [sharding.shard_samples(n, w, r) for r in range(w)]
            flat = [k for p in parts for k in p]
            assert flat == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_world_size_2_gloo_gather(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
