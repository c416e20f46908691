"""Worker of tests/test_multi_gpu.py (launched with torch.distributed.run, one rank per GPU):
shards a batch of samples over the ranks with sharding.shard_samples, solves each shard with the
real CUDA solver, gathers the kept bitmaps on rank 0 (sharding.gather_bitmaps_to_root over NCCL) and
compares them with rank 0's own one-GPU solve of the whole batch and with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

N_SAMPLES, PAIRS, L, R, M = 8, 100_032, 30_000, 150, 100  # 200 064 reads per sample (multiple of 128)
PRM = (64, 150, 1, 0)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    from __graft_entry__ import load_package
    pkg = load_package()
    from genome_downsampler_b200 import sharding
    import pyoracle as O
    n_per = 2 * PAIRS
    mine = sharding.shard_samples(N_SAMPLES, world, rank)
    parts = [O.gen_reads(12345 + k, PAIRS, L, R) for k in mine]
    s = torch.from_numpy(np.concatenate([p[0] for p in parts]).view(np.int32)).to(dev)
    e = torch.from_numpy(np.concatenate([p[1] for p in parts]).view(np.int32)).to(dev)
    solver = pkg.Solver(local_rank)
    out, res = sharding.solve_sharded(solver, s, e, n_per, L, M, N_SAMPLES, mine, dev, params=PRM,
                                      len_hint=(R, R))
    assert res.n_components == len(mine) and res.flow_value == res.fstar
    kept_total = sharding.gather_scalars([int(res.n_kept)], dev)
    if rank == 0:
        allp = [O.gen_reads(12345 + k, PAIRS, L, R) for k in range(N_SAMPLES)]
        sa = np.concatenate([p[0] for p in allp]); ea = np.concatenate([p[1] for p in allp])
        off = np.arange(N_SAMPLES + 1, dtype=np.uint64) * np.uint64(n_per)
        one = solver.solve(sa, ea, [L] * N_SAMPLES, M, read_off=off, params=PRM, verify=True)
        got = out.cpu().numpy().view(np.uint32).reshape(-1)
        assert np.array_equal(got, one.kept_bitmap), "gathered bitmap differs from the 1-GPU solve"
        assert int(kept_total.sum()) == one.n_kept
        bm, st = O.sync_solve(sa, ea, [L] * N_SAMPLES, off, M, params=PRM)
        assert np.array_equal(got, bm), "gathered bitmap differs from the oracle"
        open(os.environ["MGPU_OK_FILE"], "w").write("ok %d ranks kept %d" % (world, one.n_kept))
    else:
        assert out is None
    solver.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
