"""Probe: does running chunks of a batch on two device contexts (two streams, two host threads)
hide the latency-bound max-flow kernel behind the streaming kernels of the other chunk?
Device-resident inputs, CUDA-event timing across both streams."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_oracle, load_package  # noqa: E402

pkg = load_package()
O = load_oracle()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pairs, L, R, M = 1_000_000, 30_000, 150, 100
n_per = 2 * pairs
dev = torch.device("cuda", 0)
s_all = torch.empty(S * n_per, dtype=torch.int32, pin_memory=True)
e_all = torch.empty(S * n_per, dtype=torch.int32, pin_memory=True)
base = O.gen_reads(12345, pairs, L, R)
for k in range(S):  # same sample replicated with a rotation: cheap to make, same statistics
    sh = (k * 7919) % n_per
    s_all.numpy().view(np.uint32)[k * n_per:(k + 1) * n_per] = np.roll(base[0], sh)
    e_all.numpy().view(np.uint32)[k * n_per:(k + 1) * n_per] = np.roll(base[1], sh)
d_s, d_e = s_all.to(dev), e_all.to(dev)
read_off = np.arange(S + 1, dtype=np.uint64) * np.uint64(n_per)
ref_len = np.full(S, L, np.uint32)
bm = torch.zeros(S * n_per // 32 + 4, dtype=torch.int32, device=dev)
main = torch.cuda.Stream(device=dev)
one = pkg.Solver(0)
one.set_stream(main.cuda_stream)


def run_one():
    return one.solve_device(d_s.data_ptr(), d_e.data_ptr(), S * n_per, ref_len, M, bm.data_ptr(),
                            read_off=read_off, len_hint=(R, R))


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(reps):
        fn()
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.cuda.stream(main):
    r = run_one()
    ref_bm = bm.clone()
    print("one call: %.3f ms  (kept %d)" % (timed(run_one), r.n_kept), flush=True)
    for nctx in (2, 3):
        ch = pkg.ChunkedSolver(0, n_contexts=nctx)
        streams = [torch.cuda.Stream(device=dev) for _ in range(nctx)]
        for sv, st in zip(ch.solvers, streams):
            sv.set_stream(st.cuda_stream)
        for chunk in (32, 64, 128, 256):
            if chunk * nctx > S:
                continue

            def run_chunked():
                ev = torch.cuda.Event()
                ev.record(main)
                for st in streams:
                    st.wait_event(ev)
                ch.solve_host_batch(d_s.data_ptr(), d_e.data_ptr(), read_off, ref_len, M,
                                    bm.data_ptr(), chunk_samples=chunk, len_hint=(R, R),
                                    input_on_device=True)
                for st in streams:
                    e = torch.cuda.Event()
                    e.record(st)
                    main.wait_event(e)

            bm.zero_()
            run_chunked()
            torch.cuda.synchronize()
            same = bool(torch.equal(bm, ref_bm))
            print("%d contexts, chunks of %3d samples: %.3f ms  same bitmap %s" %
                  (nctx, chunk, timed(run_chunked), same), flush=True)
        ch.close()
