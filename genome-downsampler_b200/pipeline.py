"""Host-buffer batches larger than one PCIe transfer is worth waiting for: solve them in chunks of
independent samples on two device contexts, so the host->device copy of chunk c+1 overlaps the
kernels of chunk c.  Samples are independent, so chunking never changes results:
every chunk writes its slice of the one kept bitmap.

This is host-side orchestration over the C ABI (include/gds.h); it adds no device code.
"""
import threading

import numpy as np

from .binding import Solver


class ChunkedSolver:
    def __init__(self, device=0, n_contexts=2):
        self.solvers = [Solver(device) for _ in range(n_contexts)]

    def close(self):
        for s in self.solvers:
            s.close()

    def solve_host_batch(self, start_ptr, end_ptr, read_off, ref_len, max_coverage, bitmap_ptr,
                         chunk_samples=64, params=None, len_hint=None, start16_ptr=None,
                         input_on_device=False, encode16_ptr=None, encode_threads=8):
        """start_ptr/end_ptr: HOST pointers (pinned for full PCIe speed) of the concatenated reads;
        read_off [ns+1] (every chunk boundary must be a multiple of 32 reads so bitmap slices are
        word-aligned), ref_len [ns]; bitmap_ptr: DEVICE pointer of ceil(n/32) words.
        Compact transport (include/gds.h gds_reads.start16): start16_ptr = 16-bit starts instead
        of start_ptr, end_ptr = None for fixed-length reads (len_hint[0] == len_hint[1]).
        input_on_device=True: the pointers are device pointers (two contexts still overlap one
        chunk's max-flow with the next chunk's streaming kernels).
        encode16_ptr: a HOST scratch buffer of n uint16 — every chunk's 32-bit columns are narrowed
        into it (hostlib.encode_compact: 16-bit starts + exact read-length range) by an encoder thread
        that runs ahead of the device, and the chunk travels compact when that is legal (one
        read length, starts below 65536), as 32-bit columns otherwise: what the C++ adapter does
        with the reference's size_t columns, for callers that hold uint32 arrays.
        Returns the list of per-chunk results (in chunk order)."""
        read_off = np.ascontiguousarray(read_off, np.uint64)
        ref_len = np.ascontiguousarray(ref_len, np.uint32)
        ns = len(ref_len)
        chunks = [(a, min(a + chunk_samples, ns)) for a in range(0, ns, chunk_samples)]
        for a, _ in chunks:
            if int(read_off[a]) % 32:
                raise ValueError("chunk boundaries must fall on multiples of 32 reads")
        results = [None] * len(chunks)
        errors = []
        # Narrowing runs AHEAD of the device on its own thread (the encoder's host threads), chunk by
        # chunk in order; a worker picks a chunk up as soon as it is narrowed.  Chunk k+1 is narrowed
        # while chunk k crosses PCIe and chunk k-1 is in its kernels: the batch takes
        # max(narrowing, transfer) + one chunk instead of their sum.
        encoding = bool(encode16_ptr)
        enc_done = [threading.Event() for _ in chunks]
        enc_info = [None] * len(chunks)
        known = len_hint is not None and len_hint[0] == len_hint[1] and len_hint[0] > 0

        def encoder():
            from . import hostlib
            try:
                for ci, (a, b) in enumerate(chunks):
                    r0 = int(read_off[a])
                    n = int(read_off[b]) - r0
                    if n and int(ref_len[a:b].max()) <= 65536:
                        # a caller that KNOWS the one read length (len_hint) spares the end column
                        enc_info[ci] = hostlib.encode_compact(
                            start_ptr + 4 * r0, None if known else end_ptr + 4 * r0, n,
                            encode16_ptr + 2 * r0, threads=encode_threads)
                    enc_done[ci].set()
            except Exception as ex:
                errors.append(ex)
                for ev in enc_done:
                    ev.set()

        def worker(w):
            sv = self.solvers[w]
            try:
                for ci in range(w, len(chunks), len(self.solvers)):
                    a, b = chunks[ci]
                    r0 = int(read_off[a])
                    n = int(read_off[b]) - r0
                    sp = start_ptr + 4 * r0 if start_ptr else None
                    ep = end_ptr + 4 * r0 if end_ptr else None
                    s16 = start16_ptr + 2 * r0 if start16_ptr else None
                    hint = len_hint
                    if encoding:
                        enc_done[ci].wait()
                        if errors:
                            return
                        if enc_info[ci] is not None:
                            fits, lo, hi = enc_info[ci]
                            if not known:
                                hint = (lo, hi)
                            if fits and hint[0] == hint[1]:
                                sp, ep, s16 = None, None, encode16_ptr + 2 * r0
                    results[ci] = sv.solve_device(
                        sp, ep, n, ref_len[a:b], max_coverage,
                        bitmap_ptr + 4 * (r0 // 32), read_off=read_off[a:b + 1] - np.uint64(r0),
                        params=params, input_on_device=input_on_device, len_hint=hint,
                        start16_ptr=s16)
            except Exception as ex:  # surfaced to the caller after the join
                errors.append(ex)

        threads = [threading.Thread(target=worker, args=(w,)) for w in range(len(self.solvers))]
        if encoding:
            threads.append(threading.Thread(target=encoder))
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results
