"""Generates tests/golden/golden.json from the REFERENCE'S OWN code (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run here (the reference is not on the GPU box):
    python tests/golden/make_golden.py
Every entry is an output of unmodified reference sources: reads_gen.cpp streams,
BamApi::find_input_cover / find_filtered_cover / find_pairs, and BamApi::read_bam's pair filter
over a fake in-memory BAM.  Large arrays are stored as sha256 of their little-endian bytes."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as O  # noqa: E402  (only for the synthetic ARTIC scheme text + amplicon-aware reads)
import pyref as R  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    g = {"_comment": "outputs of the unmodified reference (see make_golden.py)"}
    # generator streams: the reference's 5 test inputs (coverage_tester.cpp:120-175) + C1 + C3
    gens = {"test_uniform": (12345, 1_000_000, 30_000, 150, 0),
            "test_low_sides": (12345, 1_000_000, 30_000, 150, 1),
            "test_hole": (12345, 1_000_000, 30_000, 150, 2),
            "test_zero_sides": (12345, 1_000_000, 30_000, 150, 3),
            "c1": (12345, 500_000, 30_000, 150, 0), "c3": (12345, 50_000, 30_000, 150, 0),
            "small_uniform": (7, 2_000, 3_000, 50, 0), "small_hole": (7, 2_000, 3_000, 50, 2)}
    g["generators"] = {}
    for name, (seed, pairs, L, Rl, shape) in gens.items():
        s, e, q, l = R.gen_reads(seed, pairs, L, Rl, shape)
        cov = R.input_cover(s, e, L)
        g["generators"][name] = dict(seed=seed, pairs=pairs, L=L, R=Rl, shape=shape, start=sha(s),
                                     end=sha(e), quality=sha(q), seq_len=sha(l), cover=sha(cov),
                                     cover_head=cov[:8].tolist(), cover_sum=int(cov.sum()))
    # the 16-read example (coverage_tester.cpp:72-93)
    ex = O.SMALL_EXAMPLE
    g["small16"] = dict(cover=R.input_cover(ex["start"], ex["end"], ex["L"]).tolist())
    ids = np.array([0, 2, 3, 5, 6, 9, 12], np.uint64)
    g["small16"]["filtered_cover_ids"] = ids.tolist()
    g["small16"]["filtered_cover"] = R.filtered_cover(ex["start"], ex["end"], ex["L"], ids).tolist()
    g["small16"]["find_pairs"] = R.find_pairs(ex["start"], ex["end"], ex["L"], ids).tolist()
    # the pair filter through BamApi::read_bam
    bed, tsv = O.artic_scheme()
    a0, a1 = O.parse_amplicons(bed, tsv)
    s, e, q, l = O.gen_reads_amplicon(12345, 20_000, 30_000, a0, a1)
    g["filter"] = dict(bed=bed, tsv=tsv, seed=12345, pairs=20_000, L=30_000, cases={})
    for name, (ml, mq, use_bed, use_tsv) in {"l90_q30_tsv": (90, 30, True, True),
                                             "l90_q30_bed_only": (90, 30, True, False),
                                             "l90_q30_noamp": (90, 30, False, False),
                                             "l0_q0_tsv": (0, 0, True, True),
                                             "l151_q0": (151, 0, False, False)}.items():
        r = R.read_bam(s, e, q, l, 30_000, ml, mq, bed if use_bed else None,
                       tsv if use_tsv else None)
        g["filter"]["cases"][name] = dict(min_len=ml, min_mapq=mq, use_bed=use_bed, use_tsv=use_tsv,
                                          n_kept=int(len(r["start"])), bam_id=sha(r["bam_id"]),
                                          start=sha(r["start"]), end=sha(r["end"]),
                                          filtered_out=sha(r["filtered_out"]),
                                          first_ids=r["bam_id"][:10].tolist())
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    json.dump(g, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)


if __name__ == "__main__":
    main()
