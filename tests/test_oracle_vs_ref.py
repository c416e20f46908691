"""The oracle against the REFERENCE'S OWN code, live (oracle/_ref/libgds_ref.so built from
/root/reference).  Skipped when the .so is absent.  The same outputs are frozen in
tests/golden/golden.json for boxes that cannot build it."""
import numpy as np

from conftest import PRM


def test_generators_bit_exact(O, R):
    for shape, name in enumerate(["uniform", "low_sides", "hole", "zero_sides"]):
        for seed, pairs, L, Rl in [(12345, 20_000, 30_000, 150), (9, 500, 700, 31)]:
            a = O.gen_reads(seed, pairs, L, Rl, name)
            b = R.gen_reads(seed, pairs, L, Rl, shape)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


def test_cover_helpers_bit_exact(O, R):
    s, e, q, l = O.gen_reads(3, 5_000, 4_000, 80, "hole")
    assert np.array_equal(O.coverage(s, e, 4_000), R.input_cover(s, e, 4_000))
    ids = np.arange(0, len(s), 7, dtype=np.uint64)
    mask = np.zeros(len(s), np.uint8); mask[ids] = 1
    assert np.array_equal(O.coverage(s, e, 4_000, mask), R.filtered_cover(s, e, 4_000, ids))


def test_filter_matches_read_bam(O, R):
    bed, tsv = O.artic_scheme()
    a0, a1 = O.parse_amplicons(bed, tsv)
    s, e, q, l = O.gen_reads_amplicon(77, 15_000, 30_000, a0, a1)
    for ml, mq, ub, ut in [(90, 30, True, True), (90, 30, True, False), (0, 0, True, True),
                           (120, 50, False, False)]:
        ref = R.read_bam(s, e, q, l, 30_000, ml, mq, bed if ub else None, tsv if ut else None)
        x0, x1 = O.parse_amplicons(bed, tsv if ut else None)
        pp, kept = O.filter_pairs(s, e, q, l, ml, mq, x0 if ub else None, x1 if ub else None)
        mask = np.repeat(pp, 2).astype(bool)
        assert kept == len(ref["start"])
        assert np.array_equal(np.nonzero(mask)[0], ref["bam_id"])
        assert np.array_equal(s[mask], ref["start"]) and np.array_equal(e[mask], ref["end"])
        assert np.array_equal(np.nonzero(~mask)[0], ref["filtered_out"])


def test_find_pairs_bitmap_equals_reference_set(O, R):
    s, e, q, l = O.gen_reads(5, 300, 900, 40)
    rng = np.random.default_rng(0)
    ids = np.sort(rng.choice(len(s), 150, replace=False)).astype(np.uint64)
    ref = R.find_pairs(s, e, 900, ids)
    bm = np.zeros((len(s) + 31) // 32, np.uint32)
    for i in ids:
        bm[int(i) >> 5] |= np.uint32(1 << (int(i) & 31))
    got = np.nonzero(O.bitmap_to_mask(O.find_pairs_bitmap(bm, len(s)), len(s)))[0]
    assert sorted(ref.tolist()) == got.tolist()


def test_reference_coverage_tester_accepts_oracle_solvers(O, R):
    # CoverageTester::test (5 cases, asserts live) through qmcp::Solver
    def sync(M, L, s, e):
        bm, st = O.sync_solve(s, e, [L], [0, len(s)], M, params=PRM)
        return np.nonzero(O.bitmap_to_mask(bm, len(s)))[0]
    assert R.run_coverage_tests(sync) == 5
