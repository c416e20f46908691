"""CPU tests of the oracle itself: known answers, the golden vectors produced by the reference's
own code (tests/golden/make_golden.py), independent max-flow / min-cost cross-checks."""
import hashlib

import numpy as np
import pytest

from conftest import PRM


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SHAPE_NAMES = ["uniform", "low_sides", "hole", "zero_sides"]


def test_small_example_known_answers(O, golden):
    # SURVEY App. A.6 / coverage_tester.cpp:72-93
    ex = O.SMALL_EXAMPLE
    cov = O.coverage(ex["start"], ex["end"], ex["L"])
    assert cov.tolist() == [3, 5, 6, 5, 6, 3, 6, 5, 5, 5, 3] == golden["small16"]["cover"]
    d, f = O.demand(cov, ex["M"])
    assert d.tolist() == [-3, -1, 0, 0, 0, 1, -1, 0, 0, 0, 1, 3]
    assert f == 5
    kept, st = O.ref_solve(ex["start"], ex["end"], ex["L"], ex["M"])
    assert st.flow_value == 5
    bm, ss = O.sync_solve(ex["start"], ex["end"], [ex["L"]], [0, 16], ex["M"], params=PRM)
    assert ss.flow_value == 5 and ss.fstar == 5
    for mask in (kept, O.bitmap_to_mask(bm, 16)):
        cout = O.coverage(ex["start"], ex["end"], ex["L"], mask)
        assert np.array_equal(np.minimum(cout, 4), np.minimum(cov, 4))
    _, nopt = O.greedy_multicover(ex["start"], ex["end"], ex["L"], ex["M"])
    assert nopt == 14
    assert ss.n_kept >= 14


@pytest.mark.parametrize("name", ["small_uniform", "small_hole", "c3", "c1", "test_uniform",
                                  "test_low_sides", "test_hole", "test_zero_sides"])
def test_generator_streams_match_reference_golden(O, golden, name):
    g = golden["generators"][name]
    s, e, q, l = O.gen_reads(g["seed"], g["pairs"], g["L"], g["R"], SHAPE_NAMES[g["shape"]])
    assert sha(s) == g["start"] and sha(e) == g["end"]
    assert sha(q) == g["quality"] and sha(l) == g["seq_len"]
    cov = O.coverage_fast(s, e, g["L"])
    assert sha(cov) == g["cover"] and int(cov.sum()) == g["cover_sum"]


def test_cover_helpers_match_reference_golden(O, golden):
    ex = O.SMALL_EXAMPLE
    ids = np.array(golden["small16"]["filtered_cover_ids"])
    mask = np.zeros(16, np.uint8)
    mask[ids] = 1
    assert O.coverage(ex["start"], ex["end"], ex["L"], mask).tolist() == \
        golden["small16"]["filtered_cover"]
    # find_pairs on a bitmap == reference's find_pairs as a set
    bm = np.zeros(1, np.uint32)
    for i in ids:
        bm[0] |= np.uint32(1 << int(i))
    paired = O.bitmap_to_mask(O.find_pairs_bitmap(bm, 16), 16)
    assert sorted(np.nonzero(paired)[0].tolist()) == sorted(golden["small16"]["find_pairs"])


@pytest.mark.parametrize("case", ["l90_q30_tsv", "l90_q30_bed_only", "l90_q30_noamp", "l0_q0_tsv",
                                  "l151_q0"])
def test_filter_matches_reference_read_bam_golden(O, golden, case):
    gf = golden["filter"]
    c = gf["cases"][case]
    a0, a1 = O.parse_amplicons(gf["bed"], gf["tsv"] if c["use_tsv"] else None)
    s, e, q, l = O.gen_reads_amplicon(gf["seed"], gf["pairs"], gf["L"],
                                      *O.parse_amplicons(gf["bed"], gf["tsv"]))
    pp, kept = O.filter_pairs(s, e, q, l, c["min_len"], c["min_mapq"],
                              a0 if c["use_bed"] else None, a1 if c["use_bed"] else None)
    mask = np.repeat(pp, 2).astype(bool)
    assert kept == c["n_kept"]
    ids = np.nonzero(mask)[0].astype(np.uint64)
    assert sha(ids) == c["bam_id"] and ids[:10].tolist() == c["first_ids"]
    assert sha(s[mask]) == c["start"] and sha(e[mask]) == c["end"]
    assert sha(np.nonzero(~mask)[0].astype(np.uint64)) == c["filtered_out"]


def test_filter_truth_table(O):
    # pair-level predicate, inclusive amplicon bounds (amplicon.cpp:5-7)
    s = np.array([10, 50, 10, 50, 10, 50, 10, 61, 10, 50, 0, 50], np.uint32)
    e = np.array([40, 90, 40, 90, 40, 90, 40, 100, 40, 101, 40, 90], np.uint32)
    q = np.array([30, 30, 29, 60, 30, 30, 30, 30, 30, 30, 30, 30], np.uint32)
    l = np.array([90, 90, 90, 90, 89, 90, 90, 90, 90, 90, 90, 90], np.uint32)
    a0 = np.array([10, 200], np.uint32)
    a1 = np.array([100, 300], np.uint32)
    pp, kept = O.filter_pairs(s, e, q, l, 90, 30, a0, a1)
    #            ok  mapq  len  ok(edge) end>100 start<10
    assert pp.tolist() == [1, 0, 0, 1, 0, 0] and kept == 4
    pp, kept = O.filter_pairs(s, e, q, l, 90, 30)
    assert pp.tolist() == [1, 0, 0, 1, 1, 1]


def test_parse_amplicons_rules(O):
    bed = "c\t30\t52\tb_LEFT\nc\t400\t420\tb_RIGHT\nc\t0\t22\ta_LEFT\nc\t380\t399\ta_RIGHT\nbad line\n"
    a0, a1 = O.parse_amplicons(bed, "a_LEFT\ta_RIGHT\nb_RIGHT\tb_LEFT\n")
    assert a0.tolist() == [0, 30] and a1.tolist() == [399, 420]  # swapped pair is re-ordered
    a0, a1 = O.parse_amplicons(bed, None)  # name-sorted consecutive pairing (bam_api.cpp:75-89)
    assert a0.tolist() == [0, 30] and a1.tolist() == [399, 420]
    with pytest.raises(ValueError):
        O.parse_amplicons("c\t0\t5\tx\nc\t9\t12\ty\nc\t20\t25\tz\n", None)  # odd count (App. B10)


def _random_instance(rng, max_L=60, max_n=40):
    L = int(rng.integers(1, max_L))
    n = 2 * int(rng.integers(0, max_n))
    s = rng.integers(0, L, size=n)
    ln = rng.integers(1, max(2, L // 2 + 1), size=n)
    e = np.minimum(s + ln - 1, L - 1)
    return s.astype(np.uint32), e.astype(np.uint32), L, int(rng.integers(1, 8))


def test_flow_value_against_scipy_dinic(O):
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import maximum_flow
    rng = np.random.default_rng(3)
    for _ in range(150):
        s, e, L, M = _random_instance(rng)
        cov = O.coverage(s, e, L)
        d, fstar = O.demand(cov, M)
        BIG = 10 ** 6
        rows, cols, caps = [], [], []
        for a, b in zip(s, e):
            rows.append(int(a)); cols.append(int(b) + 1); caps.append(1)
        for i in range(L):
            rows.append(i + 1); cols.append(i); caps.append(BIG)
        for i in range(L + 1):
            if d[i] > 0:
                rows.append(i); cols.append(L + 2); caps.append(int(d[i]))
            elif d[i] < 0:
                rows.append(L + 1); cols.append(i); caps.append(int(-d[i]))
        g = csr_matrix((np.array(caps, np.int32), (rows, cols)), shape=(L + 3, L + 3))
        g.sum_duplicates()
        f = maximum_flow(g, L + 1, L + 2).flow_value if len(caps) else 0
        kept, st = O.ref_solve(s, e, L, M)
        bm, ss = O.sync_solve(s, e, [L], [0, len(s)], M, params=PRM)
        assert f == fstar == st.flow_value == ss.flow_value == ss.fstar


def test_greedy_multicover_is_min_cost_flow_optimum(O):
    import networkx as nx
    rng = np.random.default_rng(5)
    for _ in range(40):
        s, e, L, M = _random_instance(rng, 25, 14)
        cov = O.coverage(s, e, L)
        d, _ = O.demand(cov, M)
        G = nx.MultiDiGraph()
        for i in range(L + 1):
            G.add_node(i, demand=int(d[i]))  # mcp_cpu_cost_scaling_solver.cpp:60-66
        for a, b in zip(s, e):
            G.add_edge(int(a), int(b) + 1, capacity=1, weight=1)
        for i in range(L):
            G.add_edge(i + 1, i, capacity=10 ** 6, weight=0)
        cost, _ = nx.network_simplex(G)
        kept, nopt = O.greedy_multicover(s, e, L, M)
        assert nopt == cost
        cout = O.coverage(s, e, L, kept)
        assert np.all(cout >= np.minimum(cov, M))


def test_sync_schedule_fuzz_invariants(O):
    rng = np.random.default_rng(11)
    for it in range(400):
        ns = int(rng.integers(1, 4))
        ss, ee, off, Ls = [], [], [0], []
        for _ in range(ns):
            s, e, L, _m = _random_instance(rng)
            ss.append(s); ee.append(e); off.append(off[-1] + len(s)); Ls.append(L)
        s = np.concatenate(ss); e = np.concatenate(ee)
        M = int(rng.integers(1, 8))
        # every other case cuts the references into short segments: all invariants
        # below are stated on the ORIGINAL network and must not notice
        seg = int(rng.integers(1, 60)) if it % 2 else 0
        prm = (int(rng.integers(1, 8)), int(rng.integers(0, 300)), int(rng.integers(0, 5)), 0, seg)
        bm, st, dem, cov = O.sync_solve(s, e, Ls, off, M, params=prm, want_vectors=True)
        assert st.flow_value == st.fstar
        mask = O.bitmap_to_mask(bm, len(s))
        assert mask.sum() == st.n_kept
        pos = 0
        for k, L in enumerate(Ls):
            a, b = off[k], off[k + 1]
            cin = O.coverage(s[a:b].copy(), e[a:b].copy(), L)
            cout = O.coverage(s[a:b].copy(), e[a:b].copy(), L, mask[a:b].copy())
            assert np.array_equal(np.minimum(cin, M), np.minimum(cout, M))
            dk, _ = O.demand(cin, M)
            assert np.array_equal(dem[pos:pos + L + 1], dk)
            assert np.array_equal(cov[pos:pos + L], cin) and cov[pos + L] == 0
            pos += L + 1


def test_segment_split_is_exact_and_costs_at_most_M_reads_per_cut(O):
    # config-3 input cut into ever shorter segments: F*, demand and coverage never change, every
    # result is a valid downsample, the number of kept reads grows by <= M per cut
    s, e, q, l = O.gen_reads(12345, 50_000, 30_000, 150)
    M, L = 100, 30_000
    base = None
    for seg in (0xffffffff, 8192, 2048, 512, 100):   # 100 < read length: split disabled
        bm, st, dem, cov = O.sync_solve(s, e, [L], [0, len(s)], M, params=(64, 150, 1, 0, seg),
                                        want_vectors=True)
        mask = O.bitmap_to_mask(bm, len(s))
        cout = O.coverage_fast(s, e, L, mask)
        assert np.array_equal(np.minimum(cov[:L], M), np.minimum(cout, M))
        assert st.flow_value == st.fstar == 100 and int(mask.sum()) == st.n_kept
        if base is None:
            base = (st.n_kept, dem.copy(), cov.copy())
            assert st.n_components == 1
            continue
        assert np.array_equal(dem, base[1]) and np.array_equal(cov, base[2])
        cuts = (L + seg - 1) // seg - 1 if seg >= 150 else 0
        assert st.n_components == cuts + 1
        assert base[0] <= st.n_kept <= base[0] + M * cuts
        if cuts == 0:
            assert st.n_kept == base[0]


def test_express_schedule_is_a_maximum_flow_with_fewer_rounds(O):
    # segments of a cut reference take the EXPRESS schedule (zero-length back arcs in the labels):
    # the same F*, a valid cover, and far fewer rounds and relabel levels than the classic schedule
    L, M = 200_000, 500
    s, e, _, _ = O.gen_reads(7, int(L * 1500 / 150 / 2), L, 150)
    off = [0, len(s)]

    def capped(ss, ee):
        d = np.zeros(L + 1, np.int64)
        np.add.at(d, ss, 1)
        np.add.at(d, ee + 1, -1)
        return np.minimum(np.cumsum(d)[:L], M)

    want = capped(s, e)
    out = {}
    for sched in (0, 1):
        bm, st = O.sync_solve(s, e, [L], off, M, params=(64, 150, 1, 0, 0, sched))
        assert st.flow_value == st.fstar == M
        keep = O.bitmap_to_mask(bm, len(s)) == 1
        assert np.array_equal(capped(s[keep], e[keep]), want)
        out[sched] = st.as_dict()
    assert out[0]["n_express"] == out[0]["n_components"] == 13 and out[1]["n_express"] == 0
    assert out[0]["rounds_total"] * 2 < out[1]["rounds_total"]
    assert out[0]["bfs_levels"] * 2 < out[1]["bfs_levels"]
    # an uncut reference never takes it (results of whole samples do not depend on the batch) ...
    s2, e2, _, _ = O.gen_reads(8, 100_000, 30_000, 150)
    _, st2 = O.sync_solve(s2, e2, [30_000], [0, len(s2)], 100, params=(64, 150, 1, 0, 0, 0))
    assert st2.n_express == 0
    # ... and a cut one with a hole inside a segment falls back to the classic schedule there
    s3, e3, _, _ = O.gen_reads(9, 400_000, 100_000, 150, "hole")
    bm3, st3 = O.sync_solve(s3, e3, [100_000], [0, len(s3)], 1000, params=(64, 150, 1, 0, 0, 0))
    assert 0 < st3.n_express < st3.n_components and st3.flow_value == st3.fstar


def test_forced_reads_and_cuts_make_every_shape_a_transport(O):
    # from M = 128 on, bundles that cover a position with cov <= M are taken out with their fixed
    # flow and the components are cut at those positions: the same F*, a valid cover,
    # and the reference's hard shapes need hops, not thousands of rounds
    L, M = 30_000, 1600
    for shape in ("hole", "low_sides", "zero_sides"):
        s, e, _, _ = O.gen_reads(77, 200_000, L, 150, shape)
        off = [0, len(s)]

        def capped(ss, ee):
            d = np.zeros(L + 1, np.int64)
            np.add.at(d, ss, 1)
            np.add.at(d, ee + 1, -1)
            return np.minimum(np.cumsum(d)[:L], M)

        res = {}
        for sched in (1, 0):
            bm, st = O.sync_solve(s, e, [L], off, M, params=(64, 150, 1, 0, 0, sched))
            assert st.flow_value == st.fstar
            keep = O.bitmap_to_mask(bm, len(s)) == 1
            assert np.array_equal(capped(s[keep], e[keep]), capped(s, e))
            res[sched] = st.as_dict()
        assert res[1]["n_express"] == 0 and res[0]["n_express"] >= 1, shape
        assert res[0]["rounds_max"] * 3 < res[1]["rounds_max"], (shape, res[0]["rounds_max"], res[1]["rounds_max"])
        assert res[0]["n_kept"] <= 1.001 * res[1]["n_kept"]
    # below M = 128 nothing changes: the uncut graph, the classic schedule
    s, e, _, _ = O.gen_reads(78, 100_000, L, 150, "hole")
    a = O.sync_solve(s, e, [L], [0, len(s)], 100, params=(64, 150, 1, 0, 0, 0))
    b = O.sync_solve(s, e, [L], [0, len(s)], 100, params=(64, 150, 1, 0, 0, 1))
    assert np.array_equal(a[0], b[0]) and a[1].rounds_total == b[1].rounds_total


def test_batch_equals_per_sample(O):
    # a batch is exactly the concatenation of independent per-sample solves
    parts = [O.gen_reads(100 + k, 3000, 3000, 50) for k in range(3)]
    s = np.concatenate([p[0] for p in parts]); e = np.concatenate([p[1] for p in parts])
    off = [0, 6000, 12000, 18000]
    bm, st = O.sync_solve(s, e, [3000] * 3, off, 20, params=PRM)
    mask = O.bitmap_to_mask(bm, len(s))
    for k, p in enumerate(parts):
        bk, sk = O.sync_solve(p[0], p[1], [3000], [0, 6000], 20, params=PRM)
        assert np.array_equal(O.bitmap_to_mask(bk, 6000), mask[off[k]:off[k + 1]])


def test_reference_shapes_invariant_and_quality(O):
    # the reference's own 5 cases at 1/10 size (full size runs in the gpu suite)
    for shape, M in [("uniform", 100), ("low_sides", 800), ("hole", 800), ("zero_sides", 800)]:
        s, e, q, l = O.gen_reads(12345, 100_000, 30_000, 150, shape)
        bm, st = O.sync_solve(s, e, [30_000], [0, len(s)], M, params=PRM)
        mask = O.bitmap_to_mask(bm, len(s))
        cin = O.coverage_fast(s, e, 30_000)
        cout = O.coverage_fast(s, e, 30_000, mask)
        assert np.array_equal(np.minimum(cin, M), np.minimum(cout, M))  # implies cov_out >= capped
        assert st.flow_value == st.fstar
        _, nopt = O.greedy_multicover(s, e, 30_000, M)
        assert nopt <= st.n_kept <= 1.25 * nopt  # reported quality: close to the mcp-cpu optimum


def test_ref_restatement_matches_sync_on_value_and_capped_coverage(O):
    s, e, q, l = O.gen_reads(12345, 50_000, 30_000, 150)
    kept, st = O.ref_solve(s, e, 30_000, 100)
    bm, ss = O.sync_solve(s, e, [30_000], [0, len(s)], 100, params=PRM)
    assert st.flow_value == ss.flow_value == 100
    cin = O.coverage_fast(s, e, 30_000)
    c1 = O.coverage_fast(s, e, 30_000, kept)
    c2 = O.coverage_fast(s, e, 30_000, O.bitmap_to_mask(bm, len(s)))
    assert np.array_equal(np.minimum(c1, 100), np.minimum(cin, 100))
    assert np.array_equal(np.minimum(c2, 100), np.minimum(cin, 100))


def test_bad_coordinates_rejected(O):
    s = np.array([0, 5], np.uint32); e = np.array([3, 10], np.uint32)
    with pytest.raises(ValueError):
        O.sync_solve(s, e, [10], [0, 2], 2, params=PRM)  # end >= L
    with pytest.raises(ValueError):
        O.coverage(np.array([4], np.uint32), np.array([2], np.uint32), 10)  # start > end
