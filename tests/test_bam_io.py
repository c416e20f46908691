"""BAM front and back ends of the path: the zlib-only scanner + pairing
behind BamApi(input_filepath, config) against oracle/pybam.py's statement-by-statement
restatement of bam_api.cpp:359-507, and the selective copy against bam_api.cpp:534-656.  Input
files come from pybam's independent encoder; output files are read back with python's gzip."""
import os
import random
import struct
import subprocess
import sys

import numpy as np
import pytest

import pybam


@pytest.fixture(scope="module")
def hostlib(pkg):
    from genome_downsampler_b200 import hostlib
    hostlib.load()
    return hostlib


HEADER = pybam.encode_header("@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr\tLN:5000\n@SQ\tSN:other\tLN:77\n",
                             [("chr", 5000), ("other", 77)])


def handmade_records():
    R = pybam.encode_record
    return [
        R("a", 0x41, 10, 60, [(50, "M")], 50),                                   # 0 first of a
        R("lonely", 0x41, 5, 60, [(20, "M")], 20),                               # 1 never paired
        R("b", 0x81, 100, 20, [(10, "S"), (30, "M"), (5, "I"), (20, "M")], 65),  # 2 second of b comes first
        R("a", 0x81, 40, 59, [(25, "M"), (100, "N"), (25, "M")], 50),            # 3 second of a: spliced
        R("b", 0x41, 90, 30, [(5, "H"), (40, "="), (3, "D"), (2, "X"), (4, "P"), (8, "S")], 50),  # 4 first of b
        R("c", 0x41, 300, 0, [(100, "M")], 100, tags=b"NMC\x03RGZgrp\0"),        # 5
        R("c", 0x81, 350, 255, [(100, "M")], 100),                               # 6
        R("a", 0x181, 700, 60, [(50, "M")], 50),                                 # 7 third record named a
        R("x" * 254, 0x41, 1000, 60, [(1, "M")], 1),                             # 8 longest legal QNAME
        R("x" * 254, 0x81, 1000, 60, [(1, "M")], 1),                             # 9
        R("nocigar", 0x45, 2000, 0, [], 30),                                     # 10 end = pos - 1
        R("nocigar", 0x85, 2000, 0, [], 30),                                     # 11
        R("bigtag", 0x41, 3000, 60, [(150, "M")], 150, tags=b"XXZ" + b"t" * 70000 + b"\0"),  # 12 > one member
        R("bigtag", 0x81, 3100, 60, [(150, "M")], 150),                          # 13
        R("d", 0xc1, 4000, 60, [(10, "M")], 10),                                 # 14 both flags set
        R("d", 0xc1, 4100, 60, [(10, "M")], 10),                                 # 15
    ]


def columns_of(reads):
    return {"bam_id": [r["bam_id"] for r in reads], "start": [r["start"] for r in reads],
            "end": [r["end"] for r in reads], "quality": [r["quality"] for r in reads],
            "seq_length": [r["seq_length"] for r in reads], "is_first": [int(r["is_first"]) for r in reads]}


def assert_same_reads(got, want_reads):
    want = columns_of(want_reads)
    for k, v in want.items():
        assert got[k].tolist() == v, k


@pytest.mark.parametrize("member_payload", [37, 1000, 0xff00])
@pytest.mark.parametrize("threads", [1, 4])
def test_read_bam_handmade(hostlib, tmp_path, member_payload, threads):
    recs = handmade_records()
    path = tmp_path / "in.bam"
    pybam.write_bam(path, HEADER, recs, member_payload=member_payload, empty_member_every=3)
    b = hostlib.BamFile(path, threads=threads)
    assert b.ref_length == 5000 and b.record_count == len(recs)
    want, want_out = pybam.ref_read_bam(recs)
    assert_same_reads(b.unfiltered(), want)
    # no filter configured: the filter pass keeps every pair
    assert_same_reads(b.reads(), want)
    assert b.filtered_out().tolist() == want_out == [1]
    b.close()


@pytest.mark.parametrize("cfg", [dict(min_len=50), dict(min_mapq=30), dict(min_len=40, min_mapq=1),
                                 dict(amplicons=[(0, 400), (80, 170), (3000, 3300)])])
def test_read_bam_filter_matches_reference_loop(hostlib, tmp_path, cfg):
    recs = handmade_records()
    path = tmp_path / "in.bam"
    pybam.write_bam(path, HEADER, recs, member_payload=4096)
    kw = dict(min_len=cfg.get("min_len", 0), min_mapq=cfg.get("min_mapq", 0))
    if "amplicons" in cfg:
        bed, tsv = tmp_path / "p.bed", tmp_path / "p.tsv"
        with open(bed, "w") as f, open(tsv, "w") as g:
            for k, (a, e) in enumerate(cfg["amplicons"]):
                f.write("chr\t%d\t%d\tL%d\nchr\t%d\t%d\tR%d\n" % (a, a + 5, k, e - 5, e, k))
                g.write("L%d\tR%d\n" % (k, k))
        kw.update(bed=bed, tsv=tsv)
    b = hostlib.BamFile(path, **kw)
    want, want_out = pybam.ref_read_bam(recs, cfg.get("min_len", 0), cfg.get("min_mapq", 0), cfg.get("amplicons"))
    assert_same_reads(b.reads(), want)
    assert b.filtered_out().tolist() == want_out
    assert b.unfiltered()["bam_id"].size == 0  # the filter has been applied
    b.close()


def random_records(rng, n_records, n_names):
    recs = []
    for _ in range(n_records):
        ops = []
        for _ in range(rng.randint(0, 4)):
            ops.append((rng.randint(1, 60), rng.choice(pybam.CIGAR_OPS[:9])))
        flag = 0x1 | rng.choice([0x40, 0x80, 0x40, 0x80, 0xc0, 0])
        recs.append(pybam.encode_record("q%d" % rng.randrange(n_names), flag, rng.randint(0, 4000),
                                        rng.randint(0, 60), ops, rng.randint(0, 200),
                                        tags=bytes(rng.randrange(256) for _ in range(rng.randint(0, 9)))))
    return recs


@pytest.mark.parametrize("seed", range(12))
def test_read_bam_fuzz_repeated_qnames(hostlib, tmp_path, seed):
    # few names for many records: QNAMEs seen three and more times exercise the reference's
    # never-erased map and its filter-dependent entry swap (bam_api.cpp:434-466)
    rng = random.Random(seed)
    recs = random_records(rng, 400, rng.choice([40, 150, 400]))
    path = tmp_path / "in.bam"
    pybam.write_bam(path, HEADER, recs, member_payload=rng.choice([64, 777, 0xff00]))
    min_len, min_mapq = rng.choice([0, 60, 120]), rng.choice([0, 20, 40])
    amplicons = None if seed % 3 else [(0, 2500), (2000, 4500)]
    kw = {}
    if amplicons:
        bed = tmp_path / "p.bed"
        with open(bed, "w") as f:  # no TSV: name-sorted primers are paired consecutively
            for k, (a, e) in enumerate(amplicons):
                f.write("chr\t%d\t%d\tamp%d_LEFT\nchr\t%d\t%d\tamp%d_RIGHT\n" % (a, a + 9, k, e - 9, e, k))
        kw = dict(bed=bed)
    b = hostlib.BamFile(path, min_len=min_len, min_mapq=min_mapq, threads=1 + seed % 3, **kw)
    want, want_out = pybam.ref_read_bam(recs, min_len, min_mapq, amplicons)
    assert_same_reads(b.reads(), want)
    assert b.filtered_out().tolist() == want_out
    b.close()


def test_write_bam_copies_selected_records_verbatim(hostlib, tmp_path):
    recs = handmade_records() + random_records(random.Random(5), 3000, 1500)
    src = tmp_path / "in.bam"
    pybam.write_bam(src, HEADER, recs, member_payload=5000)
    rng = random.Random(9)
    ids = rng.sample(range(len(recs)), 900) + [12]
    ids = list(dict.fromkeys(ids))
    dst = tmp_path / "out.bam"
    n = hostlib.write_bam(src, dst, ids, threads=3)
    want = pybam.ref_write_bam(recs, ids)
    assert n == len(want) == len(ids)
    header, refs, got = pybam.read_bam(dst)  # python's gzip reads what the C++ writer wrote
    assert header == HEADER and refs == [("chr", 5000), ("other", 77)]
    assert got == want
    # BGZF shape as htslib leaves it: header in its own member(s), payloads <= 0xff00, a record
    # that fits a member is never split, the 28-byte EOF member closes the file
    members = pybam.split_members(dst)
    assert members[-1][0] == pybam.EOF_MEMBER
    assert members[0][1] == HEADER
    assert all(len(p) <= 0xff00 for _, p in members)
    # member payload sizes replayed from htslib's rule (bam_write1: bgzf_flush_try(record size),
    # then bgzf_write, which cuts at 0xff00)
    sizes, cur = [], 0
    for rec in want:
        if cur + len(rec) > 0xff00 and cur:
            sizes.append(cur)
            cur = 0
        left = len(rec)
        while left:
            take = min(left, 0xff00 - cur)
            cur += take
            left -= take
            if cur == 0xff00:
                sizes.append(cur)
                cur = 0
    if cur:
        sizes.append(cur)
    assert [len(p) for _, p in members[1:-1]] == sizes


def test_write_bam_reference_loop_edge_cases(hostlib, tmp_path):
    recs = random_records(random.Random(2), 50, 25)
    src = tmp_path / "in.bam"
    pybam.write_bam(src, HEADER, recs)
    for ids in ([], [49], [0], [3, 3, 7], [10, 99999], list(range(50))):
        dst = tmp_path / "o.bam"
        n = hostlib.write_bam(src, dst, ids)
        want = pybam.ref_write_bam(recs, ids)
        assert n == len(want)
        assert pybam.read_bam(dst)[2] == want


def test_synthetic_bam_roundtrip_and_solution_write(hostlib, O, tmp_path):
    s, e, q, l = O.gen_reads(77, 20_000, 30_000, 150)
    l = np.where(np.arange(len(l)) % 5 == 0, 100, l).astype(np.uint32)  # some reads: 100M50D
    path = tmp_path / "syn.bam"
    hostlib.write_synthetic_bam(path, 30_000, s, e, q, l, coordinate_sorted=False, threads=4)
    b = hostlib.BamFile(path, threads=4)
    u = b.unfiltered()
    assert b.record_count == len(s) and b.ref_length == 30_000
    assert np.array_equal(u["start"], s) and np.array_equal(u["end"], e)
    assert np.array_equal(u["quality"], q) and np.array_equal(u["seq_length"], l)
    assert np.array_equal(u["is_first"], (np.arange(len(s)) % 2 == 0))
    assert np.array_equal(u["bam_id"], np.arange(len(s)))
    # python reads the same file to the same records
    _, refs, recs = pybam.read_bam(path)
    assert refs == [("synthetic", 30_000)] and len(recs) == len(s)
    want, _ = pybam.ref_read_bam(recs[:2000])
    assert [r["start"] for r in want] == s[:2000].tolist()
    # find_pairs + write_paired_reads (src/app.cpp:141-147): kept reads and their mates
    assert np.array_equal(b.reads()["bam_id"], np.arange(len(s)))  # no filter configured
    kept = np.array([0, 5, 6, 1001], np.uint64)
    out = tmp_path / "sol.bam"
    assert b.write_solution(out, kept, with_pairs=True) == 8
    assert pybam.read_bam(out)[2] == [recs[i] for i in (0, 1, 4, 5, 6, 7, 1000, 1001)]
    b.close()


def test_coordinate_sorted_bam_pairs_in_completion_order(hostlib, O, tmp_path):
    s, e, q, l = O.gen_reads(3, 30_000, 30_000, 150)
    path = tmp_path / "sorted.bam"
    hostlib.write_synthetic_bam(path, 30_000, s, e, q, l, coordinate_sorted=True, threads=4)
    _, _, recs = pybam.read_bam(path)
    want, want_out = pybam.ref_read_bam(recs, 0, 30)
    results = []
    for threads in (1, 8):
        b = hostlib.BamFile(path, min_mapq=30, threads=threads)
        got = b.reads()
        assert_same_reads(got, want)
        assert b.filtered_out().tolist() == want_out
        results.append(got)
        b.close()
    # mates adjacent, first mate at the even index
    assert results[0]["is_first"][0::2].all() and not results[0]["is_first"][1::2].any()


BROKEN = {
    "truncated": lambda raw: raw[:len(raw) // 2],
    "crc": lambda raw: raw[:-40] + bytes([raw[-40] ^ 0xff]) + raw[-39:],
    "magic": lambda raw: b"\x1f\x8b\x08\x00" + raw[4:],
}


@pytest.mark.parametrize("kind", sorted(BROKEN))
def test_broken_bam_logs_and_exits_like_the_reference(hostlib, tmp_path, kind):
    # bam-api errors are log + std::exit(EXIT_FAILURE) (bam_api.cpp:373-405), never an exception
    recs = random_records(random.Random(1), 300, 150)
    good = tmp_path / "good.bam"
    pybam.write_bam(good, HEADER, recs, eof=False)
    bad = tmp_path / "bad.bam"
    open(bad, "wb").write(BROKEN[kind](open(good, "rb").read()))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); from __graft_entry__ import load_package; load_package();"
            "from genome_downsampler_b200 import hostlib; b = hostlib.BamFile(%r); print(b.record_count)"
            % (root, str(bad)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1, (r.returncode, r.stdout, r.stderr)
    assert "300" not in r.stdout


@pytest.mark.gpu
def test_bam_to_bam_through_the_b200_plugin(hostlib, O, tmp_path):
    # the whole App::execute chain (src/app.cpp:113-147) on a file: read_bam -> device filter +
    # quasi-mcp-b200 -> find_pairs -> write_paired_reads / write_bam_api_filtered_out_reads
    pairs, L, M = 60_000, 30_000, 50
    a0, a1 = hostlib.artic_amplicons(L)
    s = np.empty(2 * pairs, np.uint32); e = np.empty_like(s); l = np.empty_like(s)
    q = np.empty(2 * pairs, np.uint8)
    hostlib.gen_reads_amplicon_into(99, pairs, L, a0, a1, s, e, q, l)
    path = tmp_path / "in.bam"
    hostlib.write_synthetic_bam(path, L, s, e, q, l, coordinate_sorted=True, threads=8)
    bed, tsv = tmp_path / "p.bed", tmp_path / "p.tsv"
    with open(bed, "w") as f, open(tsv, "w") as g:
        for k in range(len(a0)):
            f.write("synthetic\t%d\t%d\tL%d\nsynthetic\t%d\t%d\tR%d\n" % (a0[k], a0[k] + 20, k, a1[k] - 20, a1[k], k))
            g.write("L%d\tR%d\n" % (k, k))
    _, _, recs = pybam.read_bam(path)
    amplicons = [(int(x), int(y)) for x, y in zip(a0, a1)]
    want, want_out = pybam.ref_read_bam(recs, 90, 30, amplicons)
    b = hostlib.BamFile(path, min_len=90, min_mapq=30, bed=bed, tsv=tsv, threads=8)
    kept = b.solve("quasi-mcp-b200", M)
    assert_same_reads(b.reads(), want)           # device filter == the reference's loop
    assert b.filtered_out().tolist() == want_out
    ws = np.array([r["start"] for r in want], np.uint32); we = np.array([r["end"] for r in want], np.uint32)
    mask = np.zeros(len(want), np.uint8); mask[kept] = 1
    cin = O.coverage_fast(ws, we, L); cout = O.coverage_fast(ws, we, L, mask)
    assert np.array_equal(np.minimum(cin, M), np.minimum(cout, M))
    # (the device filter carries the amplicon table: the graph reduction applies at every M)
    bm, _ = O.sync_solve(ws, we, [L], [0, len(ws)], M, params=(64, 150, 1, 0, 0, 3))
    assert np.array_equal(O.bitmap_to_mask(bm, len(ws)), mask)
    out, fo = tmp_path / "out.bam", tmp_path / "filtered.bam"
    n = b.write_solution(out, kept, with_pairs=True)
    with_mates = sorted(set(int(k) for k in kept) | set(int(k) ^ 1 for k in kept))
    want_ids = sorted(want[i]["bam_id"] for i in with_mates)
    assert n == len(want_ids)
    assert pybam.read_bam(out)[2] == [recs[i] for i in want_ids]
    assert b.write_filtered_out(fo) == len(want_out)
    assert pybam.read_bam(fo)[2] == [recs[i] for i in want_out]
    b.close()


def test_non_bam_extension_gets_sam_text(hostlib, tmp_path):
    # open_mode = extension == ".bam" ? "wb" : "w" (bam_api.cpp:566): anything else is SAM text
    R = pybam.encode_record
    tags = (b"NMC\x03" + b"XAA!" + b"XcC\xfe" + b"Xsc\xfe" + b"XSS\x10\x27" + b"Xhs\xf0\xd8" + b"XII\xff\xff\xff\xff" +
            b"Xii\x00\x00\x00\x80" + b"Xff" + struct.pack("<f", 0.5) + b"RGZgroup one\0" + b"XHH1AE301\0" +
            b"XBBs\x03\x00\x00\x00" + struct.pack("<3h", -1, 0, 300) + b"XFBf\x02\x00\x00\x00" + struct.pack("<2f", 1.5, -2e-7) +
            b"XEBC\x00\x00\x00\x00")
    recs = [
        R("r1", 99, 10, 60, [(5, "S"), (20, "M"), (2, "I"), (3, "D"), (8, "=")], 35, next_ref=0, next_pos=200, tlen=240,
          tags=tags, seq_fill=0x28, qual_fill=40),
        R("r1", 147, 200, 0, [(50, "M")], 50, next_ref=1, next_pos=7, tlen=-240, seq_fill=0xf1, qual_fill=0xff),
        R("unmapped", 77, -1, 0, [], 0, ref_id=-1, next_ref=-1, next_pos=-1),
        R("odd", 65, 7, 255, [(7, "M")], 7, seq_fill=0x84, qual_fill=0),
    ] + handmade_records()
    src = tmp_path / "in.bam"
    text = "@HD\tVN:1.6\n@SQ\tSN:chr\tLN:5000\n@SQ\tSN:other\tLN:77\n@CO\tfree text\n"
    header = pybam.encode_header(text, [("chr", 5000), ("other", 77)])
    pybam.write_bam(src, header, recs, member_payload=900)
    ids = [0, 1, 2, 3, 5, 9, 16]
    dst = tmp_path / "out.sam"
    assert hostlib.write_bam(src, dst, ids, threads=2) == len(ids)
    want = text + "".join(pybam.format_sam(recs[i], ["chr", "other"]) for i in ids)
    got = open(dst, "rb").read().decode("latin-1")
    assert got == want
    lines = got.splitlines()
    assert lines[4].startswith("r1\t99\tchr\t11\t60\t5S20M2I3D8=\t=\t201\t240\t" + "CT" * 17 + "C\t" + "I" * 35 + "\tNM:i:3\tXA:A:!\tXc:i:254")
    assert "\tXB:B:s,-1,0,300\tXF:B:f,1.5,-2e-07\tXE:B:C" in lines[4]
    assert lines[5].split("\t")[6:11] == ["other", "8", "-240", "NA" * 25, "*"]
    assert lines[6].split("\t")[:11] == ["unmapped", "77", "*", "0", "0", "*", "*", "0", "0", "*", "*"]


def test_header_only_bam_and_more_threads_than_records(hostlib, tmp_path):
    empty = tmp_path / "empty.bam"
    pybam.write_bam(empty, HEADER, [])
    b = hostlib.BamFile(empty, threads=8)
    assert b.record_count == 0 and b.ref_length == 5000
    assert b.reads()["bam_id"].size == 0 and b.filtered_out().size == 0
    out = tmp_path / "o.bam"
    assert b.write_solution(out, [], with_pairs=True) == 0
    header, refs, recs = pybam.read_bam(out)
    assert header == HEADER and recs == []
    b.close()
    # two records, sixteen threads; no EOF member at the end of the input
    two = tmp_path / "two.bam"
    recs = handmade_records()[5:7]
    pybam.write_bam(two, HEADER, recs, eof=False)
    b = hostlib.BamFile(two, threads=16)
    want, want_out = pybam.ref_read_bam(recs)
    assert_same_reads(b.reads(), want)
    assert b.filtered_out().tolist() == want_out == []
    assert b.write_solution(out, [1], with_pairs=True) == 2
    assert pybam.read_bam(out)[2] == recs
    b.close()


@pytest.mark.parametrize("chunk_bytes", [1, 70_000, 1_000_000])
def test_many_chunks_alternate_buffers_and_carry_records(hostlib, O, tmp_path, monkeypatch, chunk_bytes):
    # GDS_BAM_CHUNK_BYTES (testing knob): a chunk per BGZF member / per few members, so records
    # straddle chunks, the scanner's two buffers alternate hundreds of times while the pairing
    # runs one chunk behind, and the 70 KB record needs several refills inside one call
    monkeypatch.setenv("GDS_BAM_CHUNK_BYTES", str(chunk_bytes))
    s, e, q, l = O.gen_reads(5, 15_000, 30_000, 150)
    path = tmp_path / "sorted.bam"
    hostlib.write_synthetic_bam(path, 30_000, s, e, q, l, coordinate_sorted=True, threads=4)
    _, _, recs = pybam.read_bam(path)
    want, want_out = pybam.ref_read_bam(recs, 0, 25)
    for threads in (1, 5):
        b = hostlib.BamFile(path, min_mapq=25, threads=threads)
        assert_same_reads(b.reads(), want)
        assert b.filtered_out().tolist() == want_out
        ids = list(range(0, len(recs), 7))
        out = tmp_path / "o.bam"
        assert hostlib.write_bam(path, out, ids, threads=threads) == len(ids)
        assert pybam.read_bam(out)[2] == [recs[i] for i in ids]
        b.close()
    recs = handmade_records()
    small = tmp_path / "handmade.bam"
    pybam.write_bam(small, HEADER, recs, member_payload=3000)
    b = hostlib.BamFile(small, threads=3)
    want, want_out = pybam.ref_read_bam(recs)
    assert_same_reads(b.reads(), want)
    assert b.filtered_out().tolist() == want_out
    b.close()


def test_grade_mode_matches_the_references_read_bam(hostlib, O, R, tmp_path):
    # AmpliconBehaviour::GRADE (bam_api.cpp:334-347, 444-454, 480-483): no pair is dropped for its
    # amplicons; qualities are shifted to start at 0 and pairs inside one amplicon are lifted by the
    # quality range.  The host mirror against the reference's own read_bam (oracle/_ref).
    bed, tsv = O.artic_scheme()
    a0, a1 = O.parse_amplicons(bed, tsv)
    s, e, q, l = O.gen_reads_amplicon(31, 6_000, 30_000, a0, a1)
    bedp, tsvp = tmp_path / "scheme.bed", tmp_path / "pairs.tsv"
    bedp.write_text(bed); tsvp.write_text(tsv)
    path = tmp_path / "grade.bam"
    hostlib.write_synthetic_bam(path, 30_000, s, e, q.astype(np.uint8), l, coordinate_sorted=False, threads=2)
    for min_len, min_mapq in ((0, 0), (90, 30)):
        b = hostlib.BamFile(path, min_len=min_len, min_mapq=min_mapq, bed=bedp, tsv=tsvp,
                            amplicon_behaviour="grade", threads=2)
        got = b.reads()
        b.close()
        ref = R.read_bam(s, e, q, l, 30_000, min_len, min_mapq, bed, tsv, grade=True)
        assert np.array_equal(got["start"], ref["start"]) and np.array_equal(got["bam_id"], ref["bam_id"])
        assert np.array_equal(got["quality"], ref["quality"])
        assert got["quality"].max() > q.max()  # pairs inside an amplicon were lifted
