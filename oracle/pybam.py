"""CPU restatement of the reference's BAM front and back ends — TEST INFRASTRUCTURE ONLY (imported
by tests/ alone, never by the product path).

Restates, in plain Python over the SAM/BAM specification (SAMv1 §4.1 BGZF, §4.2 BAM):
  * BamApi::read_bam           /root/reference libs/bam-api/src/bam_api.cpp:359-507
  * Read::Read(id, bam1_t*)    libs/bam-api/src/read.cpp:5-14  (end = pos + bam_cigar2rlen - 1)
  * should_be_filtered_out     bam_api.cpp:311-332, amplicon.cpp:5-7, amplicon_set.cpp:5-9
  * BamApi::write_bam          bam_api.cpp:534-656 (which records, in which order)
and holds an INDEPENDENT BAM encoder (struct + zlib) so the C++ scanner is never checked against
files produced by its own writer only.

PARITY UNPINNED: the reference reads and writes through htslib (pinned 1.22.1 by
scripts/install_libs.sh:14,140), which is absent from this image, and the reference ships no BAM
fixture and no test for this code.  What is pinned is the file format (python's gzip module reads
every file the C++ writer produces; the C++ reader reads files this module produces) and the
call-site logic restated below.
"""
import gzip
import struct
import zlib

CIGAR_OPS = "MIDNSHP=XB"
CONSUMES_REF = {0, 2, 3, 7, 8}  # bam_cigar_type bit 2: M D N = X
EOF_MEMBER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


# ----------------------------------------------------------------------------- encoding

def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def encode_record(qname, flag, pos, mapq, cigar, l_seq, ref_id=0, next_ref=0, next_pos=0, tlen=0,
                  tags=b"", seq_fill=0x12, qual_fill=30):
    """cigar: list of (length, op char).  Returns the record including its block_size field."""
    name = qname.encode() + b"\0"
    cig = b"".join(struct.pack("<I", (n << 4) | CIGAR_OPS.index(op)) for n, op in cigar)
    rlen = sum(n for n, op in cigar if CIGAR_OPS.index(op) in CONSUMES_REF)
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(name), mapq,
                       reg2bin(max(pos, 0), max(pos, 0) + max(rlen, 1)), len(cigar), flag, l_seq,
                       next_ref, next_pos, tlen)
    body += name + cig + bytes([seq_fill]) * ((l_seq + 1) // 2) + bytes([qual_fill]) * l_seq + tags
    return struct.pack("<I", len(body)) + body


def encode_header(text, refs):
    """refs: list of (name, length)."""
    t = text.encode()
    out = b"BAM\1" + struct.pack("<I", len(t)) + t + struct.pack("<I", len(refs))
    for name, length in refs:
        nm = name.encode() + b"\0"
        out += struct.pack("<I", len(nm)) + nm + struct.pack("<I", length)
    return out


def bgzf_member(payload, level=6):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    data = c.compress(payload) + c.flush()
    total = 18 + len(data) + 8
    assert total <= 65536
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", total - 1) + data +
            struct.pack("<II", zlib.crc32(payload), len(payload)))


def write_bam(path, header_bytes, records, member_payload=0xff00, eof=True, empty_member_every=0):
    """Cuts the uncompressed stream every `member_payload` bytes regardless of record boundaries
    (legal BGZF; exercises the reader's carry-over), optionally sprinkling empty members."""
    stream = header_bytes + b"".join(records)
    with open(path, "wb") as f:
        k = 0
        for off in range(0, len(stream), member_payload):
            f.write(bgzf_member(stream[off:off + member_payload]))
            k += 1
            if empty_member_every and k % empty_member_every == 0:
                f.write(bgzf_member(b""))
        if eof:
            f.write(EOF_MEMBER)


# ----------------------------------------------------------------------------- decoding

def split_members(path):
    """[(member bytes, inflated payload)] — walks BSIZE like a BGZF reader, checks CRC and ISIZE."""
    raw = open(path, "rb").read()
    out = []
    p = 0
    while p < len(raw):
        assert raw[p:p + 4] == b"\x1f\x8b\x08\x04", "not a BGZF member at %d" % p
        xlen = struct.unpack_from("<H", raw, p + 10)[0]
        bsize = None
        q = p + 12
        while q < p + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", raw, q)
            if (si1, si2) == (66, 67):
                bsize = struct.unpack_from("<H", raw, q + 4)[0]
            q += 4 + slen
        assert bsize is not None
        member = raw[p:p + bsize + 1]
        payload = zlib.decompress(member[12 + xlen:-8], -15)
        crc, isize = struct.unpack("<II", member[-8:])
        assert crc == zlib.crc32(payload) and isize == len(payload)
        out.append((member, payload))
        p += bsize + 1
    return out


def read_bam(path):
    """(header_bytes, refs, [record bytes incl. block_size]) via python's own gzip reader."""
    data = gzip.decompress(open(path, "rb").read())
    assert data[:4] == b"BAM\1"
    l_text = struct.unpack_from("<I", data, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", data, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<I", data, p)[0]
        name = data[p + 4:p + 4 + l_name].split(b"\0")[0].decode()
        refs.append((name, struct.unpack_from("<I", data, p + 4 + l_name)[0]))
        p += 8 + l_name
    header = data[:p]
    records = []
    while p < len(data):
        bs = struct.unpack_from("<I", data, p)[0]
        records.append(data[p:p + 4 + bs])
        p += 4 + bs
    assert p == len(data)
    return header, refs, records


def parse_record(rec, bam_id):
    """The fields Read::Read(id, bam1_t*) takes (read.cpp:5-14) plus the QNAME."""
    ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHI", rec, 4)
    name = rec[36:36 + l_name].split(b"\0")[0]
    rlen = 0
    for k in range(n_cig):
        v = struct.unpack_from("<I", rec, 36 + l_name + 4 * k)[0]
        if (v & 15) in CONSUMES_REF:
            rlen += v >> 4
    end = (pos + rlen - 1) % (1 << 64)  # hts_pos_t + uint64_t, cast to size_t
    return {"bam_id": bam_id, "start": pos % (1 << 64), "end": end, "quality": mapq,
            "seq_length": l_seq % (1 << 32), "is_first": bool(flag & 0x40), "qname": name}


# ----------------------------------------------------------------------------- call-site logic

def should_be_filtered_out(r1, r2, min_len=0, min_mapq=0, amplicons=None):
    """bam_api.cpp:311-332; amplicons = list of inclusive (start, end) or None (not FILTER mode)."""
    drop = not (r1["quality"] >= min_mapq and r2["quality"] >= min_mapq)
    drop = drop or not (r1["seq_length"] >= min_len and r2["seq_length"] >= min_len)
    if amplicons is not None:
        inc = lambda a, r: a[0] <= r["start"] and r["end"] <= a[1]
        drop = drop or not any(inc(a, r1) and inc(a, r2) for a in amplicons)
    return drop


def ref_read_bam(records, min_len=0, min_mapq=0, amplicons=None):
    """The loop of bam_api.cpp:425-478, statement by statement: returns (paired reads in
    pair-completion order, filtered_out bam ids ascending)."""
    paired, is_accepted, read_map = [], [], {}
    for bam_id, rec in enumerate(records):
        cur = parse_record(rec, bam_id)
        is_accepted.append(False)
        if cur["qname"] in read_map:
            r1, r2 = read_map[cur["qname"]], cur
            if should_be_filtered_out(r1, r2, min_len, min_mapq, amplicons):
                continue
            if r2["is_first"]:
                # std::swap(r1, r2) swaps the MAP ENTRY with the current read (:456-458)
                read_map[cur["qname"]] = r2
                r1, r2 = r2, r1
            paired += [r1, r2]
            is_accepted[r1["bam_id"]] = True
            is_accepted[r2["bam_id"]] = True
        else:
            read_map[cur["qname"]] = cur
    filtered_out = [i for i, a in enumerate(is_accepted) if not a]
    return paired, filtered_out


def ref_write_bam(records, bam_ids):
    """bam_api.cpp:603-623: sort ids, walk the file once, copy a record when its ordinal equals
    the next wanted id (a duplicated id stalls the iterator for good)."""
    ids = sorted(bam_ids)
    out, k = [], 0
    for i, rec in enumerate(records):
        if k == len(ids):
            break
        if i == ids[k]:
            out.append(rec)
            k += 1
    return out


# ----------------------------------------------------------------------------- SAM text

def format_sam(rec, ref_names):
    """One alignment line of SAM text for a BAM record (SAMv1 §1.4, §4.2.4) — what htslib's text
    writer emits when the reference opens a non-.bam output with mode "w" (bam_api.cpp:566)."""
    tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHIiii", rec, 4)
    p = 36
    name = rec[p:p + l_name].split(b"\0")[0].decode("latin-1")
    p += l_name
    cigar = "".join("%d%s" % (v >> 4, CIGAR_OPS[v & 15]) for v in struct.unpack_from("<%dI" % n_cig, rec, p)) or "*"
    p += 4 * n_cig
    seq = "".join("=ACMGRSVTWYHKDBN"[(rec[p + k // 2] >> (0 if k & 1 else 4)) & 15] for k in range(l_seq)) or "*"
    p += (l_seq + 1) // 2
    qual = "*" if l_seq == 0 or rec[p] == 0xff else "".join(chr(q + 33) for q in rec[p:p + l_seq])
    p += l_seq
    rname = ref_names[tid] if 0 <= tid < len(ref_names) else "*"
    rnext = "*" if mtid < 0 else "=" if mtid == tid else ref_names[mtid] if mtid < len(ref_names) else "*"
    cols = [name, str(flag), rname, str(pos + 1), str(mapq), cigar, rnext, str(mpos + 1), str(tlen), seq, qual]
    fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f", "d": "<d"}

    def scalar(t, p):
        v = struct.unpack_from(fmt[t], rec, p)[0]
        return ("%g" % v if t in "fd" else str(v)), p + struct.calcsize(fmt[t])

    while p + 3 <= len(rec):
        tag, t = rec[p:p + 2].decode("latin-1"), chr(rec[p + 2])
        p += 3
        if t == "A":
            cols.append("%s:A:%s" % (tag, chr(rec[p])))
            p += 1
        elif t in "ZH":
            e = rec.index(b"\0", p)
            cols.append("%s:%s:%s" % (tag, t, rec[p:e].decode("latin-1")))
            p = e + 1
        elif t == "B":
            sub, cnt = chr(rec[p]), struct.unpack_from("<I", rec, p + 1)[0]
            p += 5
            vals = []
            for _ in range(cnt):
                v, p = scalar(sub, p)
                vals.append(v)
            cols.append("%s:B:%s%s" % (tag, sub, "".join("," + v for v in vals)))
        else:
            v, p = scalar(t, p)
            cols.append("%s:%s:%s" % (tag, t if t in "fd" else "i", v))
    return "\t".join(cols) + "\n"
