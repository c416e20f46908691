// zlib-only BGZF/BAM scanner and selective copy (see bam-api/bgzf_bam.hpp).  Formats follow the
// SAM/BAM specification (SAMv1 §4.1 BGZF, §4.2 BAM); behaviour at the call sites follows
// libs/bam-api/src/bam_api.cpp:359-507 (read) and :534-656 (write) of the reference.
#include "bam-api/bgzf_bam.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

namespace bam_api::bgzf {

namespace {

inline std::uint16_t le16(const std::uint8_t* p) { return std::uint16_t(p[0] | (p[1] << 8)); }
inline std::uint32_t le32(const std::uint8_t* p) {
    return std::uint32_t(p[0]) | (std::uint32_t(p[1]) << 8) | (std::uint32_t(p[2]) << 16) |
           (std::uint32_t(p[3]) << 24);
}
inline void put16(std::uint8_t* p, std::uint32_t v) {
    p[0] = std::uint8_t(v);
    p[1] = std::uint8_t(v >> 8);
}
inline void put32(std::uint8_t* p, std::uint32_t v) {
    put16(p, v);
    put16(p + 2, v >> 16);
}

[[noreturn]] void fail(const std::string& what) { throw std::runtime_error(what); }

// one BGZF member located in the mapped file
struct Member {
    const std::uint8_t* cdata;
    std::uint32_t clen;
    std::uint32_t isize;
    std::uint32_t crc;
    std::size_t out_off;
};

// Parses the gzip member header at p (SAMv1 §4.1): returns the total member size, fills m.
std::size_t parse_member(const std::uint8_t* p, std::size_t avail, Member& m) {
    if (avail < 18) fail("truncated BGZF member header");
    if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) fail("not a BGZF member (bad gzip magic/flags)");
    std::uint32_t xlen = le16(p + 10);
    if (avail < 12 + std::size_t(xlen)) fail("truncated BGZF extra field");
    std::uint32_t bsize = 0;
    bool found = false;
    for (std::uint32_t off = 0; off + 4 <= xlen;) {
        const std::uint8_t* s = p + 12 + off;
        std::uint32_t slen = le16(s + 2);
        if (s[0] == 'B' && s[1] == 'C' && slen == 2 && off + 6 <= xlen) {
            bsize = le16(s + 4);
            found = true;
        }
        off += 4 + slen;
    }
    if (!found) fail("gzip member without the BGZF BC subfield");
    std::size_t total = std::size_t(bsize) + 1;
    if (total > avail) fail("truncated BGZF member");
    if (total < 12 + std::size_t(xlen) + 8) fail("corrupt BGZF member size");
    m.cdata = p + 12 + xlen;
    m.clen = std::uint32_t(total - 12 - xlen - 8);
    m.crc = le32(p + total - 8);
    m.isize = le32(p + total - 4);
    if (m.isize > kMaxBlock) fail("BGZF member larger than 64 KiB");
    return total;
}

void inflate_member(z_stream& zs, const Member& m, std::uint8_t* out) {
    if (m.isize == 0) return;
    if (inflateReset(&zs) != Z_OK) fail("inflateReset failed");
    zs.next_in = const_cast<Bytef*>(m.cdata);
    zs.avail_in = m.clen;
    zs.next_out = out;
    zs.avail_out = m.isize;
    int rc = inflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END || zs.avail_out != 0) fail("BGZF inflate failed");
    if (crc32(crc32(0L, Z_NULL, 0), out, m.isize) != m.crc) fail("BGZF CRC32 mismatch");
}

// bam_cigar2rlen: operations M(0) D(2) N(3) =(7) X(8) consume the reference
inline std::uint32_t cigar_ref_len(const std::uint8_t* cig, std::uint32_t n) {
    constexpr std::uint32_t consumes = (1u << 0) | (1u << 2) | (1u << 3) | (1u << 7) | (1u << 8);
    std::uint32_t len = 0;
    for (std::uint32_t k = 0; k < n; ++k) {
        std::uint32_t v = le32(cig + 4 * k);
        if ((consumes >> (v & 15)) & 1) len += v >> 4;
    }
    return len;
}

inline std::uint64_t hash_bytes(const std::uint8_t* p, std::size_t n) {
    // 8 bytes at a time, multiply-xorshift mix (only has to spread QNAMEs over a table; the
    // pairing compares the bytes as well)
    std::uint64_t h = 0x9e3779b97f4a7c15ull ^ (n * 0xff51afd7ed558ccdull);
    while (n >= 8) {
        std::uint64_t w;
        std::memcpy(&w, p, 8);
        h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
        h ^= h >> 29;
        p += 8;
        n -= 8;
    }
    std::uint64_t w = 0;
    std::memcpy(&w, p, n);
    h = (h ^ w) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    return h;
}

const std::uint8_t kEofMember[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                     0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

}  // namespace

void parallel_for(std::size_t n, std::uint32_t threads, const std::function<void(std::size_t)>& fn) {
    std::size_t t = std::min<std::size_t>(std::max<std::uint32_t>(threads, 1), n);
    if (t <= 1) {
        for (std::size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<std::size_t> next{0};
    std::exception_ptr err;
    std::atomic<bool> failed{false};
    auto body = [&]() {
        try {
            for (std::size_t i; (i = next.fetch_add(1)) < n && !failed.load();) fn(i);
        } catch (...) {
            if (!failed.exchange(true)) err = std::current_exception();
        }
    };
    std::vector<std::thread> pool;
    for (std::size_t k = 1; k < t; ++k) pool.emplace_back(body);
    body();
    for (auto& th : pool) th.join();
    if (failed.load()) std::rethrow_exception(err);
}

// ------------------------------------------------------------------ scanner

BamScanner::BamScanner(const std::filesystem::path& path, std::uint32_t threads, std::size_t chunk_bytes)
    : threads_(std::max<std::uint32_t>(threads, 1)), chunk_bytes_(std::max<std::size_t>(chunk_bytes, kMaxBlock)) {
    // testing knob: small chunks make short files cross many chunk boundaries
    if (const char* e = std::getenv("GDS_BAM_CHUNK_BYTES")) chunk_bytes_ = std::max<std::size_t>(std::strtoull(e, nullptr, 10), 1);
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) fail("Could not open " + path.string());
    struct stat st {};
    if (fstat(fd_, &st) != 0 || st.st_size <= 0) {
        ::close(fd_);
        fd_ = -1;
        fail("Could not stat " + path.string());
    }
    file_size_ = std::size_t(st.st_size);
    void* p = mmap(nullptr, file_size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (p == MAP_FAILED) {
        ::close(fd_);
        fd_ = -1;
        fail("Could not map " + path.string());
    }
    file_ = static_cast<const std::uint8_t*>(p);
    madvise(p, file_size_, MADV_SEQUENTIAL);
    try {
        read_header();
    } catch (...) {
        munmap(const_cast<std::uint8_t*>(file_), file_size_);
        ::close(fd_);
        throw;
    }
}

BamScanner::~BamScanner() {
    if (file_) munmap(const_cast<std::uint8_t*>(file_), file_size_);
    if (fd_ >= 0) ::close(fd_);
}

bool BamScanner::refill(bool in_place) {
    if (file_pos_ >= file_size_) return false;
    std::vector<Member> members;
    std::size_t add = 0;
    while (file_pos_ < file_size_ && add < chunk_bytes_) {
        Member m{};
        std::size_t total = parse_member(file_ + file_pos_, file_size_ - file_pos_, m);
        saw_eof_marker_ = (m.isize == 0 && file_pos_ + total == file_size_);
        file_pos_ += total;
        m.out_off = buf_len_ + add;
        add += m.isize;
        members.push_back(m);
    }
    // the unconsumed tail goes to the front of the target buffer, the new members behind it
    const std::size_t tail = buf_len_ - consumed_;
    std::vector<std::uint8_t>& src = bufs_[cur_];
    std::vector<std::uint8_t>& dst = in_place ? src : bufs_[1 - cur_];
    if (dst.size() < tail + add) dst.resize(std::max(tail + add, dst.size() + dst.size() / 2));
    if (tail > 0 && (consumed_ > 0 || &dst != &src)) std::memmove(dst.data(), src.data() + consumed_, tail);
    for (Member& m : members) m.out_off -= consumed_;
    if (!in_place) cur_ = 1 - cur_;
    buf_len_ = tail;
    consumed_ = 0;
    std::uint8_t* base = dst.data();
    // a few members per task so threads do not meet on the counter for every 64 KiB
    constexpr std::size_t kGroup = 8;
    std::size_t groups = (members.size() + kGroup - 1) / kGroup;
    parallel_for(groups, threads_, [&](std::size_t g) {
        z_stream zs{};
        if (inflateInit2(&zs, -15) != Z_OK) fail("inflateInit2 failed");
        try {
            for (std::size_t i = g * kGroup; i < std::min(members.size(), (g + 1) * kGroup); ++i)
                inflate_member(zs, members[i], base + members[i].out_off);
        } catch (...) {
            inflateEnd(&zs);
            throw;
        }
        inflateEnd(&zs);
    });
    buf_len_ += add;
    total_out_ += add;
    return true;
}

void BamScanner::read_header() {
    // bytes needed so far are pulled in chunk by chunk; headers are small next to a chunk
    auto need = [&](std::size_t n) {
        while (buf_len_ - consumed_ < n)
            if (!refill(true)) fail("Failed to read header from file!");
    };
    std::size_t start = 0;
    need(12);
    const std::uint8_t* b = bufs_[cur_].data() + consumed_;
    if (std::memcmp(b, "BAM\1", 4) != 0) fail("Failed to read header from file! (not a BAM file)");
    std::uint32_t l_text = le32(b + 4);
    need(12 + std::size_t(l_text));
    b = bufs_[cur_].data() + consumed_;
    header_.text.assign(reinterpret_cast<const char*>(b + 8), l_text);
    std::uint32_t n_ref = le32(b + 8 + l_text);
    std::size_t off = 12 + std::size_t(l_text);
    for (std::uint32_t r = 0; r < n_ref; ++r) {
        need(off + 4);
        b = bufs_[cur_].data() + consumed_;
        std::uint32_t l_name = le32(b + off);
        need(off + 8 + std::size_t(l_name));
        b = bufs_[cur_].data() + consumed_;
        const char* nm = reinterpret_cast<const char*>(b + off + 4);
        header_.ref_names.emplace_back(nm, l_name ? strnlen(nm, l_name) : 0);
        header_.ref_lengths.push_back(le32(b + off + 4 + l_name));
        off += 8 + std::size_t(l_name);
    }
    b = bufs_[cur_].data() + consumed_;
    header_.raw.assign(reinterpret_cast<const char*>(b + start), off);
    consumed_ += off;
}

bool BamScanner::next(RecordChunk& chunk, bool fields) {
    chunk.records.clear();
    bool switched = false;  // the previous chunk's buffer must survive this call
    for (;;) {
        // serial index of the whole records now in the buffer
        std::size_t p = consumed_;
        const std::uint8_t* base = bufs_[cur_].data();
        while (p + 4 <= buf_len_) {
            std::uint32_t bs = le32(base + p);
            if (bs < 32) fail("corrupt BAM record (block_size < 32)");
            if (p + 4 + std::size_t(bs) > buf_len_) break;
            RecordFields r{};
            r.offset = p;
            r.size = 4 + bs;
            chunk.records.push_back(r);
            p += 4 + std::size_t(bs);
        }
        if (!chunk.records.empty()) {
            chunk.data = base;
            chunk.first_id = next_id_;
            next_id_ += chunk.records.size();
            consumed_ = p;
            break;
        }
        if (!refill(switched)) {
            if (buf_len_ != consumed_) fail("truncated BAM record at end of file");
            return false;
        }
        switched = true;
    }
    if (fields) {
        const std::uint8_t* base = chunk.data;
        constexpr std::size_t kGrain = 16384;
        std::size_t n = chunk.records.size();
        parallel_for((n + kGrain - 1) / kGrain, threads_, [&](std::size_t g) {
            for (std::size_t i = g * kGrain; i < std::min(n, (g + 1) * kGrain); ++i) {
                RecordFields& r = chunk.records[i];
                const std::uint8_t* q = base + r.offset + 4;  // refID
                std::uint32_t l_read_name = q[8];
                std::uint32_t n_cigar = le16(q + 12);
                std::size_t var = 32 + std::size_t(l_read_name) + 4 * std::size_t(n_cigar);
                if (var > r.size - 4) fail("corrupt BAM record (name/CIGAR exceed block_size)");
                r.pos = std::int32_t(le32(q + 4));
                r.mapq = q[9];
                r.flag = le16(q + 14);
                r.l_seq = std::int32_t(le32(q + 16));
                const std::uint8_t* name = q + 32;
                std::uint32_t nl = l_read_name ? std::uint32_t(strnlen(reinterpret_cast<const char*>(name), l_read_name)) : 0;
                r.l_qname = std::uint8_t(nl);
                r.qname_hash = hash_bytes(name, nl);
                r.ref_len = cigar_ref_len(name + l_read_name, n_cigar);
            }
        });
    }
    return true;
}

// ------------------------------------------------------------------ writer

BgzfWriter::BgzfWriter(const std::filesystem::path& path, std::uint32_t threads, int level)
    : threads_(std::max<std::uint32_t>(threads, 1)), level_(level) {
    f_ = std::fopen(path.c_str(), "wb");
    if (!f_) fail("Could not open " + path.string());
    cur_.reserve(kBlockPayload);
}

BgzfWriter::~BgzfWriter() {
    if (f_) std::fclose(f_);  // close() not called: error path, the file is left incomplete
}

void BgzfWriter::flush() {
    if (cur_.empty()) return;
    pending_.emplace_back(std::move(cur_));
    cur_.clear();
    cur_.reserve(kBlockPayload);
    if (pending_.size() >= std::size_t(threads_) * 32) compress_pending();
}

void BgzfWriter::flush_try(std::size_t n) {
    if (cur_.size() + n > kBlockPayload) flush();
}

void BgzfWriter::write(const void* data, std::size_t n) {
    const std::uint8_t* p = static_cast<const std::uint8_t*>(data);
    while (n > 0) {
        std::size_t take = std::min(n, kBlockPayload - cur_.size());
        cur_.insert(cur_.end(), p, p + take);
        p += take;
        n -= take;
        if (cur_.size() == kBlockPayload) flush();
    }
}

void BgzfWriter::compress_pending() {
    if (pending_.empty()) return;
    std::vector<std::vector<std::uint8_t>> out(pending_.size());
    parallel_for(pending_.size(), threads_, [&](std::size_t i) {
        const auto& in = pending_[i];
        std::vector<std::uint8_t>& o = out[i];
        o.resize(kMaxBlock);
        int level = level_;
        for (;;) {
            z_stream zs{};
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) fail("deflateInit2 failed");
            zs.next_in = const_cast<Bytef*>(in.data());
            zs.avail_in = uInt(in.size());
            zs.next_out = o.data() + 18;
            zs.avail_out = uInt(kMaxBlock - 18 - 8);
            int rc = deflate(&zs, Z_FINISH);
            std::size_t clen = zs.total_out;
            deflateEnd(&zs);
            if (rc == Z_STREAM_END) {
                static const std::uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
                std::memcpy(o.data(), head, 16);
                std::size_t total = 18 + clen + 8;
                put16(o.data() + 16, std::uint32_t(total - 1));
                put32(o.data() + 18 + clen, std::uint32_t(crc32(crc32(0L, Z_NULL, 0), in.data(), uInt(in.size()))));
                put32(o.data() + 22 + clen, std::uint32_t(in.size()));
                o.resize(total);
                return;
            }
            // incompressible payload at this level: stored blocks always fit (0xff00 + 5 < 64 KiB - 26)
            if (level == 0) fail("BGZF deflate failed");
            level = 0;
        }
    });
    for (const auto& o : out) {
        if (std::fwrite(o.data(), 1, o.size(), f_) != o.size()) fail("short write to BAM output");
        bytes_written_ += o.size();
    }
    pending_.clear();
}

void BgzfWriter::close() {
    if (!f_) return;
    flush();
    compress_pending();
    bool ok = std::fwrite(kEofMember, 1, sizeof kEofMember, f_) == sizeof kEofMember;
    bytes_written_ += sizeof kEofMember;
    ok = (std::fclose(f_) == 0) && ok;
    f_ = nullptr;
    if (!ok) fail("could not finish BAM output");
}

// ------------------------------------------------------------------ SAM text

namespace {

void put_int(std::string& o, long long v) { o += std::to_string(v); }

void put_real(std::string& o, double v) {
    char b[32];
    std::snprintf(b, sizeof b, "%g", v);
    o += b;
}

// one alignment line as sam_format1 prints it (SAMv1 §1.4 / §4.2.4); rec points at block_size
void format_sam_record(const std::uint8_t* rec, std::uint32_t size, const BamHeader& h, std::string& o) {
    const std::uint8_t* q = rec + 4;
    const std::int32_t tid = std::int32_t(le32(q)), pos = std::int32_t(le32(q + 4));
    const std::uint32_t l_name = q[8], mapq = q[9], n_cigar = le16(q + 12), flag = le16(q + 14);
    const std::uint32_t l_seq = le32(q + 16);
    const std::int32_t mtid = std::int32_t(le32(q + 20)), mpos = std::int32_t(le32(q + 24)), tlen = std::int32_t(le32(q + 28));
    const std::uint8_t* name = q + 32;
    const std::uint8_t* cig = name + l_name;
    const std::uint8_t* seq = cig + 4 * std::size_t(n_cigar);
    const std::uint8_t* qual = seq + (l_seq + 1) / 2;
    const std::uint8_t* aux = qual + l_seq;
    const std::uint8_t* end = rec + size;
    if (aux > end) fail("corrupt BAM record (fields exceed block_size)");
    auto ref_name = [&](std::int32_t t) -> std::string {
        return (t >= 0 && std::size_t(t) < h.ref_names.size()) ? h.ref_names[std::size_t(t)] : std::string("*");
    };
    o.append(reinterpret_cast<const char*>(name), l_name ? strnlen(reinterpret_cast<const char*>(name), l_name) : 0);
    o += '\t';
    put_int(o, flag);
    o += '\t';
    o += ref_name(tid);
    o += '\t';
    put_int(o, (long long)pos + 1);
    o += '\t';
    put_int(o, mapq);
    o += '\t';
    if (n_cigar == 0) o += '*';
    for (std::uint32_t k = 0; k < n_cigar; ++k) {
        std::uint32_t v = le32(cig + 4 * k);
        put_int(o, v >> 4);
        o += "MIDNSHP=XB??????"[v & 15];
    }
    o += '\t';
    if (mtid < 0) o += '*';
    else if (mtid == tid) o += '=';
    else o += ref_name(mtid);
    o += '\t';
    put_int(o, (long long)mpos + 1);
    o += '\t';
    put_int(o, tlen);
    o += '\t';
    if (l_seq == 0) o += '*';
    for (std::uint32_t k = 0; k < l_seq; ++k) o += "=ACMGRSVTWYHKDBN"[(seq[k >> 1] >> ((~k & 1) << 2)) & 15];
    o += '\t';
    if (l_seq == 0 || qual[0] == 0xff) o += '*';
    else
        for (std::uint32_t k = 0; k < l_seq; ++k) o += char(qual[k] + 33);
    auto scalar = [&](char type, const std::uint8_t*& p) {
        auto need = [&](std::size_t n) {
            if (p + n > end) fail("corrupt BAM record (truncated tag)");
        };
        switch (type) {
            case 'c': need(1); put_int(o, std::int8_t(*p)); p += 1; break;
            case 'C': need(1); put_int(o, *p); p += 1; break;
            case 's': need(2); put_int(o, std::int16_t(le16(p))); p += 2; break;
            case 'S': need(2); put_int(o, le16(p)); p += 2; break;
            case 'i': need(4); put_int(o, std::int32_t(le32(p))); p += 4; break;
            case 'I': need(4); put_int(o, le32(p)); p += 4; break;
            case 'f': {
                need(4);
                float f;
                std::uint32_t u = le32(p);
                std::memcpy(&f, &u, 4);
                put_real(o, f);
                p += 4;
                break;
            }
            case 'd': {
                need(8);
                double d;
                std::uint64_t u = std::uint64_t(le32(p)) | (std::uint64_t(le32(p + 4)) << 32);
                std::memcpy(&d, &u, 8);
                put_real(o, d);
                p += 8;
                break;
            }
            default: fail("corrupt BAM record (unknown tag type)");
        }
    };
    for (const std::uint8_t* p = aux; p + 3 <= end;) {
        o += '\t';
        o += char(p[0]);
        o += char(p[1]);
        o += ':';
        const char type = char(p[2]);
        p += 3;
        if (type == 'A') {
            if (p >= end) fail("corrupt BAM record (truncated tag)");
            o += "A:";
            o += char(*p++);
        } else if (type == 'Z' || type == 'H') {
            o += type;
            o += ':';
            while (p < end && *p) o += char(*p++);
            if (p >= end) fail("corrupt BAM record (unterminated tag)");
            ++p;
        } else if (type == 'B') {
            if (p + 5 > end) fail("corrupt BAM record (truncated tag)");
            const char sub = char(*p);
            std::uint32_t cnt = le32(p + 1);
            p += 5;
            o += "B:";
            o += sub;
            for (std::uint32_t k = 0; k < cnt; ++k) {
                o += ',';
                scalar(sub, p);
            }
        } else if (type == 'f' || type == 'd') {
            o += type;
            o += ':';
            scalar(type, p);
        } else {
            o += "i:";
            scalar(type, p);
        }
    }
    o += '\n';
}

}  // namespace

// ------------------------------------------------------------------ selective copy

namespace {
// the same loop with htslib's SAM text writer on the other end (any extension but .bam)
std::uint32_t copy_as_sam(BamScanner& in, const std::filesystem::path& output, std::vector<std::size_t>& bam_ids) {
    std::FILE* f = std::fopen(output.c_str(), "wb");
    if (!f) fail("Could not open " + output.string());
    std::string text = in.header().text;
    text.resize(strnlen(text.c_str(), text.size()));  // NUL padding is not part of the text
    std::string buf = text;
    std::sort(bam_ids.begin(), bam_ids.end());
    auto want = bam_ids.begin();
    std::uint32_t written = 0;
    RecordChunk chunk;
    bool ok = true;
    while (ok && want != bam_ids.end() && in.next(chunk, false)) {
        std::uint64_t id = chunk.first_id;
        for (const RecordFields& r : chunk.records) {
            if (want == bam_ids.end()) break;
            if (id == *want) {
                format_sam_record(chunk.data + r.offset, r.size, in.header(), buf);
                ++written;
                ++want;
                if (buf.size() > (1u << 20)) {
                    ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
                    buf.clear();
                }
            }
            ++id;
        }
    }
    ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) fail("short write to " + output.string());
    return written;
}
}  // namespace

std::uint32_t copy_bam_records(const std::filesystem::path& input, const std::filesystem::path& output,
                               std::vector<std::size_t>& bam_ids, std::uint32_t threads) {
    BamScanner in(input, threads);
    if (output.extension() != ".bam") return copy_as_sam(in, output, bam_ids);  // mode "w", :566
    BgzfWriter out(output, threads);
    // sam_hdr_write → bam_hdr_write: header bytes, then bgzf_flush
    out.write(in.header().raw.data(), in.header().raw.size());
    out.flush();
    std::sort(bam_ids.begin(), bam_ids.end());  // bam_api.cpp:604
    auto want = bam_ids.begin();
    std::uint32_t written = 0;
    RecordChunk chunk;
    // the reference's loop reads while ids remain (bam_api.cpp:608-623); an id listed twice
    // stalls its iterator and ends the copy at end of file — same here
    while (want != bam_ids.end() && in.next(chunk, false)) {
        std::uint64_t id = chunk.first_id;
        for (const RecordFields& r : chunk.records) {
            if (want == bam_ids.end()) break;
            if (id == *want) {
                // bam_write1: bgzf_flush_try(4 + block_len) then bgzf_write
                out.flush_try(r.size);
                out.write(chunk.data + r.offset, r.size);
                ++written;
                ++want;
            }
            ++id;
        }
    }
    out.close();
    return written;
}

}  // namespace bam_api::bgzf
