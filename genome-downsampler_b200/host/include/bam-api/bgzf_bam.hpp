// BGZF/BAM record scanner and selective BAM copy — the front and back of the path
// (SURVEY §8(f) rows 2 and 3).  Replaces what the reference does through htslib in
// BamApi::read_bam (libs/bam-api/src/bam_api.cpp:359-507, sam_open/sam_hdr_read/sam_read1 and
// Read::Read(id, bam1_t*), libs/bam-api/src/read.cpp:5-14) and BamApi::write_bam
// (bam_api.cpp:534-656, sam_read1/sam_write1) with a zlib-only implementation: htslib is not
// in this image, zlib is.  Only the fields the path needs are decoded (pos, mapq, flag, l_seq,
// CIGAR reference length, QNAME); records are copied byte for byte on the way out.
//
// Layout of the work: the file is mapped, BGZF members are located by their BSIZE fields,
// inflated on `threads` host threads a chunk at a time (CRC32 checked as htslib does), a serial
// walk over the 4-byte block_size fields indexes the records of the chunk, and the per-record
// field extraction runs on the same threads into columns.  Nothing here touches the GPU.
#pragma once
#include <cstddef>
#include <cstdint>
#include <filesystem>
#include <functional>
#include <string>
#include <vector>

namespace bam_api::bgzf {

constexpr std::size_t kBlockPayload = 0xff00;  // htslib's BGZF_BLOCK_SIZE: payload bytes per member
constexpr std::size_t kMaxBlock = 0x10000;     // BGZF_MAX_BLOCK_SIZE

struct BamHeader {
    std::string raw;  // uncompressed bytes from "BAM\1" to the last l_ref, verbatim
    std::string text;
    std::vector<std::string> ref_names;
    std::vector<std::uint32_t> ref_lengths;
};

// one record's fields as Read::Read(id, bam1_t*) sees them, plus where it lies in the chunk
struct RecordFields {
    std::int32_t pos;
    std::uint32_t ref_len;  // bam_cigar2rlen: M, D, N, =, X consume the reference
    std::int32_t l_seq;
    std::uint16_t flag;
    std::uint8_t mapq;
    std::uint8_t l_qname;  // without the terminating NUL
    std::uint64_t qname_hash;
    std::uint64_t offset;  // of the record's block_size field in the chunk buffer
    std::uint32_t size;    // 4 + block_size
};

struct RecordChunk {
    const std::uint8_t* data = nullptr;      // uncompressed bytes of this chunk (carry + new blocks)
    std::uint64_t first_id = 0;              // BAM ordinal of records[0]
    std::vector<RecordFields> records;       // whole records only
    const char* qname(const RecordFields& r) const {
        return reinterpret_cast<const char*>(data + r.offset + 36);
    }
};

// Streams the records of a BAM file in file order.  Errors (unreadable file, bad magic, CRC
// mismatch, truncated member) throw std::runtime_error; the BamApi layer turns them into the
// reference's log-and-exit.
class BamScanner {
   public:
    BamScanner(const std::filesystem::path& path, std::uint32_t threads,
               std::size_t chunk_bytes = std::size_t(64) << 20);
    ~BamScanner();
    BamScanner(const BamScanner&) = delete;
    BamScanner& operator=(const BamScanner&) = delete;

    const BamHeader& header() const { return header_; }
    // next batch of whole records; false at end of file.  With fields == false only offset and
    // size are filled (the copy pass needs nothing else).  Chunks alternate between two buffers:
    // the bytes of a chunk stay valid until the call AFTER the next one, so a caller can work on
    // one chunk while another thread is inside next() for the following one (one next() at a time).
    bool next(RecordChunk& chunk, bool fields = true);
    bool saw_eof_marker() const { return saw_eof_marker_; }
    std::uint64_t compressed_bytes() const { return file_size_; }
    std::uint64_t uncompressed_bytes() const { return total_out_; }

   private:
    // inflate the next group of members behind the carried tail: into the other buffer, or (when
    // one next() needs more than one group) appended to the current one
    bool refill(bool in_place);
    void read_header();
    std::uint32_t threads_;
    std::size_t chunk_bytes_;
    int fd_ = -1;
    const std::uint8_t* file_ = nullptr;
    std::size_t file_size_ = 0;
    std::size_t file_pos_ = 0;
    std::vector<std::uint8_t> bufs_[2];
    int cur_ = 0;
    std::size_t buf_len_ = 0;   // valid bytes in buf_
    std::size_t consumed_ = 0;  // bytes of buf_ already handed out
    std::uint64_t next_id_ = 0;
    std::uint64_t total_out_ = 0;
    bool saw_eof_marker_ = false;
    BamHeader header_;
};

// Compressing side: payload is cut into members of at most kBlockPayload bytes exactly where
// htslib's bgzf_write/bgzf_flush_try cut them, groups of members are deflated on `threads`
// threads and written in order; close() appends the 28-byte EOF member.
class BgzfWriter {
   public:
    BgzfWriter(const std::filesystem::path& path, std::uint32_t threads, int level = -1);
    ~BgzfWriter();
    BgzfWriter(const BgzfWriter&) = delete;
    BgzfWriter& operator=(const BgzfWriter&) = delete;
    void write(const void* data, std::size_t n);  // bgzf_write: splits at the member size
    void flush_try(std::size_t n);                // bgzf_flush_try: new member if n does not fit
    void flush();                                 // bgzf_flush: close the current member
    void close();
    std::uint64_t bytes_written() const { return bytes_written_; }

   private:
    void compress_pending();  // deflate the closed members on the pool and write them in order
    std::FILE* f_ = nullptr;
    std::uint32_t threads_;
    int level_;
    std::vector<std::uint8_t> cur_;                  // the open member's payload
    std::vector<std::vector<std::uint8_t>> pending_;  // closed members waiting for deflate
    std::uint64_t bytes_written_ = 0;
};

// BamApi::write_bam (bam_api.cpp:534-656): header of `input` followed by the records whose file
// ordinals are in `bam_ids` (sorted in place, as the reference does), reading stops once the last
// requested record is out.  An output path that does not end in ".bam" gets SAM text, as the
// reference's open mode "w" does (:566).  Returns the number of records written.
std::uint32_t copy_bam_records(const std::filesystem::path& input,
                               const std::filesystem::path& output,
                               std::vector<std::size_t>& bam_ids, std::uint32_t threads);

// helper shared with the synthetic writer: run fn(i) for i in [0, n) on up to `threads` threads
void parallel_for(std::size_t n, std::uint32_t threads, const std::function<void(std::size_t)>& fn);

}  // namespace bam_api::bgzf
