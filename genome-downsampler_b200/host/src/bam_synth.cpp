// reads_gen::write_synthetic_bam — synthetic reads as a BAM file (see reads_gen.hpp).  Record
// layout per SAMv1 §4.2; members are cut and compressed by bam_api::bgzf::BgzfWriter.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>

#include "bam-api/bgzf_bam.hpp"
#include "reads_gen.hpp"

namespace reads_gen {

namespace {
inline void put32(std::vector<uint8_t>& v, uint32_t x) {
    for (int k = 0; k < 4; ++k) v.push_back(uint8_t(x >> (8 * k)));
}
// SAMv1 §5.3 reg2bin for [beg, end)
inline uint32_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return uint32_t(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return uint32_t(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return uint32_t(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return uint32_t(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return uint32_t(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
}  // namespace

uint64_t write_synthetic_bam(const std::filesystem::path& path, uint64_t n, uint32_t genome_length,
                             const uint32_t* start, const uint32_t* end, const uint8_t* mapq,
                             const uint32_t* seq_len, bool coordinate_sorted, uint32_t threads,
                             uint32_t seed) {
    bam_api::bgzf::BgzfWriter out(path, threads);
    std::vector<uint8_t> rec;
    {
        std::string text = "@HD\tVN:1.6\tSO:" + std::string(coordinate_sorted ? "coordinate" : "unsorted") +
                           "\n@SQ\tSN:synthetic\tLN:" + std::to_string(genome_length) + "\n";
        rec.insert(rec.end(), {'B', 'A', 'M', 1});
        put32(rec, uint32_t(text.size()));
        rec.insert(rec.end(), text.begin(), text.end());
        put32(rec, 1);
        const char name[] = "synthetic";
        put32(rec, sizeof name);
        rec.insert(rec.end(), name, name + sizeof name);
        put32(rec, genome_length);
        out.write(rec.data(), rec.size());
        out.flush();
    }
    std::vector<uint32_t> order;
    if (coordinate_sorted) {
        order.resize(n);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return start[a] < start[b]; });
    }
    uint64_t x = 0x9e3779b97f4a7c15ull ^ seed;
    auto rnd = [&x]() {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        return x;
    };
    for (uint64_t k = 0; k < n; ++k) {
        const uint64_t i = coordinate_sorted ? order[k] : k;
        const uint32_t span = end[i] - start[i] + 1;
        const uint32_t lseq = seq_len ? seq_len[i] : span;
        uint32_t cigar[2];
        uint32_t n_cigar = 1;
        if (lseq == span) {
            cigar[0] = span << 4;  // M
        } else if (lseq > span) {
            cigar[0] = span << 4;
            cigar[1] = ((lseq - span) << 4) | 4;  // S
            n_cigar = 2;
        } else {
            cigar[0] = lseq << 4;
            cigar[1] = ((span - lseq) << 4) | 2;  // D
            n_cigar = 2;
        }
        const std::string qname = "p" + std::to_string(i / 2);
        const uint32_t l_read_name = uint32_t(qname.size()) + 1;
        const uint64_t mate = i ^ 1;
        const uint32_t block = 32 + l_read_name + 4 * n_cigar + (lseq + 1) / 2 + lseq;
        rec.clear();
        put32(rec, block);
        put32(rec, 0);                                       // refID
        put32(rec, start[i]);                                // pos
        rec.push_back(uint8_t(l_read_name));
        rec.push_back(mapq ? mapq[i] : 60);
        uint32_t bin = reg2bin(start[i], int64_t(end[i]) + 1);
        rec.push_back(uint8_t(bin));
        rec.push_back(uint8_t(bin >> 8));
        rec.push_back(uint8_t(n_cigar));
        rec.push_back(0);
        uint32_t flag = 0x1 | 0x2 | ((i & 1) ? 0x80 : 0x40) | ((i & 1) ? 0x10 : 0x20);
        rec.push_back(uint8_t(flag));
        rec.push_back(uint8_t(flag >> 8));
        put32(rec, lseq);
        put32(rec, 0);                                       // next refID
        put32(rec, mate < n ? start[mate] : 0xffffffffu);    // next pos
        put32(rec, 0);                                       // tlen
        rec.insert(rec.end(), qname.begin(), qname.end());
        rec.push_back(0);
        for (uint32_t c = 0; c < n_cigar; ++c) put32(rec, cigar[c]);
        static const uint8_t nib[4] = {1, 2, 4, 8};  // A C G T
        for (uint32_t b = 0; b < (lseq + 1) / 2; ++b) {
            uint64_t r = rnd();
            rec.push_back(uint8_t(nib[r & 3] << 4 | nib[(r >> 2) & 3]));
        }
        // binned qualities as current instruments report them: mostly 37, some 25 and 11
        for (uint32_t b = 0; b < lseq; ++b) {
            uint32_t r = uint32_t(rnd() >> 40) % 100;
            rec.push_back(uint8_t(r < 90 ? 37 : r < 97 ? 25 : 11));
        }
        out.flush_try(rec.size());
        out.write(rec.data(), rec.size());
    }
    out.close();
    return n;
}

}  // namespace reads_gen
