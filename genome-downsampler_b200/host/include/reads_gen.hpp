// reads_gen — synthetic paired reads with the laws of libs/reads-gen/src/reads_gen.cpp:5-86
// (same libstdc++ engines and distributions, same draw order => same streams for equal seeds).
#pragma once
#include <cstdint>
#include <filesystem>
#include <functional>
#include <random>
#include <vector>

#include "bam-api/paired_reads.hpp"

namespace reads_gen {
constexpr int32_t kMaxGenQuality = 100;

bam_api::AOSPairedReads rand_reads(std::mt19937& generator, bam_api::ReadIndex pairs_count,
                                   bam_api::Index genome_length, uint32_t read_length,
                                   const std::function<double(double)>& dist_func,
                                   int32_t max_quality = kMaxGenQuality);

bam_api::AOSPairedReads rand_reads_uniform(std::mt19937& generator, bam_api::ReadIndex pairs_count,
                                           bam_api::Index genome_length, uint32_t read_length,
                                           int32_t max_quality = kMaxGenQuality);

// Same streams written straight into caller-owned SoA arrays (index = read index; mapq/seq_len
// may be null).  Allocation-free, so independent samples can be generated on many host threads.
struct SoaOut {
    uint32_t* start;
    uint32_t* end;
    uint8_t* mapq;
    uint32_t* seq_len;
};
void rand_reads_uniform_soa(std::mt19937& generator, bam_api::ReadIndex pairs_count,
                            bam_api::Index genome_length, uint32_t read_length, const SoaOut& out,
                            int32_t max_quality = kMaxGenQuality);
void rand_reads_soa(std::mt19937& generator, bam_api::ReadIndex pairs_count,
                    bam_api::Index genome_length, uint32_t read_length,
                    const std::function<double(double)>& dist_func, const SoaOut& out,
                    int32_t max_quality = kMaxGenQuality);

// Amplicon-aware extension for BASELINE config 2 (the reference's generator cannot express it:
// constant seq_length and mates placed over the whole genome, reads_gen.cpp:46-49,79-82).
// Per pair, in this draw order: len1, len2 ~ U{min_len..max_len}; u ~ U[0,1); amplicon k ~ U;
// if u < p_inside and the amplicon holds both mates, both starts ~ U inside amplicon k, else
// both starts ~ U over the genome; then quality1, quality2 ~ U{0..max_quality}.
void rand_reads_amplicon_soa(std::mt19937& generator, bam_api::ReadIndex pairs_count,
                             bam_api::Index genome_length, const std::vector<uint32_t>& amp_start,
                             const std::vector<uint32_t>& amp_end, double p_inside,
                             uint32_t min_len, uint32_t max_len, const SoaOut& out,
                             int32_t max_quality = kMaxGenQuality);

// Synthetic reads as a single-contig BAM file, so the BAM front end of the path can be measured
// and tested on the same inputs (the reference has no such writer: its generator only feeds the
// in-memory BamApi constructors).  Read i becomes one record: QNAME "p<i/2>", FLAG paired +
// first/second mate, CIGAR <seq_len>M (plus D or S so that it spans start..end exactly), random
// bases, binned qualities (37/25/11), no tags.  coordinate_sorted orders the records by start (stable), which
// separates mates the way a sorted BAM does; otherwise mates are adjacent in file order.
// Returns the number of records written.
uint64_t write_synthetic_bam(const std::filesystem::path& path, uint64_t n, uint32_t genome_length,
                             const uint32_t* start, const uint32_t* end, const uint8_t* mapq,
                             const uint32_t* seq_len, bool coordinate_sorted, uint32_t threads,
                             uint32_t seed = 1);
}  // namespace reads_gen
