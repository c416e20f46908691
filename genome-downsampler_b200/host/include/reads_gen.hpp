// reads_gen — synthetic paired reads with the laws of libs/reads-gen/src/reads_gen.cpp:5-86
// (same libstdc++ engines and distributions, same draw order => same streams for equal seeds).
#pragma once
#include <cstdint>
#include <functional>
#include <random>

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
}  // namespace reads_gen
