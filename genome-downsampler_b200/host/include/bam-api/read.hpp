// bam_api::Read and index typedefs — same names and field meaning as the reference
// (libs/bam-api/include/bam-api/read.hpp:11-30), without the htslib dependency.
#pragma once
#include <cstddef>
#include <cstdint>

namespace bam_api {
using BAMReadId = std::size_t;   // record ordinal in the BAM file
using ReadIndex = std::size_t;   // index in the in-memory arrays
using Index = std::size_t;       // 0-based reference position
using ReadQuality = std::uint32_t;

struct Read {
    BAMReadId bam_id = 0;
    Index start_ind = 0;
    Index end_ind = 0;  // inclusive
    ReadQuality quality = 0;
    std::uint32_t seq_length = 0;
    bool is_first_read = false;

    Read() = default;
    Read(BAMReadId id, Index start, Index end, ReadQuality q, std::uint32_t len, bool first)
        : bam_id(id), start_ind(start), end_ind(end), quality(q), seq_length(len),
          is_first_read(first) {}
};
}  // namespace bam_api
