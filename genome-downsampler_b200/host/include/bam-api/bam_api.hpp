// bam_api::BamApi — host-side data access mirror of libs/bam-api/include/bam-api/bam_api.hpp for
// the quasi-MCP path.  In-memory construction (bam_api.cpp:44-47) carries the solver tests and
// benchmarks; BAM files are read and written (bam_api.cpp:359-656) through the zlib-only scanner
// of bam-api/bgzf_bam.hpp, since htslib is absent from this image.
//
// Addition for the B200 path: a BamApi can hold UNFILTERED pair-ordered reads plus the filter
// settings ("pending filter").  A device solver then runs the filter on the GPU and hands the
// per-pair verdicts back through apply_pair_filter(), after which the object is in exactly the
// state BamApi::read_bam leaves it in (filtered arrays + filtered_out_reads_).  Any CPU consumer
// that asks for the reads first triggers the same filter on the host.
#pragma once
#include <cstdint>
#include <filesystem>
#include <string>
#include <utility>
#include <vector>

#include "bam-api/paired_reads.hpp"

namespace bam_api {

enum class AmpliconBehaviour { IGNORE, FILTER, GRADE };

struct BamApiConfig {
    std::filesystem::path bed_filepath;
    std::filesystem::path tsv_filepath;
    std::uint32_t hts_thread_count = 1;
    std::uint32_t min_seq_length = 0;
    std::uint32_t min_mapq = 0;
    AmpliconBehaviour amplicon_behaviour = AmpliconBehaviour::IGNORE;
};

struct Amplicon {
    Index start;
    Index end;  // inclusive, as the reference treats it (amplicon.cpp:5-7)
    bool includes(const Read& r) const { return start <= r.start_ind && r.end_ind <= end; }
};

struct AmpliconSet {
    std::vector<Amplicon> amplicons;
    bool member_includes_both(const Read& a, const Read& b) const;
};

class BamApi {
   public:
    // file-backed, as in the reference: the BAM is read on the first request for its reads.  The
    // reads stay UNFILTERED ("pending filter", below) until a device solver or a CPU consumer
    // asks, so the filter can run on the GPU without changing what any caller observes.
    BamApi(const std::filesystem::path& input_filepath, const BamApiConfig& config);
    explicit BamApi(const AOSPairedReads& paired_reads);
    explicit BamApi(const SOAPairedReads& paired_reads);
    // unfiltered pair-ordered reads + filter settings (BED/TSV parsed on the host, bam_api.cpp:53-187)
    BamApi(const SOAPairedReads& unfiltered, const BamApiConfig& config);

    void set_amplicon_behaviour(AmpliconBehaviour b) { amplicon_behaviour_ = b; }

    const AOSPairedReads& get_paired_reads_aos();
    const SOAPairedReads& get_paired_reads_soa();
    const PairedReads& get_paired_reads() const;
    const std::vector<BAMReadId>& get_filtered_out_reads() const { return filtered_out_reads_; }
    std::vector<ReadIndex> find_pairs(const std::vector<ReadIndex>& ids) const;
    // returns number of reads written (bam_api.cpp:509-532)
    std::uint32_t write_paired_reads(const std::filesystem::path& output_filepath,
                                     std::vector<ReadIndex>& active_ids) const;
    std::uint32_t write_bam_api_filtered_out_reads(const std::filesystem::path& output_filepath);
    // private in the reference (bam_api.cpp:534-656); public here so tests can reach it
    static std::uint32_t write_bam(const std::filesystem::path& input_filepath,
                                   const std::filesystem::path& output_filepath,
                                   std::vector<BAMReadId>& bam_ids, std::uint32_t hts_thread_count);

    std::vector<std::uint32_t> find_input_cover();
    std::vector<std::uint32_t> find_filtered_cover(const std::vector<ReadIndex>& active_ids);

    // ---- B200 path hooks ----
    bool has_pending_filter() {
        ensure_loaded();
        return pending_filter_;
    }
    const SOAPairedReads& unfiltered_soa() {
        ensure_loaded();
        return soa_paired_reads_;
    }
    // records in the input BAM (0 for in-memory construction) and timings of the last read_bam
    std::uint64_t bam_record_count() const { return bam_record_count_; }
    double read_bam_seconds() const { return read_bam_seconds_; }
    std::uint32_t min_seq_length() const { return min_seq_length_; }
    std::uint32_t min_mapq() const { return min_mapq_; }
    AmpliconBehaviour amplicon_behaviour() const { return amplicon_behaviour_; }
    const AmpliconSet& amplicon_set() const { return amplicon_set_; }
    // pair_pass[p] != 0 keeps reads 2p and 2p+1; leaves the read_bam post-state
    void apply_pair_filter(const std::vector<std::uint8_t>& pair_pass);
    bool should_be_filtered_out(const Read& r1, const Read& r2) const;  // bam_api.cpp:311-319

    static AmpliconSet load_amplicons(const std::filesystem::path& bed,
                                      const std::filesystem::path& tsv);

   private:
    void ensure_loaded();
    void read_bam(const std::filesystem::path& input_filepath, SOAPairedReads& unfiltered);
    void run_host_filter();
    std::filesystem::path input_filepath_;
    bool file_pending_ = false;
    std::uint32_t hts_thread_count_ = 1;
    std::uint64_t bam_record_count_ = 0;
    double read_bam_seconds_ = 0;
    SOAPairedReads soa_paired_reads_;
    bool is_soa_loaded_ = false;
    AOSPairedReads aos_paired_reads_;
    bool is_aos_loaded_ = false;
    bool pending_filter_ = false;
    AmpliconSet amplicon_set_;
    AmpliconBehaviour amplicon_behaviour_ = AmpliconBehaviour::IGNORE;
    std::vector<BAMReadId> filtered_out_reads_;
    std::uint32_t min_seq_length_ = 0;
    std::uint32_t min_mapq_ = 0;
};

}  // namespace bam_api
