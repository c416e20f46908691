// Host-side BamApi for the quasi-MCP path (in-memory construction, filter, cover helpers,
// find_pairs).  Semantics follow libs/bam-api/src/bam_api.cpp; cited per function.
#include "bam-api/bam_api.hpp"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>

#include "logging/log.hpp"

namespace bam_api {

bool AmpliconSet::member_includes_both(const Read& a, const Read& b) const {
    // amplicon_set.cpp:5-9 — linear scan, any amplicon containing both mates
    for (const Amplicon& amp : amplicons)
        if (amp.includes(a) && amp.includes(b)) return true;
    return false;
}

BamApi::BamApi(const AOSPairedReads& paired_reads)
    : aos_paired_reads_(paired_reads), is_aos_loaded_(true) {}

BamApi::BamApi(const SOAPairedReads& paired_reads)
    : soa_paired_reads_(paired_reads), is_soa_loaded_(true) {}

BamApi::BamApi(const SOAPairedReads& unfiltered, const BamApiConfig& config)
    : soa_paired_reads_(unfiltered), is_soa_loaded_(true), pending_filter_(true),
      min_seq_length_(config.min_seq_length), min_mapq_(config.min_mapq) {
    // bam_api.cpp:32-43: amplicons only matter when a BED file is given
    if (!config.bed_filepath.empty()) {
        amplicon_set_ = load_amplicons(config.bed_filepath, config.tsv_filepath);
        amplicon_behaviour_ = config.amplicon_behaviour;
    }
}

// BED rows -> primer map, TSV rows -> (left.start, right.end) amplicons; without TSV the
// name-sorted primers are paired consecutively (bam_api.cpp:53-95, :101-187).  An odd primer
// count without TSV is an error here (the reference walks past end()).
AmpliconSet BamApi::load_amplicons(const std::filesystem::path& bed,
                                   const std::filesystem::path& tsv) {
    std::map<std::string, std::pair<Index, Index>> primers;
    std::ifstream bf(bed);
    if (!bf.is_open()) {
        LOG_WITH_LEVEL(logging::ERROR) << "Error opening .bed file: " << bed;
        std::exit(EXIT_FAILURE);
    }
    std::string line;
    while (std::getline(bf, line)) {
        std::istringstream ss(line);
        std::string f[4];
        for (auto& x : f) std::getline(ss, x, '\t');
        Index a = 0, b = 0;
        try {
            a = std::stoull(f[1]);
            b = std::stoull(f[2]);
        } catch (...) {
            LOG_WITH_LEVEL(logging::ERROR) << "Invalid BED line: " << line;
            continue;
        }
        if (f[0].empty() || f[3].empty()) {
            LOG_WITH_LEVEL(logging::ERROR) << "Invalid BED line: " << line;
            continue;
        }
        primers.emplace(f[3], std::make_pair(a, b));
    }
    AmpliconSet set;
    auto add = [&set](std::pair<Index, Index>& l, std::pair<Index, Index>& r) {
        if (l.first > r.first) std::swap(l, r);
        set.amplicons.push_back(Amplicon{l.first, r.second});
    };
    if (!tsv.empty()) {
        std::ifstream tf(tsv);
        if (!tf.is_open()) {
            LOG_WITH_LEVEL(logging::ERROR) << "Error opening .tsv file: " << tsv;
            std::exit(EXIT_FAILURE);
        }
        while (std::getline(tf, line)) {
            std::istringstream ss(line);
            std::string l, r;
            std::getline(ss, l, '\t');
            std::getline(ss, r, '\t');
            if (l.empty() || r.empty()) continue;
            add(primers[l], primers[r]);
        }
    } else {
        if (primers.size() % 2) {
            LOG_WITH_LEVEL(logging::ERROR) << "Odd number of primers in " << bed << " and no TSV";
            std::exit(EXIT_FAILURE);
        }
        for (auto it = primers.begin(); it != primers.end(); ++it) {
            auto& l = it->second;
            ++it;
            add(l, it->second);
        }
    }
    return set;
}

bool BamApi::should_be_filtered_out(const Read& r1, const Read& r2) const {
    bool drop = !(r1.quality >= min_mapq_ && r2.quality >= min_mapq_) ||
                !(r1.seq_length >= min_seq_length_ && r2.seq_length >= min_seq_length_);
    if (amplicon_behaviour_ == AmpliconBehaviour::FILTER)
        drop = drop || !amplicon_set_.member_includes_both(r1, r2);
    return drop;
}

void BamApi::apply_pair_filter(const std::vector<std::uint8_t>& pair_pass) {
    // leaves the state read_bam produces: survivors in pair order, the rest in
    // filtered_out_reads_ by BAM id (bam_api.cpp:456-478)
    SOAPairedReads kept;
    kept.ref_genome_length = soa_paired_reads_.ref_genome_length;
    const ReadIndex n = soa_paired_reads_.get_reads_count();
    filtered_out_reads_.clear();
    for (ReadIndex i = 0; i < n; ++i) {
        ReadIndex p = i / 2;
        if (p < pair_pass.size() && pair_pass[p]) kept.push_back(soa_paired_reads_.get_read_by_index(i));
        else filtered_out_reads_.push_back(soa_paired_reads_.ids[i]);
    }
    soa_paired_reads_ = std::move(kept);
    is_aos_loaded_ = false;
    pending_filter_ = false;
}

void BamApi::run_host_filter() {
    const ReadIndex n = soa_paired_reads_.get_reads_count();
    std::vector<std::uint8_t> pass(n / 2, 0);
    for (ReadIndex p = 0; p < n / 2; ++p)
        pass[p] = !should_be_filtered_out(soa_paired_reads_.get_read_by_index(2 * p),
                                          soa_paired_reads_.get_read_by_index(2 * p + 1));
    apply_pair_filter(pass);
}

const SOAPairedReads& BamApi::get_paired_reads_soa() {
    if (pending_filter_) run_host_filter();
    if (!is_soa_loaded_) {
        soa_paired_reads_.from(aos_paired_reads_);
        is_soa_loaded_ = true;
    }
    return soa_paired_reads_;
}

const AOSPairedReads& BamApi::get_paired_reads_aos() {
    if (pending_filter_) run_host_filter();
    if (!is_aos_loaded_) {
        aos_paired_reads_.from(soa_paired_reads_);
        is_aos_loaded_ = true;
    }
    return aos_paired_reads_;
}

const PairedReads& BamApi::get_paired_reads() const {
    if (is_soa_loaded_) return soa_paired_reads_;
    return aos_paired_reads_;
}

std::vector<ReadIndex> BamApi::find_pairs(const std::vector<ReadIndex>& ids) const {
    // bam_api.cpp:239-273: each kept read followed by its mate (idx+1 if first, else idx-1)
    const PairedReads& pr = get_paired_reads();
    std::vector<bool> seen(pr.get_reads_count(), false);
    std::vector<ReadIndex> out;
    out.reserve(2 * ids.size());
    for (ReadIndex id : ids) {
        ReadIndex mate = pr.get_read_by_index(id).is_first_read ? id + 1 : id - 1;
        for (ReadIndex x : {id, mate})
            if (!seen[x]) {
                seen[x] = true;
                out.push_back(x);
            }
    }
    return out;
}

std::vector<std::uint32_t> BamApi::find_input_cover() {
    if (pending_filter_) run_host_filter();
    const PairedReads& pr = get_paired_reads();
    std::vector<std::uint32_t> cov(pr.ref_genome_length, 0);
    for (ReadIndex i = 0; i < pr.get_reads_count(); ++i) {
        Read r = pr.get_read_by_index(i);
        for (Index j = r.start_ind; j <= r.end_ind; ++j) ++cov[j];
    }
    return cov;
}

std::vector<std::uint32_t> BamApi::find_filtered_cover(const std::vector<ReadIndex>& active_ids) {
    const PairedReads& pr = get_paired_reads();
    std::vector<std::uint32_t> cov(pr.ref_genome_length, 0);
    for (ReadIndex id : active_ids) {
        Read r = pr.get_read_by_index(id);
        for (Index j = r.start_ind; j <= r.end_ind; ++j) ++cov[j];
    }
    return cov;
}

}  // namespace bam_api
