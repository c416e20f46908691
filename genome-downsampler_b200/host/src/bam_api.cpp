// Host-side BamApi for the quasi-MCP path (in-memory construction, filter, cover helpers,
// find_pairs).  Semantics follow libs/bam-api/src/bam_api.cpp; cited per function.
#include "bam-api/bam_api.hpp"

#include <algorithm>
#include <limits>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <map>
#include <sstream>
#include <stdexcept>

#include "bam-api/bgzf_bam.hpp"
#include "logging/log.hpp"

namespace bam_api {

bool AmpliconSet::member_includes_both(const Read& a, const Read& b) const {
    // amplicon_set.cpp:5-9 — linear scan, any amplicon containing both mates
    for (const Amplicon& amp : amplicons)
        if (amp.includes(a) && amp.includes(b)) return true;
    return false;
}

BamApi::BamApi(const std::filesystem::path& input_filepath, const BamApiConfig& config)
    : input_filepath_(input_filepath), file_pending_(true),
      hts_thread_count_(config.hts_thread_count), min_seq_length_(config.min_seq_length),
      min_mapq_(config.min_mapq) {
    // bam_api.cpp:30-42
    if (!config.bed_filepath.empty()) {
        amplicon_set_ = load_amplicons(config.bed_filepath, config.tsv_filepath);
        amplicon_behaviour_ = config.amplicon_behaviour;
    }
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
    // file-backed: every record of the BAM that is not in a surviving pair is filtered out,
    // never-paired records included (is_accepted, bam_api.cpp:430,461-478)
    std::vector<bool> accepted(bam_record_count_, false);
    for (ReadIndex i = 0; i < n; ++i) {
        ReadIndex p = i / 2;
        if (p < pair_pass.size() && pair_pass[p]) {
            kept.push_back(soa_paired_reads_.get_read_by_index(i));
            if (bam_record_count_) accepted[soa_paired_reads_.ids[i]] = true;
        } else if (!bam_record_count_) {
            filtered_out_reads_.push_back(soa_paired_reads_.ids[i]);
        }
    }
    for (BAMReadId id = 0; id < bam_record_count_; ++id)
        if (!accepted[id]) filtered_out_reads_.push_back(id);
    if (amplicon_behaviour_ == AmpliconBehaviour::GRADE) {
        // GRADE (bam_api.cpp:334-347, 444-454, 480-483): nothing is dropped for its amplicons;
        // the qualities of the surviving pairs are shifted to start at 0 and a pair that lies
        // inside one amplicon is lifted by the whole quality range, so a quality-aware solver
        // prefers it
        ReadQuality lo = std::numeric_limits<ReadQuality>::max(), hi = 0;
        const ReadIndex m = kept.get_reads_count();
        for (ReadIndex i = 0; i < m; ++i) {
            lo = std::min(lo, kept.qualities[i]);
            hi = std::max(hi, kept.qualities[i]);
        }
        if (hi > 0 && lo < std::numeric_limits<ReadQuality>::max()) {
            for (ReadIndex i = 0; i + 1 < m; i += 2) {
                const bool inside = amplicon_set_.member_includes_both(kept.get_read_by_index(i),
                                                                       kept.get_read_by_index(i + 1));
                for (ReadIndex j = i; j < i + 2; ++j) {
                    ReadQuality q = kept.qualities[j] - lo;
                    if (inside) q += hi - lo;
                    kept.qualities[j] = q;
                }
            }
        }
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
    ensure_loaded();
    if (pending_filter_) run_host_filter();
    if (!is_soa_loaded_) {
        soa_paired_reads_.from(aos_paired_reads_);
        is_soa_loaded_ = true;
    }
    return soa_paired_reads_;
}

const AOSPairedReads& BamApi::get_paired_reads_aos() {
    ensure_loaded();
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
    ensure_loaded();
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

// ---------------------------------------------------------------- BAM files

namespace {

// QNAME -> first read seen with it.  The reference keeps a std::map<std::string, Read> that is
// never erased from (bam_api.cpp:425-466); an open-addressed table over a byte arena does the
// same job without a node allocation per read.
class QnameTable {
   public:
    // what the map holds for a name; `partner` indexes partners_ while a swap is pending
    struct Entry {
        std::uint64_t name_off = 0;
        Read read;
        std::uint32_t partner = kNone;  // the later read whose arrival may have swapped the entry
        std::uint8_t name_len = 0;
    };
    static constexpr std::uint32_t kNone = 0xffffffffu;
    QnameTable() : slots_(1u << 10) {}
    Entry* find(std::uint64_t h, const char* name, std::uint8_t len) {
        for (std::size_t i = h & (slots_.size() - 1);; i = (i + 1) & (slots_.size() - 1)) {
            const Slot& s = slots_[i];
            if (!s.entry_plus1) return nullptr;
            if (s.hash != h) continue;
            Entry& e = entries_[s.entry_plus1 - 1];
            if (e.name_len == len && std::memcmp(arena_.data() + e.name_off, name, len) == 0) return &e;
        }
    }
    void insert(std::uint64_t h, const char* name, std::uint8_t len, const Read& r) {
        if ((entries_.size() + 1) * 10 > slots_.size() * 6) grow();
        Entry e;
        e.name_off = arena_.size();
        e.name_len = len;
        e.read = r;
        arena_.insert(arena_.end(), name, name + len);
        entries_.push_back(e);
        place(slots_, Slot{h, static_cast<std::uint32_t>(entries_.size())});
    }
    // the entry becomes `partner` iff that pair passed the filter (resolved when the name returns)
    bool swap_pending(const Entry& e) const { return e.partner != kNone; }
    const Read& partner(const Entry& e) const { return partners_[e.partner]; }
    void set_partner(Entry& e, const Read& r) {
        if (e.partner == kNone) {
            e.partner = static_cast<std::uint32_t>(partners_.size());
            partners_.push_back(r);
        } else {
            partners_[e.partner] = r;
        }
    }
    void clear_partner(Entry& e) { e.partner = kNone; }  // the slot in partners_ is simply left behind

   private:
    // 16-byte probe records: the reads and names live in side arrays, so a table of a million
    // names probes 32 MB instead of 200
    struct Slot {
        std::uint64_t hash = 0;
        std::uint32_t entry_plus1 = 0;  // 0 = empty
    };
    static void place(std::vector<Slot>& t, const Slot& s) {
        std::size_t i = s.hash & (t.size() - 1);
        while (t[i].entry_plus1) i = (i + 1) & (t.size() - 1);
        t[i] = s;
    }
    void grow() {
        std::vector<Slot> bigger(slots_.size() * 2);
        for (const Slot& s : slots_)
            if (s.entry_plus1) place(bigger, s);
        slots_.swap(bigger);
    }
    std::vector<Slot> slots_;
    std::vector<Entry> entries_;
    std::vector<Read> partners_;
    std::vector<char> arena_;
};

}  // namespace

void BamApi::ensure_loaded() {
    if (!file_pending_) return;
    file_pending_ = false;
    read_bam(input_filepath_, soa_paired_reads_);
    is_soa_loaded_ = true;
    pending_filter_ = true;
}

// bam_api.cpp:359-507 without the filter: every QNAME-matched pair lands in `unfiltered` in
// pair-completion order, first mate at the even index (the swap of :456-458).  The filter the
// reference applies inside this loop (:438-441) runs afterwards over the pairs — on the device
// for quasi-mcp-b200, on the host otherwise — and apply_pair_filter() leaves the same arrays and
// the same filtered_out_reads_ the reference's loop does.
void BamApi::read_bam(const std::filesystem::path& input_filepath, SOAPairedReads& unfiltered) {
    LOG_WITH_LEVEL(logging::INFO) << "Reading " << input_filepath.filename() << " input file...";
    auto t0 = std::chrono::high_resolution_clock::now();
    try {
        bgzf::BamScanner scanner(input_filepath, hts_thread_count_);
        if (scanner.header().ref_lengths.empty()) throw std::runtime_error("BAM header lists no reference sequence");
        unfiltered.clear();
        unfiltered.ref_genome_length = scanner.header().ref_lengths[0];  // :421 target_len[0]
        // QNAMEs are split over one table per thread by hash, so the lookups of a chunk run in
        // parallel while every name still meets its earlier records in file order; the pairs are
        // then appended serially in the order their second record appears (pair-completion order)
        const std::uint32_t parts = std::min<std::uint32_t>(std::max<std::uint32_t>(hts_thread_count_, 1), 4096);
        std::vector<QnameTable> tables(parts);
        std::vector<Read> mate;
        std::vector<std::uint8_t> completes;  // 1: (map entry, record)  2: (record, map entry)
        std::vector<std::uint16_t> part_of;
        std::vector<std::uint32_t> order, counts, part_begin;
        bgzf::RecordChunk chunks[2];
        BAMReadId id = 0;
        auto make_read = [](BAMReadId rid, const bgzf::RecordFields& f) {
            // Read::Read(id, bam1_t*), read.cpp:5-14: end = pos + bam_cigar2rlen - 1
            return Read(rid, static_cast<Index>(static_cast<std::int64_t>(f.pos)),
                        static_cast<Index>(static_cast<std::uint64_t>(static_cast<std::int64_t>(f.pos)) + f.ref_len - 1),
                        f.mapq, static_cast<std::uint32_t>(f.l_seq), (f.flag & 0x40) != 0);
        };
        // the scanner inflates and indexes chunk k+1 on its own threads while chunk k is paired
        int cur = 0;
        double t_pair = 0, t_append = 0, t_wait = 0;
        bool have = scanner.next(chunks[0]);
        while (have) {
            const bgzf::RecordChunk& chunk = chunks[cur];
            auto ahead = std::async(std::launch::async, [&scanner, &chunks, cur] { return scanner.next(chunks[1 - cur]); });
            const std::size_t n = chunk.records.size();
            auto tp0 = std::chrono::steady_clock::now();
            mate.resize(n);
            completes.assign(n, 0);
            // counting sort of the chunk's record numbers by table (stable, so every table sees
            // its records in file order): each pairing thread then touches only its own records
            constexpr std::size_t kGrain = 8192;
            const std::size_t grains = (n + kGrain - 1) / kGrain;
            part_of.resize(n);
            order.resize(n);
            counts.assign(grains * parts + 1, 0);
            bgzf::parallel_for(grains, parts, [&](std::size_t g) {
                std::uint32_t* c = counts.data() + g * parts;
                for (std::size_t i = g * kGrain; i < std::min(n, (g + 1) * kGrain); ++i) {
                    const std::uint32_t pt = static_cast<std::uint32_t>((chunk.records[i].qname_hash >> 40) % parts);
                    part_of[i] = static_cast<std::uint16_t>(pt);
                    ++c[pt];
                }
            });
            part_begin.assign(parts + 1, 0);
            {
                std::uint32_t run = 0;
                for (std::uint32_t pt = 0; pt < parts; ++pt) {
                    part_begin[pt] = run;
                    for (std::size_t g = 0; g < grains; ++g) {
                        const std::uint32_t c = counts[g * parts + pt];
                        counts[g * parts + pt] = run;
                        run += c;
                    }
                }
                part_begin[parts] = run;
            }
            bgzf::parallel_for(grains, parts, [&](std::size_t g) {
                std::uint32_t* cursor = counts.data() + g * parts;
                for (std::size_t i = g * kGrain; i < std::min(n, (g + 1) * kGrain); ++i)
                    order[cursor[part_of[i]]++] = static_cast<std::uint32_t>(i);
            });
            bgzf::parallel_for(parts, parts, [&](std::size_t part) {
                QnameTable& table = tables[part];
                for (std::uint32_t k = part_begin[part]; k < part_begin[part + 1]; ++k) {
                    const std::size_t i = order[k];
                    const bgzf::RecordFields& f = chunk.records[i];
                    Read cur = make_read(id + i, f);
                    const char* name = chunk.qname(f);
                    if (QnameTable::Entry* e = table.find(f.qname_hash, name, f.l_qname)) {
                        // a QNAME seen a third time pairs with whatever the map holds by then: the
                        // reference swaps the map entry with the second read only when that pair
                        // survived the filter (the `continue` of :438-441 skips the swap)
                        if (table.swap_pending(*e)) {
                            const Read& p = table.partner(*e);
                            if (!should_be_filtered_out(e->read, p)) e->read = p;
                            table.clear_partner(*e);
                        }
                        mate[i] = e->read;
                        if (cur.is_first_read) {
                            completes[i] = 2;
                            table.set_partner(*e, cur);
                        } else {
                            completes[i] = 1;
                        }
                    } else {
                        table.insert(f.qname_hash, name, f.l_qname, cur);
                    }
                }
            });
            auto tp1 = std::chrono::steady_clock::now();
            for (std::size_t i = 0; i < n; ++i) {
                if (!completes[i]) continue;
                Read cur = make_read(id + i, chunk.records[i]);
                unfiltered.push_back(completes[i] == 2 ? cur : mate[i]);
                unfiltered.push_back(completes[i] == 2 ? mate[i] : cur);
            }
            id += n;
            auto tp2 = std::chrono::steady_clock::now();
            have = ahead.get();
            auto tp3 = std::chrono::steady_clock::now();
            t_pair += std::chrono::duration<double>(tp1 - tp0).count();
            t_append += std::chrono::duration<double>(tp2 - tp1).count();
            t_wait += std::chrono::duration<double>(tp3 - tp2).count();
            cur = 1 - cur;
        }
        bam_record_count_ = id;
        LOG_WITH_LEVEL(logging::DEBUG) << "BamApi: pairing " << t_pair << " s, pair append " << t_append
                                       << " s, waiting for the scanner " << t_wait << " s";
    } catch (const std::exception& e) {
        LOG_WITH_LEVEL(logging::ERROR) << e.what();
        std::exit(EXIT_FAILURE);
    }
    read_bam_seconds_ = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    LOG_WITH_LEVEL(logging::DEBUG) << "BamApi: " << bam_record_count_ << " reads have been read, "
                                   << unfiltered.get_reads_count() << " in pairs; read_bam took "
                                   << read_bam_seconds_ << " seconds";
}

std::uint32_t BamApi::write_paired_reads(const std::filesystem::path& output_filepath,
                                         std::vector<ReadIndex>& active_ids) const {
    // bam_api.cpp:509-524: in-memory indices -> BAM ordinals
    LOG_WITH_LEVEL(logging::INFO) << "Writing solution of size " << active_ids.size() << " reads "
                                  << output_filepath.filename() << "...";
    const PairedReads& paired_reads = get_paired_reads();
    std::vector<BAMReadId> active_bam_ids;
    active_bam_ids.reserve(active_ids.size());
    for (const auto& id : active_ids) active_bam_ids.push_back(paired_reads.get_read_by_index(id).bam_id);
    return write_bam(input_filepath_, output_filepath, active_bam_ids, hts_thread_count_);
}

std::uint32_t BamApi::write_bam_api_filtered_out_reads(const std::filesystem::path& output_filepath) {
    // bam_api.cpp:526-532
    LOG_WITH_LEVEL(logging::INFO) << "Writing " << filtered_out_reads_.size()
                                  << " preprocessing filtered out reads to "
                                  << output_filepath.filename() << "...";
    return write_bam(input_filepath_, output_filepath, filtered_out_reads_, hts_thread_count_);
}

std::uint32_t BamApi::write_bam(const std::filesystem::path& input_filepath,
                                const std::filesystem::path& output_filepath,
                                std::vector<BAMReadId>& bam_ids, std::uint32_t hts_thread_count) {
    // ".bam" -> BGZF-compressed BAM, anything else -> SAM text (open_mode, bam_api.cpp:566)
    try {
        return bgzf::copy_bam_records(input_filepath, output_filepath, bam_ids, hts_thread_count);
    } catch (const std::exception& e) {
        LOG_WITH_LEVEL(logging::ERROR) << e.what();
        std::exit(EXIT_FAILURE);
    }
}

}  // namespace bam_api
