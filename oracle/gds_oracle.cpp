// gds_oracle.cpp — CPU ORACLE (test infrastructure only; see gds_oracle.h for the rules and the
// parity-pin status).  Each block cites the reference file:line it restates.
#include "gds_oracle.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

namespace {
using clk = std::chrono::steady_clock;
inline double secs(clk::time_point a, clk::time_point b) {
    return std::chrono::duration<double>(b - a).count();
}
}  // namespace

// =====================================================================================
// Generators — restates libs/reads-gen/src/reads_gen.cpp:5-53 (histogram law) and :55-86
// (uniform law).  Draw order per pair: first, second, quality(first), quality(second); the
// libstdc++ mt19937 / uniform_int_distribution<> / discrete_distribution<> are used directly so
// the stream is the reference's stream (checked against oracle/_ref in tests).
// =====================================================================================
static double shape_value(int shape, double x) {
    switch (shape) {
        case 1:  // coverage_tester.cpp:157-160
            return x - x * x;
        case 2:  // coverage_tester.cpp:162-170
            if (x > 0.3684 && x < 0.6316) {
                return 1000.0 * (x * x - x + 0.25) * (x * x - x + 0.25) + 0.2;
            }
            return 0.5;
        case 3:  // coverage_tester.cpp:172-175
            return -10.0 * (x - 0.5) * (x - 0.5) + 1.0;
        default:
            return 1.0;
    }
}

extern "C" int orc_gen_reads(uint32_t seed, uint64_t pairs, uint32_t genome_len, uint32_t read_len,
                             int shape, int32_t max_quality, uint32_t* start, uint32_t* end,
                             uint32_t* quality, uint32_t* seq_len) {
    if (genome_len < 2 * read_len || read_len == 0) return -1;
    std::mt19937 gen(seed);
    std::uniform_int_distribution<> qd(0, max_quality);
    auto emit = [&](uint64_t i, uint64_t a, uint64_t b) {
        start[i] = (uint32_t)a;
        end[i] = (uint32_t)(a + read_len - 1);
        quality[i] = (uint32_t)qd(gen);
        seq_len[i] = read_len;
        start[i + 1] = (uint32_t)b;
        end[i + 1] = (uint32_t)(b + read_len - 1);
        quality[i + 1] = (uint32_t)qd(gen);
        seq_len[i + 1] = read_len;
    };
    if (shape == 0) {
        // reads_gen.cpp:59-63
        std::uniform_int_distribution<> d1(0, (int32_t)(genome_len - 2 * read_len));
        std::uniform_int_distribution<> d2(0, (int32_t)(genome_len - read_len));
        for (uint64_t p = 0; p < pairs; ++p) {
            uint64_t a = (uint64_t)d1(gen);
            uint64_t b = (uint64_t)d2(gen);
            if (a > b) std::swap(a, b);
            if (a + read_len > b) b = a + read_len;  // reads_gen.cpp:75-77
            emit(2 * p, a, b);
        }
        return 0;
    }
    // reads_gen.cpp:10-27
    uint32_t starts_count = genome_len - read_len + 1;
    std::vector<double> w(starts_count, 0.0);
    double sum = 0;
    for (uint32_t i = 0; i < starts_count; ++i) {
        double v = shape_value(shape, (double)i / (double)(starts_count - 1));
        if (v < 0.0) v = 0.0;
        w[i] = v;
        sum += v;
    }
    for (uint32_t i = 0; i < starts_count; ++i) w[i] /= sum;
    std::discrete_distribution<> dd(w.begin(), w.end());
    const uint64_t G = genome_len, R = read_len;
    for (uint64_t p = 0; p < pairs; ++p) {
        uint64_t a = (uint64_t)dd(gen);
        uint64_t b = (uint64_t)dd(gen);
        if (a > b) std::swap(a, b);
        if (a > G - 2 * R && b > G - 2 * R) {  // reads_gen.cpp:39-41
            a = G - 2 * R;
            b = G - R;
        } else if (a + R > b) {  // reads_gen.cpp:42-44
            b = a + R;
        }
        emit(2 * p, a, b);
    }
    return 0;
}

// Synthetic ARTIC-like scheme: amplicon k covers
// [k*(amp_len-overlap), k*(amp_len-overlap)+amp_len-1]; its LEFT primer is the first primer_len
// bases and its RIGHT primer the last primer_len bases.  BED rows: chrom, start, end, name.
extern "C" int orc_gen_artic_scheme(uint32_t genome_len, uint32_t n_amplicons, uint32_t amp_len,
                                    uint32_t overlap, uint32_t primer_len, char* bed,
                                    uint64_t bed_cap, char* tsv, uint64_t tsv_cap) {
    std::ostringstream b, t;
    uint32_t step = amp_len - overlap;
    for (uint32_t k = 0; k < n_amplicons; ++k) {
        uint64_t a0 = (uint64_t)k * step;
        uint64_t a1 = a0 + amp_len - 1;
        if (a1 >= genome_len) return -1;
        char ln[64], rn[64];
        snprintf(ln, sizeof ln, "amp_%03u_LEFT", k + 1);
        snprintf(rn, sizeof rn, "amp_%03u_RIGHT", k + 1);
        b << "synth\t" << a0 << "\t" << (a0 + primer_len - 1) << "\t" << ln << "\t" << (k % 2 + 1)
          << "\t+\n";
        b << "synth\t" << (a1 - primer_len + 1) << "\t" << a1 << "\t" << rn << "\t" << (k % 2 + 1)
          << "\t-\n";
        t << ln << "\t" << rn << "\n";
    }
    std::string bs = b.str(), ts = t.str();
    if (bs.size() + 1 > bed_cap || ts.size() + 1 > tsv_cap) return -2;
    memcpy(bed, bs.c_str(), bs.size() + 1);
    memcpy(tsv, ts.c_str(), ts.size() + 1);
    return 0;
}

extern "C" int orc_gen_reads_amplicon(uint32_t seed, uint64_t pairs, uint32_t genome_len,
                                      uint32_t n_amp, const uint32_t* amp_start,
                                      const uint32_t* amp_end, double p_inside, uint32_t min_len,
                                      uint32_t max_len, int32_t max_quality, uint32_t* start,
                                      uint32_t* end, uint32_t* quality, uint32_t* seq_len) {
    if (n_amp == 0 || max_len < min_len || genome_len < 2 * max_len) return -1;
    std::mt19937 gen(seed);
    std::uniform_int_distribution<> qd(0, max_quality);
    std::uniform_int_distribution<> ld((int)min_len, (int)max_len);
    std::uniform_int_distribution<> ad(0, (int)n_amp - 1);
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    for (uint64_t p = 0; p < pairs; ++p) {
        uint32_t l1 = (uint32_t)ld(gen), l2 = (uint32_t)ld(gen);
        uint64_t a, b;
        bool inside = ud(gen) < p_inside;
        uint32_t k = (uint32_t)ad(gen);
        uint64_t A0 = amp_start[k], A1 = amp_end[k];
        if (inside && A1 - A0 + 1 >= std::max(l1, l2)) {
            // first mate flush left region, second mate flush right region, both inside amplicon k
            std::uniform_int_distribution<uint64_t> s1(A0, A1 + 1 - l1);
            std::uniform_int_distribution<uint64_t> s2(A0, A1 + 1 - l2);
            a = s1(gen);
            b = s2(gen);
        } else {
            std::uniform_int_distribution<uint64_t> s1(0, genome_len - l1);
            std::uniform_int_distribution<uint64_t> s2(0, genome_len - l2);
            a = s1(gen);
            b = s2(gen);
        }
        uint64_t i = 2 * p;
        start[i] = (uint32_t)a;
        end[i] = (uint32_t)(a + l1 - 1);
        quality[i] = (uint32_t)qd(gen);
        seq_len[i] = l1;
        start[i + 1] = (uint32_t)b;
        end[i + 1] = (uint32_t)(b + l2 - 1);
        quality[i + 1] = (uint32_t)qd(gen);
        seq_len[i + 1] = l2;
    }
    return 0;
}

// =====================================================================================
// Amplicon table — restates BamApi::set_amplicon_filter (bam_api.cpp:53-95),
// process_bed_file (:101-152) and process_tsv_file (:154-187), reading from text buffers.
// Quirks kept as written: BED end treated as inclusive (App. B11); unknown primer names
// default-insert (0,0) (B9); swap mutates the map entries in place (B9).  Without a TSV the
// name-sorted rows are paired consecutively; an odd count is rejected here (-1) instead of
// walking past end() (B10 is undefined behaviour in the reference).
// =====================================================================================
static bool parse_u64(const std::string& s, uint64_t& out) {
    try {
        size_t pos = 0;
        out = std::stoull(s, &pos);
        (void)pos;
        return true;
    } catch (...) {
        return false;
    }
}

extern "C" int64_t orc_parse_amplicons(const char* bed_text, const char* tsv_text,
                                       uint32_t* amp_start, uint32_t* amp_end, uint64_t cap) {
    std::map<std::string, std::pair<uint64_t, uint64_t>> primers;
    {
        std::istringstream in(bed_text);
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ss(line);
            std::string chrom, s, e, name;
            std::getline(ss, chrom, '\t');
            std::getline(ss, s, '\t');
            std::getline(ss, e, '\t');
            std::getline(ss, name, '\t');
            uint64_t si = 0, ei = 0;
            if (!parse_u64(s, si) || !parse_u64(e, ei)) continue;  // bam_api.cpp:128-137
            if (!chrom.empty() && !s.empty() && !e.empty() && !name.empty()) {
                primers.emplace(name, std::make_pair(si, ei));  // first occurrence wins
            }
        }
    }
    std::vector<std::pair<uint64_t, uint64_t>> amps;
    if (tsv_text != nullptr && tsv_text[0] != '\0') {
        std::istringstream in(tsv_text);
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ss(line);
            std::string l, r;
            std::getline(ss, l, '\t');
            std::getline(ss, r, '\t');
            if (l.empty() || r.empty()) continue;  // bam_api.cpp:175-179
            auto& lp = primers[l];
            auto& rp = primers[r];
            if (lp.first > rp.first) std::swap(lp, rp);  // bam_api.cpp:68-70
            amps.emplace_back(lp.first, rp.second);
        }
    } else {
        if (primers.size() % 2 != 0) return -1;
        for (auto it = primers.begin(); it != primers.end();) {
            auto& lp = it->second;
            ++it;
            auto& rp = it->second;
            if (lp.first > rp.first) std::swap(lp, rp);
            amps.emplace_back(lp.first, rp.second);
            ++it;
        }
    }
    if (amps.size() > cap) return -2;
    for (size_t i = 0; i < amps.size(); ++i) {
        amp_start[i] = (uint32_t)amps[i].first;
        amp_end[i] = (uint32_t)amps[i].second;
    }
    return (int64_t)amps.size();
}

// =====================================================================================
// Filter — restates should_be_filtered_out (bam_api.cpp:311-319), have_min_length (:321-323),
// have_min_mapq (:325-327), are_from_single_amplicon (:329-332), Amplicon::includes
// (amplicon.cpp:5-7) and AmpliconSet::member_includes_both (amplicon_set.cpp:5-9): linear scan.
// =====================================================================================
extern "C" uint64_t orc_filter_pairs(uint64_t n_reads, const uint32_t* start, const uint32_t* end,
                                     const uint32_t* quality, const uint32_t* seq_len,
                                     uint32_t min_len, uint32_t min_mapq, int use_amplicons,
                                     uint32_t n_amp, const uint32_t* amp_start,
                                     const uint32_t* amp_end, uint8_t* pair_pass) {
    uint64_t kept = 0;
    for (uint64_t p = 0; 2 * p + 1 < n_reads; ++p) {
        uint64_t i = 2 * p, j = i + 1;
        bool drop = !(quality[i] >= min_mapq && quality[j] >= min_mapq) ||
                    !(seq_len[i] >= min_len && seq_len[j] >= min_len);
        if (use_amplicons) {
            bool any = false;
            for (uint32_t a = 0; a < n_amp && !any; ++a) {
                bool in1 = amp_start[a] <= start[i] && end[i] <= amp_end[a];
                bool in2 = amp_start[a] <= start[j] && end[j] <= amp_end[a];
                any = in1 && in2;
            }
            drop = drop || !any;
        }
        pair_pass[p] = drop ? 0 : 1;
        if (!drop) kept += 2;
    }
    return kept;
}

// =====================================================================================
// Coverage and demand — restates create_b_function (quasi_mcp_cpu_max_flow_solver.cpp:58-73,
// un-shifted: cov[j] == b[j+1]), find_input_cover / find_filtered_cover (bam_api.cpp:275-301)
// and create_demand_function (:75-87).
// =====================================================================================
// a read that consumes no reference (end == start - 1, CIGAR '*') is legal: Read::Read gives it
// end = pos + rlen - 1 (read.cpp:5-14), the coverage loops below never run for it and its arc
// start -> end + 1 is a self-loop without flow
static inline bool read_ok(uint32_t s, uint32_t e, uint32_t L) {
    return (e + 1u == s) ? s < L : (s <= e && e < L);
}

extern "C" int orc_coverage_ref(uint64_t n, const uint32_t* start, const uint32_t* end, uint32_t L,
                                uint32_t* cov) {
    std::fill(cov, cov + L, 0u);
    for (uint64_t i = 0; i < n; ++i) {
        if (!read_ok(start[i], end[i], L)) return -1;
        if (end[i] + 1u == start[i]) continue;
        for (uint32_t j = start[i]; j <= end[i]; ++j) ++cov[j];
    }
    return 0;
}

extern "C" int orc_coverage_subset(uint64_t n, const uint32_t* start, const uint32_t* end,
                                   const uint8_t* kept, uint32_t L, uint32_t* cov) {
    std::fill(cov, cov + L, 0u);
    for (uint64_t i = 0; i < n; ++i) {
        if (!kept[i]) continue;
        if (!read_ok(start[i], end[i], L)) return -1;
        if (end[i] + 1u == start[i]) continue;
        for (uint32_t j = start[i]; j <= end[i]; ++j) ++cov[j];
    }
    return 0;
}

extern "C" int64_t orc_demand(const uint32_t* cov, uint32_t L, uint32_t M, int32_t* demand) {
    // b[j+1] = min(cov[j], M), b[0] = 0; d[0] = -b[1]; d[i] = b[i]-b[i+1] (1<=i<L); d[L] = b[L]
    int64_t fstar = 0;
    auto capd = [&](uint32_t j) -> int32_t { return (int32_t)std::min(cov[j], M); };
    if (L == 0) {
        demand[0] = 0;
        return 0;
    }
    demand[0] = -capd(0);
    for (uint32_t i = 1; i < L; ++i) demand[i] = capd(i - 1) - capd(i);
    demand[L] = capd(L - 1);
    for (uint32_t i = 0; i <= L; ++i)
        if (demand[i] < 0) fstar += -(int64_t)demand[i];
    return fstar;
}

// =====================================================================================
// Sequential max flow (stand-in for operations_research::SimpleMaxFlow, absent here).
// Generic FIFO push-relabel with exact initial labels, periodic global relabelling and a second
// BFS from the source so that stranded excess can return (single-phase variant, labels < 2n).
// =====================================================================================
namespace {
struct SeqMaxFlow {
    struct Arc {
        uint32_t head;
        int64_t res;  // residual capacity
    };
    uint32_t n = 0;
    std::vector<uint32_t> tail_;  // per arc (forward arcs only, insertion order)
    std::vector<uint32_t> head_;
    std::vector<int64_t> cap_;
    // CSR over 2*m residual arcs
    std::vector<uint32_t> ptr, adj_arc;  // adj_arc: residual arc id (2*a = forward, 2*a+1 = reverse)
    std::vector<int64_t> res;
    std::vector<int64_t> excess;
    std::vector<uint32_t> label, cur;
    uint64_t pushes = 0, relabels = 0, grs = 0;

    void add_arc(uint32_t u, uint32_t v, int64_t c) {
        tail_.push_back(u);
        head_.push_back(v);
        cap_.push_back(c);
    }
    inline uint32_t arc_head(uint32_t ra) const { return (ra & 1) ? tail_[ra >> 1] : head_[ra >> 1]; }

    void finalize(uint32_t n_nodes) {
        n = n_nodes;
        size_t m = tail_.size();
        ptr.assign(n + 1, 0);
        for (size_t a = 0; a < m; ++a) {
            ++ptr[tail_[a] + 1];
            ++ptr[head_[a] + 1];
        }
        for (uint32_t v = 0; v < n; ++v) ptr[v + 1] += ptr[v];
        adj_arc.resize(2 * m);
        std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
        for (size_t a = 0; a < m; ++a) {
            adj_arc[fill[tail_[a]]++] = (uint32_t)(2 * a);
            adj_arc[fill[head_[a]]++] = (uint32_t)(2 * a + 1);
        }
        res.resize(2 * m);
        for (size_t a = 0; a < m; ++a) {
            res[2 * a] = cap_[a];
            res[2 * a + 1] = 0;
        }
    }

    void global_relabel(uint32_t s, uint32_t t) {
        ++grs;
        const uint32_t UNSET = 0xffffffffu;
        std::fill(label.begin(), label.end(), UNSET);
        std::vector<uint32_t> q;
        q.reserve(n);
        auto bfs = [&](uint32_t root, uint32_t base) {
            size_t qh = q.size();
            label[root] = base;
            q.push_back(root);
            while (qh < q.size()) {
                uint32_t w = q[qh++];
                for (uint32_t k = ptr[w]; k < ptr[w + 1]; ++k) {
                    uint32_t ra = adj_arc[k];  // arc w -> x ; we need residual on x -> w = ra^1
                    uint32_t x = arc_head(ra);
                    if (label[x] != UNSET || res[ra ^ 1] <= 0) continue;
                    label[x] = label[w] + 1;
                    q.push_back(x);
                }
            }
        };
        bfs(t, 0);
        if (label[s] == UNSET) bfs(s, n);
        else label[s] = n;
        for (uint32_t v = 0; v < n; ++v)
            if (label[v] == UNSET) label[v] = 2 * n;
        label[s] = n;
        std::copy(ptr.begin(), ptr.end() - 1, cur.begin());
    }

    int64_t solve(uint32_t s, uint32_t t) {
        excess.assign(n, 0);
        label.assign(n, 0);
        cur.assign(n, 0);
        std::deque<uint32_t> fifo;
        std::vector<uint8_t> inq(n, 0);
        // preflow
        for (uint32_t k = ptr[s]; k < ptr[s + 1]; ++k) {
            uint32_t ra = adj_arc[k];
            if (ra & 1) continue;
            int64_t c = res[ra];
            if (c <= 0) continue;
            uint32_t v = arc_head(ra);
            res[ra] -= c;
            res[ra ^ 1] += c;
            excess[v] += c;
            excess[s] -= c;
        }
        global_relabel(s, t);
        for (uint32_t v = 0; v < n; ++v)
            if (v != s && v != t && excess[v] > 0) {
                fifo.push_back(v);
                inq[v] = 1;
            }
        uint64_t relabels_since = 0;
        while (!fifo.empty()) {
            uint32_t v = fifo.front();
            fifo.pop_front();
            inq[v] = 0;
            // discharge
            while (excess[v] > 0 && label[v] < 2 * n) {
                if (cur[v] == ptr[v + 1]) {
                    uint32_t mn = 2 * n;
                    for (uint32_t k = ptr[v]; k < ptr[v + 1]; ++k) {
                        uint32_t ra = adj_arc[k];
                        if (res[ra] > 0) mn = std::min(mn, label[arc_head(ra)] + 1);
                    }
                    label[v] = mn;
                    cur[v] = ptr[v];
                    ++relabels;
                    ++relabels_since;
                    if (relabels_since >= n) break;
                    continue;
                }
                uint32_t ra = adj_arc[cur[v]];
                uint32_t w = arc_head(ra);
                if (res[ra] > 0 && label[v] == label[w] + 1) {
                    int64_t d = std::min(excess[v], res[ra]);
                    res[ra] -= d;
                    res[ra ^ 1] += d;
                    excess[v] -= d;
                    excess[w] += d;
                    ++pushes;
                    if (w != s && w != t && !inq[w]) {
                        fifo.push_back(w);
                        inq[w] = 1;
                    }
                } else {
                    ++cur[v];
                }
            }
            if (relabels_since >= n) {
                relabels_since = 0;
                global_relabel(s, t);
                if (excess[v] > 0 && !inq[v] && label[v] < 2 * n) {
                    fifo.push_back(v);
                    inq[v] = 1;
                }
            }
        }
        return excess[t];
    }
    inline int64_t flow(size_t a) const { return res[2 * a + 1]; }
};
}  // namespace

// quasi_mcp_cpu_max_flow_solver.cpp:11-28 (solve), :30-56 (graph in the reference's arc order:
// N read arcs with arc id == read index, L back arcs i+1 -> i with cap INT64_MAX, then
// sink/source arcs), :89-100 (keep read i iff Flow(i) > 0).
extern "C" int orc_ref_solve(uint64_t n, const uint32_t* start, const uint32_t* end, uint32_t L,
                             uint32_t M, uint8_t* kept, orc_ref_stats* st) {
    auto t0 = clk::now();
    std::vector<uint32_t> cov(L);
    if (orc_coverage_ref(n, start, end, L, cov.data()) != 0) return -1;
    std::vector<int32_t> demand(L + 1);
    orc_demand(cov.data(), L, M, demand.data());
    auto t1 = clk::now();
    SeqMaxFlow mf;
    mf.tail_.reserve(n + 2 * (size_t)L + 2);
    mf.head_.reserve(n + 2 * (size_t)L + 2);
    mf.cap_.reserve(n + 2 * (size_t)L + 2);
    for (uint64_t i = 0; i < n; ++i) mf.add_arc(start[i], end[i] + 1, 1);
    for (uint32_t i = 0; i < L; ++i) mf.add_arc(i + 1, i, INT64_MAX / 4);
    uint32_t s = L + 1, t = L + 2;
    for (uint32_t i = 0; i <= L; ++i) {
        if (demand[i] > 0) mf.add_arc(i, t, demand[i]);
        else if (demand[i] < 0) mf.add_arc(s, i, -(int64_t)demand[i]);
    }
    mf.finalize(L + 3);
    auto t2 = clk::now();
    int64_t F = mf.solve(s, t);
    auto t3 = clk::now();
    uint64_t nk = 0;
    for (uint64_t i = 0; i < n; ++i) {
        kept[i] = mf.flow(i) > 0 ? 1 : 0;
        nk += kept[i];
    }
    auto t4 = clk::now();
    if (st) {
        st->flow_value = F;
        st->n_kept = nk;
        st->n_arcs = mf.tail_.size();
        st->pushes = mf.pushes;
        st->relabels = mf.relabels;
        st->global_relabels = mf.grs;
        st->t_coverage_s = secs(t0, t1);
        st->t_graph_s = secs(t1, t2);
        st->t_maxflow_s = secs(t2, t3);
        st->t_select_s = secs(t3, t4);
        st->t_total_s = secs(t0, t4);
    }
    return 0;
}

// =====================================================================================
// Deterministic bulk-synchronous push-relabel on the BUNDLED graph.  This is the
// schedule the CUDA kernel executes; every round is a pure function of the previous state, so a
// sequential replay gives the same flows and therefore the same kept bitmap.
//
// Network (same as quasi_mcp_cpu_max_flow_solver.cpp:30-56, bundled per SURVEY App. A.2):
//   nodes: per sample k, ref_len[k]+1 consecutive nodes; read (s,e) -> arc s -> e+1
//   bundle b = all reads with equal (s,t): capacity mult[b], flow f[b]
//   back arc v -> v-1 (infinite) exists iff position v-1 is covered (App. A.3 cut rule); g[v]=flow
//   source arcs are consumed by the preflow (initial excess = -demand), sink arcs are snk[v]
// =====================================================================================
namespace {
constexpr uint32_t LBL_INF = 0x3fffffffu;

struct VSample {  // per-sample layout of the virtual node space
    uint32_t obase, vbase, L, nseg, P, W, vn;
};

struct SyncGraph {
    uint32_t n_nodes = 0;   // virtual nodes (segments + phantom ids)
    uint32_t n_onodes = 0;  // original nodes = sum(ref_len + 1)
    std::vector<VSample> vs;
    std::vector<uint32_t> covR;    // virtual: coverage of the position between node v and v+1
    std::vector<int32_t> demand;   // virtual, per node
    std::vector<uint32_t> ocov;    // original space covR
    std::vector<int32_t> odemand;  // original space demand
    int64_t through = 0;           // flow that passes straight through cut nodes when stitched
    std::vector<uint32_t> b_s, b_t, b_mult, b_first;
    std::vector<uint32_t> sorted_idx;  // owner read of every sorted arc item
    std::vector<uint32_t> out_ptr, in_ptr, in_bid;
    std::vector<uint32_t> comp_lo, comp_hi;
    std::vector<uint32_t> b_forced;  // forced bundles: their multiplicity (b_mult is 0 during the solve)
    int64_t vsupply = 0;             // supply of the virtual network before the forced flows are taken out
    uint32_t M = 0;
};

// seg_len == 0: the library's default rule (csrc/gds_api.cu default_seg_len) — 16 384 positions,
// stretched in steps of 128 by up to a quarter when that brings the batch's segment count down to
// 296 (two resident components per SM of a B200).  A constant of the schedule, not a device query.
constexpr uint32_t kDefaultSegLen = 16384, kSegResident = 296;
static uint32_t default_seg_len(uint32_t n_samples, const uint32_t* ref_len) {
    auto count = [&](uint32_t seg) {
        uint64_t n = 0;
        for (uint32_t k = 0; k < n_samples; ++k)
            n += (uint64_t)ref_len[k] > 2ull * seg ? ((uint64_t)ref_len[k] + seg - 1) / seg : 1;
        return n;
    };
    const uint64_t n0 = count(kDefaultSegLen);
    if (n0 <= kSegResident || n0 > kSegResident + kSegResident / 4) return kDefaultSegLen;
    for (uint32_t seg = kDefaultSegLen + 128; seg <= kDefaultSegLen + kDefaultSegLen / 4; seg += 128)
        if (count(seg) <= kSegResident) return seg;
    return kDefaultSegLen;
}

// Long references are cut into segments of seg positions (a generalisation of the zero-coverage
// split): a read crossing a cut becomes two arcs, truncated at the cut node, one
// per segment, and is kept if either part carries flow.  Each segment is then an independent
// max-flow problem; stitched together (surplus on the back arcs) they form a maximum flow of the
// whole network, so F*, the demand vector and the capped coverage are unchanged.
//
// Virtual node ids: sample k, segment j, original node x  ->  vbase + j*W + P + (x - j*seg), with
// W = P + seg + 1 and P = maxlen-1 phantom ids in front of every segment.  Every arc item is
// keyed by (fake start id, original length): the right part of a crossing read gets the fake
// start  (its end node - its length), which falls on a phantom id, so the key keeps the original
// length and the key width does not grow.  Decoding clamps to the segment's real node range.
int build_sync_graph(uint32_t n_samples, const uint64_t* read_off, const uint32_t* ref_len,
                     const uint32_t* start, const uint32_t* end, uint32_t M, uint32_t seg_len,
                     SyncGraph& G, bool forced_cuts = false) {
    const uint64_t N = read_off[n_samples];
    uint32_t minlen = 0xffffffffu, maxlen = 0;
    for (uint32_t k = 0; k < n_samples; ++k)
        for (uint64_t i = read_off[k]; i < read_off[k + 1]; ++i) {
            if (!read_ok(start[i], end[i], ref_len[k])) return -1;
            uint32_t len = end[i] - start[i] + 1;
            minlen = std::min(minlen, len);
            maxlen = std::max(maxlen, len);
        }
    if (N == 0) minlen = maxlen = 1;
    if (seg_len == 0) {
        seg_len = default_seg_len(n_samples, ref_len);
        // one read length: a whole number of reads per segment (csrc/gds_api.cu, same rounding)
        if (minlen == maxlen && maxlen > 1) seg_len = (seg_len + maxlen - 1) / maxlen * maxlen;
    }
    const uint32_t seg = seg_len >= maxlen ? seg_len : 0xffffffffu;  // a read crosses <= 1 cut
    G.vs.resize(n_samples);
    uint32_t ob = 0, vb = 0;
    for (uint32_t k = 0; k < n_samples; ++k) {
        VSample& v = G.vs[k];
        v.obase = ob;
        v.vbase = vb;
        v.L = ref_len[k];
        // only references longer than two segments are cut (a 30 kb sample stays whole)
        v.nseg = (uint64_t)v.L > 2ull * seg ? (uint32_t)(((uint64_t)v.L + seg - 1) / seg) : 1;
        v.P = v.nseg > 1 ? maxlen - 1 : 0;
        v.W = v.nseg > 1 ? v.P + seg + 1 : 0;
        v.vn = v.nseg == 1 ? v.L + 1 : (v.nseg - 1) * v.W + v.P + (v.L - (v.nseg - 1) * seg) + 1;
        ob += v.L + 1;
        vb += v.vn;
    }
    G.n_onodes = ob;
    G.n_nodes = vb;
    const uint32_t nn = G.n_nodes;

    struct Item {
        uint32_t fake, len, owner, sample;
    };
    std::vector<Item> items;
    items.reserve(N + N / 64);
    std::vector<Item> extra;
    for (uint32_t k = 0; k < n_samples; ++k) {
        const VSample& v = G.vs[k];
        for (uint64_t i = read_off[k]; i < read_off[k + 1]; ++i) {
            const uint32_t s = start[i], e = end[i], len = e - s + 1;
            if (v.nseg == 1) {
                items.push_back({v.vbase + s, len, (uint32_t)i, k});
                continue;
            }
            const uint32_t js = s / seg, je = len ? e / seg : js;  // length 0 crosses nothing
            items.push_back({v.vbase + js * v.W + v.P + (s - js * seg), len, (uint32_t)i, k});
            if (je > js) {
                uint32_t vt = v.vbase + je * v.W + v.P + (e + 1 - je * seg);
                extra.push_back({vt - len, len, (uint32_t)i, k});
            }
        }
    }
    items.insert(items.end(), extra.begin(), extra.end());  // whole reads first, then right parts
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
        return a.fake != b.fake ? a.fake < b.fake : a.len < b.len;
    });
    const uint64_t NI = items.size();
    G.sorted_idx.resize(NI);
    std::vector<int64_t> diff(nn + 1, 0), odiff((size_t)G.n_onodes + 1, 0);
    auto to_orig = [&](const VSample& v, uint32_t vid) -> uint32_t {
        if (v.nseg == 1) return v.obase + (vid - v.vbase);
        uint32_t j = (vid - v.vbase) / v.W;
        return v.obase + j * seg + ((vid - v.vbase) - j * v.W - v.P);
    };
    for (uint64_t q = 0; q < NI; ++q) {
        const Item& it = items[q];
        G.sorted_idx[q] = it.owner;
        if (q == 0 || it.fake != items[q - 1].fake || it.len != items[q - 1].len) {
            const VSample& v = G.vs[it.sample];
            uint32_t sr = it.fake, tr = it.fake + it.len;
            if (v.nseg > 1) {
                uint32_t j = (it.fake - v.vbase) / v.W;
                uint32_t first = v.vbase + j * v.W + v.P;
                uint32_t last = first + std::min(seg, v.L - j * seg);
                sr = std::max(sr, first);
                tr = std::min(tr, last);
            }
            G.b_s.push_back(sr);
            G.b_t.push_back(tr);
            G.b_mult.push_back(0);
            G.b_first.push_back((uint32_t)q);
        }
        ++G.b_mult.back();
    }
    const uint32_t B = (uint32_t)G.b_s.size();
    for (uint32_t b = 0; b < B; ++b) {
        const VSample& v = G.vs[items[G.b_first[b]].sample];
        diff[G.b_s[b]] += G.b_mult[b];
        diff[G.b_t[b]] -= G.b_mult[b];
        odiff[to_orig(v, G.b_s[b])] += G.b_mult[b];
        odiff[to_orig(v, G.b_t[b])] -= G.b_mult[b];
    }
    G.covR.assign(nn, 0);
    G.demand.assign(nn, 0);
    {
        int64_t run = 0;
        uint32_t prev_capped = 0;
        for (uint32_t v = 0; v < nn; ++v) {
            run += diff[v];
            G.covR[v] = (uint32_t)run;
            uint32_t c = std::min((uint32_t)run, M);
            G.demand[v] = (int32_t)prev_capped - (int32_t)c;
            prev_capped = c;
        }
    }
    G.ocov.assign(G.n_onodes, 0);
    G.odemand.assign(G.n_onodes, 0);
    {
        int64_t run = 0;
        uint32_t prev_capped = 0;
        for (uint32_t v = 0; v < G.n_onodes; ++v) {
            run += odiff[v];
            G.ocov[v] = (uint32_t)run;
            uint32_t c = std::min((uint32_t)run, M);
            G.odemand[v] = (int32_t)prev_capped - (int32_t)c;
            prev_capped = c;
        }
    }
    G.through = 0;
    for (const VSample& v : G.vs)
        for (uint32_t j = 1; j < v.nseg; ++j) {
            uint32_t x = v.obase + j * seg;  // cut node: left position x-1, right position x
            G.through += std::min(std::min(G.ocov[x - 1], M), std::min(G.ocov[x], M));
        }
    // out-CSR: bundles are sorted by key and the real start is monotone in the key
    G.out_ptr.assign(nn + 1, 0);
    for (uint32_t b = 0; b < B; ++b) ++G.out_ptr[G.b_s[b] + 1];
    for (uint32_t v = 0; v < nn; ++v) G.out_ptr[v + 1] += G.out_ptr[v];
    G.in_ptr.assign(nn + 1, 0);
    for (uint32_t b = 0; b < B; ++b) ++G.in_ptr[G.b_t[b] + 1];
    for (uint32_t v = 0; v < nn; ++v) G.in_ptr[v + 1] += G.in_ptr[v];
    G.in_bid.resize(B);
    {
        std::vector<uint32_t> cursor(G.in_ptr.begin(), G.in_ptr.end() - 1);
        for (uint32_t b = 0; b < B; ++b) G.in_bid[cursor[G.b_t[b]]++] = b;  // ascending bundle id
    }
    // Forced reads out, cuts in: a bundle that covers a position with cov <= M is in
    // every valid answer — its flow is fixed at its multiplicity, its ends' demands absorb it and
    // it leaves the residual graph (capacity 0 during the solve, b_forced remembers the reads) —
    // and the back arc over such a position carries cov_S - min(cov, M) = 0 in every valid answer,
    // so the component rule becomes "v and v+1 joined iff covR[v] > M".
    G.b_forced.assign(B, 0);
    G.M = M;
    G.vsupply = 0;
    for (uint32_t v = 0; v < nn; ++v)
        if (G.demand[v] < 0) G.vsupply += -(int64_t)G.demand[v];
    const uint32_t cut_at = forced_cuts ? M : 0;
    if (forced_cuts) {
        std::vector<uint32_t> unc(nn + 1, 0);  // positions p < v with covR[p] <= M
        for (uint32_t v = 0; v < nn; ++v) unc[v + 1] = unc[v] + (G.covR[v] <= M ? 1u : 0u);
        for (uint32_t b = 0; b < B; ++b)
            if (unc[G.b_t[b]] - unc[G.b_s[b]] > 0) {
                G.b_forced[b] = G.b_mult[b];
                G.demand[G.b_s[b]] += (int32_t)G.b_mult[b];
                G.demand[G.b_t[b]] -= (int32_t)G.b_mult[b];
                G.b_mult[b] = 0;
            }
    }
    // components: v and v+1 joined iff covR[v] > cut_at (App. A.3: 0); segment ends have covR == 0
    uint32_t lo = 0;
    for (uint32_t v = 0; v < nn; ++v) {
        if (G.covR[v] <= cut_at) {
            if (v > lo) {
                G.comp_lo.push_back(lo);
                G.comp_hi.push_back(v);
            }
            lo = v + 1;
        }
    }
    return 0;
}

struct SyncState {
    std::vector<uint32_t> d;      // labels
    std::vector<int32_t> e, eadd, snk, g;
    std::vector<uint32_t> f;      // bundle flows
    std::vector<uint32_t> stamp;
    std::vector<uint8_t> sat, satn;
};

struct CompStats {
    uint64_t rounds = 0, pushes = 0, relabels = 0, grs = 0, bfs_levels = 0, max_frontier = 0;
    int64_t sink_flow = 0;
    int64_t stuck = 0;
    uint64_t express = 0;
};

// The EXPRESS schedule (csrc/maxflow_sm.cuh): back arcs v -> v-1 have length 0 in
// the distance labels, every other residual arc length 1.  Excess then changes lane (moves left
// past saturated nodes) at no label cost and in one round, which is what a segment of a long
// reference needs: all its supply sits at the left end and all its sinks at the right end, so M
// units must be spread over the read "lanes" and gathered again.  Chosen per component from the
// data alone (express_component below), never from the device or the batch.
constexpr uint32_t kExpressMaxNodes = 40961, kExpressEdge = 512, kHeavyDegOracle = 6;

uint32_t sync_global_relabel(const SyncGraph& G, SyncState& S, uint32_t lo, uint32_t hi,
                             CompStats& cs, bool express) {
    ++cs.grs;
    for (uint32_t v = lo; v <= hi; ++v) S.d[v] = LBL_INF;
    std::vector<uint32_t> cur, nxt;
    for (uint32_t v = lo; v <= hi; ++v)
        if (S.snk[v] > 0) {
            S.d[v] = 1;
            cur.push_back(v);
        }
    uint32_t level = 1;
    while (!cur.empty()) {
        ++cs.bfs_levels;
        if (express) {  // a level is closed under "right neighbour" (zero-length back arcs) first
            const size_t seeds = cur.size();
            for (size_t i = 0; i < seeds; ++i)
                for (uint32_t u = cur[i] + 1; u <= hi && S.d[u] == LBL_INF; ++u) {
                    S.d[u] = level;
                    cur.push_back(u);
                }
        }
        nxt.clear();
        auto visit = [&](uint32_t u) {
            if (S.d[u] == LBL_INF) {
                S.d[u] = level + 1;
                nxt.push_back(u);
            }
        };
        for (uint32_t w : cur) {
            if (!express && w < hi) visit(w + 1);           // back arc (w+1) -> w, always residual
            if (w > lo && S.g[w] > 0) visit(w - 1);         // reverse of back arc w -> w-1
            for (uint32_t k = G.in_ptr[w]; k < G.in_ptr[w + 1]; ++k) {
                uint32_t b = G.in_bid[k];
                if (S.f[b] < G.b_mult[b]) visit(G.b_s[b]);  // bundle arc s -> w has residual
            }
            for (uint32_t b = G.out_ptr[w]; b < G.out_ptr[w + 1]; ++b)
                if (S.f[b] > 0) visit(G.b_t[b]);            // reverse arc t -> w has residual
        }
        cur.swap(nxt);
        ++level;
    }
    return level;
}

// Which components run the express schedule — decided from the component's data alone: few
// bundles per node, a supply of at least kExpressMinSupply (below that the classic schedule needs
// no more rounds than hops: config 5's M = 100) that fits 16 bits, at most kExpressMaxNodes nodes
// (what one SM holds), every supply within kExpressEdge nodes of the left end and every sink within
// kExpressEdge of the right end (true by construction once the forced reads are out, for reads of up
// to kExpressEdge positions).  schedule 2 drops the lower bound on the supply (experiments).
constexpr uint32_t kExpressMinSupply = 128;
bool express_component(const SyncGraph& G, uint32_t lo, uint32_t hi, uint32_t schedule) {
    if (schedule == 1 || (schedule != 2 && G.M < kExpressMinSupply)) return false;  // (the host's gate)
    const uint32_t n = hi - lo + 1;
    if (n > kExpressMaxNodes) return false;
    const uint64_t n_bund = G.out_ptr[hi + 1] - G.out_ptr[lo];
    if (2 * n_bund > (uint64_t)kHeavyDegOracle * n) return false;
    uint64_t supply = 0;
    for (uint32_t v = lo; v <= hi; ++v) {
        const int32_t dm = G.demand[v];
        if (dm < 0) {
            supply += (uint32_t)(-dm);
            if (v - lo >= kExpressEdge) return false;
        } else if (dm > 0 && hi - v >= kExpressEdge) {
            return false;
        }
    }
    return supply <= 0xffffu && (schedule == 2 || supply >= kExpressMinSupply);
}

void sync_solve_component(const SyncGraph& G, SyncState& S, uint32_t lo, uint32_t hi,
                          const orc_sync_params& P, CompStats& cs) {
    const uint32_t ncomp = hi - lo + 1;
    const bool express = express_component(G, lo, hi, P.schedule);
    const uint32_t bl = express ? 0u : 1u;  // length of a back arc in the labels
    if (express) ++cs.express;
    // express: "saturated" bits — no sink capacity and no residual on any own bundle, as far as
    // the rules below know (0 is always safe).  Walkers read the snapshot of the round start.
    std::vector<uint8_t>& sat = S.sat;
    std::vector<uint8_t>& satn = S.satn;
    std::vector<uint32_t> sat_touched;
    if (express)
        for (uint32_t v = lo; v <= hi; ++v) sat[v] = satn[v] = 0;
    for (uint32_t v = lo; v <= hi; ++v) {
        int32_t dm = G.demand[v];
        S.e[v] = dm < 0 ? -dm : 0;
        S.snk[v] = dm > 0 ? dm : 0;
        S.g[v] = 0;
        S.eadd[v] = 0;
    }
    uint32_t last_levels = sync_global_relabel(G, S, lo, hi, cs, express);
    std::vector<uint32_t> F, T, NF;
    uint32_t round = 0;
    for (uint32_t v = lo; v <= hi; ++v)
        if (S.e[v] > 0) {
            F.push_back(v);
            S.stamp[v] = 1;
        }
    uint64_t relabels_since = 0, rounds_since = 0;
    std::vector<std::pair<uint32_t, uint32_t>> newlab;
    while (!F.empty()) {
        if (P.max_rounds && cs.rounds >= P.max_rounds) break;
        // deterministic global-relabel trigger (state-only)
        // (express: a level is a whole read hop, and lane changes cost rounds but no levels: twice the interval)
        uint64_t interval = std::max<uint64_t>(P.gr_interval_min,
                                               (uint64_t)last_levels * P.gr_levels_pct / 100 * (2u - bl));
        if (rounds_since >= interval &&
            relabels_since * 100 >= (uint64_t)P.gr_relabel_pct * ncomp) {
            last_levels = sync_global_relabel(G, S, lo, hi, cs, express);
            relabels_since = 0;
            rounds_since = 0;
        }
        ++round;
        ++cs.rounds;
        ++rounds_since;
        cs.max_frontier = std::max<uint64_t>(cs.max_frontier, F.size());
        T.clear();
        auto give = [&](uint32_t w, int32_t delta) {
            S.eadd[w] += delta;
            if (S.stamp[w] != round) {
                S.stamp[w] = round;
                T.push_back(w);
            }
            ++cs.pushes;
        };
        // ---- phase A: pushes decided from the labels at the start of the round ----
        for (uint32_t v : F) {
            const uint32_t dv = S.d[v];
            if (dv >= LBL_INF) continue;
            int32_t ex = S.e[v];
            // 1. sink
            if (dv == 1 && S.snk[v] > 0) {
                int32_t dl = std::min(ex, S.snk[v]);
                S.snk[v] -= dl;
                ex -= dl;
                cs.sink_flow += dl;
                ++cs.pushes;
                // U4: the sink of a node without a live bundle of its own (all forced, or none) is
                // full — from now on walkers pass it
                if (express && S.snk[v] == 0) {
                    bool live = false;
                    for (uint32_t b = G.out_ptr[v]; !live && b < G.out_ptr[v + 1]; ++b) live = G.b_mult[b] != 0;
                    if (!live) {
                        satn[v] = 1;
                        sat_touched.push_back(v);
                    }
                }
            }
            // 2. own bundles, farthest end first
            for (uint32_t b = G.out_ptr[v + 1]; ex > 0 && b-- > G.out_ptr[v];) {
                uint32_t t = G.b_t[b];
                if (S.d[t] + 1 != dv) continue;
                uint32_t r = G.b_mult[b] - S.f[b];
                if (r == 0) continue;
                int32_t dl = (int32_t)std::min<uint32_t>((uint32_t)ex, r);
                S.f[b] += dl;
                ex -= dl;
                give(t, dl);
                // U1: the owner filled its only bundle (nobody else can touch it this round)
                if (express && (uint32_t)dl == r && G.out_ptr[v + 1] - G.out_ptr[v] == 1 && S.snk[v] == 0) {
                    satn[v] = 1;
                    sat_touched.push_back(v);
                }
            }
            // 3. cancel back-flow towards the right neighbour
            if (ex > 0 && v < hi && S.d[v + 1] + 1 == dv && S.g[v + 1] > 0) {
                int32_t dl = std::min(ex, S.g[v + 1]);
                S.g[v + 1] -= dl;
                ex -= dl;
                give(v + 1, dl);
            }
            // 4. back arc to the left neighbour (infinite capacity)
            if (ex > 0 && v > lo && S.d[v - 1] + bl == dv) {
                uint32_t u = v;
                if (express) {  // walk the zero-length chain past nodes known to be saturated
                    for (;;) {
                        S.g[u] += ex;
                        --u;
                        if (u == lo || !sat[u] || S.d[u - 1] != S.d[u]) break;
                    }
                } else {
                    S.g[v] += ex;
                    --u;
                }
                give(u, ex);
                ex = 0;
            }
            // 5. cancel flow on incoming bundles, nearest start first
            for (uint32_t k = G.in_ptr[v + 1]; ex > 0 && k-- > G.in_ptr[v];) {
                uint32_t b = G.in_bid[k];
                uint32_t s = G.b_s[b];
                if (S.d[s] + 1 != dv || S.f[b] == 0) continue;
                int32_t dl = (int32_t)std::min<uint32_t>((uint32_t)ex, S.f[b]);
                S.f[b] -= dl;
                ex -= dl;
                give(s, dl);
                if (express) {  // U2: s has residual capacity again
                    satn[s] = 0;
                    sat_touched.push_back(s);
                }
            }
            S.e[v] = ex;
        }
        // ---- phase B: merge received excess, relabel from the label snapshot, next frontier ----
        NF.clear();
        newlab.clear();
        auto phase_b = [&](uint32_t w, bool in_front) {
            int32_t left = S.e[w];
            if (in_front && left > 0 && S.d[w] < LBL_INF) {
                uint32_t mn = LBL_INF;
                if (S.snk[w] > 0) mn = 0;
                for (uint32_t b = G.out_ptr[w]; b < G.out_ptr[w + 1]; ++b)
                    if (S.f[b] < G.b_mult[b]) mn = std::min(mn, S.d[G.b_t[b]]);
                if (w < hi && S.g[w + 1] > 0) mn = std::min(mn, S.d[w + 1]);
                if (w > lo) mn = std::min(mn, S.d[w - 1] - (1u - bl));  // labels are >= 1
                for (uint32_t k = G.in_ptr[w]; k < G.in_ptr[w + 1]; ++k) {
                    uint32_t b = G.in_bid[k];
                    if (S.f[b] > 0) mn = std::min(mn, S.d[G.b_s[b]]);
                }
                uint32_t nl = mn >= LBL_INF ? LBL_INF : mn + 1;
                newlab.emplace_back(w, nl);
                if (express) {  // U3: the relabel has just looked at every own bundle
                    bool any = S.snk[w] > 0;
                    for (uint32_t b = G.out_ptr[w]; !any && b < G.out_ptr[w + 1]; ++b)
                        any = S.f[b] < G.b_mult[b];
                    satn[w] = any ? 0 : 1;
                    sat_touched.push_back(w);
                }
                ++cs.relabels;
                ++relabels_since;
            }
            int32_t tot = left + S.eadd[w];
            S.eadd[w] = 0;
            S.e[w] = tot;
            if (tot > 0) {
                NF.push_back(w);
                S.stamp[w] = round + 1;
            }
        };
        for (uint32_t v : F) phase_b(v, true);
        for (uint32_t w : T) phase_b(w, false);
        for (auto& pr : newlab) S.d[pr.first] = pr.second;  // applied after all reads (snapshot)
        for (uint32_t w : sat_touched) sat[w] = satn[w];    // the next round's snapshot
        sat_touched.clear();
        // drop frozen nodes (cannot happen on this network — SURVEY App. A.1 — but stay safe)
        F.clear();
        for (uint32_t w : NF) {
            if (S.d[w] >= LBL_INF) {
                cs.stuck += S.e[w];
                continue;
            }
            F.push_back(w);
        }
    }
}
}  // namespace

extern "C" int orc_sync_solve(uint32_t n_samples, const uint64_t* read_off, const uint32_t* ref_len,
                              const uint32_t* start, const uint32_t* end, uint32_t M,
                              const orc_sync_params* prm, uint32_t* kept_bitmap,
                              int32_t* demand_out, uint32_t* cov_out, orc_sync_stats* st) {
    orc_sync_params P = prm ? *prm : orc_sync_params{64, 150, 1, 0, 0, 0};
    auto t0 = clk::now();
    SyncGraph G;
    // forced reads out + cuts: from the supply at which the express schedule starts (the host's gate)
    const bool forced_cuts = P.schedule >= 2 || (P.schedule != 1 && M >= kExpressMinSupply);
    if (build_sync_graph(n_samples, read_off, ref_len, start, end, M, P.seg_len, G, forced_cuts) != 0)
        return -1;
    auto t1 = clk::now();
    const uint32_t nn = G.n_nodes;
    const uint32_t B = (uint32_t)G.b_s.size();
    SyncState S;
    S.d.assign(nn, LBL_INF);
    S.e.assign(nn, 0);
    S.eadd.assign(nn, 0);
    S.snk.assign(nn, 0);
    S.g.assign(nn, 0);
    S.f.assign(B, 0);
    S.stamp.assign(nn, 0);
    S.sat.assign(nn, 0);
    S.satn.assign(nn, 0);
    orc_sync_stats out{};
    for (size_t c = 0; c < G.comp_lo.size(); ++c) {
        CompStats cs;
        sync_solve_component(G, S, G.comp_lo[c], G.comp_hi[c], P, cs);
        out.flow_value += cs.sink_flow;
        out.rounds_total += cs.rounds;
        out.rounds_max = std::max(out.rounds_max, cs.rounds);
        out.pushes += cs.pushes;
        out.relabels += cs.relabels;
        out.global_relabels += cs.grs;
        out.bfs_levels += cs.bfs_levels;
        out.max_frontier = std::max(out.max_frontier, cs.max_frontier);
        out.n_express += (uint32_t)cs.express;
    }
    // the forced bundles come back with their fixed flow; what the residual problem left
    // undelivered (nothing, on valid input) is missing from the flow value
    int64_t res_supply = 0;
    for (uint32_t v = 0; v < nn; ++v)
        if (G.demand[v] < 0) res_supply += -(int64_t)G.demand[v];
    out.flow_value = G.vsupply - (res_supply - out.flow_value);
    for (uint32_t b = 0; b < B; ++b)
        if (G.b_forced[b]) {
            G.b_mult[b] = G.b_forced[b];
            S.f[b] = G.b_forced[b];
        }
    auto t2 = clk::now();
    const uint64_t N = read_off[n_samples];
    std::fill(kept_bitmap, kept_bitmap + (N + 31) / 32, 0u);
    uint64_t nk = 0;
    for (uint32_t b = 0; b < B; ++b) {
        for (uint32_t r = 0; r < S.f[b]; ++r) {
            uint32_t i = G.sorted_idx[G.b_first[b] + r];  // a crossing read may be chosen twice
            if (!(kept_bitmap[i >> 5] >> (i & 31) & 1u)) ++nk;
            kept_bitmap[i >> 5] |= 1u << (i & 31);
        }
    }
    auto t3 = clk::now();
    int64_t fstar = 0;  // closed form on the ORIGINAL network
    for (uint32_t v = 0; v < G.n_onodes; ++v)
        if (G.odemand[v] < 0) fstar += -(int64_t)G.odemand[v];
    if (demand_out) std::copy(G.odemand.begin(), G.odemand.end(), demand_out);
    if (cov_out) std::copy(G.ocov.begin(), G.ocov.end(), cov_out);
    out.flow_value -= G.through;  // value of the stitched flow on the original network
    out.fstar = fstar;
    out.n_kept = nk;
    out.n_bundles = B;
    out.n_components = (uint32_t)G.comp_lo.size();
    out.t_build_s = secs(t0, t1);
    out.t_solve_s = secs(t1, t2);
    out.t_select_s = secs(t2, t3);
    if (st) *st = out;
    return 0;
}

// =====================================================================================
// Minimum-cardinality solve on the bundled graph = what `mcp-cpu` computes
// (mcp_cpu_cost_scaling_solver.cpp:33-67: every read arc costs 1, so the optimum is the fewest reads
// with cov_S >= min(cov, M)), restated as the deterministic sweep the device runs (csrc/sweep.cuh):
// per component, positions left to right; reads that have started are pooled BY END NODE; on a
// deficit take from the farthest end first (greedy interval multicover, exact for this objective);
// afterwards the reads taken with end node t are handed to the bundles ending at t in in-CSR order
// (earliest start first — a read of the same end with an earlier start covers a superset).
// Segmented references: every segment is swept on its own (at most M extra reads per cut).
// =====================================================================================
extern "C" int orc_sweep_solve(uint32_t n_samples, const uint64_t* read_off, const uint32_t* ref_len,
                               const uint32_t* start, const uint32_t* end, uint32_t M,
                               const orc_sync_params* prm, uint32_t* kept_bitmap,
                               int32_t* demand_out, uint32_t* cov_out, orc_sync_stats* st) {
    orc_sync_params P = prm ? *prm : orc_sync_params{64, 150, 1, 0, 0, 0};
    SyncGraph G;
    if (build_sync_graph(n_samples, read_off, ref_len, start, end, M, P.seg_len, G) != 0) return -1;
    const uint32_t nn = G.n_nodes;
    const uint32_t B = (uint32_t)G.b_s.size();
    std::vector<uint32_t> taken(nn + 1, 0), pool(nn + 1, 0), f(B, 0);
    orc_sync_stats out{};
    for (size_t c = 0; c < G.comp_lo.size(); ++c) {
        const uint32_t lo = G.comp_lo[c], hi = G.comp_hi[c];
        uint64_t have = 0;
        uint32_t top = lo;
        for (uint32_t p = lo; p < hi; ++p) {
            have -= taken[p];  // reads whose end node is p stop covering here
            pool[p] = 0;
            for (uint32_t b = G.out_ptr[p]; b < G.out_ptr[p + 1]; ++b) {
                if (G.b_t[b] == p) continue;  // length 0: covers nothing
                pool[G.b_t[b]] += G.b_mult[b];
                top = std::max(top, G.b_t[b]);
            }
            const uint64_t need = std::min(G.covR[p], M);
            while (have < need) {
                while (top > p && pool[top] == 0) --top;
                if (top <= p) return -2;  // cannot happen: need <= coverage
                const uint64_t k = std::min<uint64_t>(need - have, pool[top]);
                pool[top] -= (uint32_t)k;
                taken[top] += (uint32_t)k;
                have += k;
            }
        }
        out.flow_value += 0;
    }
    for (uint32_t t = 0; t < nn; ++t) {  // hand the reads taken per end node to the bundles ending there
        uint32_t rem = taken[t];
        for (uint32_t k = G.in_ptr[t]; k < G.in_ptr[t + 1] && rem; ++k) {
            const uint32_t b = G.in_bid[k];
            if (G.b_s[b] == G.b_t[b]) continue;
            const uint32_t x = std::min(rem, G.b_mult[b]);
            f[b] = x;
            rem -= x;
        }
        if (rem) return -3;
    }
    const uint64_t N = read_off[n_samples];
    std::fill(kept_bitmap, kept_bitmap + (N + 31) / 32, 0u);
    uint64_t nk = 0;
    for (uint32_t b = 0; b < B; ++b)
        for (uint32_t r = 0; r < f[b]; ++r) {
            uint32_t i = G.sorted_idx[G.b_first[b] + r];  // a crossing read may be chosen twice
            if (!(kept_bitmap[i >> 5] >> (i & 31) & 1u)) ++nk;
            kept_bitmap[i >> 5] |= 1u << (i & 31);
        }
    int64_t fstar = 0;
    for (uint32_t v = 0; v < G.n_onodes; ++v)
        if (G.odemand[v] < 0) fstar += -(int64_t)G.odemand[v];
    if (demand_out) std::copy(G.odemand.begin(), G.odemand.end(), demand_out);
    if (cov_out) std::copy(G.ocov.begin(), G.ocov.end(), cov_out);
    out.flow_value = fstar;
    out.fstar = fstar;
    out.n_kept = nk;
    out.n_bundles = B;
    out.n_components = (uint32_t)G.comp_lo.size();
    if (st) *st = out;
    return 0;
}

// =====================================================================================
// Greedy interval multicover = minimum number of kept reads subject to cov_S >= min(cov, M)
// (the objective of mcp-cpu: read arcs cost 1, mcp_cpu_cost_scaling_solver.cpp:45-48).
// Sweep left to right; on a deficit at position p take the available read covering p with the
// farthest end.
// =====================================================================================
extern "C" uint64_t orc_greedy_multicover(uint64_t n, const uint32_t* start, const uint32_t* end,
                                          uint32_t L, uint32_t M, uint8_t* kept) {
    std::vector<uint32_t> cov(L);
    if (orc_coverage_ref(n, start, end, L, cov.data()) != 0) return (uint64_t)-1;
    std::fill(kept, kept + n, 0);
    std::vector<uint32_t> order(n);
    for (uint64_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return start[a] != start[b] ? start[a] < start[b] : a < b;
    });
    // max-heap on end of reads started so far and not taken
    std::vector<std::pair<uint32_t, uint32_t>> heap;  // (end, ~idx) so lowest idx wins ties
    std::vector<int32_t> drop(L + 1, 0);
    int64_t have = 0;
    uint64_t nk = 0, ptr = 0;
    for (uint32_t p = 0; p < L; ++p) {
        have -= drop[p];
        while (ptr < n && start[order[ptr]] <= p) {
            uint32_t i = order[ptr++];
            heap.emplace_back(end[i], ~i);
            std::push_heap(heap.begin(), heap.end());
        }
        int64_t need = std::min(cov[p], M);
        while (have < need) {
            std::pop_heap(heap.begin(), heap.end());
            auto top = heap.back();
            heap.pop_back();
            if (top.first < p) continue;  // stale (cannot happen while need <= cov[p])
            uint32_t i = ~top.second;
            kept[i] = 1;
            ++nk;
            ++have;
            drop[top.first + 1] += 1;
        }
    }
    return nk;
}

// find_pairs (bam_api.cpp:239-273): mates are adjacent (first at even index), so on a bitmap
// "add each kept read's mate" is: both bits of a pair = OR of the two.
extern "C" void orc_find_pairs_bitmap(uint64_t n, uint32_t* bitmap) {
    uint64_t words = (n + 31) / 32;
    for (uint64_t w = 0; w < words; ++w) {
        uint32_t x = bitmap[w];
        uint32_t evens = x & 0x55555555u, odds = x & 0xaaaaaaaau;
        uint32_t any = evens | (odds >> 1);
        bitmap[w] = any | (any << 1);
    }
    if (n % 32) bitmap[words - 1] &= (1u << (n % 32)) - 1;
}
