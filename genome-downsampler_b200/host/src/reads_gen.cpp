// Synthetic paired reads.  Law (not code) of libs/reads-gen/src/reads_gen.cpp:
//   uniform  (:55-86): first ~ U{0..G-2R}, second ~ U{0..G-R}; order them; if they overlap the
//                      second is moved to first+R; qualities ~ U{0..max_quality}
//   weighted (:5-53) : both starts from a discrete distribution over G-R+1 start positions built
//                      from dist_func on [0,1] (negative weights clamp to 0); if both fall in the
//                      last 2R bases they are pinned to G-2R and G-R, else overlap is resolved as
//                      above.
// Draw order per pair: first, second, quality(first), quality(second).
#include "reads_gen.hpp"

#include <algorithm>
#include <vector>

namespace {
// Sink = where generated reads go: the AoS container, or caller-owned SoA arrays (no allocation,
// so many samples can be generated concurrently on host threads).
struct AosSink {
    bam_api::AOSPairedReads out;
    void push_back(const bam_api::Read& r) { out.push_back(r); }
};
struct SoaSink {
    reads_gen::SoaOut o;
    void push_back(const bam_api::Read& r) {
        o.start[r.bam_id] = static_cast<uint32_t>(r.start_ind);
        o.end[r.bam_id] = static_cast<uint32_t>(r.end_ind);
        if (o.mapq) o.mapq[r.bam_id] = static_cast<uint8_t>(r.quality);
        if (o.seq_len) o.seq_len[r.bam_id] = r.seq_length;
    }
};

template <typename Sink, typename StartDraw>
void generate(Sink& out, std::mt19937& gen, bam_api::ReadIndex pairs, bam_api::Index G,
              uint32_t R, int32_t max_quality, bool pin_tail, StartDraw draw) {
    std::uniform_int_distribution<> quality(0, max_quality);
    for (bam_api::ReadIndex p = 0; p < pairs; ++p) {
        auto [a, b] = draw(gen);
        if (a > b) std::swap(a, b);
        if (pin_tail && a > G - 2 * R && b > G - 2 * R) {
            a = G - 2 * R;
            b = G - R;
        } else if (a + R > b) {
            b = a + R;
        }
        uint32_t qa = static_cast<uint32_t>(quality(gen));
        out.push_back(bam_api::Read(2 * p, a, a + R - 1, qa, R, true));
        uint32_t qb = static_cast<uint32_t>(quality(gen));
        out.push_back(bam_api::Read(2 * p + 1, b, b + R - 1, qb, R, false));
    }
}

template <typename Sink>
void uniform_into(Sink& sink, std::mt19937& generator, bam_api::ReadIndex pairs_count,
                  bam_api::Index genome_length, uint32_t read_length, int32_t max_quality) {
    std::uniform_int_distribution<> d1(0, static_cast<int32_t>(genome_length - 2 * read_length));
    std::uniform_int_distribution<> d2(0, static_cast<int32_t>(genome_length - read_length));
    generate(sink, generator, pairs_count, genome_length, read_length, max_quality, false,
             [&](std::mt19937& g) {
                 bam_api::Index a = d1(g);
                 bam_api::Index b = d2(g);
                 return std::pair<bam_api::Index, bam_api::Index>(a, b);
             });
}

template <typename Sink>
void weighted_into(Sink& sink, std::mt19937& generator, bam_api::ReadIndex pairs_count,
                   bam_api::Index genome_length, uint32_t read_length,
                   const std::function<double(double)>& dist_func, int32_t max_quality) {
    const uint32_t n_starts = static_cast<uint32_t>(genome_length - read_length + 1);
    std::vector<double> w(n_starts);
    double total = 0;
    for (uint32_t i = 0; i < n_starts; ++i) {
        w[i] = std::max(0.0, dist_func(static_cast<double>(i) / static_cast<double>(n_starts - 1)));
        total += w[i];
    }
    for (double& x : w) x /= total;
    std::discrete_distribution<> dd(w.begin(), w.end());
    generate(sink, generator, pairs_count, genome_length, read_length, max_quality, true,
             [&](std::mt19937& g) {
                 bam_api::Index a = dd(g);
                 bam_api::Index b = dd(g);
                 return std::pair<bam_api::Index, bam_api::Index>(a, b);
             });
}
}  // namespace

bam_api::AOSPairedReads reads_gen::rand_reads_uniform(std::mt19937& generator,
                                                      bam_api::ReadIndex pairs_count,
                                                      bam_api::Index genome_length,
                                                      uint32_t read_length, int32_t max_quality) {
    AosSink sink;
    sink.out.ref_genome_length = genome_length;
    sink.out.reserve(2 * pairs_count);
    uniform_into(sink, generator, pairs_count, genome_length, read_length, max_quality);
    return std::move(sink.out);
}

bam_api::AOSPairedReads reads_gen::rand_reads(std::mt19937& generator,
                                              bam_api::ReadIndex pairs_count,
                                              bam_api::Index genome_length, uint32_t read_length,
                                              const std::function<double(double)>& dist_func,
                                              int32_t max_quality) {
    AosSink sink;
    sink.out.ref_genome_length = genome_length;
    sink.out.reserve(2 * pairs_count);
    weighted_into(sink, generator, pairs_count, genome_length, read_length, dist_func, max_quality);
    return std::move(sink.out);
}

void reads_gen::rand_reads_uniform_soa(std::mt19937& generator, bam_api::ReadIndex pairs_count,
                                       bam_api::Index genome_length, uint32_t read_length,
                                       const SoaOut& out, int32_t max_quality) {
    SoaSink sink{out};
    uniform_into(sink, generator, pairs_count, genome_length, read_length, max_quality);
}

void reads_gen::rand_reads_soa(std::mt19937& generator, bam_api::ReadIndex pairs_count,
                               bam_api::Index genome_length, uint32_t read_length,
                               const std::function<double(double)>& dist_func, const SoaOut& out,
                               int32_t max_quality) {
    SoaSink sink{out};
    weighted_into(sink, generator, pairs_count, genome_length, read_length, dist_func, max_quality);
}

void reads_gen::rand_reads_amplicon_soa(std::mt19937& gen, bam_api::ReadIndex pairs,
                                        bam_api::Index genome_length,
                                        const std::vector<uint32_t>& amp_start,
                                        const std::vector<uint32_t>& amp_end, double p_inside,
                                        uint32_t min_len, uint32_t max_len, const SoaOut& out,
                                        int32_t max_quality) {
    std::uniform_int_distribution<> quality(0, max_quality);
    std::uniform_int_distribution<> length(static_cast<int>(min_len), static_cast<int>(max_len));
    std::uniform_int_distribution<> pick(0, static_cast<int>(amp_start.size()) - 1);
    std::uniform_real_distribution<double> unit(0.0, 1.0);
    for (bam_api::ReadIndex p = 0; p < pairs; ++p) {
        const uint32_t l1 = static_cast<uint32_t>(length(gen));
        const uint32_t l2 = static_cast<uint32_t>(length(gen));
        const bool inside = unit(gen) < p_inside;
        const uint32_t k = static_cast<uint32_t>(pick(gen));
        const uint64_t a0 = amp_start[k], a1 = amp_end[k];
        uint64_t s1, s2;
        if (inside && a1 - a0 + 1 >= std::max(l1, l2)) {
            std::uniform_int_distribution<uint64_t> d1(a0, a1 + 1 - l1), d2(a0, a1 + 1 - l2);
            s1 = d1(gen);
            s2 = d2(gen);
        } else {
            std::uniform_int_distribution<uint64_t> d1(0, genome_length - l1), d2(0, genome_length - l2);
            s1 = d1(gen);
            s2 = d2(gen);
        }
        const uint64_t i = 2 * p;
        out.start[i] = static_cast<uint32_t>(s1);
        out.end[i] = static_cast<uint32_t>(s1 + l1 - 1);
        const uint32_t q1 = static_cast<uint32_t>(quality(gen));
        out.start[i + 1] = static_cast<uint32_t>(s2);
        out.end[i + 1] = static_cast<uint32_t>(s2 + l2 - 1);
        const uint32_t q2 = static_cast<uint32_t>(quality(gen));
        if (out.mapq) {
            out.mapq[i] = static_cast<uint8_t>(q1);
            out.mapq[i + 1] = static_cast<uint8_t>(q2);
        }
        if (out.seq_len) {
            out.seq_len[i] = l1;
            out.seq_len[i + 1] = l2;
        }
    }
}
