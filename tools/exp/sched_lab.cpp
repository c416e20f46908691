// Schedule lab (research tool, CPU): the bulk-synchronous push-relabel schedule of
// oracle/gds_oracle.cpp on a single-length, single-component instance, with switches for
// experimental variants.  Prints rounds / pushes / relabels and optionally a per-round trace.
// build: g++ -O2 -std=c++17 -o tools/exp/sched_lab tools/exp/sched_lab.cpp
// usage: sched_lab [L=..] [R=..] [cov=..] [M=..] [seed=..] [shape=0|1] [file=starts.u32] [trace=N]
//                  [gri=..] [grl=..] [grr=..] [K=gate spacing] [variant=bits]
// variant bits: 1 relabel count alone may trigger a global relabel; 2 walk the chain of admissible
// back arcs (unit lengths); 4 back arcs have length 0 (+ walk); 8 no walk; 16 walk stops on
// label-admissible bundles regardless of saturation; 32 walk stops on the saturation bit only (what
// the device does: variant 36 = the express schedule); 64 warm start: bundles covering a position
// with cov <= M start saturated; 128 the same bundles leave the graph and the back arcs over such
// positions are cut (variant 164 = that + the express schedule).
// file= takes the reference generator's reads (oracle: gen_reads(...)[0].tofile(path)).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
using namespace std;
static const uint32_t INF = 0x3fffffff;
struct G {
    uint32_t n;  // nodes 0..n-1 (n = L+1)
    vector<uint32_t> bs, bt, bm, out_ptr, in_ptr, in_bid;
    vector<int32_t> dem;
    vector<uint32_t> unc;  // unc[v] = number of positions p < v with cov[p] <= M
    vector<uint8_t> cut;   // cut[v]: the back arc v -> v-1 does not exist (position v-1 is not capped)
};
struct Opt {
    int K = 0;
    int gri = 64, grl = 150, grr = 1;
    int variant = 0;
    int trace = 0;
    int twophase = 0;
};
struct St {
    uint64_t rounds = 0, pushes = 0, relabels = 0, grs = 0, levels = 0, maxf = 0;
};
// variant 4: back arcs (w+1 -> w) have length 0, every other residual arc length 1.  A level is
// closed under "right neighbour" before the next one starts.
static int g_K = 0;
static inline uint32_t blen(uint32_t v) { return g_K && v % g_K == 0 ? 1u : 0u; }  // length of back arc v -> v-1
static uint32_t global_relabel0(const G& g, vector<uint32_t>& d, const vector<int32_t>& snk,
                                const vector<int32_t>& gb, const vector<uint32_t>& f, St& st) {
    ++st.grs;
    fill(d.begin(), d.end(), INF);
    vector<uint32_t> cur, nxt;
    for (uint32_t v = 0; v < g.n; ++v)
        if (snk[v] > 0) { d[v] = 1; cur.push_back(v); }
    uint32_t level = 1;
    while (!cur.empty()) {
        ++st.levels;
        // closure: everything to the right of a level node, up to the next labelled node
        size_t k0 = cur.size();
        for (size_t i = 0; i < k0; ++i)
            for (uint32_t u = cur[i] + 1; u < g.n && d[u] == INF && blen(u) == 0 && !g.cut[u]; ++u) { d[u] = level; cur.push_back(u); }
        nxt.clear();
        auto visit = [&](uint32_t u) { if (d[u] == INF) { d[u] = level + 1; nxt.push_back(u); } };
        for (uint32_t w : cur) {
            if (w + 1 < g.n && blen(w + 1) && !g.cut[w + 1]) visit(w + 1);  // gate arc (w+1) -> w, length 1
            if (w > 0 && gb[w] > 0) visit(w - 1);
            for (uint32_t k = g.in_ptr[w]; k < g.in_ptr[w + 1]; ++k) { uint32_t b = g.in_bid[k]; if (f[b] < g.bm[b]) visit(g.bs[b]); }
            for (uint32_t b = g.out_ptr[w]; b < g.out_ptr[w + 1]; ++b) if (f[b] > 0) visit(g.bt[b]);
        }
        cur.swap(nxt);
        ++level;
    }
    return level;
}
static uint32_t global_relabel(const G& g, vector<uint32_t>& d, const vector<int32_t>& snk,
                               const vector<int32_t>& gb, const vector<uint32_t>& f, St& st) {
    ++st.grs;
    fill(d.begin(), d.end(), INF);
    vector<uint32_t> cur, nxt;
    for (uint32_t v = 0; v < g.n; ++v)
        if (snk[v] > 0) { d[v] = 1; cur.push_back(v); }
    uint32_t level = 1;
    while (!cur.empty()) {
        ++st.levels;
        nxt.clear();
        auto visit = [&](uint32_t u) { if (d[u] == INF) { d[u] = level + 1; nxt.push_back(u); } };
        for (uint32_t w : cur) {
            if (w + 1 < g.n && !g.cut[w + 1]) visit(w + 1);
            if (w > 0 && gb[w] > 0) visit(w - 1);
            for (uint32_t k = g.in_ptr[w]; k < g.in_ptr[w + 1]; ++k) { uint32_t b = g.in_bid[k]; if (f[b] < g.bm[b]) visit(g.bs[b]); }
            for (uint32_t b = g.out_ptr[w]; b < g.out_ptr[w + 1]; ++b) if (f[b] > 0) visit(g.bt[b]);
        }
        cur.swap(nxt);
        ++level;
    }
    return level;
}
static St solve(const G& g, const Opt& o) {
    St st;
    const uint32_t n = g.n;
    vector<uint32_t> d(n), f(g.bs.size(), 0), stamp(n, 0);
    vector<int32_t> e(n), eadd(n, 0), snk(n), gb(n, 0);
    vector<int32_t> dem = g.dem;
    if (o.variant & 64) {  // warm start: bundles that cover a position with cov <= M are saturated up front
        uint64_t forced = 0, forced_reads = 0;
        for (size_t b = 0; b < g.bs.size(); ++b)
            if (g.unc[g.bt[b]] - g.unc[g.bs[b]] > 0) {
                f[b] = g.bm[b];
                dem[g.bs[b]] += (int32_t)g.bm[b];
                dem[g.bt[b]] -= (int32_t)g.bm[b];
                ++forced; forced_reads += g.bm[b];
            }
        printf("  warm start: %lu of %zu bundles forced (%lu reads)\n", (unsigned long)forced, g.bs.size(), (unsigned long)forced_reads);
    }

    for (uint32_t v = 0; v < n; ++v) { e[v] = dem[v] < 0 ? -dem[v] : 0; snk[v] = dem[v] > 0 ? dem[v] : 0; }
    const bool z = o.variant & 4;
    g_K = o.K;
    auto BL = [&](uint32_t v) -> uint32_t { return z ? blen(v) : 1u; };  // length of back arc v -> v-1
    uint32_t last_levels = z ? global_relabel0(g, d, snk, gb, f, st) : global_relabel(g, d, snk, gb, f, st);
    vector<uint32_t> F, T, NF;
    for (uint32_t v = 0; v < n; ++v) if (e[v] > 0) { F.push_back(v); stamp[v] = 1; }
    uint64_t rel_since = 0, rounds_since = 0;
    uint32_t round = 0;
    vector<pair<uint32_t, uint32_t>> newlab;
    int64_t sink_flow = 0;
    vector<uint32_t> f0;
    vector<int32_t> snk0;
    uint64_t walk_steps = 0;
    while (!F.empty()) {
        uint64_t interval = max<uint64_t>(o.gri, (uint64_t)last_levels * o.grl / 100);
        bool trig = rounds_since >= interval && rel_since * 100 >= (uint64_t)o.grr * n;
        if (o.variant & 1) {  // variant 1: relabel count alone may trigger (work-based, like hi_pr)
            trig = trig || rel_since >= n / 4;
        }
        if (trig) { last_levels = z ? global_relabel0(g, d, snk, gb, f, st) : global_relabel(g, d, snk, gb, f, st); rel_since = 0; rounds_since = 0; }
        ++round; ++st.rounds; ++rounds_since;
        st.maxf = max<uint64_t>(st.maxf, F.size());
        T.clear();
        if (o.variant & 6) { f0 = f; snk0 = snk; }
        auto give = [&](uint32_t w, int32_t dl) { eadd[w] += dl; if (stamp[w] != round) { stamp[w] = round; T.push_back(w); } ++st.pushes; };
        for (uint32_t v : F) {
            uint32_t dv = d[v];
            if (dv >= INF) continue;
            int32_t ex = e[v];
            if (dv == 1 && snk[v] > 0) { int32_t dl = min(ex, snk[v]); snk[v] -= dl; ex -= dl; sink_flow += dl; ++st.pushes; }
            for (uint32_t b = g.out_ptr[v + 1]; ex > 0 && b-- > g.out_ptr[v];) {
                uint32_t t = g.bt[b];
                if (d[t] + 1 != dv) continue;
                uint32_t r = g.bm[b] - f[b];
                if (!r) continue;
                int32_t dl = min<uint32_t>(ex, r); f[b] += dl; ex -= dl; give(t, dl);
            }
            if (ex > 0 && v + 1 < n && d[v + 1] + 1 == dv && gb[v + 1] > 0) { int32_t dl = min(ex, gb[v + 1]); gb[v + 1] -= dl; ex -= dl; give(v + 1, dl); }
            if (ex > 0 && v > 0 && !g.cut[v] && d[v - 1] + BL(v) == dv) {
                if ((o.variant & 6) && !(o.variant & 8)) {
                    // walk the chain of admissible back arcs until a node that could use the excess
                    // (decided from labels and the start-of-round saturation snapshot only)
                    uint32_t u = v;
                    for (;;) {
                        gb[u] += ex;
                        --u;
                        ++walk_steps;
                        bool stop = u == 0 || g.cut[u] || (d[u] == 1 && snk0[u] > 0) || d[u - 1] + BL(u) != d[u];
                        if (o.variant & 32) {  // saturation bit only: sink capacity or any unsaturated own bundle
                            stop = u == 0 || g.cut[u] || snk0[u] > 0 || d[u - 1] + BL(u) != d[u];
                            for (uint32_t b = g.out_ptr[u]; !stop && b < g.out_ptr[u + 1]; ++b)
                                if (f0[b] < g.bm[b]) stop = true;
                        }
                        for (uint32_t b = g.out_ptr[u]; !stop && b < g.out_ptr[u + 1]; ++b)
                            if (d[g.bt[b]] + 1 == d[u] && ((o.variant & 16) || f0[b] < g.bm[b])) stop = true;
                        if (stop) break;
                    }
                    give(u, ex);
                    ex = 0;
                } else {
                    gb[v] += ex; give(v - 1, ex); ex = 0;
                }
            }
            for (uint32_t k = g.in_ptr[v + 1]; ex > 0 && k-- > g.in_ptr[v];) {
                uint32_t b = g.in_bid[k]; uint32_t s = g.bs[b];
                if (d[s] + 1 != dv || f[b] == 0) continue;
                int32_t dl = min<uint32_t>(ex, f[b]); f[b] -= dl; ex -= dl; give(s, dl);
            }
            e[v] = ex;
        }
        NF.clear(); newlab.clear();
        uint64_t rel_round = 0;
        auto phase_b = [&](uint32_t w, bool in_front) {
            int32_t left = e[w];
            if (in_front && left > 0 && d[w] < INF) {
                uint32_t mn = INF;
                if (snk[w] > 0) mn = 0;
                for (uint32_t b = g.out_ptr[w]; b < g.out_ptr[w + 1]; ++b) if (f[b] < g.bm[b]) mn = min(mn, d[g.bt[b]]);
                if (w + 1 < n && gb[w + 1] > 0) mn = min(mn, d[w + 1]);
                if (w > 0 && !g.cut[w]) mn = min(mn, d[w - 1] + BL(w) - 1);
                for (uint32_t k = g.in_ptr[w]; k < g.in_ptr[w + 1]; ++k) { uint32_t b = g.in_bid[k]; if (f[b] > 0) mn = min(mn, d[g.bs[b]]); }
                newlab.emplace_back(w, mn >= INF ? INF : mn + 1);
                ++st.relabels; ++rel_since; ++rel_round;
            }
            int32_t tot = left + eadd[w]; eadd[w] = 0; e[w] = tot;
            if (tot > 0) { NF.push_back(w); stamp[w] = round + 1; }
        };
        for (uint32_t v : F) phase_b(v, true);
        for (uint32_t w : T) phase_b(w, false);
        for (auto& pr : newlab) d[pr.first] = pr.second;
        if (o.trace && (round % o.trace == 0 || round < 5)) {
            int64_t tot = 0; uint32_t mnp = n, mxp = 0;
            for (uint32_t w : NF) { tot += e[w]; mnp = min(mnp, w); mxp = max(mxp, w); }
            printf("  r %5u |F| %5zu excess %6ld sink %6ld pos [%6u,%6u] relabels %4lu grs %lu\n", round, NF.size(), (long)tot, (long)sink_flow, mnp, mxp, (unsigned long)rel_round, (unsigned long)st.grs);
        }
        F.clear();
        for (uint32_t w : NF) if (d[w] < INF) F.push_back(w);
    }
    if (o.variant & 6) printf("  walk steps %lu\n", (unsigned long)walk_steps);
    return st;
}
// uniform single-length instance like the reference's reads_gen
static const char* g_file = nullptr;
static G make(uint32_t L, uint32_t R, uint32_t cov, uint32_t M, uint32_t seed, int shape) {
    uint64_t N = (uint64_t)L * cov / R;
    mt19937_64 rng(seed);
    vector<uint32_t> cnt(L + 1, 0);
    if (g_file) {  // u32 start positions
        FILE* fp = fopen(g_file, "rb");
        uint32_t s;
        while (fread(&s, 4, 1, fp) == 1) ++cnt[s];
        fclose(fp);
        N = 0;
    }
    for (uint64_t i = 0; i < N; ++i) {
        uint32_t s = rng() % (L - R + 1);
        if (shape == 1) {  // hole in the middle third: thin coverage there
            if (s > L / 3 && s < 2 * L / 3 && (rng() % 10)) { --i; continue; }
        }
        ++cnt[s];
    }
    G g; g.n = L + 1;
    vector<int64_t> diff(L + 2, 0);
    for (uint32_t s = 0; s + R <= L; ++s) if (cnt[s]) { g.bs.push_back(s); g.bt.push_back(s + R); g.bm.push_back(cnt[s]); diff[s] += cnt[s]; diff[s + R] -= cnt[s]; }
    g.dem.assign(g.n, 0);
    int64_t run = 0; uint32_t prev = 0;
    g.unc.assign(g.n + 1, 0);
    g.cut.assign(g.n + 1, 0);
    for (uint32_t v = 0; v < g.n; ++v) { run += diff[v]; uint32_t c = min<uint64_t>(run, M); g.dem[v] = (int32_t)prev - (int32_t)c; prev = c; g.unc[v + 1] = g.unc[v] + ((uint64_t)run <= M ? 1u : 0u); }
    uint32_t B = g.bs.size();
    g.out_ptr.assign(g.n + 1, 0); g.in_ptr.assign(g.n + 1, 0);
    for (uint32_t b = 0; b < B; ++b) { ++g.out_ptr[g.bs[b] + 1]; ++g.in_ptr[g.bt[b] + 1]; }
    for (uint32_t v = 0; v < g.n; ++v) { g.out_ptr[v + 1] += g.out_ptr[v]; g.in_ptr[v + 1] += g.in_ptr[v]; }
    g.in_bid.resize(B);
    vector<uint32_t> cur(g.in_ptr.begin(), g.in_ptr.end() - 1);
    for (uint32_t b = 0; b < B; ++b) g.in_bid[cur[g.bt[b]]++] = b;
    return g;
}
// Forced reads out, cuts in: a read that covers a position with cov <= M is in every valid answer, so
// its flow is fixed (its ends' demands absorb it) and it leaves the graph; the back arc over such a
// position carries cov_S - min(cov, M) = 0 in every valid answer, so it leaves the graph too.
static G transform(const G& g) {
    G h; h.n = g.n; h.dem = g.dem; h.unc = g.unc; h.cut.assign(g.n + 1, 0);
    uint64_t forced = 0, reads = 0;
    for (size_t b = 0; b < g.bs.size(); ++b) {
        if (g.unc[g.bt[b]] - g.unc[g.bs[b]] > 0) {
            h.dem[g.bs[b]] += (int32_t)g.bm[b];
            h.dem[g.bt[b]] -= (int32_t)g.bm[b];
            ++forced; reads += g.bm[b];
        } else { h.bs.push_back(g.bs[b]); h.bt.push_back(g.bt[b]); h.bm.push_back(g.bm[b]); }
    }
    uint32_t ncut = 0;
    for (uint32_t v = 1; v < g.n; ++v) if (g.unc[v] - g.unc[v - 1]) { h.cut[v] = 1; ++ncut; }
    uint32_t B = h.bs.size();
    h.out_ptr.assign(h.n + 1, 0); h.in_ptr.assign(h.n + 1, 0);
    for (uint32_t b = 0; b < B; ++b) { ++h.out_ptr[h.bs[b] + 1]; ++h.in_ptr[h.bt[b] + 1]; }
    for (uint32_t v = 0; v < h.n; ++v) { h.out_ptr[v + 1] += h.out_ptr[v]; h.in_ptr[v + 1] += h.in_ptr[v]; }
    h.in_bid.resize(B);
    vector<uint32_t> cur(h.in_ptr.begin(), h.in_ptr.end() - 1);
    for (uint32_t b = 0; b < B; ++b) h.in_bid[cur[h.bt[b]]++] = b;
    int64_t sup = 0; uint32_t nsup = 0, nsnk = 0;
    for (uint32_t v = 0; v < h.n; ++v) { if (h.dem[v] < 0) { sup -= h.dem[v]; ++nsup; } if (h.dem[v] > 0) ++nsnk; }
    printf("  transform: %lu bundles (%lu reads) forced, %u back arcs cut, residual supply %ld at %u nodes, %u sink nodes\n",
           (unsigned long)forced, (unsigned long)reads, ncut, (long)sup, nsup, nsnk);
    return h;
}
int main(int argc, char** argv) {
    uint32_t L = 16000, R = 150, cov = 1500, M = 500, seed = 1; int shape = 0;
    Opt o;
    for (int i = 1; i < argc; ++i) {
        if (!strncmp(argv[i], "file=", 5)) { g_file = argv[i] + 5; continue; }
        auto kv = [&](const char* k, auto& v) { size_t l = strlen(k); if (!strncmp(argv[i], k, l) && argv[i][l] == '=') { v = atoi(argv[i] + l + 1); return true; } return false; };
        kv("L", L) || kv("R", R) || kv("cov", cov) || kv("M", M) || kv("seed", seed) || kv("shape", shape) || kv("gri", o.gri) || kv("grl", o.grl) || kv("grr", o.grr) || kv("variant", o.variant) || kv("K", o.K) || kv("trace", o.trace);
    }
    G g = make(L, R, cov, M, seed, shape);
    if (o.variant & 128) g = transform(g);
    St st = solve(g, o);
    printf("L %u cov %u M %u variant %d: rounds %lu pushes %lu relabels %lu grs %lu levels %lu maxF %lu  (hops %u)\n", L, cov, M, o.variant,
           (unsigned long)st.rounds, (unsigned long)st.pushes, (unsigned long)st.relabels, (unsigned long)st.grs, (unsigned long)st.levels, (unsigned long)st.maxf, L / R);
}
