// In-memory fake of the few htslib calls the reference's bam-api makes.  TEST INFRASTRUCTURE.
#include <htslib/sam.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
struct FakeBam {
    std::string name;
    uint64_t n = 0;
    uint32_t ref_len = 0;
    std::vector<uint32_t> start, end, qual, len;
} g_bam;
std::vector<uint64_t> g_written;
}  // namespace

struct htsFile {
    bool writing;
    uint64_t cursor;
};

extern "C" {
void gds_fake_bam_set(const char* name, uint64_t n, uint32_t ref_len, const uint32_t* start,
                      const uint32_t* end, const uint32_t* quality, const uint32_t* seq_len) {
    g_bam.name = name;
    g_bam.n = n;
    g_bam.ref_len = ref_len;
    g_bam.start.assign(start, start + n);
    g_bam.end.assign(end, end + n);
    g_bam.qual.assign(quality, quality + n);
    g_bam.len.assign(seq_len, seq_len + n);
}
uint64_t gds_fake_bam_written(uint64_t* ids, uint64_t cap) {
    uint64_t c = g_written.size() < cap ? g_written.size() : cap;
    for (uint64_t i = 0; i < c; ++i) ids[i] = g_written[i];
    return g_written.size();
}
bam1_t* bam_init1(void) {
    bam1_t* b = (bam1_t*)calloc(1, sizeof(bam1_t));
    b->data = (uint8_t*)calloc(1, 64);
    return b;
}
void bam_destroy1(bam1_t* b) {
    if (!b) return;
    free(b->data);
    free(b);
}
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t* cigar) {
    hts_pos_t l = 0;
    for (int i = 0; i < n_cigar; ++i) {
        uint32_t op = cigar[i] & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += cigar[i] >> 4;
    }
    return l;
}
samFile* sam_open(const char* fn, const char* mode) {
    htsFile* f = new htsFile();
    f->writing = mode && mode[0] == 'w';
    f->cursor = 0;
    if (f->writing) g_written.clear();
    else if (g_bam.name != fn) {
        delete f;
        return nullptr;
    }
    return f;
}
int sam_close(samFile* fp) {
    delete fp;
    return 0;
}
sam_hdr_t* sam_hdr_read(samFile*) {
    sam_hdr_t* h = (sam_hdr_t*)calloc(1, sizeof(sam_hdr_t));
    h->n_targets = 1;
    h->target_len = (uint32_t*)calloc(1, sizeof(uint32_t));
    h->target_len[0] = g_bam.ref_len;
    return h;
}
int sam_hdr_write(samFile*, const sam_hdr_t*) { return 0; }
void sam_hdr_destroy(sam_hdr_t* h) {
    if (!h) return;
    free(h->target_len);
    free(h);
}
int sam_read1(samFile* fp, sam_hdr_t*, bam1_t* b) {
    if (fp->cursor >= g_bam.n) return -1;
    uint64_t i = fp->cursor++;
    b->core.pos = g_bam.start[i];
    b->core.qual = (uint8_t)g_bam.qual[i];
    b->core.l_qseq = (int32_t)g_bam.len[i];
    b->core.flag = (i % 2 == 0) ? BAM_FREAD1 : BAM_FREAD2;  // mates adjacent, first mate first
    b->core.n_cigar = 1;
    int n = snprintf((char*)b->data, 40, "pair%llu", (unsigned long long)(i / 2));
    uint16_t lq = (uint16_t)(((n + 1) + 3) & ~3);  // keep the cigar 4-byte aligned
    b->core.l_qname = lq;
    uint32_t rlen = g_bam.end[i] - g_bam.start[i] + 1;
    uint32_t cig = rlen << 4;  // <rlen>M
    memcpy(b->data + lq, &cig, 4);
    return 0;
}
int sam_write1(samFile* fp, const sam_hdr_t*, const bam1_t* b) {
    (void)b;
    (void)fp;
    return 0;
}
hts_idx_t* sam_index_load(samFile*, const char*) { return nullptr; }
int hts_idx_get_stat(const hts_idx_t*, int, uint64_t* m, uint64_t* u) {
    *m = *u = 0;
    return 0;
}
void hts_idx_destroy(hts_idx_t*) {}
hts_tpool* hts_tpool_init(int) { return nullptr; }
void hts_tpool_destroy(hts_tpool*) {}
int hts_set_thread_pool(htsFile*, htsThreadPool*) { return 0; }
}
