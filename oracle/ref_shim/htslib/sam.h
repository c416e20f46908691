/* Stand-in for <htslib/sam.h> — TEST INFRASTRUCTURE (oracle/_ref only).
 * htslib is absent from this image.  This header declares just enough of its API for the
 * reference's libs/bam-api sources to compile UNMODIFIED from /root/reference; hts_stub.cpp backs
 * it with an in-memory "BAM" so that BamApi::read_bam (pair matching + filter) really executes on
 * synthetic records.  Nothing here is derived from htslib sources: only the public names/fields
 * the reference touches (bam_api.cpp:359-656, read.cpp:5-14) are provided. */
#ifndef GDS_FAKE_HTSLIB_SAM_H
#define GDS_FAKE_HTSLIB_SAM_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int64_t hts_pos_t;
typedef struct {
    hts_pos_t pos;
    uint8_t qual;
    uint16_t l_qname;
    uint16_t flag;
    uint32_t n_cigar;
    int32_t l_qseq;
} bam1_core_t;
typedef struct {
    bam1_core_t core;
    uint8_t* data; /* qname (NUL terminated, l_qname bytes) then cigar */
} bam1_t;
typedef struct {
    int32_t n_targets;
    uint32_t* target_len;
} sam_hdr_t;
typedef struct htsFile htsFile;
typedef htsFile samFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_tpool hts_tpool;
typedef struct {
    hts_tpool* pool;
    int qsize;
} htsThreadPool;
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
bam1_t* bam_init1(void);
void bam_destroy1(bam1_t* b);
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t* cigar);
samFile* sam_open(const char* fn, const char* mode);
int sam_close(samFile* fp);
sam_hdr_t* sam_hdr_read(samFile* fp);
int sam_hdr_write(samFile* fp, const sam_hdr_t* h);
void sam_hdr_destroy(sam_hdr_t* h);
int sam_read1(samFile* fp, sam_hdr_t* h, bam1_t* b);
int sam_write1(samFile* fp, const sam_hdr_t* h, const bam1_t* b);
hts_idx_t* sam_index_load(samFile* fp, const char* fn);
int hts_idx_get_stat(const hts_idx_t* idx, int tid, uint64_t* mapped, uint64_t* unmapped);
void hts_idx_destroy(hts_idx_t* idx);
hts_tpool* hts_tpool_init(int n);
void hts_tpool_destroy(hts_tpool* p);
int hts_set_thread_pool(htsFile* fp, htsThreadPool* p);
#define hts_log_info(...) ((void)0)

/* ---- fake-BAM registry (not htslib): sam_open(name) serves these records ---- */
void gds_fake_bam_set(const char* name, uint64_t n, uint32_t ref_len, const uint32_t* start,
                      const uint32_t* end, const uint32_t* quality, const uint32_t* seq_len);
/* records passed to sam_write1 on the last output file, by 0-based record ordinal */
uint64_t gds_fake_bam_written(uint64_t* ids, uint64_t cap);
#ifdef __cplusplus
}
#endif
#endif
