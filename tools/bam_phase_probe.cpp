// Phase timings of BamApi::read_bam (DEBUG log): g++ -O2 -std=c++17 -pthread -Igenome-downsampler_b200/host/include -Iinclude tools/bam_phase_probe.cpp genome-downsampler_b200/host/src/bam_api.cpp genome-downsampler_b200/host/src/bgzf_bam.cpp -lz -o /tmp/bam_phase_probe
//   /tmp/bam_phase_probe in.bam THREADS
#include "bam-api/bam_api.hpp"
#include "logging/log.hpp"
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
    SET_LOG_LEVEL(logging::DEBUG);
    bam_api::BamApiConfig cfg; cfg.hts_thread_count = atoi(argv[2]);
    bam_api::BamApi api(std::filesystem::path(argv[1]), cfg);
    api.has_pending_filter();
    std::printf("records %lu read_bam %.3f s\n", (unsigned long)api.bam_record_count(), api.read_bam_seconds());
}
