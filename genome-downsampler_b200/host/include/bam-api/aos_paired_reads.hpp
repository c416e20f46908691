// same header name as the reference; the containers live in paired_reads.hpp
#pragma once
#include "bam-api/paired_reads.hpp"
