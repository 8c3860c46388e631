// SWAR form of the tokenizer (same semantics as parse.cuh; checked against it line by line).
// PLACEHOLDER until the word-at-a-time implementation lands: reports "not handled" so callers
// take the scalar path.
#pragma once
#include "common.cuh"
#include "parse.cuh"

namespace sid {

struct FastLine {
    int status;
    int32_t pos;
    uint64_t profile;
    uint32_t chrom_off, chrom_len;
};

SID_HD bool parse_line_fast_smem(const uint8_t*, uint64_t, uint32_t, uint64_t, FastLine&) { return false; }

}  // namespace sid
