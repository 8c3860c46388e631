// Pileup line tokenizer + profile builder (scalar form).
//   parsePileupLine  pileup.cpp:13-68   (strtok_r on " \t", atoi, 1-char reference)
//   parseReadBases   pileup.cpp:70-153  (bases state machine)
//   parseQualities   pileup.cpp:155-167 (only offsets/lengths are located here)
// A "line" starts at byte p and ends at the first '\n', the first NUL (the reference hands the
// std::getline buffer to C string functions) or the end of the text.
#pragma once
#include "common.cuh"

namespace sid {

struct ParsedLine {
    int status;          // LineStatus
    int32_t pos;         // atoi(column 2)
    uint64_t profile;    // packed counts
    uint32_t n_bases;    // counted bases (A,C,G,T only)
    uint32_t chrom_off;  // offsets are relative to the line start
    uint32_t chrom_len;
    uint32_t bases_off, bases_len;
    uint32_t bq_off, bq_len;
    uint32_t mq_off, mq_len;
    char ref;
};

// Byte source over a flat buffer with an exclusive limit; reads at or past the limit look like '\n'.
struct FlatSrc {
    const uint8_t* base;
    uint64_t limit;
    SID_HD uint8_t at(uint64_t off) const { return off < limit ? base[off] : (uint8_t)'\n'; }
};

SID_HD bool is_delim(uint8_t c) { return c == ' ' || c == '\t'; }   // pileup.cpp:11
SID_HD bool is_eol(uint8_t c) { return c == '\n' || c == 0; }

// glibc atoi == (int)strtol(s, NULL, 10): leading isspace() skipped, optional sign, digits,
// LONG_MAX / LONG_MIN on overflow, then truncation to 32 bits (pileup.cpp:24).
struct AtoiState {
    uint64_t acc;
    int phase;   // 0 leading space, 1 after sign / in digits, 2 done
    bool neg, ovf;
    SID_HD void init() { acc = 0; phase = 0; neg = false; ovf = false; }
    SID_HD void feed(uint8_t c) {
        if (phase == 2) return;
        if (phase == 0) {
            if (c == '\v' || c == '\f' || c == '\r') return;  // ' ', '\t', '\n' cannot occur inside a token
            phase = 1;
            if (c == '-') { neg = true; return; }
            if (c == '+') return;
        }
        uint32_t d = (uint32_t)c - (uint32_t)'0';
        if (d > 9) { phase = 2; return; }
        if (acc > (0xFFFFFFFFFFFFFFFFull - d) / 10) ovf = true; else acc = acc * 10 + d;
    }
    SID_HD int32_t value() const {
        if (neg) {
            if (ovf || acc > 0x8000000000000000ull) return 0;           // (int)LONG_MIN
            return (int32_t)(uint32_t)(0ull - acc);
        }
        if (ovf || acc > 0x7FFFFFFFFFFFFFFFull) return -1;              // (int)LONG_MAX
        return (int32_t)(uint32_t)acc;
    }
};

// toupper / tolower of the "C" locale (pileup.cpp:79,82 run under it: the reference never calls setlocale).
SID_HD uint8_t ascii_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
SID_HD uint8_t ascii_lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

// Bases state machine, one byte at a time (pileup.cpp:76-150).  '.' and ',' are replaced by
// toupper / tolower of the reference character BEFORE the switch (pileup.cpp:78-83), so a reference
// column of '^', '+' or '-' turns every '.' into that control character; the digit look-ahead of an
// indel (pileup.cpp:131,136) reads the raw text.
struct BasesState {
    uint32_t cnt[4];
    uint64_t skip;           // bytes still to be skipped ('^' -> 1, indel -> N)
    uint64_t num;            // indel length being read
    int mode;                // 0 normal, 1 just saw '+'/'-', 2 reading the indel length
    uint8_t dot_as, comma_as;
    SID_HD void init(uint8_t ref) {
        cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; skip = 0; num = 0; mode = 0;
        dot_as = ascii_upper(ref); comma_as = ascii_lower(ref);
    }
    // returns the index 0..3 of the counted base, -1 when nothing is counted
    SID_HD int feed(uint8_t c) {
        if (mode == 1) {                        // pileup.cpp:131-133: sign not followed by a digit is ignored
            uint32_t d = (uint32_t)c - (uint32_t)'0';
            if (d <= 9) { mode = 2; num = d; return -1; }
            mode = 0;
        } else if (mode == 2) {                 // strtol over the digits (pileup.cpp:136), saturating
            uint32_t d = (uint32_t)c - (uint32_t)'0';
            if (d <= 9) { if (num < (1ull << 40)) num = num * 10 + d; return -1; }
            mode = 0;
            skip = num;                         // pileup.cpp:144: skip that many bytes after the number
        }
        if (skip) { --skip; return -1; }
        if (c == '.') c = dot_as; else if (c == ',') c = comma_as;    // pileup.cpp:78-83
        switch (c) {
            case 'A': case 'a': ++cnt[0]; return 0;
            case 'C': case 'c': ++cnt[1]; return 1;
            case 'G': case 'g': ++cnt[2]; return 2;
            case 'T': case 't': ++cnt[3]; return 3;
            case '^': skip = 1; return -1;      // pileup.cpp:125-127
            case '+': case '-': mode = 1; return -1;
            default: return -1;
        }
    }
};

// Reference characters for which '.' / ',' become control characters of the bases grammar: the
// bit-parallel tokenizer leaves such lines to the byte-wise state machine.
SID_HD bool ref_is_control(uint8_t ref) { return ref == '^' || ref == '+' || ref == '-'; }

// Index of the reference base for '.' (toupper) and ',' (tolower) substitution, or -1 when the
// substituted character is not one of ACGTacgt (pileup.cpp:78-83 then the switch default).
SID_HD int ref_index(uint8_t ref) {
    switch (ref) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

template <class Src>
SID_HD void parse_line(const Src& src, uint64_t p, bool want_qual, ParsedLine& o) {
    o.status = LINE_MALFORMED;
    o.pos = -1;
    o.profile = 0;
    o.n_bases = 0;
    o.ref = 'N';
    o.chrom_off = o.chrom_len = 0;
    o.bases_off = o.bases_len = o.bq_off = o.bq_len = o.mq_off = o.mq_len = 0;
    uint64_t q = p;
    uint8_t c = src.at(q);
    // token 0: chromosome name (pileup.cpp:17-18)
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    o.chrom_off = (uint32_t)(q - p);
    while (!is_delim(c) && !is_eol(c)) c = src.at(++q);
    o.chrom_len = (uint32_t)(q - p) - o.chrom_off;
    // token 1: position (pileup.cpp:20-24)
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    AtoiState a;
    a.init();
    while (!is_delim(c) && !is_eol(c)) { a.feed(c); c = src.at(++q); }
    o.pos = a.value();
    // token 2: reference base, exactly one character (pileup.cpp:26-30)
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    o.ref = (char)c;
    c = src.at(++q);
    if (!is_delim(c) && !is_eol(c)) return;
    // token 3: coverage (pileup.cpp:32-36; value unused beyond a reserve() hint)
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    while (!is_delim(c) && !is_eol(c)) c = src.at(++q);
    // token 4: read bases (pileup.cpp:38-45)
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    o.bases_off = (uint32_t)(q - p);
    BasesState b;
    b.init((uint8_t)o.ref);
    while (!is_delim(c) && !is_eol(c)) { b.feed(c); c = src.at(++q); }
    o.bases_len = (uint32_t)(q - p) - o.bases_off;
    o.profile = pack_profile(b.cnt[0], b.cnt[1], b.cnt[2], b.cnt[3]);
    o.n_bases = b.cnt[0] + b.cnt[1] + b.cnt[2] + b.cnt[3];
    if (!want_qual) { o.status = LINE_OK; return; }
    // token 5: base qualities (pileup.cpp:49-57; the reference dereferences NULL when it is missing)
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    o.bq_off = (uint32_t)(q - p);
    while (!is_delim(c) && !is_eol(c)) c = src.at(++q);
    o.bq_len = (uint32_t)(q - p) - o.bq_off;
    // token 6: mapping qualities (pileup.cpp:60-66)
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) { o.status = LINE_MISSING_MAPQ; return; }
    o.mq_off = (uint32_t)(q - p);
    while (!is_delim(c) && !is_eol(c)) c = src.at(++q);
    o.mq_len = (uint32_t)(q - p) - o.mq_off;
    o.status = (o.n_bases > o.bq_len || o.n_bases > o.mq_len) ? LINE_QUAL_SHORT : LINE_OK;
}

// Flat text with word access for the field scan below: base is 8-byte aligned, reads at or past
// `limit` look like '\n'.
struct WordSrc {
    const uint8_t* base;
    uint64_t limit;
    SID_HD uint8_t at(uint64_t off) const { return off < limit ? base[off] : (uint8_t)'\n'; }
    SID_HD uint64_t word(uint64_t k) const {                    // bytes 8k .. 8k+7, little endian
        if (8 * k + 8 <= limit) return *reinterpret_cast<const uint64_t*>(base + 8 * k);
        uint64_t w = 0;
        for (int i = 0; i < 8; ++i) w |= (uint64_t)at(8 * k + i) << (8 * i);
        return w;
    }
    // smallest q >= from whose byte is <= 0x20 (every delimiter and line end is); the text ends with one
    SID_HD uint64_t next_low(uint64_t from) const {
        uint64_t k = from >> 3;
        uint64_t keep = ~0ull << (8 * (from & 7));              // bytes of the first word at or after `from`
        for (;;) {
            const uint64_t w = word(k);
            // bit 7 of a byte of `ge` is set iff the byte is >= 0x21
            const uint64_t ge = (((w & 0x7F7F7F7F7F7F7F7Full) + 0x5F5F5F5F5F5F5F5Full) | w) & 0x8080808080808080ull;
            const uint64_t low = (ge ^ 0x8080808080808080ull) & keep;
            if (low) return 8 * k + (uint64_t)(ctz64(low) >> 3);
            keep = ~0ull;
            ++k;
        }
    }
    // end of the token that contains `q`: the first delimiter or line end at or after q
    SID_HD uint64_t token_end(uint64_t q) const {
        for (;;) {
            q = next_low(q);
            const uint8_t c = at(q);
            if (is_delim(c) || is_eol(c)) return q;
            ++q;                                                // a control byte inside the token
        }
    }
};

// The fields of a line whose profile is already known (the tokenizer stored it): the same offsets,
// lengths and status as parse_line(..., want_qual = true), without counting the bases again and with
// the token ends found eight bytes at a time.  Only for lines whose bases field is shorter than 65536
// bytes (no 16-bit count can have wrapped, so the number of counted bases is the sum of the profile);
// longer ones go through parse_line.
SID_HD void quality_fields(const WordSrc& src, uint64_t p, uint64_t profile, ParsedLine& o) {
    o.status = LINE_MALFORMED;
    o.pos = -1;
    o.profile = profile;
    o.n_bases = 0;
    o.ref = 'N';
    o.chrom_off = o.chrom_len = 0;
    o.bases_off = o.bases_len = o.bq_off = o.bq_len = o.mq_off = o.mq_len = 0;
    uint64_t q = p;
    uint8_t c = src.at(q);
    // tokens 0 and 1: chromosome name, position
    for (int t = 0; t < 2; ++t) {
        while (is_delim(c)) c = src.at(++q);
        if (is_eol(c)) return;
        if (t == 0) o.chrom_off = (uint32_t)(q - p);
        q = src.token_end(q);
        c = src.at(q);
        if (t == 0) o.chrom_len = (uint32_t)(q - p) - o.chrom_off;
    }
    // token 2: reference base, exactly one character
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    o.ref = (char)c;
    c = src.at(++q);
    if (!is_delim(c) && !is_eol(c)) return;
    // token 3: coverage
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    q = src.token_end(q);
    c = src.at(q);
    // token 4: read bases
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    o.bases_off = (uint32_t)(q - p);
    q = src.token_end(q);
    c = src.at(q);
    o.bases_len = (uint32_t)(q - p) - o.bases_off;
    if (o.bases_len >= 65536u) { parse_line(src, p, true, o); return; }
    o.n_bases = profile_count(profile, 0) + profile_count(profile, 1) + profile_count(profile, 2) + profile_count(profile, 3);
    // token 5: base qualities
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) return;
    o.bq_off = (uint32_t)(q - p);
    q = src.token_end(q);
    c = src.at(q);
    o.bq_len = (uint32_t)(q - p) - o.bq_off;
    // token 6: mapping qualities
    while (is_delim(c)) c = src.at(++q);
    if (is_eol(c)) { o.status = LINE_MISSING_MAPQ; return; }
    o.mq_off = (uint32_t)(q - p);
    q = src.token_end(q);
    o.mq_len = (uint32_t)(q - p) - o.mq_off;
    o.status = (o.n_bases > o.bq_len || o.n_bases > o.mq_len) ? LINE_QUAL_SHORT : LINE_OK;
}

}  // namespace sid
