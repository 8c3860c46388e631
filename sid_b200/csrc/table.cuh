// Device-resident lookup structures shared by the kernels:
//   * the unique-profile table: open addressing on the packed 64-bit profile; replaces the
//     sort + std::map of countUniqueProfiles (pileup.cpp:169-196) and call.cpp:217-221 / 276-285.
//   * the chromosome-name dictionary: interns the first column so a site carries a 4-byte name
//     reference instead of a std::string (pileup.hpp:10, call.hpp:23-27).
#pragma once
#include "common.cuh"

namespace sid {

constexpr uint64_t TABLE_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int SUFFIX_BYTES = 48;   // ",het,AC,1.23457e-308,1.23457e-308,probability\n" is 46 bytes

struct TableView {
    unsigned long long* keys;     // cap + 1 entries; entry [cap] serves the all-ones profile
    unsigned long long* counts;   // cap + 1
    uint32_t* entry_list;         // slots in insertion order
    unsigned int* n_entries;
    unsigned int* special_used;   // 0/1: the all-ones profile has been seen
    unsigned int* overflow;       // set when probing wrapped the whole table
    uint32_t cap, mask;
    // per-slot classification written by the calling kernels
    char* suffix;                 // SUFFIX_BYTES per slot, last byte = length
    uint8_t* label;               // 0 hom, 1 het, 255 dropped
    char* gt;                     // 2 per slot
    double* hom;
    double* het;
};

// Hash of a packed profile: two 32-bit multiplies and a finaliser (the FMA pipe; the 64-bit mix64 costs three times
// as many instructions on the tokenizer's critical pipe).
SID_HD uint32_t table_hash(uint64_t key) {
    uint32_t h = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA77u);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

constexpr uint32_t SUFFIX_READY = 0x80u;   // bit 7 of a suffix record's last byte: the record is complete (low bits: length)

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t table_find_or_insert(const TableView& t, uint64_t key) {
    if (key == TABLE_EMPTY) {
        if (atomicCAS(t.special_used, 0u, 1u) == 0u) {
            t.keys[t.cap] = key;
            t.entry_list[atomicAdd(t.n_entries, 1u)] = t.cap;
        }
        return t.cap;
    }
    uint32_t h = table_hash(key) & t.mask;
    for (uint32_t probes = 0; probes <= t.mask; ++probes) {
        unsigned long long k = *((volatile unsigned long long*)&t.keys[h]);
        if (k == key) return h;
        if (k == TABLE_EMPTY) {
            const unsigned long long old = atomicCAS(&t.keys[h], (unsigned long long)TABLE_EMPTY, (unsigned long long)key);
            if (old == TABLE_EMPTY) {
                t.entry_list[atomicAdd(t.n_entries, 1u)] = h;
                return h;
            }
            if (old == key) return h;
        }
        h = (h + 1) & t.mask;
    }
    atomicExch(t.overflow, 1u);
    return 0;
}

__device__ __forceinline__ uint32_t table_find(const TableView& t, uint64_t key) {
    if (key == TABLE_EMPTY) return t.cap;
    uint32_t h = table_hash(key) & t.mask;
    for (uint32_t probes = 0; probes <= t.mask; ++probes) {
        const unsigned long long k = t.keys[h];
        if (k == key) return h;
        if (k == TABLE_EMPTY) break;
        h = (h + 1) & t.mask;
    }
    return 0xFFFFFFFFu;
}

#endif  // __CUDACC__

// ---- chromosome-name dictionary ----------------------------------------------------------------
// slot word = pool_offset << 32 | length << 16 | tag16; 0 = empty (pool offsets start at 4).
// pool record at pool_offset: 2-byte little-endian length, then the bytes.
struct NameDict {
    unsigned long long* slots;
    char* pool;
    unsigned int* cursor;     // next free pool byte
    unsigned int* overflow;
    uint32_t mask;
    uint32_t pool_cap;
};

#if defined(__CUDACC__)

template <class Src>
__device__ __forceinline__ uint64_t name_hash(const Src& src, uint64_t off, uint32_t len) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint32_t i = 0; i < len; ++i) { h ^= src.at(off + i); h *= 0x100000001B3ull; }
    return mix64(h ^ len);
}

template <class Src>
__device__ __forceinline__ bool name_equals(const NameDict& d, uint32_t pool_off, const Src& src, uint64_t off, uint32_t len) {
    // __ldcg: read at L2, the name may have been appended by another SM after this SM cached the line
    const uint8_t* p = (const uint8_t*)d.pool + pool_off;
    const uint32_t plen = (uint32_t)__ldcg(p) | ((uint32_t)__ldcg(p + 1) << 8);
    if (plen != len) return false;
    for (uint32_t i = 0; i < len; ++i) if (__ldcg(p + 2 + i) != src.at(off + i)) return false;
    return true;
}

// Returns the pool offset of the interned name (0 on overflow).
template <class Src>
__device__ __forceinline__ uint32_t name_intern(const NameDict& d, const Src& src, uint64_t off, uint32_t len) {
    if (len > 0xFFFFu) { atomicExch(d.overflow, 2u); return 0; }
    if (*((volatile unsigned int*)d.overflow)) return 0;      // full: the host reports it; do not advance the cursor further
    const uint64_t hv = name_hash(src, off, len);
    const uint32_t tag = (uint32_t)(hv >> 48) & 0xFFFFu;
    uint32_t h = (uint32_t)hv & d.mask;
    uint32_t mine = 0;   // pool offset of our speculative copy
    for (uint32_t probes = 0; probes <= d.mask; ++probes) {
        unsigned long long w = *((volatile unsigned long long*)&d.slots[h]);
        if (w == 0ull) {
            if (mine == 0) {
                const uint32_t need = (2u + len + 3u) & ~3u;
                mine = atomicAdd(d.cursor, need);
                if ((uint64_t)mine + need > d.pool_cap) { atomicExch(d.overflow, 1u); return 0; }
                uint8_t* p = (uint8_t*)d.pool + mine;
                p[0] = (uint8_t)(len & 0xFF);
                p[1] = (uint8_t)(len >> 8);
                for (uint32_t i = 0; i < len; ++i) p[2 + i] = src.at(off + i);
                __threadfence();
            }
            const unsigned long long word = ((unsigned long long)mine << 32) | ((unsigned long long)len << 16) | tag;
            const unsigned long long old = atomicCAS(&d.slots[h], 0ull, word);
            if (old == 0ull) return mine;
            w = old;
        }
        if ((uint32_t)(w & 0xFFFFu) == tag && (uint32_t)((w >> 16) & 0xFFFFu) == len) {
            const uint32_t po = (uint32_t)(w >> 32);
            __threadfence();
            if (name_equals(d, po, src, off, len)) return po;
        }
        h = (h + 1) & d.mask;
    }
    atomicExch(d.overflow, 1u);
    return 0;
}

#endif  // __CUDACC__

}  // namespace sid
