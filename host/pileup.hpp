// Host mirror of the reference's pileup interface (pileup.hpp:7-44): same type and function names,
// same argument meaning.  Every function runs on the GPU through libsidgpu (include/sidgpu.h);
// nothing is parsed on the CPU.
#pragma once
#include <array>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

using profile_t = std::array<uint16_t, 4>;          // pileup.hpp:7

typedef struct {                                      // pileup.hpp:9-18
    std::string chromosome_name;
    int position {-1};
    char reference_base {'N'};
    profile_t base_counts;
    // Per-read vectors (pileup.hpp:14-17).  The calling path never needs them (`-m quality` consumes the text on
    // the device, k_quality.cuh); parsePileupLine / readFile fill bases and strands always, like the reference, and
    // the two quality vectors when asked (sidgpu_read_counts / sidgpu_read_fill).
    std::vector<char> bases;
    std::vector<bool> strands;
    std::vector<uint8_t> base_qualities;
    std::vector<uint8_t> mapping_qualities;
} PileupLine;

// pileup.cpp:13-68.  Unlike the reference the buffer is not modified.  Throws
// std::invalid_argument("Malformed pileup line" / "... or missing mapping qualities").
PileupLine parsePileupLine(char* line, bool parse_base_qualities, bool parse_mapping_qualities);

typedef struct {                                      // pileup.hpp:22-26
    std::vector<char> bases;
    std::vector<bool> strands;
    profile_t counts;
} ReadStack;

// pileup.cpp:70-153: counts, letters and strands of a bases string against a reference base.
ReadStack parseReadBases(const char* read_bases, char reference, int coverage);

// pileup.cpp:155-167: character - 33, at least 1, up to the first NUL, tab or line end.
std::vector<uint8_t> parseQualities(const char* base_qualities, int coverage);

typedef struct UniqueProfile {                        // pileup.hpp:32-40
    profile_t profile;
    uint32_t count;
    uint32_t coverage;
    UniqueProfile() : profile {0, 0, 0, 0}, count {0}, coverage {0} {}
    UniqueProfile(profile_t p, uint32_t count) : profile {p}, count {count} {
        coverage = std::accumulate(profile.begin(), profile.end(), 0);
    }
} UniqueProfile;

// pileup.cpp:169-196 and :198-217
std::vector<UniqueProfile> countUniqueProfiles(const std::vector<PileupLine>&);
std::array<double, 4> computeNucleotideDistribution(const std::vector<UniqueProfile>&);

// Whole-text form of readFile (call.cpp:11-20): every line of a pileup text in one GPU pass.
std::vector<PileupLine> parsePileupText(const char* text, size_t len, bool parse_base_qualities, bool parse_mapping_qualities);
