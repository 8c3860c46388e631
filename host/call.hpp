// Host mirror of the reference's calling interface (call.hpp:12-43): same names, argument order
// and meaning, same error behaviour; the work runs on the GPU through libsidgpu.
#pragma once
#include <array>
#include <iostream>
#include <string>
#include <vector>

#include "pileup.hpp"

std::vector<PileupLine> readFile(std::istream& in, bool parse_base_qualities, bool parse_mapping_qualities);   // call.hpp:12

typedef struct {                                      // call.hpp:14-21
    std::string label {"none"};
    std::string genotype {"NN"};
    double confidence_homozygous;
    double confidence_heterozygous;
    std::string confidence_type {"unspecified"};
    std::vector<std::string> additional_data;
} Classification;

typedef struct {                                      // call.hpp:23-27
    std::string chromosome_name;
    int position;
    Classification classification;
} OutputRecord;

inline std::ostream& operator<<(std::ostream& os, const OutputRecord& r) {   // call.hpp:29-38
    os << r.chromosome_name;
    os << ',' << r.position;
    os << ',' << r.classification.label;
    os << ',' << r.classification.genotype;
    os << ',' << r.classification.confidence_homozygous;
    os << ',' << r.classification.confidence_heterozygous;
    os << ',' << r.classification.confidence_type;
    return os;
}

// call.hpp:40-43
std::vector<OutputRecord> callLikelihoodRatio(std::istream& in, const bool use_prior, const double significance_level);
std::vector<OutputRecord> callBayes(std::istream& in);
std::vector<OutputRecord> callSiteMLError(std::istream& in, const bool estimate_prior, double prior, double error_threshold, const double significance_level);
std::vector<OutputRecord> callQualityBasedSimple(std::istream& in, const bool estimate_prior, double prior, const double significance_level);

// ---- beyond the reference: the streaming form the `sid` binary uses --------------------------------
struct SidRunInfo {
    uint64_t n_sites = 0, n_rows = 0;
    bool has_fit = false;
    double heterozygosity = 0, error_rate = 0;
    int iterations = 0;
    bool converged = false;
    uint64_t unique_profiles = 0;
};
// Runs `method` ("local", "bayes", "likelihood_ratio", "quality") over a whole pileup text and writes
// `header` (when not null, followed by std::endl as in sid.cpp:102) and then the CSV rows to `out`
// -- like the reference nothing is printed when the input is malformed.  Prints the reference's
// `# ...` progress lines to `log`.
SidRunInfo sidCallToStream(const std::string& method, const char* text, size_t len, bool estimate_prior, double prior,
                           double error_threshold, double significance_level, std::ostream& out, std::ostream& log,
                           const char* header = nullptr);
// The streaming form the `sid` binary uses for one GPU (sid.cpp:85-105 without the whole file and the whole result
// in memory): text is pulled from `fd_in` chunk by chunk (a gzip stream is inflated on the fly), rows are written to
// `fd_out` as they come back from the device.  The header line goes out with the first rows, so an input that fails
// within its first chunk prints nothing, like the reference; a later failure leaves the rows of the chunks before
// it on fd_out (the exit status and the message are the reference's).  read_threads > 1 reads a regular file with
// that many parallel preads per chunk.
SidRunInfo sidCallFile(const std::string& method, int fd_in, bool gzip, bool estimate_prior, double prior, double error_threshold,
                       double significance_level, int fd_out, std::ostream& log, const char* header = nullptr, int read_threads = 1);
// Selects the GPU (default 0) and the chunk size of the host path for subsequent calls.
void sidSetDevice(int device, size_t max_chunk_bytes = 0);
// The same over several GPUs of one node, one host thread and one position shard (a line-aligned byte range
// of the text) per device, rows written in shard order.  `local` and `quality` without -R need
// no exchange between shards.  The methods with a genome-wide fit share it: the host sums the shards' integer
// nucleotide counts and, per optimiser step, their objective values (one double each), and merges their
// unique-profile lists for the BH ranks of likelihood_ratio.
SidRunInfo sidCallToStreamSharded(const std::string& method, const char* text, size_t len, bool estimate_prior, double prior,
                                  double error_threshold, double significance_level, const std::vector<int>& devices, std::ostream& out,
                                  std::ostream& log, const char* header = nullptr);
// Only rows labelled "het" from now on: the `grep ',het,'` of scripts/sid-pipeline/run-sid.sh:16-17 done
// before the rows leave the GPU.  Applies to sidCallToStream and to the four call* functions.
void sidSetHetOnly(bool het_only);
// BGZF input (sidCallFile): true = its members are inflated by the host's threads (host/bgzf.hpp), false (default) = on the device.
void sidSetHostInflate(bool on);
