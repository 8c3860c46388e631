/* sidgpu.h -- C ABI of the B200-native genotype-calling path (libsidgpu.so).
 *
 * The reference (EvolBioInf/sid) has no FFI layer: its boundary is a set of C++ free functions
 * (call.hpp:12,40-43; pileup.hpp:20-44; lynch.hpp:44-46; stats.hpp:6-13).  Each entry point below
 * names the reference interface it replaces.  Signatures use plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns an int status (SIDGPU_OK == 0); sidgpu_last_error() gives the text.
 *   - one ctx per GPU; calls on one ctx are not thread safe; different ctxs are independent.
 *   - "d_" pointers are device pointers on the ctx's device, "h_" pointers are host pointers.
 *   - the caller owns every buffer it passes in; the ctx owns scratch, the profile table,
 *     the site store and its stream.  Pointers handed OUT by the ctx (sidgpu_sites_view,
 *     sidgpu_unique_view) stay valid until the next call that mutates the session.
 *   - work is enqueued on the ctx's stream; functions that return host-visible numbers
 *     synchronise that stream before returning.
 *   - there is no CPU fallback: every entry point fails with SIDGPU_ECUDA when no sm_100 device
 *     is usable.
 */
#ifndef SIDGPU_H
#define SIDGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    SIDGPU_OK = 0,
    SIDGPU_EINVAL = 1,        /* bad argument */
    SIDGPU_ECUDA = 2,         /* CUDA runtime / driver error, or no usable device */
    SIDGPU_EMALFORMED = 3,    /* pileup.cpp:9   "Malformed pileup line" */
    SIDGPU_EMISSING_MAPQ = 4, /* pileup.cpp:10  "Malformed pileup line or missing mapping qualities" */
    SIDGPU_EQUAL_SHORT = 5,   /* fewer quality characters than counted bases (call.cpp:330-331 reads
                                 past the vectors there: undefined behaviour in the reference) */
    SIDGPU_ECAPACITY = 6,     /* an output buffer passed by the caller is too small */
    SIDGPU_ENOMEM = 7,
    SIDGPU_ESTATE = 8,        /* call sequence error (e.g. emit before finish) */
    SIDGPU_EINTERNAL = 9
};

/* Calling methods: the `-m` strings of sid.cpp:92-100 and the functions of call.hpp:40-43. */
enum {
    SIDGPU_METHOD_LOCAL = 0,            /* callSiteMLError        call.cpp:213-289 */
    SIDGPU_METHOD_BAYES = 1,            /* callBayes              call.cpp:145-211 */
    SIDGPU_METHOD_LIKELIHOOD_RATIO = 2, /* callLikelihoodRatio    call.cpp:62-143  */
    SIDGPU_METHOD_QUALITY = 3           /* callQualityBasedSimple call.cpp:291-372 */
};

typedef struct sidgpu_ctx sidgpu_ctx;

typedef struct {
    int device;               /* CUDA device ordinal */
    size_t max_chunk_bytes;   /* largest text chunk one feed/tokenize call may carry (0: 256 MiB) */
    size_t max_sites;         /* site-store capacity for methods that keep all sites (0: grow on demand) */
    int table_log2;           /* initial log2 capacity of the unique-profile table (0: 20); grows */
    void* stream;             /* cudaStream_t to run on (NULL: the ctx creates its own) */
} sidgpu_config;

/* Parameters of one calling session: GlobalOptions of sid.cpp:11-17.
 * (Environment: SIDGPU_SLICE_LINES=<lines per tokenizer slice, default 31> is a tuning knob read once
 * per process; results do not depend on it.) */
typedef struct {
    int method;                 /* SIDGPU_METHOD_* */
    int estimate_prior;         /* -R */
    double prior;               /* -r   (<= 0: none) */
    double error_threshold;     /* -E   default 0.1  */
    double significance_level;  /* -p   default 0.05 */
    /* Lynch fit override for tests and multi-GPU: when fit_given != 0 the session uses these
     * (pi, eps, nucleotide distribution) instead of running the optimiser. */
    int fit_given;
    double fit_pi, fit_eps, fit_nd[4];
    /* Emit only the rows whose label is "het": what the reference's pipeline keeps of the CSV
     * (scripts/sid-pipeline/run-sid.sh:16-17 pipes it through grep ',het,').  Sites, fits and
     * sidgpu_emit_records are unaffected; *rows_out counts the rows written. */
    int het_only;
    /* Keep, per site, the profile of the bases read on the forward strand (ReadStack::strands,
     * pileup.hpp:15, summed per letter; a by-product of the tokenizer pass): sidgpu_emit_columns can then deliver
     * d_profile and d_fwd beside the call, which is what a strand-bias filter needs.  Costs 16 bytes per site of the store. */
    int want_strands;
} sidgpu_params;

/* ------------------------------------------------------------------------------------------------
 * Lifetime, memory, errors
 * --------------------------------------------------------------------------------------------- */
int sidgpu_create(const sidgpu_config* cfg, sidgpu_ctx** out);
void sidgpu_destroy(sidgpu_ctx* ctx);
const char* sidgpu_last_error(const sidgpu_ctx* ctx);   /* ctx may be NULL: error of a failed create */
const char* sidgpu_version(void);
int sidgpu_synchronize(sidgpu_ctx* ctx);
/* Thin allocation/copy helpers so that hosts without a CUDA toolchain (cgo, ctypes, the C++ host)
 * can stage buffers.  Pinned host memory for the streaming path. */
int sidgpu_malloc(sidgpu_ctx* ctx, size_t bytes, void** d_ptr);
int sidgpu_free(sidgpu_ctx* ctx, void* d_ptr);
int sidgpu_malloc_host(sidgpu_ctx* ctx, size_t bytes, void** h_ptr);
int sidgpu_free_host(sidgpu_ctx* ctx, void* h_ptr);
int sidgpu_memcpy_h2d(sidgpu_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int sidgpu_memcpy_d2h(sidgpu_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int sidgpu_memcpy_d2d(sidgpu_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);
/* The same, only enqueued on the ctx's stream (for hosts that run their own work on that stream). */
int sidgpu_memcpy_d2d_async(sidgpu_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);

/* ------------------------------------------------------------------------------------------------
 * K1: tokenizer + profile builder
 *   replaces readFile (call.cpp:11-20), parsePileupLine (pileup.cpp:13-68),
 *   parseReadBases (pileup.cpp:70-153) and, with want_qual, parseQualities (pileup.cpp:155-167).
 * Text layout: `d_text` is the base of a device buffer holding `text_len` valid bytes of mpileup
 * text.  The call owns the lines whose FIRST byte lies in [range_begin, range_end): a byte p starts a
 * line iff text[p] != '\n' and (p == 0 or text[p-1] == '\n') -- the sharding rule of SURVEY.md 8(e);
 * a line that straddles range_end is read to its end (up to text_len).  Empty lines are skipped
 * (call.cpp:14).  d_text must be 16-byte aligned.
 * want_qual: 0 = five columns suffice; 1 = the quality columns are required as `quality` requires them (pileup.cpp:42-66)
 * and d_line_off is filled; 2 = d_line_off is filled, the quality columns are not looked at; 3 = as 2, and the tokenizer
 * also counts the strands (a by-product of the same pass): d_fwd is filled.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t n_sites;          /* lines parsed by the call */
    const uint64_t* d_profile; /* per site: A | C<<16 | G<<32 | T<<48, each count mod 65536 (pileup.hpp:7) */
    const int32_t* d_pos;      /* per site: atoi(position column) */
    const uint32_t* d_slot;    /* per site: slot of its profile in the unique-profile table */
    const uint64_t* d_line_off;/* per site: byte offset of the line in d_text (only with want_qual != 0) */
    /* chromosome names are interned on the device: d_name_ref[i] is a byte offset into d_names,
     * where a 2-byte little-endian length is followed by the name bytes */
    const uint32_t* d_name_ref;
    const char* d_names;
    uint64_t names_bytes;
    /* only with want_qual == 3: per site, the profile of the bases read on the FORWARD strand (upper-case characters and
     * '.'; pileup.cpp:78-124), packed like d_profile; the reverse strand is d_profile - d_fwd, count by count */
    const uint64_t* d_fwd;
} sidgpu_sites_view;

int sidgpu_tokenize(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin,
                    size_t range_end, int want_qual, sidgpu_sites_view* out);

/* The per-read side of the parser interface, for hosts that ask for it (readFile(in, true, true), call.cpp:11-20;
 * PileupLine::bases / strands / base_qualities / mapping_qualities, pileup.hpp:14-17).  The calling path never builds
 * these vectors (`quality` consumes the text on the device); they exist for inspection and for the reference's tests.
 *   sidgpu_qualities    parseQualities (pileup.cpp:155-167) of one string: character - 33, at least 1, up to the first
 *                       NUL, tab or line end; *n_out = number of qualities written to d_out.
 *   sidgpu_read_counts  per line (d_line_off[i] = offset of its first byte in d_text): the lengths of its three vectors.
 *                       A line with fewer than five columns is SIDGPU_EMALFORMED.  The quality columns matter only
 *                       when asked for: a missing sixth column is SIDGPU_EMALFORMED with want_baseq != 0, a missing
 *                       seventh SIDGPU_EMISSING_MAPQ with want_mapq != 0 (pileup.cpp:60-66); else their counts are 0.
 *   sidgpu_read_fill    the vectors themselves at the offsets the caller derived from the counts (exclusive prefix sums):
 *                       bases as upper-case letters, strands 1 = upper case / forward (pileup.cpp:84-123), qualities
 *                       as numbers.  Any output pointer may be NULL.
 *   sidgpu_strand_counts   the strands the reference parses and never reads (pileup.hpp:15,
 *                       pileup.cpp:87-123), summed per line: d_fwd[i] / d_rev[i] = the counted bases of line i read on the
 *                       forward strand (upper-case characters and '.') / the reverse strand (lower case and ','), packed
 *                       like a profile (A | C<<16 | G<<32 | T<<48, each mod 65536); d_fwd[i] + d_rev[i] is the profile,
 *                       count by count.  Either pointer may be NULL.  Line offsets: sidgpu_tokenize(..., want_qual = 2). */
int sidgpu_qualities(sidgpu_ctx* ctx, const char* d_quals, size_t n, uint8_t* d_out, uint64_t* n_out);
int sidgpu_strand_counts(sidgpu_ctx* ctx, const char* d_text, size_t text_len, const uint64_t* d_line_off, uint64_t n_lines,
                         uint64_t* d_fwd, uint64_t* d_rev);
int sidgpu_read_counts(sidgpu_ctx* ctx, const char* d_text, size_t text_len, const uint64_t* d_line_off, uint64_t n_lines,
                       int want_baseq, int want_mapq, uint32_t* d_n_bases, uint32_t* d_n_bq, uint32_t* d_n_mq);
int sidgpu_read_fill(sidgpu_ctx* ctx, const char* d_text, size_t text_len, const uint64_t* d_line_off, uint64_t n_lines,
                     const uint64_t* d_base_off, const uint64_t* d_bq_off, const uint64_t* d_mq_off, char* d_bases,
                     uint8_t* d_strands, uint8_t* d_bq, uint8_t* d_mq);

/* ------------------------------------------------------------------------------------------------
 * BGZF input inflated on the device (the pipeline stores its pileups gzipped, scripts/prepare-data.sh:14,
 * and `zcat`s them to a temporary file on one core, scripts/sid-pipeline/run-sid.sh:15).  A BGZF file (`bgzip`) is a
 * series of independent gzip members of at most 64 KiB of text; only the compressed bytes cross the link.
 *   sidgpu_bgzf_scan     host only: walks the member headers of h_comp[0, len) and lists the members whose text fits
 *                        text_cap (offsets relative to h_comp and to the start of the text); *consumed = bytes of the
 *                        whole members it accepted (a member cut by `len` is left for the next call), *text_bytes =
 *                        their text.  SIDGPU_EINVAL for bytes that are not a BGZF member header.
 *   sidgpu_inflate_bgzf  inflates the listed members of the device buffer d_comp (4-byte aligned, readable 8 bytes
 *                        past comp_len) into d_text, one warp per member; checks every member's text against the ISIZE and
 *                        the CRC-32 of its trailer; SIDGPU_EINVAL with the member's index in sidgpu_last_error for a
 *                        damaged member.
 *   sidgpu_call_io_bgzf  sidgpu_call_io for a BGZF file: the read callback delivers the FILE's bytes (compressed).
 *   sidgpu_call_host_bgzf  the same from a host buffer to a host buffer.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t c_off;    /* offset of the member's deflate stream in the compressed buffer */
    uint64_t out_off;  /* offset of its text in the text buffer */
    uint32_t c_len;    /* bytes of the deflate stream */
    uint32_t isize;    /* bytes of text (trailer) */
    uint32_t crc;      /* CRC-32 of the text (trailer) */
    uint32_t reserved;
} sidgpu_bgzf_block;
int sidgpu_bgzf_scan(const void* h_comp, size_t len, sidgpu_bgzf_block* blocks, size_t max_blocks, size_t text_cap,
                     size_t* n_blocks, size_t* consumed, size_t* text_bytes);
int sidgpu_inflate_bgzf(sidgpu_ctx* ctx, const void* d_comp, size_t comp_len, const sidgpu_bgzf_block* h_blocks, size_t n_blocks,
                        char* d_text, size_t text_cap);
/* sidgpu_call_host (below) for a BGZF file that lies in host memory: h_comp holds the FILE's bytes (pin it for speed), the
 * members are inflated on the device chunk by chunk; everything else as sidgpu_call_host. */
int sidgpu_call_host_bgzf(sidgpu_ctx* ctx, const sidgpu_params* params, const void* h_comp, size_t comp_len, char* h_csv,
                          size_t csv_cap, uint64_t* csv_bytes, uint64_t* n_sites, uint64_t* n_rows);

/* ------------------------------------------------------------------------------------------------
 * Calling sessions: the four functions of call.hpp:40-43, streamed.
 *
 *   sidgpu_begin(params)
 *   sidgpu_feed(d_text, ...)      any number of chunks, in file order; chunks must end on a line end
 *                                 (or be the last one).  For SIDGPU_METHOD_LOCAL without -R and for
 *                                 SIDGPU_METHOD_QUALITY without -R every feed can be followed at
 *                                 once by sidgpu_emit_csv for the sites of that chunk.
 *   sidgpu_finish()               global step: Lynch fit (lynch.cpp:17-35) when the method needs it,
 *                                 per-unique-profile classification (call.cpp:93-127,176-194,238-273),
 *                                 Benjamini-Hochberg (stats.cpp:58-80).
 *   sidgpu_emit_csv(...)          CSV rows (call.hpp:29-38), in site order, sites whose profile has
 *                                 coverage < 4 dropped for bayes / likelihood_ratio (call.cpp:131-140).
 * --------------------------------------------------------------------------------------------- */
int sidgpu_begin(sidgpu_ctx* ctx, const sidgpu_params* params);
/* Tokenizes one chunk and joins its sites against the unique-profile table (K1 + K3).
 * n_sites_out (optional) receives the number of sites the chunk added.
 * Limits: sessions that keep their sites (bayes, likelihood_ratio, -R) index them with 32 bits: at most 2^32 - 2 sites
 * per ctx (SIDGPU_ECAPACITY beyond; shard larger genomes over several GPUs or sessions).  One call takes at most
 * 2^31 tokenizer slices (about 5 TB of text). */
int sidgpu_feed(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin,
                size_t range_end, uint64_t* n_sites_out);
int sidgpu_finish(sidgpu_ctx* ctx);
/* Text in, CSV rows out in ONE pass for sessions whose rows need no genome-wide step (SIDGPU_METHOD_LOCAL without -R):
 * readFile + callSiteMLError + operator<< (call.cpp:11-20,213-289; call.hpp:29-38) of the lines that start in
 * [range_begin, range_end).  The tokenizer kernel classifies a profile when it first meets it and writes the rows
 * itself; nothing is stored per site, so the rows exist only in d_out (sidgpu_emit_* address the sites of sidgpu_feed,
 * not of this call).  *bytes_out: CSV bytes written (no header line); *rows_out: rows (fewer than sites with het_only);
 * *n_sites_out: lines parsed.  SIDGPU_ECAPACITY (needed size in *bytes_out) when out_cap is too small. */
int sidgpu_feed_rows(sidgpu_ctx* ctx, const char* d_text, size_t text_len, size_t range_begin, size_t range_end,
                     char* d_out, size_t out_cap, uint64_t* bytes_out, uint64_t* rows_out, uint64_t* n_sites_out);
/* Formats sites [site_begin, site_begin + n_sites) of the session into d_out (K2 for profiles not yet
 * classified + K6).  In streaming mode (local/quality without -R) only the sites of the most recent
 * feed are addressable.  *bytes_out receives the CSV bytes written (no header line). */
int sidgpu_emit_csv(sidgpu_ctx* ctx, uint64_t site_begin, uint64_t n_sites, char* d_out, size_t out_cap,
                    uint64_t* bytes_out, uint64_t* rows_out);
/* Per-site results as arrays instead of text (what OutputRecord carries, call.hpp:14-27):
 * label 0 hom / 1 het / 255 dropped; gt two chars; confidences as doubles.  Any pointer may be NULL. */
int sidgpu_emit_records(sidgpu_ctx* ctx, uint64_t site_begin, uint64_t n_sites, uint8_t* d_label,
                        char* d_gt, double* d_hom_conf, double* d_het_conf);

/* The same per-site results as columns, with the position and the chromosome name of every site: the
 * columnar form downstream consumers read (scripts/nonsynonymous.py:10-12,36 re-parses the CSV into
 * exactly these fields).  Any pointer may be NULL; arrays of n_sites elements (d_gt: 2 per site), file
 * order.  d_name_ref[i] is the byte offset of a record (uint16 length, then the bytes) in the names
 * pool that sidgpu_names returns.  (`quality` sessions: the per-site call runs for the range, as for sidgpu_emit_records.) */
typedef struct {
    int32_t* d_pos;
    uint32_t* d_name_ref;
    uint8_t* d_label;
    char* d_gt;
    double* d_hom_conf;
    double* d_het_conf;
    uint64_t* d_profile;    /* the site's counts, A | C<<16 | G<<32 | T<<48 */
    uint64_t* d_fwd;        /* those of the forward strand (sessions begun with want_strands; SIDGPU_EINVAL otherwise) */
} sidgpu_columns;
int sidgpu_emit_columns(sidgpu_ctx* ctx, uint64_t site_begin, uint64_t n_sites, const sidgpu_columns* cols);
/* The chromosome-name pool of the ctx (device memory, valid until the ctx is destroyed or grown). */
int sidgpu_names(sidgpu_ctx* ctx, const char** d_names, uint64_t* names_bytes);

/* One-call host-buffer path (what the `sid` binary and the call.hpp wrappers use): chunked,
 * double-buffered pinned H2D of `h_text`, the session above on the device, D2H of the CSV.
 * h_csv receives the rows (no header); *csv_bytes the size.  Returns SIDGPU_ECAPACITY (with the
 * needed size in *csv_bytes) when csv_cap is too small. */
int sidgpu_call_host(sidgpu_ctx* ctx, const sidgpu_params* params, const char* h_text, size_t text_len,
                     char* h_csv, size_t csv_cap, uint64_t* csv_bytes, uint64_t* n_sites, uint64_t* n_rows);

/* The two halves of sidgpu_call_host for sessions that keep their sites (bayes, likelihood_ratio, local -R)
 * when something has to happen between them: several shards of one genome, one ctx each, that share one fit
 * (sidgpu_histogram / sidgpu_lynch_objective per shard, summed by the host; sidgpu_set_fit; sidgpu_finish or
 * sidgpu_finish_global).  sidgpu_feed_host: after sidgpu_begin, chunked upload + K1/K3 of a host text.
 * sidgpu_emit_host: after the finish call, the rows of all stored sites into a host buffer. */
int sidgpu_feed_host(sidgpu_ctx* ctx, const char* h_text, size_t text_len, uint64_t* n_sites);
int sidgpu_emit_host(sidgpu_ctx* ctx, char* h_csv, size_t csv_cap, uint64_t* csv_bytes, uint64_t* n_rows);
/* Text in, rows out chunk by chunk on an open session: the second pass of `quality -R` (its sites are not kept;
 * call it with the same text after sidgpu_finish), or a streaming session (local / quality without -R). */
int sidgpu_stream_host(sidgpu_ctx* ctx, const char* h_text, size_t text_len, char* h_csv, size_t csv_cap,
                       uint64_t* csv_bytes, uint64_t* n_sites, uint64_t* n_rows);

/* The same session for inputs and outputs of any size (what the `sid` binary uses: sid.cpp:85-105 reads the whole
 * file into a vector of lines and prints a vector of records; here neither exists).  The ctx owns a ring of pinned
 * text slots (max_chunk_bytes each, 64 MiB at most) and a ring of pinned CSV slots; an internal reader thread fills
 * the text slots through `read` and cuts them at line ends, the calling thread keeps the copies and kernels going,
 * an internal writer thread hands finished rows to `write` in file order.  Methods with a genome-wide step read the
 * whole input first (their sites stay in device memory) and emit afterwards; `quality` with -R reads the input
 * twice and needs `rewind`.
 *   read(user, dst, cap)  up to cap bytes of text into dst; returns the count, 0 at the end of the input, < 0 on error.
 *   write(user, rows, n)  n bytes of CSV rows (whole lines, no header); returns 0, anything else aborts the call.
 *   rewind(user)          restarts the input; may be NULL; returns 0 on success.
 * read and rewind are called from one thread, write from another, never concurrently with themselves.
 * *csv_bytes / *n_sites / *n_rows (optional) receive the totals. */
typedef struct {
    int64_t (*read)(void* user, char* dst, size_t cap);
    int (*write)(void* user, const char* rows, size_t n);
    int (*rewind)(void* user);
    void* user;
} sidgpu_io;
int sidgpu_call_io(sidgpu_ctx* ctx, const sidgpu_params* params, const sidgpu_io* io, uint64_t* csv_bytes,
                   uint64_t* n_sites, uint64_t* n_rows);
int sidgpu_call_io_bgzf(sidgpu_ctx* ctx, const sidgpu_params* params, const sidgpu_io* io, uint64_t* csv_bytes,
                        uint64_t* n_sites, uint64_t* n_rows);

/* ------------------------------------------------------------------------------------------------
 * K3: unique-profile histogram   (countUniqueProfiles pileup.cpp:169-196,
 *                                 computeNucleotideDistribution pileup.cpp:198-217)
 * --------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t n_unique;          /* entries below */
    const uint64_t* d_profile;  /* packed profile, lexicographic order of (A,C,G,T) like the reference */
    const uint64_t* d_count;    /* sites with that profile (64-bit: the reference's uint32 would wrap) */
    double nd[4];               /* nucleotide distribution over the entries */
    uint64_t nd_sums[5];        /* its integer numerators (A,C,G,T) and denominator: what a multi-GPU host all-reduces */
} sidgpu_unique_view;
/* Compacts and sorts the session's table.  min_coverage = 4 gives the Lynch input (call.cpp:66-70). */
int sidgpu_histogram(sidgpu_ctx* ctx, uint32_t min_coverage, sidgpu_unique_view* out);
/* countUniqueProfiles / computeNucleotideDistribution on caller-supplied profiles (packed as in
 * sidgpu_sites_view), outside a session: n profiles, optionally weighted by d_counts. */
int sidgpu_count_unique(sidgpu_ctx* ctx, const uint64_t* d_profiles, uint64_t n, uint32_t min_coverage, sidgpu_unique_view* out);
int sidgpu_count_unique_weighted(sidgpu_ctx* ctx, const uint64_t* d_profiles, const uint64_t* d_counts, uint64_t n,
                                 uint32_t min_coverage, sidgpu_unique_view* out);

/* ------------------------------------------------------------------------------------------------
 * K4: Lynch objective   (compoundLikelihood lynch.cpp:37-61 with lynch.hpp:57-74,82-90)
 * Evaluates -sum_u count_u * log((1-pi) L_hom,u + pi L_het,u) over the last sidgpu_histogram().
 * _partial leaves the (rank-local) sum in device memory at d_out so that a multi-GPU host can
 * all-reduce it (NCCL) before reading; sidgpu_lynch_objective returns the local value to the host.
 * Out-of-box arguments give DBL_MAX (lynch.cpp:41-43).
 * --------------------------------------------------------------------------------------------- */
int sidgpu_lynch_objective_partial(sidgpu_ctx* ctx, const double nd[4], double pi, double eps, double* d_out);
int sidgpu_lynch_objective(sidgpu_ctx* ctx, const double nd[4], double pi, double eps, double* value);
/* estimateProfileGenotypeLikelihoods (lynch.cpp:17-35) + FunctionMinimizer<2>::run
 * (optimization.hpp:51-89): Nelder-Mead from (1e-3,1e-3), steps 1e-4, stop at size < 1e-5 or
 * 1000 iterations; every objective evaluation runs on the device. */
typedef struct {
    double pi, eps, fval;
    int iterations, evaluations, converged;
} sidgpu_fit;
int sidgpu_lynch_fit(sidgpu_ctx* ctx, const double nd[4], sidgpu_fit* out);
/* Injects (pi, eps, nucleotide distribution) into the running session before sidgpu_finish, instead
 * of the local optimiser: a multi-GPU host fits on the all-reduced objective and hands every rank
 * the same result. */
int sidgpu_set_fit(sidgpu_ctx* ctx, double pi, double eps, const double nd[4]);
/* sidgpu_finish for a position-sharded `likelihood_ratio` session: Benjamini-Hochberg ranks over the
 * unique profiles of ALL shards (call.cpp:105-106, m = number of unique profiles), so the host merges
 * the shards' histograms (sidgpu_histogram) and hands every rank the merged profile list, sorted in
 * the reference's lexicographic order (n_global packed profiles, host memory).  Requires sidgpu_set_fit.
 * p-values, the BH adjustment and the classification of this rank's own profiles run on the device. */
int sidgpu_finish_global(sidgpu_ctx* ctx, const uint64_t* h_profiles_sorted, uint64_t n_global);
/* Position shards with a genome-wide step, the collective form: every rank passes the histograms of ALL shards
 * (device memory: n (profile, count) pairs, e.g. the all-gather of each rank's sidgpu_histogram, padded with
 * count 0) before sidgpu_finish.  The device merges them (counts summed, lexicographic order = countUniqueProfiles of
 * the whole genome, pileup.cpp:169-196), derives the nucleotide distribution from the merged integers and maps this
 * rank's own profiles onto the merged list; sidgpu_finish then fits and classifies exactly as a single GPU holding
 * the whole genome would, so every rank gets bit-identical (pi, eps) without any per-evaluation exchange. */
int sidgpu_set_global_histogram(sidgpu_ctx* ctx, const uint64_t* d_profiles, const uint64_t* d_counts, uint64_t n);
/* The fit the session used (after sidgpu_finish). */
int sidgpu_session_fit(sidgpu_ctx* ctx, sidgpu_fit* out, double nd[4], uint64_t* n_unique);

/* ------------------------------------------------------------------------------------------------
 * Device-side statistics kernels exposed for tests
 *   likelihoodRatioTest stats.cpp:29-37, adjustBenjaminiHochberg stats.cpp:58-80,
 *   `%g` formatting of operator<< call.hpp:29-38.
 * --------------------------------------------------------------------------------------------- */
int sidgpu_bh_adjust(sidgpu_ctx* ctx, const double* d_p, uint64_t n, double* d_adjusted);
/* likelihoodRatioTest (stats.cpp:29-37) on n pairs of LOG likelihoods (-inf stands for l == 0): p-values out. */
int sidgpu_lr_test(sidgpu_ctx* ctx, const double* d_log_h0, const double* d_log_h1, uint64_t n, double* d_p);
/* log homozygousLikelihood / log heterozygousLikelihood (lynch.hpp:57-74,82-90, multinomial coefficient included) of n
 * packed profiles under (nucleotide distribution, eps): what estimateProfileGenotypeLikelihoods returns per profile
 * (lynch.cpp:27-34), in log space. */
int sidgpu_profile_loglik(sidgpu_ctx* ctx, const uint64_t* d_profiles, uint64_t n, const double nd[4], double eps,
                          double* d_log_hom, double* d_log_het);
/* Formats n doubles like printf("%g"); out is n fixed 16-byte cells, NUL padded. */
int sidgpu_format_g(sidgpu_ctx* ctx, const double* d_values, uint64_t n, char* d_out16);

/* Counters for benchmarking: kernels launched by this ctx since creation; and, when enabled,
 * device time per kernel family measured with CUDA events on the ctx's stream around each launch:
 * ms[0]/launches[0] tokenizer (K1), [1] classification (K2), [2] CSV (K6, or the compaction of the fused row writer's
 * regions), [3] the two small kernels that turn the tokenizer's block table into the file order of the sites,
 * [4] Lynch fit (K4: one k_lynch_fit launch, or the objective evaluations of the host loop), [5] histogram (K3),
 * [6] quality, [7] reserved.  sidgpu_profile(ctx, enable) resets the accumulators. */
#define SIDGPU_N_TIMERS 8
uint64_t sidgpu_launch_count(const sidgpu_ctx* ctx);
int sidgpu_profile(sidgpu_ctx* ctx, int enable);
int sidgpu_kernel_times(sidgpu_ctx* ctx, double ms[SIDGPU_N_TIMERS], uint64_t launches[SIDGPU_N_TIMERS]);

#ifdef __cplusplus
}
#endif
#endif
