/* TEST INFRASTRUCTURE ONLY -- `sid`-compatible command line around the C restatement
 * (sid.cpp:11-17 defaults, :26-58 flags, :92-105 dispatch and CSV). */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "sid_oracle.h"

int main(int argc, char** argv) {
    const char* method = "local";
    int estimate_prior = 0;
    double prior = -1, alpha = 0.05, E = 0.1;
    int flag;
    while ((flag = getopt(argc, argv, "E:Rhm:p:r:")) != -1) {
        switch (flag) {
            case 'm': method = optarg; break;
            case 'r': prior = atof(optarg); break;
            case 'R': estimate_prior = 1; break;
            case 'p': alpha = atof(optarg); break;
            case 'E': E = atof(optarg); break;
            case 'h': printf("sid_oracle [flags] input_file\n"); break;
            default: return EXIT_FAILURE;
        }
    }
    if (optind >= argc) { fprintf(stderr, "No file name given!\n"); return EXIT_FAILURE; }
    const char* path = argv[optind];
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "Could not open file: %s\n", path); return EXIT_FAILURE; }
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* text = (char*)malloc((size_t)len + 1);
    if (fread(text, 1, (size_t)len, f) != (size_t)len) { fprintf(stderr, "read error\n"); return EXIT_FAILURE; }
    fclose(f);
    int m = -1;
    if (!strcmp(method, "local")) m = ORC_LOCAL;
    else if (!strcmp(method, "bayes")) m = ORC_BAYES;
    else if (!strcmp(method, "likelihood_ratio")) m = ORC_LIKELIHOOD_RATIO;
    else if (!strcmp(method, "quality")) m = ORC_QUALITY;
    if (m < 0) { printf("chrom,pos,label,gt,hom_conf,het_conf,conf_type\n"); return 0; } /* sid.cpp:92-102 */
    orc_result r;
    int st = orc_call(text, (size_t)len, m, estimate_prior, prior, E, alpha, &r);
    if (st != ORC_OK) {
        fprintf(stderr, "terminate: %s\n", st == ORC_MALFORMED_OR_MISSING
                ? "Malformed pileup line or missing mapping qualities" : "Malformed pileup line");
        return 134;
    }
    if (m != ORC_QUALITY && (m != ORC_LOCAL || estimate_prior)) {
        if (m != ORC_LOCAL) fprintf(stderr, "# unique profiles: %zu\n", r.n_unique);
        fprintf(stderr, "# GSL function minimization converged in %d iterations.\n", r.iterations);
        if (m != ORC_LOCAL) {
            fprintf(stderr, "# heterozygosity: %e\n# error: %e\n", r.heterozygosity, r.error_rate);
        }
    }
    size_t need = orc_write_csv(text, &r, NULL, 0);
    char* out = (char*)malloc(need);
    orc_write_csv(text, &r, out, need);
    fwrite(out, 1, need, stdout);
    free(out); free(text);
    orc_free_result(&r);
    return 0;
}
