// Exercises the reference-compatible C++ API (call.hpp / pileup.hpp) the way the reference's own
// unit tests do (test/test-pileup_parser.cpp, test/test-profiles.cpp, test/test-call.cpp), but
// against the GPU implementation.  Prints one line per check; exit status 0 iff all pass.
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>

#include "call.hpp"
#include "lynch.hpp"
#include "stats.hpp"

static int failures = 0;
#define CHECK(cond)                                                    \
    do {                                                               \
        if (!(cond)) { ++failures; std::cout << "FAIL " #cond << "\n"; } \
        else std::cout << "ok   " #cond << "\n";                       \
    } while (0)

static bool counts(const char* bases, char ref, uint16_t a, uint16_t c, uint16_t g, uint16_t t) {
    const ReadStack s = parseReadBases(bases, ref, 0);
    return s.counts == profile_t {a, c, g, t};
}

int main(int argc, char** argv) {
    // test/test-profiles.cpp:16-55
    CHECK(counts("aA", 'n', 2, 0, 0, 0));
    CHECK(counts("tT", 'n', 0, 0, 0, 2));
    CHECK(counts("", 'n', 0, 0, 0, 0));
    CHECK(counts("a$", 'n', 1, 0, 0, 0));
    CHECK(counts("a^a", 'n', 1, 0, 0, 0));
    CHECK(counts("^aa", 'n', 1, 0, 0, 0));
    CHECK(counts("a+3act", 'n', 1, 0, 0, 0));
    CHECK(counts("-3acta", 'n', 1, 0, 0, 0));
    CHECK(counts("a.", 'g', 1, 0, 1, 0));
    CHECK(counts(",g", 'a', 1, 0, 1, 0));
    CHECK(counts("--3ggga", 'n', 1, 0, 0, 0));
    // test/test-pileup_parser.cpp:37-56
    {
        char line[] = "chr19\t1337\tA\t6\tAgACgt\t++5D5D\tDD55DD";
        const PileupLine l = parsePileupLine(line, true, true);
        CHECK(l.chromosome_name == "chr19");
        CHECK(l.position == 1337);
        CHECK(l.reference_base == 'A');
        CHECK((l.base_counts == profile_t {2, 1, 2, 1}));
    }
    {
        char bad[] = "chr19\t1337";
        bool threw = false;
        try { parsePileupLine(bad, false, false); } catch (const std::invalid_argument& e) { threw = std::string(e.what()) == "Malformed pileup line"; }
        CHECK(threw);
    }
    // test/test-call.cpp:16-35
    {
        std::vector<PileupLine> lines(3);
        lines[0].base_counts = {1, 1, 1, 1};
        lines[1].base_counts = {2, 2, 2, 2};
        lines[2].base_counts = {1, 1, 1, 1};
        const auto u = countUniqueProfiles(lines);
        CHECK(u.size() == 2 && u[0].profile == (profile_t {1, 1, 1, 1}) && u[0].count == 2 && u[1].count == 1 && u[1].coverage == 8);
        CHECK(countUniqueProfiles({}).empty());
        // test/test-likelihoods.cpp:51-83
        const auto nd0 = computeNucleotideDistribution({});
        CHECK(nd0[0] == 0.25 && nd0[3] == 0.25);
        std::vector<UniqueProfile> w {{{1, 0, 0, 0}, 4}, {{1, 1, 0, 0}, 2}, {{0, 0, 0, 1}, 2}};
        const auto nd = computeNucleotideDistribution(w);
        CHECK(std::fabs(nd[0] - 0.6) < 1e-12 && std::fabs(nd[1] - 0.2) < 1e-12 && nd[2] == 0 && std::fabs(nd[3] - 0.2) < 1e-12);
    }
    // test/test-pileup_parser.cpp:8-56: qualities, letters, strands; a line with six bases and five qualities
    CHECK((parseQualities("+5D", 3) == std::vector<uint8_t> {10, 20, 35}));
    CHECK(parseQualities("", 0).empty());
    CHECK((parseQualities("!\"I\tJJ", 3) == std::vector<uint8_t> {1, 1, 40}));      // '!' is 0 -> 1 (pileup.cpp:160-162); stops at the tab
    {
        const ReadStack s = parseReadBases("AgACgt", 'N', 6);
        CHECK((s.bases == std::vector<char> {'A', 'G', 'A', 'C', 'G', 'T'}));
        CHECK((s.strands == std::vector<bool> {1, 0, 1, 1, 0, 0}));
        const ReadStack d = parseReadBases(".,^F.+2ac,$", 't', 4);
        CHECK((d.bases == std::vector<char> {'T', 'T', 'T', 'T'}) && (d.strands == std::vector<bool> {1, 0, 1, 0}));
    }
    {
        char line[] = "chr19\t1337\tA\t6\tAgACgt\t++5D5\tDD55D";
        const PileupLine l = parsePileupLine(line, true, true);
        CHECK((l.base_counts == profile_t {2, 1, 2, 1}));
        CHECK((l.bases == std::vector<char> {'A', 'G', 'A', 'C', 'G', 'T'}));
        CHECK((l.strands == std::vector<bool> {1, 0, 1, 1, 0, 0}));
        CHECK((l.base_qualities == std::vector<uint8_t> {10, 10, 20, 35, 20}));
        CHECK((l.mapping_qualities == std::vector<uint8_t> {35, 35, 20, 20, 35}));
        char six[] = "chr19\t1337\tA\t6\tAgACgt\t++5D5";
        bool threw = false;
        try { parsePileupLine(six, true, true); } catch (const std::invalid_argument& e) { threw = std::string(e.what()) == "Malformed pileup line or missing mapping qualities"; }
        CHECK(threw);
        const PileupLine only_bq = parsePileupLine(six, true, false);
        CHECK(only_bq.base_qualities.size() == 5 && only_bq.mapping_qualities.empty());
    }
    // stats.hpp:8,11 (stats.cpp:29-37,58-80)
    CHECK(std::fabs(likelihoodRatioTest(0.5L, 1.0L) - std::erfc(std::sqrt(std::log(2.0)))) < 1e-15);
    CHECK(likelihoodRatioTest(1.0L, 0.5L) == 1.0 && likelihoodRatioTest(0.0L, 0.5L) == 0.0);
    {
        const std::vector<double> adj = adjustBenjaminiHochberg({0.01, 0.04, 0.03, 0.005});
        CHECK(adj.size() == 4 && std::fabs(adj[0] - 0.02) < 1e-15 && std::fabs(adj[1] - 0.04) < 1e-15 && std::fabs(adj[2] - 0.04) < 1e-15 && std::fabs(adj[3] - 0.02) < 1e-15);
    }
    // lynch.hpp:44-46: the fit, and its per-profile likelihoods reproduce the objective at the fitted point
    {
        std::vector<UniqueProfile> u;
        for (uint16_t n = 8; n < 40; ++n) {
            u.emplace_back(profile_t {n, 0, 0, 0}, 400u + n);
            u.emplace_back(profile_t {0, 0, n, 0}, 380u);
            u.emplace_back(profile_t {uint16_t(n - 1), 1, 0, 0}, 40u);
            u.emplace_back(profile_t {uint16_t(n / 2), 0, 0, uint16_t(n - n / 2)}, 3u);
        }
        const std::array<double, 4> nd = computeNucleotideDistribution(u);
        const ProfileGenotypeLikelihoods g = estimateProfileGenotypeLikelihoods(u, nd);
        CHECK(g.heterozygosity > 0 && g.heterozygosity < 0.1 && g.error_rate > 0 && g.error_rate < 0.1 && g.profile_likelihoods.size() == u.size());
        long double f = 0;
        for (size_t i = 0; i < u.size(); ++i)
            f -= u[i].count * logl((1 - g.heterozygosity) * g.profile_likelihoods[i].L_homozygous + g.heterozygosity * g.profile_likelihoods[i].L_heterozygous);
        const double dev = compoundLikelihood(g.heterozygosity, g.error_rate, u, nd);
        CHECK(std::fabs((double)f - dev) <= 1e-9 * std::fabs(dev));
        CHECK(compoundLikelihood(-0.1, 0.01, u, nd) == std::numeric_limits<double>::max());
        CHECK(compoundLikelihood(g.heterozygosity * 3, g.error_rate, u, nd) > dev);
    }
    // call.hpp:40-43 through an istream
    {
        std::istringstream in("chr1\t1\tA\t6\tAgACgt\tIIIIII\nchr1\t2\tA\t10\t..........\tIIIIIIIIII\n");
        const auto r = callSiteMLError(in, false, -1, 0.1, 0.05);
        CHECK(r.size() == 2);
        std::ostringstream os;
        for (const auto& x : r) os << x << '\n';
        CHECK(os.str() == "chr1,1,het,GA,1,0.00486457,p_value\nchr1,2,hom,AA,0.000196638,1,p_value\n");
    }
    if (argc > 1) {                                   // a pileup file: all four entry points run and agree on the row count rules
        std::ifstream f1(argv[1]), f2(argv[1]), f3(argv[1]);
        const auto a = callSiteMLError(f1, false, -1, 0.1, 0.05);
        const auto b = callBayes(f2);
        const auto c = callLikelihoodRatio(f3, false, 0.05);
        CHECK(!a.empty() && b.size() <= a.size() && b.size() == c.size());
        CHECK(b.empty() || b[0].classification.confidence_type == "probability");
    }
    std::cout << (failures ? "FAILED" : "PASSED") << "\n";
    return failures ? 1 : 0;
}
