// Exercises the reference-compatible C++ API (call.hpp / pileup.hpp) the way the reference's own
// unit tests do (test/test-pileup_parser.cpp, test/test-profiles.cpp, test/test-call.cpp), but
// against the GPU implementation.  Prints one line per check; exit status 0 iff all pass.
#include <cmath>
#include <fstream>
#include <iostream>
#include <sstream>

#include "call.hpp"

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
