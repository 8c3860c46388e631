"""Generates the golden fixtures of this directory from the REFERENCE ITSELF.

Run in the build container (needs /root/reference and therefore oracle/_ref/sid_ref, see
oracle/Makefile):  python tests/golden/make_golden.py
Inputs (*.plp) are written by this script; outputs (*.csv) are what the unmodified reference
prints for them; manifest.json lists every (input, flags, output) case plus the fitted
heterozygosity / error rate the reference logs on stderr."""
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "sid_ref")

from sid_b200 import synth  # noqa: E402

EDGE = [
    "chr1\t1\tA\t6\tAgACgt\tIIIIII",
    "chr1\t2\tA\t10\t..........\tIIIIIIIIII",
    "chr1\t3\tA\t10\t.....CCCCC\tIIIIIIIIII",
    "chr1\t4\tA\t10\t.....ccccT\tIIIIIIIIII",
    "chr1\t5\tA\t0\t*\t*",
    "chr1\t6\tg\t9\t^].,$.+2AC,-3acgT*<>nN\tIIIIIIIII",
    "chr1\t7\tC\t30\t" + "C" * 15 + "T" * 15 + "\t" + "I" * 30,
    "chr1\t8\tC\t30\t" + "C" * 27 + "TtG\t" + "I" * 30,
    "chr1\t9\tN\t4\tACGT\tIIII",
    "chr1\t10\tN\t4\tAACC\tIIII",
    "chr1\t11\tN\t3\tAGG\tIII",
    "chr1\t12\tA\t5\t..^+3..\tIIIII",
    "chr1\t13\tA\t4000\t" + "A" * 2000 + "C" * 2000 + "\t" + "I" * 4000,
    "chr1\t14\tT\t3\t^^A^+3AC\tIII",
    "chr1\t15\tT\t4\t+12345678901234567890AAAA\tIIII",
    "chr1\t16\tT\t4\tA-0AC+007acgtacgT\tIIII",
    "chr1\t17\tn\t2\ta+3act--3ggga\tII",
    "chr2 18 c 4 ,,.. IIII",
    "chr2\t\t19 \t G\t\t3\t.,A\tIII",
    "  chr2\t20\tA\t2\t.G\tII",
    "chr2\t21\tA\t2\t.$,",
    "chr2\t-5\tA\t1\t.\tI",
    "chr2\t+7\tA\t1\t,\tI",
    "chr2\t12abc\tA\t1\tA\tI",
    "chr2\t99999999999999999999\tA\t1\tC\tI",
    "chr2\t-99999999999999999999\tA\t1\tC\tI",
    "chr2\t4294967301\tA\t1\tG\tI",
    "chrX\t25\t*\t3\t.,T\tIII",
    "chrX\t26\tA\t3\t.,T\tIII\r",
    "",
    "",
    "chrX\t27\tA\t12\tAAAACCCCGGGG\tIIIIIIIIIIII",
    "chrX\t28\tA\t12\tAAAACCCCTTTT\tIIIIIIIIIIII",
    "chrX\t29\tA\t8\tacgtACGT\tIIIIIIII",
    "a_rather_long_scaffold_name_NW_003613580v1_random\t30\tA\t5\t..,,^~.\tIIIII",
    "chrX\t31\tA\t6\t.$.$^I.^+.+\tIIIIII",
    "chrX\t32\tA\t6\t.-2AA.+1C.^-.\tIIIIII",
    "chrX\t33\tA\t1\t*\tI",
    "chrX\t34\tA\t70000\t" + "." * 70000 + "\t" + "I" * 70000,
    "chrY\t35\tG\t7\t.,.,AAt\tIIIIIII",
    # pileup.cpp:78-83 substitutes '.' / ',' by the reference character BEFORE the switch: with '^' every '.' eats
    # the next byte, with '+' / '-' it starts an indel when a digit follows
    "chrY\t36\t^\t4\t.A.C\tIIII",
    "chrY\t37\t^\t6\t,AC.GT.\tIIIIII",
    "chrY\t38\t+\t6\t.2ACGT,1AC\tIIIIII",
    "chrY\t39\t-\t6\tA.1CG,TT.x\tIIIIII",
    "chrY\t40\t+\t5\t..3ACGTA\tIIIII",
    "chrY\t41\t-\t5\tAC,GT.\tIIIII",
]

QUAL_EDGE = [
    "chr19\t1337\tA\t6\tAgACgt\t++5D5D\tDD55DD",
    "chr19\t1338\tA\t6\t.,.,.,\tIIIIII\t]]]]]]",
    "chr19\t1339\tA\t6\t.,*N<>AC\t!!5D5D5D\tDD55DD55",
    "chr19\t1340\tA\t6\t^].,$.+2AC,-3acgT\tIIIIII\t]]]]]]",
    "chr19\t1341\tC\t0\t*\t*\t*",
    "chr19\t1342\tC\t8\tccccTTTT\t+5?I+5?I\t!!!!IIII",
    "chr19\t1343\t^\t4\t.A.CGT\tIIII\t]]]]",
    "chr19\t1344\t+\t4\t.1AC,GT\tIIII\t]]]]",
    "chr19\t1345\t-\t4\tA.2CGTT\tIIII\t]]]]",
]


def run(plp, flags):
    r = subprocess.run([REF] + flags + [plp], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout, r.stderr.decode()


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/sid_ref missing: run `make -C oracle ref` (needs /root/reference)")
    files = {}
    files["edge"] = ("\n".join(EDGE) + "\n").encode() + b"chrY\t42\tT\t3\t..,\tIII"      # last line unterminated
    files["edge_quality"] = ("\n".join(QUAL_EDGE) + "\n").encode()
    files["depth30"] = synth.generate(4000, seed=11, **synth.CONFIGS["depth30"]).tobytes()
    files["depth30_two_chroms"] = synth.generate(3000, seed=12, chroms=("chr1", "chr2"), chrom_lengths=(1700, 1300),
                                                 **synth.CONFIGS["depth30"]).tobytes()
    files["depth500"] = synth.generate(300, seed=13, **synth.CONFIGS["depth500"]).tobytes()
    files["depth5"] = synth.generate(3000, seed=14, lam=5.0, het=0.01, err=0.02, start=0.02, indel=0.01).tobytes()
    files["quality30"] = synth.generate(3000, seed=15, seven_columns=True, **synth.CONFIGS["depth30"]).tobytes()
    cases = []
    flagsets = {
        "edge": [["-m", "local"], ["-m", "local", "-r", "0.001"], ["-m", "local", "-E", "0.05", "-p", "0.01"],
                 ["-m", "bayes"], ["-m", "likelihood_ratio"], ["-m", "likelihood_ratio", "-R"], ["-m", "local", "-R"]],
        "edge_quality": [["-m", "quality"], ["-m", "quality", "-r", "0.001"], ["-m", "local"]],
        "depth30": [["-m", "local"], ["-m", "bayes"], ["-m", "likelihood_ratio"], ["-m", "likelihood_ratio", "-R"],
                    ["-m", "local", "-R"], ["-m", "local", "-r", "0.01", "-p", "0.001"]],
        "depth30_two_chroms": [["-m", "local"], ["-m", "bayes"]],
        "depth500": [["-m", "local"], ["-m", "bayes"], ["-m", "likelihood_ratio"]],
        "depth5": [["-m", "local"], ["-m", "bayes"], ["-m", "likelihood_ratio"]],
        "quality30": [["-m", "quality"], ["-m", "quality", "-R"], ["-m", "quality", "-r", "0.001", "-p", "0.01"], ["-m", "bayes"]],
    }
    for name, data in files.items():
        plp = os.path.join(HERE, name + ".plp")
        with open(plp, "wb") as f:
            f.write(data)
        for flags in flagsets[name]:
            tag = "_".join(x.strip("-") for x in flags).replace(".", "p")
            rc, out, err = run(plp, flags)
            assert rc == 0, (name, flags, rc, err)
            csv = "%s.%s.csv" % (name, tag)
            with open(os.path.join(HERE, csv), "wb") as f:
                f.write(out)
            case = {"input": name + ".plp", "flags": flags, "csv": csv}
            m = re.search(r"# heterozygosity: (\S+)", err)
            if m:
                case["heterozygosity"] = float(m.group(1))
                case["error"] = float(re.search(r"# error: (\S+)", err).group(1))
                case["unique_profiles"] = int(re.search(r"# unique profiles: (\d+)", err).group(1))
            m = re.search(r"converged in (\d+) iterations", err)
            if m:
                case["iterations"] = int(m.group(1))
            cases.append(case)
    # malformed inputs: the reference throws std::invalid_argument and aborts (exit 134)
    malformed = {
        "too_few_columns": b"chr1\t1\tA\t3\n",
        "long_reference": b"chr1\t1\tAC\t3\t...\tIII\n",
        "only_name": b"chr1\n",
        "second_line_bad": b"chr1\t1\tA\t3\t...\tIII\nchr1\t2\n",
    }
    bad = []
    for name, data in malformed.items():
        plp = os.path.join(HERE, "malformed_" + name + ".plp")
        with open(plp, "wb") as f:
            f.write(data)
        rc, out, err = run(plp, ["-m", "local"])
        bad.append({"input": "malformed_" + name + ".plp", "flags": ["-m", "local"], "returncode": rc,
                    "what": "Malformed pileup line" if "Malformed pileup line" in err else err.strip()[-80:]})
    plp = os.path.join(HERE, "malformed_missing_mapq.plp")
    with open(plp, "wb") as f:
        f.write(b"chr1\t1\tA\t3\t...\tIII\n")
    rc, out, err = run(plp, ["-m", "quality"])
    bad.append({"input": "malformed_missing_mapq.plp", "flags": ["-m", "quality"], "returncode": rc,
                "what": "missing mapping qualities" if "missing mapping qualities" in err else err.strip()[-80:]})
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump({"cases": cases, "malformed": bad,
                   "generated_by": "tests/golden/make_golden.py with oracle/_ref/sid_ref (unmodified reference + GSL stand-in)"},
                  f, indent=1)
    print("wrote", len(cases), "cases,", len(bad), "malformed cases")


if __name__ == "__main__":
    main()
