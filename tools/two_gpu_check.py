import subprocess, sys, os
sys.path.insert(0, "tests")
import oracle_py as op
G = "tests/golden"
for flags, csv in ((["-m", "bayes"], "depth30.m_bayes.csv"), (["-m", "likelihood_ratio", "-R"], "depth30.m_likelihood_ratio_R.csv"), (["-m", "local"], "depth30.m_local.csv")):
    r = subprocess.run(["host/sid", "--devices", "0,1"] + flags + [os.path.join(G, "depth30.plp")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    n, diffs = op.compare_csv(r.stdout, open(os.path.join(G, csv), "rb").read())
    print(flags, r.returncode, n, diffs, r.stderr.decode().strip().replace("\n", " | ")[:160])
