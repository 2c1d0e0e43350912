"""Open-ended run of the random-pattern differential test (tests/test_regex_fuzz_cpu.py) on the CPU host build.
usage: [CF_PAIR_FILTER=0|1] python -u tools/fuzz_regex.py [first_seed] [rounds]"""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_regex_fuzz_cpu import run_round

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 50
t0, n, nbad = time.time(), 0, 0
for rd in range(rounds):
    k, bad = run_round(seed0 * 100000 + rd)
    n += k
    for b in bad[:2]:
        print("MISMATCH", seed0, rd, b)
    nbad += bool(bad)
print("rounds", rounds, "patterns", n, "rounds with mismatches", nbad, "seconds", round(time.time() - t0, 1))
