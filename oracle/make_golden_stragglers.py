"""TEST INFRASTRUCTURE - adds to tests/golden/dense_results.json the UNMODIFIED reference's `interior` results
(replayed with its own functions, oracle/ref_harness.py) on the two generator LPs that matter for the batched
solver's straggler handling: seed 16893 (trapped by the GPU's four-pass iteration) and seed
31186 (trapped by the normal equations on the CPU).  Existing entries are left untouched.

    python oracle/make_golden_stragglers.py        (build container only: needs /root/reference)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "dense_results.json")


def main():
    from oracle import ipm_oracle as orc
    from oracle import ref_harness as rh
    rh.load_reference()
    d = json.load(open(GOLD))
    for seed in (16893, 31186):
        A, b, c = orc.synthetic_dense_lp(256, 512, seed)
        t0 = time.time()
        r = rh.replay_interior_dense(A, b, c, tol=1e-8)
        d["synthetic_256x512_seed%d" % seed] = dict(k=r["k"], obj=r["obj"], wall_s=round(time.time() - t0, 3))
        print(seed, d["synthetic_256x512_seed%d" % seed], flush=True)
    json.dump(d, open(GOLD, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
