#!/usr/bin/env python3
"""Generates tests/golden/hotpath_vectors.json from oracle/pyref.py (pure big-int restatement).

The reference holds no golden vectors for this path and cannot be run here (Rust,
un-vendored arkworks), so these vectors pin OUR two restatements and the CUDA path to one another;
bit-exactness against the Rust binary stays asserted-by-construction ("parity unpinned").
Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyref as R  # noqa: E402


def main():
    rng = random.Random(20261018)
    out = {"generator": "tests/golden/make_golden.py", "seed": 20261018, "groups": {}, "phase1": {}}
    for cname, cv in R.CURVES.items():
        for g in (cv.g1, cv.g2):
            n = 6
            pts = [g.mul(g.gen, rng.randrange(1, g.r)) for _ in range(n)]
            pts[4] = None
            tau, coeff = rng.randrange(cv.r), rng.randrange(cv.r)
            first = rng.randrange(1 << 40)
            powers = R.generate_powers_of_tau(cv, tau, first, first + n)
            rho = [rng.randrange(cv.r) for _ in range(n)]
            s, sx = R.merge_pairs(g, pts[:-1], pts[1:], rho[:-1])
            out["groups"][g.name] = {
                "generator_uncompressed": g.encode(g.gen, False).hex(),
                "generator_compressed": g.encode(g.gen, True).hex(),
                "infinity_compressed": g.encode(None, True).hex(),
                "infinity_uncompressed": g.encode(None, False).hex(),
                "in_uncompressed": g.write_batch(pts, False).hex(),
                "in_compressed": g.write_batch(pts, True).hex(),
                "tau": hex(tau), "coeff": hex(coeff), "first_power": first,
                "powers": [hex(p) for p in powers],
                "out_plain_compressed": g.write_batch(R.batch_exp(g, pts, powers), True).hex(),
                "out_coeff_uncompressed": g.write_batch(R.batch_exp(g, pts, powers, coeff), False).hex(),
                "rho": [hex(r) for r in rho[:-1]],
                "power_pairs_s": g.encode(s, False).hex(),
                "power_pairs_sx": g.encode(sx, False).hex(),
            }
    # Phase1::computation transcript: new (generators) -> contribute #1 -> contribute #2 (compressed)
    for cname, cv in R.CURVES.items():
        p = R.Phase1Parameters(cv, 2, 3)
        keys = [[rng.randrange(1, cv.r) for _ in range(3)] for _ in range(2)]
        acc0 = bytes(R.phase1_initialization(p, False))
        acc1 = bytes(R.phase1_computation(p, acc0, False, False, R.NO, *keys[0]))
        resp = bytes(R.phase1_computation(p, acc1, False, True, R.NO, *keys[1]))
        out["phase1"][cname] = {"power": 2, "batch_size": 3, "keys": [[hex(k) for k in ks] for ks in keys],
                                "windows": R.iter_chunk(p),
                                "accumulator_size": p.accumulator_size, "contribution_size": p.contribution_size,
                                "challenge1": acc1.hex(), "response2": resp.hex()}
    with open(os.path.join(HERE, "hotpath_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "hotpath_vectors.json"))


if __name__ == "__main__":
    main()
