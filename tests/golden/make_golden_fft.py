#!/usr/bin/env python3
"""Generates tests/golden/prepare_phase2_vectors.json from oracle/pyref.py: Groth16Params::new + ::write
(setup-utils/src/groth16_utils.rs:81-168) of the power-2 accumulators pinned in hotpath_vectors.json
("challenge1"), computed with the O(n^2) *definition* of the inverse DFT (pyref.group_ifft).

The reference holds no golden vectors for this path and cannot be run here; these vectors pin
the oracle's fast algorithm, the C++ oracle route and the CUDA path to the definition.  The root of unity is
ark-ff 0.4's GENERATOR^t (pyref.FR_GENERATOR) — recalled, not verifiable offline: parity unpinned.
Run:  python tests/golden/make_golden_fft.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyref as R  # noqa: E402


def main():
    hot = json.load(open(os.path.join(HERE, "hotpath_vectors.json")))
    out = {"generator": "tests/golden/make_golden_fft.py", "curves": {}}
    for cname, cv in R.CURVES.items():
        v = hot["phase1"][cname]
        p = R.Phase1Parameters(cv, v["power"], v["batch_size"])
        acc = bytes.fromhex(v["challenge1"])
        ent = {"power": v["power"], "root_of_unity_4": hex(R.get_root_of_unity(cv.r, 4)),
               "two_adic_root_of_unity": hex(R.get_root_of_unity(cv.r, 1 << R.two_adicity(cv.r))), "params": {}}
        for phase2_size in (4, 2):
            ent["params"][str(phase2_size)] = {
                "compressed": R.groth16_params_new(p, acc, False, phase2_size, True, ifft=R.group_ifft).hex(),
                "uncompressed_blake2b": __import__("hashlib").blake2b(
                    R.groth16_params_new(p, acc, False, phase2_size, False, ifft=R.group_ifft)).hexdigest(),
            }
        out["curves"][cname] = ent
    with open(os.path.join(HERE, "prepare_phase2_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote prepare_phase2_vectors.json")


if __name__ == "__main__":
    main()
