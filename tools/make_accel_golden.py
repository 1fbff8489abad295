#!/usr/bin/env python3
"""Write tests/golden/accel_hashes.json: SHA-256 of the acceleration-structure arrays that the
REFERENCE's bvh.cpp (compiled in place, oracle/_ref) produces for the fixed test scenes.
Run in the authoring container (needs oracle/_ref/libgdpt_refbvh.so).  The hashes let a
checkout without the reference still pin the product builder to the reference's bytes."""
import hashlib
import json
import os
import sys

import numpy as np

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from accel_cases import CASES, reference_buffers  # noqa: E402


def main():
    out = {}
    for name, make in CASES.items():
        sc = make()
        ref = reference_buffers(sc)
        out[name] = {k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() for k, v in ref.items()}
        out[name]["counts"] = {k: int(len(v)) for k, v in ref.items()}
        print(name, out[name]["counts"])
    path = os.path.join(REPO, "tests", "golden", "accel_hashes.json")
    json.dump({"_source": "reference src/bvh/bvh.cpp compiled in place via oracle/ref_bridge.cpp; tools/make_accel_golden.py",
               "cases": out}, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
