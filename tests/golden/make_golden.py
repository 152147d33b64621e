"""Regenerates tests/golden/*.npz from the CPU oracle:  python tests/golden/make_golden.py

The reference (Rust) cannot run in this image, so the vectors are the oracle's - which is itself pinned on
the reference's unit-test vectors (tests/test_oracle_kat.py). They freeze today's oracle results so that a
later change to the oracle or to the device path that alters any output is caught on both sides."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from scenarios import SCENARIOS, OracleBackend  # noqa: E402

if __name__ == "__main__":
    for name, fn in SCENARIOS.items():
        res = fn(OracleBackend())
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **res)
        print(f"{path}: {sum(v.nbytes for v in res.values())} bytes of arrays, {os.path.getsize(path)} on disk")
