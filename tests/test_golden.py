"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit, and so does the second, independent restatement of the path
(tests/independent.py) - the vectors are what two separately written statements of the reference agree on.
GPU: the device path matches them - f64 cursors bit-exact,
mixed samples within 1e-5 * max(|ref|, RMS) (summation order is the only freedom)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from scenarios import SCENARIOS, DeviceBackend, OracleBackend  # noqa: E402


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_reproduces_golden(name, oracle):
    gold = load(name)
    res = SCENARIOS[name](OracleBackend())
    assert set(res) == set(gold)
    for k in gold:
        np.testing.assert_array_equal(res[k], gold[k], err_msg=f"{name}:{k}")


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_independent_restatement_reproduces_golden(name):
    from independent import IndependentBackend

    gold = load(name)
    res = SCENARIOS[name](IndependentBackend())
    assert set(res) == set(gold)
    for k in gold:
        np.testing.assert_array_equal(np.asarray(res[k]).view(np.uint32 if k.startswith("out") else np.uint64),
                                      gold[k].view(np.uint32 if k.startswith("out") else np.uint64), err_msg=f"{name}:{k}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_device_matches_golden(name):
    gold = load(name)
    res = SCENARIOS[name](DeviceBackend())
    for k in gold:
        if k.startswith("t"):
            np.testing.assert_array_equal(res[k], gold[k], err_msg=f"{name}:{k} (f64 cursors must be bit-exact)")
        else:
            ref = gold[k].astype(np.float64)
            rms = float(np.sqrt(np.mean(ref ** 2)))
            tol = 1e-5 * np.maximum(np.abs(ref), rms) + 1e-30
            bad = np.abs(res[k].astype(np.float64) - ref) > tol
            assert not bad.any(), f"{name}:{k}: {bad.sum()} of {bad.size} samples outside 1e-5"
