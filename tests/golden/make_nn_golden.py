#!/usr/bin/env python
"""Generates tests/golden/nn_golden.npz: outputs of the REFERENCE's own naive operators (oracle/_ref, compiled from
/root/reference by oracle/build_ref.sh) for the case tables of tests/_nn.py, on seeded inputs.  The CPU suite pins
oracle/nn_oracle.c against these without needing the reference.   python tests/golden/make_nn_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import _nn as N  # noqa: E402
from tests.test_nn_oracle import all_cases  # noqa: E402

out = {}
for key, run in all_cases():
    out[key] = run("ref")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nn_golden.npz"), **out)
print("wrote", len(out), "golden outputs")
