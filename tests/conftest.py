import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    return sorted(os.path.basename(p)[len("farneback_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "farneback_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, "farneback_%s.npz" % name))
    d = {k: z[k] for k in z.files}
    p = d["params"]
    d["kw"] = dict(pyr_scale=float(p[0]), levels=int(p[1]), winsize=int(p[2]), iterations=int(p[3]),
                   poly_n=int(p[4]), poly_sigma=float(p[5]), flags=int(p[6]))
    return d


def epe(a, b):
    d = np.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum(-1))
    return float(d.mean()), float(d.max())


# Tolerances stated by BASELINE.json north_star.
EPE_MEAN_TOL = 1e-3
EPE_MAX_TOL = 1e-2


@pytest.fixture(scope="session")
def oracle():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle
