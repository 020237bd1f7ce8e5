import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
# larger trees are committed gzip-compressed (<name>.pb.gz) and unpacked on first use into tests/golden/_unpacked/
GOLDEN_CASES = sorted([os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.pb"))] +
                      [os.path.basename(p)[:-6] for p in glob.glob(os.path.join(GOLDEN, "*.pb.gz"))])


def golden_proto_path(name):
    """path of the proto file of a golden case (unpacking <name>.pb.gz if that is how it is stored)"""
    plain = os.path.join(GOLDEN, name + ".pb")
    if os.path.exists(plain):
        return plain
    import gzip
    out = os.path.join(GOLDEN, "_unpacked", name + ".pb")
    if not os.path.exists(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tmp = out + f".{os.getpid()}.tmp"
        with gzip.open(plain + ".gz", "rb") as f, open(tmp, "wb") as o:
            o.write(f.read())
        os.replace(tmp, out)
    return out

# fp64 value parity bar from BASELINE.json north_star: 1e-12 relative.  The reference is not
# bit-reproducible against itself (summation order = hash iteration order), and a
# relative bound is meaningless where the interpolant crosses zero, so the bound is relative to the
# magnitude of the sum's terms:  |a-b| <= 1e-12 * max(|ref|, sum_k |w_k| max_j|v_k[j]|).
VALUE_RTOL = 1e-12


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def assert_values_close(got, ref, scale, rtol=VALUE_RTOL):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    if got.ndim == 2:
        scale = scale[:, None]
    bound = rtol * np.maximum(np.abs(ref), scale)
    bad = ~(np.abs(got - ref) <= bound)
    assert not bad.any(), f"{int(bad.sum())} values beyond {rtol:g} relative; worst abs diff {np.nanmax(np.abs(got - ref)):.3e}"


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return golden_proto_path(name), np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the oracle are compiled once per session (nvcc cross-compiles without a GPU)."""
    from locointerpolate_b200 import build
    from oracle import oracleapi
    build.build()
    oracleapi.build()
