"""The C++ facade (include/loco_lc.hpp: the reference's class surface over the C ABI) driven by the reference's own
test scenarios (tests/cpp/test_facade.cpp, after TestExamples/TestExamples.cpp)."""
import os
import subprocess

import numpy as np
import pytest

import facade_build
import protoschema
from conftest import GOLDEN, assert_values_close
from oracle import oracleapi


def _run(args):
    exe = facade_build.build()
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)


def test_facade_compiles_and_host_side_works(tmp_path):
    """No GPU needed: the facade compiles with -Wall -Werror, links against the C-ABI library, the proto round trip
    (binary and ascii) preserves the model, and evaluation without a device fails loudly."""
    r = _run([os.path.join(GOLDEN, "cub2d_prototest.pb"), str(tmp_path / "dump.txt"), "--host-only"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "NOT verified" not in r.stdout and "PASSED" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cub2d_prototest", "lin2d_prototest", "cub2d_adapt6"])
def test_reference_scenarios_through_facade(tmp_path, name):
    pb = os.path.join(GOLDEN, name + ".pb")
    dump = str(tmp_path / "dump.txt")
    r = _run([pb, dump])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "NOT verified" not in r.stdout and "PASSED" in r.stdout
    # the 11x11 sweep the program evaluated one query at a time, against the oracle
    t = np.loadtxt(dump)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    o_out, _, o_status, scale = orc.eval(t[:, :2])
    assert (o_status == 0).all()
    assert_values_close(t[:, 2], o_out, scale)
