"""Compiles tests/cpp/test_facade.cpp (the reference's test scenarios against include/loco_lc.hpp) with g++."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_facade.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade")
LIBDIR = os.path.join(ROOT, "locointerpolate_b200")


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(ROOT, "include", "loco_lc.hpp"), os.path.join(ROOT, "include", "loco_grid.hpp"),
            os.path.join(ROOT, "include", "loco_b200.h")]
    if not force and os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L" + LIBDIR, "-lloco_b200", "-Wl,-rpath,$ORIGIN/../../locointerpolate_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return EXE
