"""The C++ facade (include/fbstab/*.h) against the reference's own solver-level
tests, restated in tests/cpp/facade_tests.cc (gtest is not in this image)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "facade_tests")


def _build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.join(ROOT, "fbstab_b200")
    assert os.path.exists(os.path.join(libdir, "libfbstab_b200.so")), \
        "build the engine first: python -c 'import __graft_entry__ as g; g.build()'"
    subprocess.check_call(
        ["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror",
         "-I" + os.path.join(ROOT, "include"),
         os.path.join(ROOT, "tests", "cpp", "facade_tests.cc"),
         "-L" + libdir, "-lfbstab_b200", "-Wl,-rpath," + libdir, "-o", BIN])


def _run(args):
    p = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-2000:]
    return p.stdout


def test_facade_host_cases():
    """Types, option clamps, size validation, loud failure without a device."""
    _build()
    out = _run(["--host"])
    assert "0 failed" in out


@pytest.mark.gpu
def test_facade_reference_cases():
    """All ten reference solver tests + batched entries through the facade."""
    _build()
    out = _run([])
    assert "0 failed" in out
    assert "FBstabMpc.CopolymerizationReactor" in out
