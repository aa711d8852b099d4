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


# ---- the REFERENCE'S OWN unit tests against the facade and the engine -------------------
REFERENCE = "/root/reference"
REF_BIN = os.path.join(ROOT, "build", "reference_unit_tests_on_engine")


def build_reference_tests(reference=REFERENCE):
    """Compiles the reference's own live unit tests -- fbstab/test/fbstab_dense_unit_tests.cc
    and fbstab_mpc_unit_tests.cc, UNMODIFIED, read where they lie -- against this
    repository's facade (include/fbstab/*.h with its stand-in Eigen types, -DFBSTAB_NO_EIGEN),
    tests/cpp/ref_compat (forwarding headers for the three include paths the tests name) and
    oracle/gtest_shim, and links them with libfbstab_b200.so.  Only where the reference tree
    exists; the binary (build/ is git-ignored but travels to the GPU box) is what the GPU
    test runs.  Returns its path, or None when there is neither a tree nor a prebuilt
    binary."""
    tests = [os.path.join(reference, "fbstab", "test", f)
             for f in ("fbstab_dense_unit_tests.cc", "fbstab_mpc_unit_tests.cc")]
    if all(os.path.exists(t) for t in tests):
        libdir = os.path.join(ROOT, "fbstab_b200")
        assert os.path.exists(os.path.join(libdir, "libfbstab_b200.so")), \
            "build the engine first: python -c 'import __graft_entry__ as g; g.build()'"
        os.makedirs(os.path.dirname(REF_BIN), exist_ok=True)
        subprocess.check_call(
            ["g++", "-std=c++14", "-O1", "-DFBSTAB_NO_EIGEN",
             "-I" + os.path.join(ROOT, "include"),
             "-I" + os.path.join(ROOT, "tests", "cpp", "ref_compat"),
             "-I" + os.path.join(ROOT, "oracle", "gtest_shim"),
             os.path.join(ROOT, "tests", "cpp", "ref_tests_main.cc")] + tests +
            ["-L" + libdir, "-lfbstab_b200", "-Wl,-rpath,$ORIGIN/../fbstab_b200", "-o", REF_BIN])
    return REF_BIN if os.path.exists(REF_BIN) else None


def test_the_references_own_unit_tests_compile_against_the_facade():
    """CPU: the two test files of the reference compile, unmodified, against the facade (the
    API surface is the reference's), and without a GPU every one of them fails loudly -- the
    engine has no CPU path."""
    import torch
    binary = build_reference_tests()
    if binary is None:
        pytest.skip("no reference tree and no prebuilt build/reference_unit_tests_on_engine")
    if not torch.cuda.is_available():
        p = subprocess.run([binary], capture_output=True, text=True, timeout=600)
        assert p.returncode != 0 and "10 tests, 10 failed" in p.stdout, p.stdout[-2000:]
        assert "no CUDA device available" in p.stdout


@pytest.mark.gpu
def test_the_references_own_unit_tests_pass_on_the_engine():
    """GPU: fbstab/test/fbstab_dense_unit_tests.cc and fbstab_mpc_unit_tests.cc of the
    reference -- its ten live tests, unmodified -- pass against the facade with the CUDA
    engine behind it (profiles/r2_reference_unit_tests_on_engine.txt)."""
    binary = build_reference_tests()
    if binary is None:
        pytest.skip("no reference tree and no prebuilt build/reference_unit_tests_on_engine")
    p = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "10 tests, 0 failed" in p.stdout


def test_facade_compiles_in_its_eigen_branch():
    """With an <Eigen/Dense> on the include path the facade takes its FBSTAB_HAVE_EIGEN branch
    (Eigen's own types instead of the stand-ins of linalg.h).  Eigen is not installed here;
    oracle/eigen_shim provides the header, which is enough to compile that branch -- the
    facade's own tests and, where the reference tree exists, the reference's unit tests."""
    srcs = [os.path.join(ROOT, "tests", "cpp", "facade_tests.cc")]
    ref = [os.path.join(REFERENCE, "fbstab", "test", f)
           for f in ("fbstab_dense_unit_tests.cc", "fbstab_mpc_unit_tests.cc")]
    if all(os.path.exists(t) for t in ref):
        srcs += ref
    for src in srcs:
        subprocess.check_call(
            ["g++", "-std=c++14", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"),
             "-I" + os.path.join(ROOT, "oracle", "eigen_shim"),
             "-I" + os.path.join(ROOT, "tests", "cpp", "ref_compat"),
             "-I" + os.path.join(ROOT, "oracle", "gtest_shim"), src])
