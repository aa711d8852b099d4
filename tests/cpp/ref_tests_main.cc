// Runs the REFERENCE'S OWN unit tests (fbstab/test/fbstab_dense_unit_tests.cc and
// fbstab_mpc_unit_tests.cc, compiled unmodified from /root/reference) against the facade of
// this repository and, behind it, the CUDA engine: the drop-in claim as an executable.
// Built by tests/test_cpp_facade.py::build_reference_tests where the reference tree exists.
#include <gtest/gtest.h>

int main() { return gtest_shim::RunAll(); }
