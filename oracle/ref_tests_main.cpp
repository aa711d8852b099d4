// Runs the reference's own unit tests (fbstab/test/*.cc, compiled unmodified against
// gtest_shim/ and eigen_shim/) -- `make -C oracle _ref_tests`.  TEST INFRASTRUCTURE.
#include <gtest/gtest.h>

int main() { return gtest_shim::RunAll(); }
