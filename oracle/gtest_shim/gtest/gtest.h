// oracle/gtest_shim/gtest/gtest.h -- the handful of googletest macros the reference's own
// unit tests use (fbstab/test/*.cc), so that those files compile and run UNMODIFIED where
// googletest is not installed (TEST INFRASTRUCTURE; `make -C oracle _ref_tests`).
#pragma once

#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace gtest_shim {
struct Case {
  std::string name;
  std::function<void()> body;
};
inline std::vector<Case>& cases() {
  static std::vector<Case> c;
  return c;
}
inline int& failures() {
  static int f = 0;
  return f;
}
struct Fatal {};
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<void()> body) {
    cases().push_back({std::string(suite) + "." + name, std::move(body)});
  }
};
inline void Report(const char* file, int line, const char* what) {
  std::printf("%s:%d: Failure: %s\n", file, line, what);
  failures()++;
}
inline int RunAll() {
  int bad = 0;
  for (auto& c : cases()) {
    const int before = failures();
    try {
      c.body();
    } catch (const Fatal&) {
    } catch (const std::exception& e) {
      std::printf("uncaught exception in %s: %s\n", c.name.c_str(), e.what());
      failures()++;
    }
    const bool ok = failures() == before;
    std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", c.name.c_str());
    bad += ok ? 0 : 1;
  }
  std::printf("%d tests, %d failed\n", (int)cases().size(), bad);
  return bad ? 1 : 0;
}
}  // namespace gtest_shim

#define GTEST_TEST(suite, name)                                                        \
  static void gtest_shim_##suite##_##name();                                           \
  static gtest_shim::Registrar gtest_shim_reg_##suite##_##name(#suite, #name,         \
                                                               gtest_shim_##suite##_##name); \
  static void gtest_shim_##suite##_##name()
#define TEST(suite, name) GTEST_TEST(suite, name)

#define GTEST_SHIM_CHECK(cond, text, fatal)                    \
  do {                                                         \
    if (!(cond)) {                                             \
      gtest_shim::Report(__FILE__, __LINE__, text);            \
      if (fatal) throw gtest_shim::Fatal();                    \
    }                                                          \
  } while (0)
#define EXPECT_TRUE(c) GTEST_SHIM_CHECK((c), "EXPECT_TRUE(" #c ")", false)
#define ASSERT_TRUE(c) GTEST_SHIM_CHECK((c), "ASSERT_TRUE(" #c ")", true)
#define EXPECT_FALSE(c) GTEST_SHIM_CHECK(!(c), "EXPECT_FALSE(" #c ")", false)
#define ASSERT_FALSE(c) GTEST_SHIM_CHECK(!(c), "ASSERT_FALSE(" #c ")", true)
#define EXPECT_EQ(a, b) GTEST_SHIM_CHECK((a) == (b), "EXPECT_EQ(" #a ", " #b ")", false)
#define ASSERT_EQ(a, b) GTEST_SHIM_CHECK((a) == (b), "ASSERT_EQ(" #a ", " #b ")", true)
#define EXPECT_NE(a, b) GTEST_SHIM_CHECK((a) != (b), "EXPECT_NE(" #a ", " #b ")", false)
#define EXPECT_LE(a, b) GTEST_SHIM_CHECK((a) <= (b), "EXPECT_LE(" #a ", " #b ")", false)
#define ASSERT_LE(a, b) GTEST_SHIM_CHECK((a) <= (b), "ASSERT_LE(" #a ", " #b ")", true)
#define EXPECT_LT(a, b) GTEST_SHIM_CHECK((a) < (b), "EXPECT_LT(" #a ", " #b ")", false)
#define EXPECT_GE(a, b) GTEST_SHIM_CHECK((a) >= (b), "EXPECT_GE(" #a ", " #b ")", false)
#define EXPECT_GT(a, b) GTEST_SHIM_CHECK((a) > (b), "EXPECT_GT(" #a ", " #b ")", false)
#define EXPECT_NEAR(a, b, tol) \
  GTEST_SHIM_CHECK(std::fabs((a) - (b)) <= (tol), "EXPECT_NEAR(" #a ", " #b ", " #tol ")", false)
#define ASSERT_NEAR(a, b, tol) \
  GTEST_SHIM_CHECK(std::fabs((a) - (b)) <= (tol), "ASSERT_NEAR(" #a ", " #b ", " #tol ")", true)
#define EXPECT_DOUBLE_EQ(a, b) EXPECT_NEAR(a, b, 4e-16 * std::fmax(std::fabs(a), std::fabs(b)))
#define EXPECT_NO_THROW(stmt)                                              \
  do {                                                                     \
    try {                                                                  \
      stmt;                                                                \
    } catch (...) {                                                        \
      gtest_shim::Report(__FILE__, __LINE__, "EXPECT_NO_THROW(" #stmt ")"); \
    }                                                                      \
  } while (0)
#define EXPECT_ANY_THROW(stmt)                                                 \
  do {                                                                         \
    bool gtest_shim_threw = false;                                             \
    try {                                                                      \
      stmt;                                                                    \
    } catch (...) {                                                            \
      gtest_shim_threw = true;                                                 \
    }                                                                          \
    if (!gtest_shim_threw) gtest_shim::Report(__FILE__, __LINE__, "EXPECT_ANY_THROW(" #stmt ")"); \
  } while (0)
#define EXPECT_THROW(stmt, type)                                                \
  do {                                                                          \
    bool gtest_shim_threw = false;                                              \
    try {                                                                       \
      stmt;                                                                     \
    } catch (const type&) {                                                     \
      gtest_shim_threw = true;                                                  \
    } catch (...) {                                                             \
    }                                                                           \
    if (!gtest_shim_threw) gtest_shim::Report(__FILE__, __LINE__, "EXPECT_THROW(" #stmt ", " #type ")"); \
  } while (0)
