// Minimal stand-in for <gtest/gtest.h> (GoogleTest is fetched from the network by the reference's
// build, cmake/gtest.cmake:5-6, and is not available here).  Just enough to compile and run the
// reference's tests/*.cpp unchanged: TEST registration, InitGoogleTest, RUN_ALL_TESTS, and the
// EXPECT/ASSERT macros (the reference's tests use none).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <exception>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

namespace testing {
struct Registry {
  struct Item { std::string name; std::function<void()> fn; };
  static std::vector<Item>& items() { static std::vector<Item> v; return v; }
  static int& failures() { static int f = 0; return f; }
};
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<void()> fn) {
    Registry::items().push_back({std::string(suite) + "." + name, std::move(fn)});
  }
};
inline void InitGoogleTest(int*, char**) {}
inline int RunAll() {
  int failed = 0;
  for (auto& it : Registry::items()) {
    std::cout << "[ RUN      ] " << it.name << std::endl;
    const int before = Registry::failures();
    try {
      it.fn();
    } catch (std::exception const& e) {
      std::cout << "exception: " << e.what() << std::endl;
      ++Registry::failures();
    }
    const bool ok = Registry::failures() == before;
    std::cout << (ok ? "[       OK ] " : "[  FAILED  ] ") << it.name << std::endl;
    failed += !ok;
  }
  std::cout << "[==========] " << Registry::items().size() << " tests ran, " << failed << " failed" << std::endl;
  return failed ? 1 : 0;
}
}  // namespace testing

#define TEST(suite, name)                                                                    \
  static void suite##_##name##_body();                                                       \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);    \
  static void suite##_##name##_body()
#define RUN_ALL_TESTS() ::testing::RunAll()
#define B200_GT_CHECK(cond, text)                                                            \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      std::cout << __FILE__ << ":" << __LINE__ << ": Failure: " << text << std::endl;        \
      ++::testing::Registry::failures();                                                     \
    }                                                                                        \
  } while (0)
#define EXPECT_TRUE(c) B200_GT_CHECK((c), #c)
#define EXPECT_FALSE(c) B200_GT_CHECK(!(c), "!(" #c ")")
#define EXPECT_EQ(a, b) B200_GT_CHECK((a) == (b), #a " == " #b)
#define EXPECT_NE(a, b) B200_GT_CHECK((a) != (b), #a " != " #b)
#define EXPECT_LT(a, b) B200_GT_CHECK((a) < (b), #a " < " #b)
#define EXPECT_LE(a, b) B200_GT_CHECK((a) <= (b), #a " <= " #b)
#define EXPECT_GT(a, b) B200_GT_CHECK((a) > (b), #a " > " #b)
#define EXPECT_NEAR(a, b, tol) B200_GT_CHECK(std::fabs((a) - (b)) <= (tol), #a " ~ " #b)
#define EXPECT_THROW(stmt, ex)                                                               \
  do {                                                                                       \
    bool caught = false;                                                                     \
    try { stmt; } catch (ex const&) { caught = true; }                                       \
    B200_GT_CHECK(caught, #stmt " throws " #ex);                                             \
  } while (0)
#define ASSERT_TRUE EXPECT_TRUE
#define ASSERT_EQ EXPECT_EQ
#define ASSERT_NEAR EXPECT_NEAR
