// Minimal stand-in for <gtest/gtest.h> (googletest is not installed in this image; the reference fetches it at configure time,
// CMakeLists.txt:46-54): just enough of TEST_F / EXPECT_* to compile and run the reference's OWN test files, unmodified, against
// the drop-in host layer (tests/test_host_drivers.py::test_reference_lattice_tests_run_against_the_host_layer).
#ifndef MGMC_REF_COMPAT_GTEST_H
#define MGMC_REF_COMPAT_GTEST_H
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <utility>
#include <vector>

namespace testing {
class Test {
 public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};
struct Registry {
  std::vector<std::pair<std::string, std::function<Test *()>>> tests;
  int failures_in_current = 0;
  static Registry &get() {
    static Registry r;
    return r;
  }
};
inline int register_test(const char *name, std::function<Test *()> make) {
  Registry::get().tests.push_back({name, std::move(make)});
  return 0;
}
inline void report_failure(const char *file, int line, const char *what) {
  std::printf("%s:%d: Failure: %s\n", file, line, what);
  Registry::get().failures_in_current++;
}
// filter: run only the tests whose name contains this substring (like --gtest_filter=*x*); empty = all
inline int run_all_tests(const std::string &filter = "") {
  int failed = 0, ran = 0;
  for (auto &t : Registry::get().tests) {
    if (!filter.empty() && t.first.find(filter) == std::string::npos) continue;
    ++ran;
    Registry::get().failures_in_current = 0;
    Test *obj = t.second();
    obj->SetUp();
    obj->TestBody();
    obj->TearDown();
    delete obj;
    const bool ok = Registry::get().failures_in_current == 0;
    std::printf("[ %s ] %s\n", ok ? "      OK" : "  FAILED", t.first.c_str());
    failed += ok ? 0 : 1;
  }
  std::printf("[  %s  ] %d tests, %d failed\n", failed ? "FAILED" : "PASSED", ran, failed);
  return failed ? 1 : 0;
}
}  // namespace testing

#define TEST_F(fixture, name)                                                                                        \
  class fixture##_##name##_Test : public fixture {                                                                   \
   public:                                                                                                           \
    void TestBody() override;                                                                                        \
    static int registered_;                                                                                          \
  };                                                                                                                 \
  int fixture##_##name##_Test::registered_ =                                                                         \
      ::testing::register_test(#fixture "." #name, [] { return static_cast<::testing::Test *>(new fixture##_##name##_Test); }); \
  void fixture##_##name##_Test::TestBody()

#define EXPECT_EQ(a, b)                                                              \
  do {                                                                               \
    if (!((a) == (b))) ::testing::report_failure(__FILE__, __LINE__, #a " == " #b); \
  } while (0)
#define EXPECT_TRUE(a)                                                    \
  do {                                                                    \
    if (!(a)) ::testing::report_failure(__FILE__, __LINE__, #a);          \
  } while (0)
#define EXPECT_NEAR(a, b, tol)                                                                      \
  do {                                                                                              \
    if (!(std::fabs((a) - (b)) <= (tol))) ::testing::report_failure(__FILE__, __LINE__, #a " ~ " #b); \
  } while (0)
#endif
