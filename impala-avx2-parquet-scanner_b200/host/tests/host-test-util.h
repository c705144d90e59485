// Minimal gtest-flavoured assertions for the host-side C++ tests (no gtest in this image).
#ifndef FLE_HOST_TEST_UTIL_H
#define FLE_HOST_TEST_UTIL_H
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

struct TestRegistry {
  typedef void (*Fn)();
  std::vector<std::pair<std::string, Fn> > tests;
  int failures = 0;
  static TestRegistry& get() { static TestRegistry r; return r; }
};
struct TestRegistrar {
  TestRegistrar(const char* name, TestRegistry::Fn fn) { TestRegistry::get().tests.push_back(std::make_pair(name, fn)); }
};
#define TEST(suite, name)                                                    \
  static void suite##_##name();                                              \
  static TestRegistrar reg_##suite##_##name(#suite "." #name, suite##_##name); \
  static void suite##_##name()
#define EXPECT_TRUE(x) do { if (!(x)) { ++TestRegistry::get().failures; std::fprintf(stderr, "%s:%d: EXPECT_TRUE(%s) failed\n", __FILE__, __LINE__, #x); } } while (0)
#define EXPECT_EQ(a, b) do { if (!((a) == (b))) { ++TestRegistry::get().failures; std::fprintf(stderr, "%s:%d: EXPECT_EQ(%s, %s) failed: %lld vs %lld\n", __FILE__, __LINE__, #a, #b, (long long)(a), (long long)(b)); } } while (0)
#define EXPECT_LE(a, b) EXPECT_TRUE((a) <= (b))
#define ASSERT_TRUE(x) do { if (!(x)) { ++TestRegistry::get().failures; std::fprintf(stderr, "%s:%d: ASSERT_TRUE(%s) failed\n", __FILE__, __LINE__, #x); return; } } while (0)

inline int RunAllTests() {
  TestRegistry& r = TestRegistry::get();
  for (size_t i = 0; i < r.tests.size(); ++i) {
    int before = r.failures;
    r.tests[i].second();
    std::printf("[%s] %s\n", r.failures == before ? "  OK  " : "FAILED", r.tests[i].first.c_str());
  }
  std::printf("%d test(s), %d failure(s)\n", (int)r.tests.size(), r.failures);
  return r.failures == 0 ? 0 : 1;
}
#endif
