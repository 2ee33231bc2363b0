// Minimal GoogleTest stand-in (gtest is not installed and there is no network): just enough of
// TEST / ASSERT_* for /root/reference/tests/*.cpp to compile UNMODIFIED.  A failed assertion prints the
// location and makes main() return non-zero.  TEST INFRASTRUCTURE (oracle/), not product.
#pragma once

#include <cstdio>
#include <exception>
#include <sstream>
#include <vector>

namespace gtest_shim {
struct Case {
  const char* suite;
  const char* name;
  void (*fn)(bool&);
};
inline std::vector<Case>& cases() {
  static std::vector<Case> all;
  return all;
}
struct Registrar {
  Registrar(const char* suite, const char* name, void (*fn)(bool&)) { cases().push_back({suite, name, fn}); }
};
struct Message {
  std::ostringstream text;
  template <class T>
  Message& operator<<(const T& v) {
    text << v;
    return *this;
  }
};
struct Failure {
  bool& ok;
  const char* file;
  int line;
  const char* what;
  void operator=(const Message& m) const {
    std::printf("  %s:%d: assertion failed: %s %s\n", file, line, what, m.text.str().c_str());
    ok = false;
  }
};
inline int runAll() {
  int failed = 0;
  for (const Case& c : cases()) {
    bool ok = true;
    std::printf("[ RUN      ] %s.%s\n", c.suite, c.name);
    try {
      c.fn(ok);
    } catch (const std::exception& e) {
      std::printf("  uncaught exception: %s\n", e.what());
      ok = false;
    } catch (...) {
      std::printf("  uncaught exception\n");
      ok = false;
    }
    std::printf("[ %s ] %s.%s\n", ok ? "      OK" : " FAILED ", c.suite, c.name);
    failed += ok ? 0 : 1;
  }
  std::printf("%zu tests, %d failed\n", cases().size(), failed);
  return failed == 0 ? 0 : 1;
}
}  // namespace gtest_shim

#define TEST(suite, name)                                                                   \
  static void suite##_##name##_body(bool& gtest_ok__);                                      \
  static gtest_shim::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);  \
  static void suite##_##name##_body(bool& gtest_ok__)

// `ASSERT_X(...) << "message"` works like gtest's: the streamed text is printed on failure
#define GTEST_SHIM_CHECK(cond, text) \
  if (cond)                          \
    ;                                \
  else                               \
    return gtest_shim::Failure {gtest_ok__, __FILE__, __LINE__, text} = gtest_shim::Message()

#define ASSERT_TRUE(c) GTEST_SHIM_CHECK((c), #c)
#define ASSERT_FALSE(c) GTEST_SHIM_CHECK(!(c), "!(" #c ")")
#define ASSERT_EQ(a, b) GTEST_SHIM_CHECK((a) == (b), #a " == " #b)
#define ASSERT_NE(a, b) GTEST_SHIM_CHECK((a) != (b), #a " != " #b)
#define ASSERT_LT(a, b) GTEST_SHIM_CHECK((a) < (b), #a " < " #b)
#define ASSERT_LE(a, b) GTEST_SHIM_CHECK((a) <= (b), #a " <= " #b)
#define ASSERT_GT(a, b) GTEST_SHIM_CHECK((a) > (b), #a " > " #b)
#define ASSERT_GE(a, b) GTEST_SHIM_CHECK((a) >= (b), #a " >= " #b)
#define EXPECT_TRUE ASSERT_TRUE
#define EXPECT_EQ ASSERT_EQ
#define EXPECT_LT ASSERT_LT

