// main() of the gtest stand-in (oracle/ref/gtest/gtest.h): runs every TEST registered by the reference's
// unmodified test sources and returns non-zero if any assertion failed.  TEST INFRASTRUCTURE.
#include <gtest/gtest.h>

int main() { return gtest_shim::runAll(); }
