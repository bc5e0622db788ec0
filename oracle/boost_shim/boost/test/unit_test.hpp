// TEST INFRASTRUCTURE: a minimal stand-in for Boost.Test (absent from this image) so that the
// reference's own fake_spectra/test.cpp can be compiled UNMODIFIED by oracle/build.py --selftest.
// Provides only the four macros that file uses, a registry and a main().
#pragma once
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace boost_shim {
struct Case { std::string name; std::function<void()> body; };
inline std::vector<Case> &cases() { static std::vector<Case> c; return c; }
inline long &n_checks() { static long n = 0; return n; }
inline long &n_failed() { static long n = 0; return n; }
struct Registrar { Registrar(const char *n, void (*f)()) { cases().push_back(Case{n, f}); } };
inline void report(bool ok, const char *file, int line, const std::string &msg) {
    ++n_checks();
    if (!ok) { ++n_failed(); std::cerr << file << "(" << line << "): check failed: " << msg << "\n"; }
}
}  // namespace boost_shim

#define BOOST_AUTO_TEST_CASE(NAME)                                            \
    static void NAME##_body();                                                \
    static boost_shim::Registrar NAME##_registrar(#NAME, &NAME##_body);       \
    static void NAME##_body()

#define BOOST_CHECK_MESSAGE(COND, MSG)                                        \
    do { std::ostringstream shim_os_; shim_os_ << MSG;                        \
         boost_shim::report(static_cast<bool>(COND), __FILE__, __LINE__, shim_os_.str()); } while (0)

#define BOOST_CHECK(COND) BOOST_CHECK_MESSAGE(COND, #COND)

// Operands are evaluated exactly once (test.cpp:279 passes "(++it)->first").
#define BOOST_CHECK_EQUAL(A, B)                                               \
    do { const auto shim_a_ = (A); const auto shim_b_ = (B);                  \
         BOOST_CHECK_MESSAGE(shim_a_ == shim_b_, #A " == " #B " [" << shim_a_ << " vs " << shim_b_ << "]"); } while (0)

int main() {
    for (auto &c : boost_shim::cases()) {
        const long before = boost_shim::n_failed();
        c.body();
        std::cout << c.name << (boost_shim::n_failed() == before ? " passed" : " FAILED") << "\n";
    }
    std::cout << boost_shim::cases().size() << " cases, " << boost_shim::n_checks() << " checks, "
              << boost_shim::n_failed() << " failures\n";
    return boost_shim::n_failed() ? 1 : 0;
}
