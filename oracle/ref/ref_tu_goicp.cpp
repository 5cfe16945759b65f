// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- never linked into the product library.
//
// Wrapper translation unit that compiles the UNMODIFIED reference driver
// /root/reference/src/goicp/jly_goicp.cpp (found through -I, not copied) with one
// macro-injected counter: every call the reference makes to intro_select() is routed
// through ref_counted_select(), which bumps ref_select_calls and forwards.  The
// reference calls intro_select exactly once per bound evaluation inside
// GoICP::InnerBnB (jly_goicp.cpp:293-299), once for the initial error
// (jly_goicp.cpp:361-367) and once per GoICP::ICP (jly_goicp.cpp:118-127), so
// bound_evals = ref_select_calls - 1 - icp_calls.
#include <cstddef>
#include "goicp/jly_sorting.hpp"

long long ref_select_calls = 0;

template <typename T>
static inline void ref_counted_select(T* data, size_t st, size_t en, size_t k)
{
    ++ref_select_calls;
    intro_select(data, st, en, k);
}

#define intro_select ref_counted_select
#include "goicp/jly_goicp.cpp"
