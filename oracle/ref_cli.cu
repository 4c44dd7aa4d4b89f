// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// The patched reference (make_ref.sh P1-P8) as a command-line program with its own main() (v4:384-474):
// the binary the drop-in (v4_shim_env_*.out, bin/solver.out) is compared with on LPs that need more than the
// stock MAX_ITER = 5 iterations.  V4_EPS / V4_MAX_ITER from the environment set the two run-time constants.
#include <cstdlib>

static int*  ref_trace = nullptr;
static long  ref_trace_cap = 0;
static long  ref_iterations = 0;
static int   ref_quiet = 0;           // print "# Iteration k" like the stock binary (v4:287)
static int   ref_always_readback = 0;

#include REF_SOURCE

static int ref_env_init = [] {
	if (const char* s = std::getenv("V4_MAX_ITER")) MAX_ITER = std::atoi(s);
	if (const char* s = std::getenv("V4_EPS")) EPS = (real)std::atof(s);
	(void)ref_trace; (void)ref_trace_cap; (void)ref_iterations;
	return 0;
}();
