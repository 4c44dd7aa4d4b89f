// CPU baseline harness with the shape of the reference's solver_glpk.cpp (/root/reference/solver_glpk.cpp:15-39):
// read a fixed-format MPS deck (tools/lp_convert.py to-mps writes one from the solver text format), run
// glp_simplex with default controls (primal simplex, one thread) and print `x[i] = ...` (1-based) followed by
// `Optimal objective: ...`, or `Problem status: <code>`.
//
// GLPK is an optional dependency: the program is compiled against it only where <glpk.h> exists
// (g++ -O2 tools/solver_glpk_harness.cpp -o bin/solver_glpk.out -lglpk).  Without the header it still builds
// and says so (exit code 3), so that the build and the tests do not depend on GLPK being installed; in that
// case `tools/lp_convert.py solve` (HiGHS dual simplex through scipy) prints the same lines instead.
#include <chrono>
#include <iostream>

#if defined(__has_include)
#if __has_include(<glpk.h>)
#define B200LP_HAVE_GLPK 1
#include <glpk.h>
#endif
#endif

int main(int argc, char* argv[]) {
	if (argc < 2) {
		std::cerr << "Usage: " << argv[0] << " file.mps\n";
		return 1;
	}
#ifdef B200LP_HAVE_GLPK
	glp_term_out(GLP_OFF);
	glp_prob* lp = glp_create_prob();
	glp_set_prob_name(lp, argv[1]);
	const int err = glp_read_mps(lp, GLP_MPS_DECK, NULL, argv[1]);
	if (err != 0) {
		std::cerr << "Error reading MPS file: " << err << "\n";
		glp_delete_prob(lp);
		return 2;
	}
	const auto t0 = std::chrono::steady_clock::now();
	glp_simplex(lp, NULL);
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	const int status = glp_get_status(lp);
	if (status == GLP_OPT) {
		const int n = glp_get_num_cols(lp);
		for (int i = 1; i <= n; ++i) std::cout << "x[" << i << "] = " << glp_get_col_prim(lp, i) << "\n";
		std::cout << "Optimal objective: " << glp_get_obj_val(lp) << "\n";
	} else {
		std::cout << "Problem status: " << status << "\n";
	}
	std::cerr << "glp_simplex: " << secs << " s on 1 core, GLPK " << glp_version() << "\n";
	glp_delete_prob(lp);
	return 0;
#else
	std::cerr << "solver_glpk: built without GLPK (<glpk.h> not found on this machine); "
	             "use `python tools/lp_convert.py solve " << argv[1] << "` (HiGHS) for the CPU optimum\n";
	return 3;
#endif
}
