/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See simplex_oracle_impl.h.
 *
 * CPU oracle for the reference's dense revised simplex (src/v4_cub_reduction.cu)
 * plus the synthetic LP generators of the benchmark configurations.
 * Parity status: pinned on the reference's only known answer
 * (input/sample.txt:15-16: optimum 9 at x0 = 1, x1 = 3) and, on the GPU box,
 * against the reference's own v4 build (oracle/_ref, see make_ref.sh).
 */
#ifndef SIMPLEX_ORACLE_H
#define SIMPLEX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status values mirror `enum class SolveStatus` at v4:49-54 */
enum { ORACLE_MAX_ITER = 0, ORACLE_OPTIMUM = 1, ORACLE_UNBOUNDED = 2, ORACLE_THETA_OVERFLOW = 3 };

typedef struct {
	int status;
	long iterations; /* number of "# Iteration k" lines v4 would print (v4:287) */
	long pivots;
	double z;
} oracle_result;

/*
 * Modes outside the reference's parity contract; all zero = the reference's loop.
 * The engine implements the same rules with the same arithmetic (include/b200lp.h: pivot_tol, ratio_mode,
 * harris_delta, pricing_rule), so engine and oracle stay comparable pivot for pivot in every mode.
 *   pivot_tol ....... ratio-test eligibility alpha > pivot_tol instead of alpha > 0 (v4:203; README.md:30
 *                     "division by a small number")
 *   ratio_mode ...... 0 textbook (v4:199-208); 1 bounded: theta = max(x_b, 0) / alpha (README.md:29 "what if
 *                     x_b_t < 0"); 2 Harris two-pass: theta_max = min (max(x_b,0) + harris_delta) / alpha over
 *                     the eligible rows, then the LARGEST alpha among rows with max(x_b,0)/alpha <= theta_max
 *                     (lowest index on ties)
 *   pricing_rule .... 0 Dantzig (v4:288-302); 1 steepest edge with the Goldfarb-Reid recurrence (README.md:16-17):
 *                     p = argmax e_j^2 / gamma_j over e_j < -eps, gamma_j = 1 + |B^-1 a_j|^2 kept up to date with
 *                     gamma_j <- max(gamma_j - 2 t_j (a_j . v) + t_j^2 gamma_p, 1 + t_j^2), t_j = (row_q . a_j) / alpha_q,
 *                     v = B^-T alpha, gamma_p = 1 + alpha . alpha exact; optimality test unchanged (min e >= -eps)
 */
typedef struct {
	double pivot_tol;
	double harris_delta;
	int ratio_mode;
	int pricing_rule;
	int nranks;      /* order = 1 with steepest edge on R ranks: v = B^-T alpha is the sum, in rank order, of the column
	                    dots over each rank's row block (the engine's row partition); 0 / 1 = one block */
	int reserved;
} oracle_opts;

/*
 * A is column-major m x n (v4:59-60), the slack/identity block is the LAST m
 * columns (v4:272-277).  order: 0 = plain left-to-right sums, 1 = the B200
 * engine's summation order.  Every output pointer may be NULL.
 * trace_gap_p / trace_gap_q: distance from the chosen minimum to the runner-up
 * (reduced cost, resp. theta) so callers can apply the "ties within 1e-12" rule.
 */
int oracle_solve_f64(const double* A, const double* b, const double* c, long m, long n,
		double eps, long max_iter, int order,
		double* x_b, int* b_ixs, double* y, double* Binv,
		int* trace_p, int* trace_q, double* trace_gap_p, double* trace_gap_q,
		long trace_cap, oracle_result* res);

int oracle_solve_f32(const float* A, const float* b, const float* c, long m, long n,
		float eps, long max_iter, int order,
		float* x_b, int* b_ixs, float* y, float* Binv,
		int* trace_p, int* trace_q, double* trace_gap_p, double* trace_gap_q,
		long trace_cap, oracle_result* res);

/* the same with the optional modes above (opts == NULL: the reference's loop) */
int oracle_solve_ex_f64(const double* A, const double* b, const double* c, long m, long n,
		double eps, long max_iter, int order, const oracle_opts* opts,
		double* x_b, int* b_ixs, double* y, double* Binv,
		int* trace_p, int* trace_q, double* trace_gap_p, double* trace_gap_q,
		long trace_cap, oracle_result* res);
int oracle_solve_ex_f32(const float* A, const float* b, const float* c, long m, long n,
		float eps, long max_iter, int order, const oracle_opts* opts,
		float* x_b, int* b_ixs, float* y, float* Binv,
		int* trace_p, int* trace_q, double* trace_gap_p, double* trace_gap_q,
		long trace_cap, oracle_result* res);

/* ---- synthetic LPs; all write the full [A_s, I_m] matrix ---- */

/* counter-based uniform in [0,1): u01(splitmix64(seed*K + idx)) */
double lpgen_u01(uint64_t seed, uint64_t stream, uint64_t idx);

/* dense random: A_s ~ U(0,1), b = (n_s/2) U(1,2), c_s ~ U(0.5,1.5); n = n_s + m */
void lpgen_dense_f64(double* A, double* b, double* c, long m, long n, uint64_t seed);
void lpgen_dense_f32(float* A, float* b, float* c, long m, long n, uint64_t seed);

/* Klee-Minty cube of dimension d: m = d, n = 2d, optimum 5^d after 2^d - 1 pivots */
void lpgen_klee_minty_f64(double* A, double* b, double* c, long d);

/* assignment LP k x k with integer weights in 1..99: m = 2k, n = k*k + 2k */
void lpgen_assignment_f64(double* A, double* b, double* c, long k, uint64_t seed, double* w_out);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
