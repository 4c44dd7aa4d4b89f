/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * CPU oracle (see simplex_oracle_impl.h for the reference lines followed)
 * and the synthetic LP generators shared by tests and bench.
 * Build: make -C oracle   (gcc -O3 -march=native -fopenmp -ffp-contract=off)
 */
#include "simplex_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL double
#define SUFFIX _f64
#define FMA(a, b, c) fma((a), (b), (c))
#include "simplex_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FMA

#define REAL float
#define SUFFIX _f32
#define FMA(a, b, c) fmaf((a), (b), (c))
#include "simplex_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FMA

int oracle_num_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* ---- generators ------------------------------------------------------ */

static inline uint64_t mix64(uint64_t z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

double lpgen_u01(uint64_t seed, uint64_t stream, uint64_t idx) {
	uint64_t z = mix64(seed * 0x9E3779B97F4A7C15ULL + stream * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL);
	z = mix64(z + idx * 0x9E3779B97F4A7C15ULL);
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

#define GEN_DENSE(NAME, T)                                                         \
	void NAME(T* A, T* b, T* c, long m, long n, uint64_t seed) {                   \
		const long ns = n - m;                                                     \
		_Pragma("omp parallel for schedule(static)")                               \
		for (long j = 0; j < n; ++j) {                                             \
			T* col = A + j * m;                                                    \
			if (j < ns) {                                                          \
				for (long i = 0; i < m; ++i)                                       \
					col[i] = (T)lpgen_u01(seed, 0, (uint64_t)(i * ns + j));        \
			} else {                                                               \
				for (long i = 0; i < m; ++i) col[i] = (T)(i == j - ns);            \
			}                                                                      \
		}                                                                          \
		for (long i = 0; i < m; ++i)                                               \
			b[i] = (T)(0.5 * (double)ns * (1.0 + lpgen_u01(seed, 1, (uint64_t)i))); \
		for (long j = 0; j < n; ++j)                                               \
			c[j] = j < ns ? (T)(0.5 + lpgen_u01(seed, 2, (uint64_t)j)) : (T)0;     \
	}

GEN_DENSE(lpgen_dense_f64, double)
GEN_DENSE(lpgen_dense_f32, float)

void lpgen_klee_minty_f64(double* A, double* b, double* c, long d) {
	const long m = d, n = 2 * d;
	memset(A, 0, sizeof(double) * (size_t)m * n);
	for (long i = 0; i < d; ++i) {           /* constraint i+1 */
		for (long j = 0; j < i; ++j)         /* 2 * 2^(i-j) x_j */
			A[i + j * m] = ldexp(1.0, (int)(i - j + 1));
		A[i + i * m] = 1.0;
		A[i + (d + i) * m] = 1.0;            /* slack */
		b[i] = pow(5.0, (double)(i + 1));
	}
	for (long j = 0; j < d; ++j) c[j] = ldexp(1.0, (int)(d - 1 - j));
	for (long j = d; j < n; ++j) c[j] = 0.0;
}

void lpgen_assignment_f64(double* A, double* b, double* c, long k, uint64_t seed, double* w_out) {
	const long m = 2 * k, ns = k * k, n = ns + m;
	memset(A, 0, sizeof(double) * (size_t)m * n);
	for (long i = 0; i < k; ++i)
		for (long j = 0; j < k; ++j) {
			const long col = i * k + j;
			A[i + col * m] = 1.0;          /* sum_j x_ij <= 1 */
			A[(k + j) + col * m] = 1.0;    /* sum_i x_ij <= 1 */
			const double w = 1.0 + floor(99.0 * lpgen_u01(seed, 3, (uint64_t)col));
			c[col] = w;
			if (w_out) w_out[col] = w;
		}
	for (long i = 0; i < m; ++i) {
		A[i + (ns + i) * m] = 1.0;
		b[i] = 1.0;
		c[ns + i] = 0.0;
	}
}
