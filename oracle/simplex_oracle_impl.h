/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of the per-pivot loop of the reference's revised simplex
 * (reference: src/v4_cub_reduction.cu, "v4" below).  This header is included
 * twice by simplex_oracle.c, once with REAL=float and once with REAL=double.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may call it; the product path (simplex_method_gpu_b200/csrc) never does.
 *
 * What is followed, line by line:
 *   initial state ........ v4:272-279  (B_inv = I, c_b = c[n-m..), x_b = b,
 *                                      b_ixs[j] = n-m+j, y = c_b; the
 *                                      `n - m` length at v4:277 is read as m,
 *                                      the only length that fits y_aug)
 *   pricing .............. v4:288-302  e = [1 y]·[-c; A] over ALL n columns,
 *                                      argmin with lowest index on ties
 *                                      (cub ArgMin), optimal iff min >= -EPS
 *   FTRAN ................ v4:306-308  alpha = B_inv · A[:,p]
 *   ratio test ........... v4:199-208, 311-326  eligibility alpha > 0 (strict),
 *                                      theta = x_b/alpha else +INF, unbounded
 *                                      iff no row is eligible, lowest index
 *   row extract / E_q .... v4:331-332, 210-215 (row taken BEFORE the update;
 *                                      the i==q entry is evaluated in double)
 *   rank-1 update ........ v4:333      B_inv += E_q (x) row_q
 *   bookkeeping .......... v4:339-342  c_b_q = c_b[q]; c_b[q] = c[p]; b_ixs[q] = p
 *   x_b .................. v4:347-348  x_b += (row_q · b) E_q
 *   y .................... v4:353-356  y += ((c_b_new · E_q) + (c_p - c_b_q)) row_q
 *   loop / MAX_ITER ...... v4:286-287, 359
 *   objective ............ v4:362-368  z = c_b · x_b
 *
 * cuBLAS does not document its summation order, so dot products / GEMV here
 * use plain left-to-right fused multiply-adds.  `order` = 1 selects the
 * summation order of the B200 engine instead (lane-strided partial sums and a
 * butterfly, 256-column chunks for the FTRAN) so that engine results can be
 * compared bit for bit; it changes nothing but the association of the sums.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* ---- dot products ---------------------------------------------------- */

/* left-to-right fma chain starting from `init` */
static REAL FN(dot_seq)(const REAL* x, const REAL* y, long n, REAL init) {
	REAL acc = init;
	for (long i = 0; i < n; ++i) acc = FMA(x[i], y[i], acc);
	return acc;
}

/*
 * Engine order of the pricing dot product (price_phase in kernels.cuh): one
 * 256-thread CTA per column, thread t owns the 16-byte vectors t, t+256, ...
 * (V = 16/sizeof(REAL) elements each), one fma chain per vector slot, slots
 * summed left to right, butterfly (xor 16,8,4,2,1) inside each warp, then the
 * 8 warp sums left to right.
 */
static REAL FN(dot_block256)(const REAL* x, const REAL* y, long n) {
	enum { V = 16 / sizeof(REAL) };
	REAL th[256];
	for (int t = 0; t < 256; ++t) {
		REAL acc[V];
		for (int v = 0; v < V; ++v) acc[v] = (REAL)0;
		for (long base = (long)t * V; base < n; base += 256 * V)
			for (int v = 0; v < V; ++v)
				if (base + v < n) acc[v] = FMA(x[base + v], y[base + v], acc[v]);
		REAL a = acc[0];
		for (int v = 1; v < V; ++v) a = a + acc[v];
		th[t] = a;
	}
	REAL tot = (REAL)0;
	for (int w = 0; w < 8; ++w) {
		REAL s[32];
		for (int l = 0; l < 32; ++l) s[l] = th[w * 32 + l];
		for (int off = 16; off >= 1; off >>= 1) {
			REAL t2[32];
			for (int l = 0; l < 32; ++l) t2[l] = s[l] + s[l ^ off];
			for (int l = 0; l < 32; ++l) s[l] = t2[l];
		}
		tot = w == 0 ? s[0] : tot + s[0];
	}
	return tot;
}

/*
 * Engine order for the O(m) bookkeeping dots: fixed slices of ORACLE_SLICE
 * elements, each reduced by one 256-thread block (thread t owns element t, warp
 * butterfly, then the 8 warp sums left to right); slice sums are added left to
 * right (book1_phase / book2_phase / objective in kernels.cuh).
 */
#ifndef ORACLE_SLICE
#define ORACLE_SLICE 256
#endif
static REAL FN(dot_sliced)(const REAL* x, const REAL* y, long n) {
	REAL total = (REAL)0;
	for (long s0 = 0; s0 < n; s0 += ORACLE_SLICE) {
		long len = n - s0 < ORACLE_SLICE ? n - s0 : ORACLE_SLICE;
		REAL th[256];
		for (int t = 0; t < 256; ++t) {
			REAL a = (REAL)0;
			for (long i = t; i < len; i += 256) a = FMA(x[s0 + i], y[s0 + i], a);
			th[t] = a;
		}
		REAL blk = (REAL)0;
		for (int w = 0; w < 8; ++w) {
			REAL s[32];
			for (int l = 0; l < 32; ++l) s[l] = th[w * 32 + l];
			for (int off = 16; off >= 1; off >>= 1) {
				REAL t2[32];
				for (int l = 0; l < 32; ++l) t2[l] = s[l] + s[l ^ off];
				for (int l = 0; l < 32; ++l) s[l] = t2[l];
			}
			blk = blk + s[0];
		}
		total = total + blk;
	}
	return total;
}

/* ---- the solver ------------------------------------------------------ */

int FN(oracle_solve_ex)(const REAL* A, const REAL* b, const REAL* c, long m, long n,
		REAL eps, long max_iter, int order, const oracle_opts* opts,
		REAL* x_b_out, int* b_ixs_out, REAL* y_out, REAL* Binv_out,
		int* trace_p, int* trace_q, double* trace_gap_p, double* trace_gap_q,
		long trace_cap, oracle_result* res) {
	if (m <= 0 || n <= 0 || m > n) return -1;
	/* modes outside the reference's contract (all zero = v4); see simplex_oracle.h */
	const REAL pivot_tol = opts ? (REAL)opts->pivot_tol : (REAL)0;
	const REAL harris_delta = opts ? (REAL)opts->harris_delta : (REAL)0;
	const int ratio_mode = opts ? opts->ratio_mode : 0;
	const int steepest = opts ? opts->pricing_rule == 1 : 0;
	const int nranks = opts && opts->nranks > 1 ? opts->nranks : 1;
	/* the engine's row partition of B^-1 (Engine ctor in engine.cu): equal blocks of whole warp-wide vector rows */
	const long rowq = 32 * (16 / (long)sizeof(REAL));
	const long ld_e = (m + rowq - 1) / rowq * rowq;
	const long rpr = ((ld_e + nranks - 1) / nranks + rowq - 1) / rowq * rowq;

	const long chunk = 256; /* engine FTRAN chunk width (order = 1) */
	REAL* Binv = (REAL*)calloc((size_t)m * m, sizeof(REAL));
	REAL* c_b = (REAL*)malloc(sizeof(REAL) * m);
	REAL* x_b = (REAL*)malloc(sizeof(REAL) * m);
	REAL* y = (REAL*)malloc(sizeof(REAL) * m);
	REAL* e = (REAL*)malloc(sizeof(REAL) * n);
	REAL* alpha = (REAL*)malloc(sizeof(REAL) * m);
	REAL* theta = (REAL*)malloc(sizeof(REAL) * m);
	REAL* row_q = (REAL*)malloc(sizeof(REAL) * m);
	REAL* E_q = (REAL*)malloc(sizeof(REAL) * m);
	int* b_ixs = (int*)malloc(sizeof(int) * m);
	REAL* gamma = steepest ? (REAL*)malloc(sizeof(REAL) * n) : NULL;   /* steepest-edge weights 1 + |B^-1 a_j|^2 */
	REAL* vbt = steepest ? (REAL*)malloc(sizeof(REAL) * m) : NULL;     /* v = B^-T alpha */
	if (!Binv || !c_b || !x_b || !y || !e || !alpha || !theta || !row_q || !E_q || !b_ixs) return -2;
	if (steepest && (!gamma || !vbt)) return -2;

	/* v4:272-277 */
	for (long i = 0; i < m; ++i) {
		Binv[i + i * m] = (REAL)1;
		c_b[i] = c[n - m + i];
		x_b[i] = b[i];
		b_ixs[i] = (int)(n - m + i);
		y[i] = c_b[i];
	}

	/* order = 1: like the engine, recognise an identity slack block (v4:272 assumes it) */
	long ns = n;
	if (order == 1) {
		int ident = 1;
		for (long j = 0; j < m && ident; ++j)
			for (long i = 0; i < m; ++i)
				if (A[i + (n - m + j) * m] != (REAL)(i == j)) { ident = 0; break; }
		if (ident) ns = n - m;
	}

	if (steepest) {
		/* reference framework = the slack basis: B^-1 = I, gamma_j = 1 + |a_j|^2 (basic columns: never looked at) */
		#pragma omp parallel for schedule(static)
		for (long j = 0; j < n; ++j) {
			const REAL* col = A + j * m;
			if (j >= ns) gamma[j] = (REAL)2;
			else gamma[j] = (REAL)1 + (order == 0 ? FN(dot_seq)(col, col, m, (REAL)0) : FN(dot_block256)(col, col, m));
		}
	}

	int status = 0; /* MaxIter */
	long it = 0, pivots = 0;

	do {
		/* ---- pricing, v4:288-302 ---- */
		#pragma omp parallel for schedule(static)
		for (long j = 0; j < n; ++j) {
			const REAL* col = A + j * m;
			if (order == 0) {
				e[j] = FN(dot_seq)(y, col, m, -c[j]);
			} else if (j < ns) {
				/* engine: block dot, then "- c_j" */
				e[j] = FN(dot_block256)(col, y, m) - c[j];
			} else {
				/* engine: recognised slack column, e_j = y_k - c_j (no matrix bytes) */
				e[j] = y[j - ns] - c[j];
			}
		}
		long p = 0;
		REAL min_val = e[0];
		for (long j = 1; j < n; ++j)
			if (e[j] < min_val) { min_val = e[j]; p = j; }
		double gap_p = INFINITY;
		for (long j = 0; j < n; ++j)
			if (j != p && (double)e[j] - (double)min_val < gap_p) gap_p = (double)e[j] - (double)min_val;

		if (min_val >= -eps) { status = 1; ++it; break; }

		if (steepest) {
			/* p = argmax e_j^2 / gamma_j over the attractive columns e_j < -eps, lowest index on ties */
			REAL best = (REAL)-1, second = (REAL)-1;
			for (long j = 0; j < n; ++j) {
				if (!(e[j] < -eps)) continue;
				const REAL sc = (e[j] * e[j]) / gamma[j];
				if (sc > best) { second = best; best = sc; p = j; }
				else if (sc > second) second = sc;
			}
			gap_p = second < (REAL)0 ? INFINITY : (double)best - (double)second;
		}

		/* ---- FTRAN, v4:307-308 ---- */
		const REAL* a_p = A + p * m;
		if (order == 0) {
			for (long i = 0; i < m; ++i) alpha[i] = (REAL)0;
			#pragma omp parallel
			{
				#pragma omp for schedule(static)
				for (long i0 = 0; i0 < m; i0 += 512) {
					long i1 = i0 + 512 < m ? i0 + 512 : m;
					for (long j = 0; j < m; ++j) {
						const REAL aj = a_p[j];
						const REAL* bc = Binv + j * m;
						for (long i = i0; i < i1; ++i) alpha[i] = FMA(bc[i], aj, alpha[i]);
					}
				}
			}
		} else {
			/* engine (update_ftran_phase): 32-column sub-blocks, one fma chain each;
			 * pairwise tree over the 8 sub-blocks of a 256-column chunk; chunks left to right */
			#pragma omp parallel for schedule(static)
			for (long i0 = 0; i0 < m; i0 += 256) {
				long i1 = i0 + 256 < m ? i0 + 256 : m;
				REAL tot[256], sub[8][256];
				for (long j0 = 0; j0 < m; j0 += chunk) {
					for (int sb = 0; sb < 8; ++sb) {
						for (long i = i0; i < i1; ++i) sub[sb][i - i0] = (REAL)0;
						long ja = j0 + sb * 32, jb = ja + 32 < m ? ja + 32 : m;
						for (long j = ja; j < jb; ++j) {
							const REAL aj = a_p[j];
							const REAL* bc = Binv + j * m;
							for (long i = i0; i < i1; ++i) sub[sb][i - i0] = FMA(bc[i], aj, sub[sb][i - i0]);
						}
					}
					for (long i = i0; i < i1; ++i) {
						const long r = i - i0;
						const REAL c8 = ((sub[0][r] + sub[1][r]) + (sub[2][r] + sub[3][r]))
						              + ((sub[4][r] + sub[5][r]) + (sub[6][r] + sub[7][r]));
						tot[r] = j0 == 0 ? c8 : tot[r] + c8;
					}
				}
				for (long i = i0; i < i1; ++i) alpha[i] = tot[i - i0];
			}
		}

		/* ---- ratio test, v4:199-208, 311-326 (+ the optional modes) ---- */
		long num_non_pos = 0;
		for (long i = 0; i < m; ++i) {
			int flag = alpha[i] > pivot_tol;
			const REAL xb = ratio_mode >= 1 && x_b[i] < (REAL)0 ? (REAL)0 : x_b[i];
			theta[i] = flag ? (xb / alpha[i]) : (REAL)INFINITY;
			num_non_pos += !flag;
		}
		if (num_non_pos == m) { status = 2; ++it; break; }
		long q = 0;
		double gap_q = INFINITY;
		if (ratio_mode == 2) {
			/* Harris: widest step the tolerance allows, then the largest pivot element inside it */
			REAL theta_max = (REAL)INFINITY;
			for (long i = 0; i < m; ++i) {
				if (!(alpha[i] > pivot_tol)) continue;
				const REAL xb = x_b[i] < (REAL)0 ? (REAL)0 : x_b[i];
				const REAL t1 = (xb + harris_delta) / alpha[i];
				if (t1 < theta_max) theta_max = t1;
			}
			REAL best_a = (REAL)-1;
			for (long i = 0; i < m; ++i)
				if (alpha[i] > pivot_tol && theta[i] <= theta_max && alpha[i] > best_a) { best_a = alpha[i]; q = i; }
		} else {
			REAL min_theta = theta[0];
			for (long i = 1; i < m; ++i)
				if (theta[i] < min_theta) { min_theta = theta[i]; q = i; }
			for (long i = 0; i < m; ++i)
				if (i != q && (double)theta[i] - (double)min_theta < gap_q) gap_q = (double)theta[i] - (double)min_theta;
		}

		if (pivots < trace_cap) {
			if (trace_p) trace_p[pivots] = (int)p;
			if (trace_q) trace_q[pivots] = (int)q;
			if (trace_gap_p) trace_gap_p[pivots] = gap_p;
			if (trace_gap_q) trace_gap_q[pivots] = gap_q;
		}

		REAL gamma_p = (REAL)0;
		if (steepest) {
			/* v = B^-T alpha with the inverse this pivot started from; gamma_p = 1 + |alpha|^2 exactly */
			#pragma omp parallel for schedule(static)
			for (long j = 0; j < m; ++j) {
				const REAL* bc = Binv + j * m;
				if (order == 0) vbt[j] = FN(dot_seq)(alpha, bc, m, (REAL)0);
				else {
					REAL acc = (REAL)0;
					for (int r = 0; r < nranks; ++r) {
						const long i0 = (long)r * rpr < m ? (long)r * rpr : m;
						const long i1 = (long)(r + 1) * rpr < m ? (long)(r + 1) * rpr : m;
						const REAL part = FN(dot_block256)(bc + i0, alpha + i0, i1 - i0);
						acc = r == 0 ? part : acc + part;
					}
					vbt[j] = acc;
				}
			}
			gamma_p = (REAL)1 + (order == 0 ? FN(dot_seq)(alpha, alpha, m, (REAL)0) : FN(dot_sliced)(alpha, alpha, m));
		}

		/* ---- row extract + E_q + rank-1 update, v4:331-333 ---- */
		const REAL alpha_q = alpha[q];
		for (long j = 0; j < m; ++j) row_q[j] = Binv[q + j * m];
		for (long i = 0; i < m; ++i)
			E_q[i] = (i != q) ? (-alpha[i] / alpha_q) : (REAL)(1.0 / (double)alpha_q - 1.0);
		#pragma omp parallel for schedule(static)
		for (long j = 0; j < m; ++j) {
			const REAL r = row_q[j];
			REAL* bc = Binv + j * m;
			for (long i = 0; i < m; ++i) bc[i] = FMA(E_q[i], r, bc[i]);
		}

		/* ---- bookkeeping, v4:339-342 ---- */
		const REAL c_b_q = c_b[q];
		const long leaving = b_ixs[q];
		c_b[q] = c[p];
		b_ixs[q] = (int)p;

		if (steepest) {
			/* Goldfarb-Reid recurrence over every column (basic ones carry values nobody reads) */
			#pragma omp parallel for schedule(static)
			for (long j = 0; j < n; ++j) {
				const REAL* col = A + j * m;
				REAL r, w;
				if (j >= ns) { r = row_q[j - ns]; w = vbt[j - ns]; }          /* recognised unit column */
				else if (order == 0) { r = FN(dot_seq)(row_q, col, m, (REAL)0); w = FN(dot_seq)(vbt, col, m, (REAL)0); }
				else { r = FN(dot_block256)(col, row_q, m); w = FN(dot_block256)(col, vbt, m); }
				const REAL t = r / alpha_q;
				const REAL g1 = FMA(t * t, gamma_p, FMA((REAL)-2 * t, w, gamma[j]));
				const REAL g2 = FMA(t, t, (REAL)1);
				gamma[j] = g1 > g2 ? g1 : g2;
			}
			{
				const REAL ia = (REAL)1 / alpha_q;
				const REAL g1 = gamma_p * (ia * ia), g2 = FMA(ia, ia, (REAL)1);
				gamma[leaving] = g1 > g2 ? g1 : g2;
				gamma[p] = (REAL)2;
			}
		}

		/* ---- x_b, v4:347-348 ---- */
		REAL s = order == 0 ? FN(dot_seq)(row_q, b, m, (REAL)0) : FN(dot_sliced)(row_q, b, m);
		for (long i = 0; i < m; ++i) x_b[i] = FMA(s, E_q[i], x_b[i]);

		/* ---- y, v4:353-356 ---- */
		s = order == 0 ? FN(dot_seq)(c_b, E_q, m, (REAL)0) : FN(dot_sliced)(c_b, E_q, m);
		s += c[p] - c_b_q;
		for (long i = 0; i < m; ++i) y[i] = FMA(s, row_q[i], y[i]);

		++pivots;
	} while (++it < max_iter);

	/* v4:362-368 (computed for every status so callers can inspect MaxIter states) */
	REAL z = order == 0 ? FN(dot_seq)(c_b, x_b, m, (REAL)0) : FN(dot_sliced)(c_b, x_b, m);

	if (res) {
		res->status = status;
		res->iterations = it;
		res->pivots = pivots;
		res->z = (double)z;
	}
	if (x_b_out) memcpy(x_b_out, x_b, sizeof(REAL) * m);
	if (b_ixs_out) memcpy(b_ixs_out, b_ixs, sizeof(int) * m);
	if (y_out) memcpy(y_out, y, sizeof(REAL) * m);
	if (Binv_out) memcpy(Binv_out, Binv, sizeof(REAL) * (size_t)m * m);

	free(Binv); free(c_b); free(x_b); free(y); free(e); free(alpha);
	free(theta); free(row_q); free(E_q); free(b_ixs); free(gamma); free(vbt);
	return 0;
}

int FN(oracle_solve)(const REAL* A, const REAL* b, const REAL* c, long m, long n,
		REAL eps, long max_iter, int order,
		REAL* x_b_out, int* b_ixs_out, REAL* y_out, REAL* Binv_out,
		int* trace_p, int* trace_q, double* trace_gap_p, double* trace_gap_q,
		long trace_cap, oracle_result* res) {
	return FN(oracle_solve_ex)(A, b, c, m, n, eps, max_iter, order, NULL, x_b_out, b_ixs_out, y_out, Binv_out,
		trace_p, trace_q, trace_gap_p, trace_gap_q, trace_cap, res);
}

#undef FN
#undef CAT
#undef CAT_
