// Device code of the B200 dense revised-simplex engine (sm_100a).
//
// One pivot of the reference loop (src/v4_cub_reduction.cu:286-359) is five
// phases separated by grid-wide barriers; all of them live in ONE persistent
// cooperative kernel (simplex_persistent), so there is no host round trip per
// pivot.  The same phase functions are also wrapped as stand-alone kernels for
// unit tests and for the sharded multi-GPU driver.
//
//   price          e_j = y.A_j - c_j fused with the argmin           (v4:289-296)
//   update_ftran   B^-1 += E_q (x) row_q  AND  alpha = B^-1_new a_p  (v4:333 + v4:307-308)
//                  -> B^-1 crosses HBM once per pivot (read + write)
//   ratio          alpha = sum of chunk partials, masked argmin      (v4:311-325)
//   book1          row_q gather, E_q, the two O(m) dot products      (v4:331-332, 347, 354)
//   book2          x_b, y, c_b, b_ixs                                (v4:339-356)
//
// Layout: everything column-major like the reference (v4:59-60).  The leading
// dimension ld is m rounded up to one warp-wide 16-byte vector row (64 doubles
// / 128 floats = 512 B) and the padding rows are zero, so every warp access is
// a full, aligned 512 B segment without predicates.
//
// Summation orders are fixed by the problem size only (never by the grid or
// the tile shape), so results are bit-identical for every launch geometry:
//   pricing dot ..... thread t of 256 owns vectors t, t+256, ...; per-slot fma
//                     chains, slots left to right, warp butterfly, 8 warp sums
//                     left to right, then "- c_j"
//   FTRAN ........... 32-column sub-blocks (one fma chain each), pairwise tree
//                     over the 8 sub-blocks of a 256-column chunk, chunks left
//                     to right
//   O(m) dots ....... 256-element slices (thread t owns element t), warp
//                     butterfly, 8 warp sums left to right, slices left to right
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

namespace b200lp {

constexpr int NT = 256;       // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int CHUNK = 256;    // FTRAN partial-sum chunk (columns)
constexpr int SUBW = 32;      // FTRAN sub-block (columns)
constexpr int SLICE = 256;    // O(m) dot slice (elements)
constexpr int PRICE_NC = 4;   // columns priced together by one CTA
constexpr int MIN_CTAS = 2;   // resident CTAs per SM the persistent kernel is compiled for
constexpr int MAXR = 8;       // ranks (GPUs) of one NVSwitch box

template <typename T> struct VecT;
template <> struct VecT<double> { using V = double2; static constexpr int N = 2; };
template <> struct VecT<float>  { using V = float4;  static constexpr int N = 4; };

// (value, index) candidate of a distributed argmin; value kept as double for
// both dtypes (float -> double is exact and order preserving)
struct Cand {
	double val;
	long long idx;
};

struct Ctl {
	unsigned long long bar;   // grid barrier arrivals (reset by the host before a launch)
	long long iter;           // iterations done ("# Iteration" lines, v4:287)
	long long pivots;
	long long it_end;         // this launch stops at iter == it_end
	int status;               // B200LP_STATUS_*
	int pending;              // rank-1 update (E_q, row_q) not yet applied to B^-1
	int done;                 // optimum / unbounded reached
	int bad;                  // peer barrier timed out (sharded mode)
	unsigned long long xepoch; // cross-GPU barrier epoch, monotonic over the engine's life
	long long p, q;           // last entering column / leaving row
	double min_e;             // last pricing minimum
	double c_b_q;             // c_b[q] before the swap (v4:339)
	double z;                 // c_b . x_b (v4:365)
};

template <typename T>
struct Dev {
	long long m, n, ns, ld;   // ns = dense (structural) columns; the n - ns others are unit vectors
	int nchunk, nslice;
	const T* A;               // ld x ns, read-only while solving
	T* B;                     // ld x m   (B^-1)
	const T* b;               // ld
	const T* c;               // n
	T *y, *x_b, *c_b, *alpha, *E_q, *row_q; // ld each
	T* alpha_part;            // nchunk x ld
	T* dpart;                 // 2 x nslice
	int* b_ixs;               // m
	Cand* cand;               // one per CTA
	long long* cnt;           // eligible rows, one per CTA
	Ctl* ctl;
	int2* trace;
	long long trace_cap;
	double eps;
	// ---- sharding over the GPUs of one box (single GPU: rank 0 of 1, row0 = 0, ldb = ld, col0 = 0, nsl = ns)
	// B^-1 is row-block sharded (B holds rows [row0, row0+ldb) with leading dimension ldb),
	// A is column-block sharded (A holds dense columns [col0, col0+nsl), ld rows each),
	// every O(m) vector is replicated.
	int rank, nranks;
	long long row0, ldb;
	long long col0, nsl;
	long long k0, k1;                 // share of the unit (slack) columns priced on this rank
	long long colstart[MAXR + 1], rowstart[MAXR + 1];
	const T* A_peer[MAXR];            // every rank's A shard (peer-mapped over NVLink)
	unsigned char* mbox_peer[MAXR];   // every rank's mailbox: [XHdr][alpha ld][row_q ld]; alpha/row_q above point into our own
	T* acol;                          // local copy of the entering column (sharded mode only)
};

// head of a rank's mailbox; peers store into it over NVLink
struct XHdr {
	unsigned long long xflag[MAXR];   // xflag[r] = last barrier epoch rank r has signalled to us
	Cand candx[MAXR];                 // candx[r] = rank r's pricing candidate
	unsigned long long pad[8];
};
static_assert(sizeof(XHdr) % 256 == 0, "mailbox vectors must stay 16-byte aligned");

// ---------------------------------------------------------------- memory ops

template <typename T> struct Mem;
template <> struct Mem<double> {
	using V = double2;
	// read-only stream (A): non-coherent path, do not allocate in L1
	static __device__ __forceinline__ V ld_nc(const double* p) {
		V v;
		asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
		return v;
	}
	// read-write stream (B^-1): coherent, do not allocate in L1
	static __device__ __forceinline__ V ld_stream(const double* p) {
		V v;
		asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
		return v;
	}
	static __device__ __forceinline__ void st_stream(double* p, V v) {
		asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
	}
	static __device__ __forceinline__ double get(const V& v, int k) { return k == 0 ? v.x : v.y; }
	static __device__ __forceinline__ void set(V& v, int k, double s) { if (k == 0) v.x = s; else v.y = s; }
};
template <> struct Mem<float> {
	using V = float4;
	static __device__ __forceinline__ V ld_nc(const float* p) {
		V v;
		asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
		return v;
	}
	static __device__ __forceinline__ V ld_stream(const float* p) {
		V v;
		asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
		return v;
	}
	static __device__ __forceinline__ void st_stream(float* p, V v) {
		asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
	}
	static __device__ __forceinline__ float get(const V& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
	static __device__ __forceinline__ void set(V& v, int k, float s) { if (k == 0) v.x = s; else if (k == 1) v.y = s; else if (k == 2) v.z = s; else v.w = s; }
};

__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }

template <typename T>
__device__ __forceinline__ T warp_butterfly_sum(T s) {
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
	return s;
}

// lexicographic (value, index): the lowest index wins ties, like cub ArgMin (v4:294, 324)
__device__ __forceinline__ bool cand_better(double v, long long i, double bv, long long bi) {
	return v < bv || (v == bv && i < bi);
}

__device__ __forceinline__ void warp_argmin(double& v, long long& i) {
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) {
		const double ov = __shfl_xor_sync(0xffffffffu, v, off);
		const long long oi = __shfl_xor_sync(0xffffffffu, i, off);
		if (cand_better(ov, oi, v, i)) { v = ov; i = oi; }
	}
}

struct Smem {
	double red_v[NWARP];
	long long red_i[NWARP];
	long long red_c[NWARP];
	double wsum[2][PRICE_NC][NWARP];  // pricing: warp sums, double buffered
	double dsum[2][NWARP];            // O(m) dots
	double stage[2][CHUNK];           // update_ftran: row_q chunk, a_p chunk (as T)
	double comb[NWARP][32][4];        // update_ftran: cross-warp combine (WC > 1)
	double bc_v;                      // broadcasts
	long long bc_i;
	long long bc_c;
	double bc_s[2];
};

// block-wide argmin; result valid in every thread
__device__ __forceinline__ void block_argmin(double& v, long long& i, Smem& sh) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	warp_argmin(v, i);
	__syncthreads();
	if (lane == 0) { sh.red_v[warp] = v; sh.red_i[warp] = i; }
	__syncthreads();
	v = sh.red_v[0]; i = sh.red_i[0];
#pragma unroll
	for (int w = 1; w < NWARP; ++w)
		if (cand_better(sh.red_v[w], sh.red_i[w], v, i)) { v = sh.red_v[w]; i = sh.red_i[w]; }
}

// ---------------------------------------------------------------- grid barrier

// All CTAs are co-resident (cooperative launch).  One arrival per CTA on a
// monotonically increasing counter; release/acquire through __threadfence,
// which also invalidates L1 so plain loads after the barrier see fresh data.
__device__ __forceinline__ void grid_barrier(Ctl* ctl, unsigned long long& epoch) {
	epoch += gridDim.x;
	if (gridDim.x == 1) { __syncthreads(); return; }
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(&ctl->bar, 1ULL);
		while (*((volatile unsigned long long*)&ctl->bar) < epoch) { }
		__threadfence();
	}
	__syncthreads();
}

// every CTA reduces the per-CTA candidates redundantly -> same answer everywhere
__device__ __forceinline__ void reduce_cands(const Cand* cand, int ncand, double& v, long long& i, Smem& sh) {
	v = CUDART_INF; i = LLONG_MAX;
	for (int k = threadIdx.x; k < ncand; k += NT) {
		const double cv = __ldcg(&cand[k].val);
		const long long ci = __ldcg(&cand[k].idx);
		if (cand_better(cv, ci, v, i)) { v = cv; i = ci; }
	}
	block_argmin(v, i, sh);
}

__device__ __forceinline__ long long reduce_counts(const long long* cnt, int n, Smem& sh) {
	long long c = 0;
	for (int k = threadIdx.x; k < n; k += NT) c += __ldcg(&cnt[k]);
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) sh.red_c[threadIdx.x >> 5] = c;
	__syncthreads();
	c = 0;
#pragma unroll
	for (int w = 0; w < NWARP; ++w) c += sh.red_c[w];
	return c;
}

// ---------------------------------------------------------------- phase: pricing

// e_j = y.A_j - c_j for the ns dense columns (CTA b owns a contiguous column
// range), e_j = y_k - c_j for the unit columns, fused with the argmin.
// Replaces cublasSgemm(M=1) + cub::DeviceReduce::ArgMin (v4:289-294).
template <typename T>
__device__ void price_phase(const Dev<T>& d, Smem& sh, int part, int nparts) {
	using M = Mem<T>;
	using V = typename VecT<T>::V;
	constexpr int VN = VecT<T>::N;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const long long ld = d.ld;

	double best_v = CUDART_INF;
	long long best_i = LLONG_MAX;

	const long long c0 = d.nsl * part / nparts, c1 = d.nsl * (part + 1) / nparts;   // local columns
	int buf = 0;
	for (long long col = c0; col < c1; col += PRICE_NC, buf ^= 1) {
		const T* ap[PRICE_NC];
#pragma unroll
		for (int k = 0; k < PRICE_NC; ++k) {
			const long long cc = col + k < c1 ? col + k : c1 - 1;   // ragged tail: re-read the last column
			ap[k] = d.A + cc * ld;
		}
		T acc[PRICE_NC][VN];
#pragma unroll
		for (int k = 0; k < PRICE_NC; ++k)
#pragma unroll
			for (int v = 0; v < VN; ++v) acc[k][v] = T(0);

#pragma unroll 4
		for (long long i = (long long)tid * VN; i < ld; i += (long long)NT * VN) {
			const V yv = *reinterpret_cast<const V*>(d.y + i);
			V av[PRICE_NC];
#pragma unroll
			for (int k = 0; k < PRICE_NC; ++k) av[k] = M::ld_nc(ap[k] + i);
#pragma unroll
			for (int k = 0; k < PRICE_NC; ++k)
#pragma unroll
				for (int v = 0; v < VN; ++v) acc[k][v] = fma_t(M::get(av[k], v), M::get(yv, v), acc[k][v]);
		}
#pragma unroll
		for (int k = 0; k < PRICE_NC; ++k) {
			T s = acc[k][0];
#pragma unroll
			for (int v = 1; v < VN; ++v) s = s + acc[k][v];
			s = warp_butterfly_sum(s);
			if (lane == 0) sh.wsum[buf][k][warp] = (double)s;
		}
		__syncthreads();
		if (tid < PRICE_NC && col + tid < c1) {
			T s = (T)sh.wsum[buf][tid][0];
#pragma unroll
			for (int w = 1; w < NWARP; ++w) s = s + (T)sh.wsum[buf][tid][w];
			const long long j = d.col0 + col + tid;     // global column index
			const double e = (double)(s - d.c[j]);
			if (cand_better(e, j, best_v, best_i)) { best_v = e; best_i = j; }
		}
	}

	// unit (slack) columns: e_j = y_k - c_j, no matrix bytes (the reference reads
	// the identity block through the same GEMM; 0*y terms vanish exactly)
	for (long long k = d.k0 + (long long)part * NT + tid; k < d.k1; k += (long long)nparts * NT) {
		const double e = (double)(d.y[k] - d.c[d.ns + k]);
		const long long j = d.ns + k;
		if (cand_better(e, j, best_v, best_i)) { best_v = e; best_i = j; }
	}

	block_argmin(best_v, best_i, sh);
	if (tid == 0) { d.cand[part].val = best_v; d.cand[part].idx = best_i; }
}

// ---------------------------------------------------------------- phase: update + FTRAN

// One pass over B^-1:  (UPDATE) B^-1 += E_q (x) row_q   [cublasSger, v4:333]
//                      (FTRAN)  alpha = B^-1_new a_p     [cublasSgemv, v4:307-308 of the NEXT iteration]
// Thread owns VN consecutive rows (one 16-byte vector), a warp 32*VN rows, the
// CTA's 8 warps are arranged WR (rows) x WC (columns) over a tile of
// WR*32*VN rows x CHUNK columns.  row_q[chunk] and a_p[chunk] are staged in
// shared memory.  alpha_part[chunk][row] receives the chunk partial.
template <typename T, int WC, bool UPDATE, bool FTRAN>
__device__ void update_ftran_phase(const Dev<T>& d, Smem& sh, const T* acol, long long uk, int part, int nparts) {
	using M = Mem<T>;
	using V = typename VecT<T>::V;
	constexpr int VN = VecT<T>::N;
	constexpr int WR = NWARP / WC;
	constexpr int TR = WR * 32 * VN;      // tile rows
	constexpr int SUBS = (CHUNK / SUBW) / WC; // sub-blocks per warp
	static_assert(SUBS >= 1, "bad WC");

	static_assert(NT == CHUNK, "staging assumes one thread per chunk column");
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wr = warp % WR, wc = warp / WR;
	const long long ld = d.ldb, m = d.m;   // local row block
	const long long ntr = (ld + TR - 1) / TR;
	const long long ntiles = ntr * d.nchunk;
	T* stage_rq = reinterpret_cast<T*>(sh.stage[0]);
	T* stage_a = reinterpret_cast<T*>(sh.stage[1]);
	const bool unit = acol == nullptr;     // entering column is the unit vector e_uk

	for (long long tile = part; tile < ntiles; tile += nparts) {
		const long long rt = tile % ntr, ck = tile / ntr;
		const long long j0 = ck * CHUNK;
		__syncthreads();
		{
			const long long j = j0 + tid;     // NT == CHUNK
			T r = T(0), a = T(0);
			if (j < m) {
				if (UPDATE) r = d.row_q[j];
				if (FTRAN) a = unit ? (j == uk ? T(1) : T(0)) : acol[j];   // plain load: acol may have been written by this kernel
			}
			stage_rq[tid] = r;
			stage_a[tid] = a;
		}
		__syncthreads();

		const long long row = rt * TR + (long long)wr * 32 * VN + (long long)lane * VN;
		const bool active = row < ld;     // warp uniform (ld is a multiple of 32*VN)
		T wpart[VN];
#pragma unroll
		for (int v = 0; v < VN; ++v) wpart[v] = T(0);

		if (active) {
			V Ev;
			if (UPDATE) Ev = *reinterpret_cast<const V*>(d.E_q + d.row0 + row);
			// pairwise tree over this warp's sub-blocks, kept as a binary-counter stack so
			// at most log2(SUBS)+1 partials are live
			T stk[4][VN];
#pragma unroll
			for (int s = 0; s < SUBS; ++s) {
				const int jb = wc * (CHUNK / WC) + s * SUBW;   // offset inside the chunk
				T acc[VN];
#pragma unroll
				for (int v = 0; v < VN; ++v) acc[v] = T(0);
				T* bp = d.B + (j0 + jb) * ld + row;
				long long ncols = m - (j0 + jb);
				if (ncols >= SUBW) {
#pragma unroll
					for (int u0 = 0; u0 < SUBW; u0 += 8) {
						V v8[8];
#pragma unroll
						for (int u = 0; u < 8; ++u) v8[u] = M::ld_stream(bp + (long long)(u0 + u) * ld);
#pragma unroll
						for (int u = 0; u < 8; ++u) {
							const T r = stage_rq[jb + u0 + u], a = stage_a[jb + u0 + u];
#pragma unroll
							for (int v = 0; v < VN; ++v) {
								T x = M::get(v8[u], v);
								if (UPDATE) { x = fma_t(M::get(Ev, v), r, x); M::set(v8[u], v, x); }
								if (FTRAN) acc[v] = fma_t(x, a, acc[v]);
							}
							if (UPDATE) M::st_stream(bp + (long long)(u0 + u) * ld, v8[u]);
						}
					}
				} else {
					for (int u = 0; u < ncols; ++u) {
						V vv = M::ld_stream(bp + (long long)u * ld);
						const T r = stage_rq[jb + u], a = stage_a[jb + u];
#pragma unroll
						for (int v = 0; v < VN; ++v) {
							T x = M::get(vv, v);
							if (UPDATE) { x = fma_t(M::get(Ev, v), r, x); M::set(vv, v, x); }
							if (FTRAN) acc[v] = fma_t(x, a, acc[v]);
						}
						if (UPDATE) M::st_stream(bp + (long long)u * ld, vv);
					}
				}
				int lvl = 0;
#pragma unroll
				for (; lvl < 3; ++lvl) {
					if (!((s >> lvl) & 1)) break;
#pragma unroll
					for (int v = 0; v < VN; ++v) acc[v] = stk[lvl][v] + acc[v];
				}
#pragma unroll
				for (int v = 0; v < VN; ++v) stk[lvl][v] = acc[v];
			}
			constexpr int TOP = SUBS == 8 ? 3 : SUBS == 4 ? 2 : SUBS == 2 ? 1 : 0;
#pragma unroll
			for (int v = 0; v < VN; ++v) wpart[v] = stk[TOP][v];
		}

		if (FTRAN) {
			if (WC == 1) {
				if (active) {
					V o;
#pragma unroll
					for (int v = 0; v < VN; ++v) M::set(o, v, wpart[v]);
					*reinterpret_cast<V*>(d.alpha_part + ck * ld + row) = o;
				}
			} else {
				// continue the same pairwise tree across the WC warps that share these rows
#pragma unroll
				for (int v = 0; v < VN; ++v) sh.comb[warp][lane][v] = (double)wpart[v];
				__syncthreads();
				if (wc == 0 && active) {
					T t[WC][VN];
#pragma unroll
					for (int c = 0; c < WC; ++c)
#pragma unroll
						for (int v = 0; v < VN; ++v) t[c][v] = (T)sh.comb[c * WR + wr][lane][v];
#pragma unroll
					for (int w = 1; w < WC; w <<= 1)
#pragma unroll
						for (int c = 0; c + w < WC; c += 2 * w)
#pragma unroll
							for (int v = 0; v < VN; ++v) t[c][v] = t[c][v] + t[c + w][v];
					V o;
#pragma unroll
					for (int v = 0; v < VN; ++v) M::set(o, v, t[0][v]);
					*reinterpret_cast<V*>(d.alpha_part + ck * ld + row) = o;
				}
			}
		}
	}
}

// ---------------------------------------------------------------- phase: ratio test

// alpha_i = sum of the chunk partials (left to right); theta_i = x_b_i/alpha_i
// over alpha_i > 0 (strict, v4:203); masked argmin + eligible-row count.
// Replaces cudaMemset + compute_theta + D2H + cub ArgMin (v4:311-325).
template <typename T, bool FROM_PARTIALS>
__device__ void ratio_phase(const Dev<T>& d, Smem& sh, int part, int nparts) {
	const int tid = threadIdx.x;
	double best_v = CUDART_INF;
	long long best_i = LLONG_MAX;
	long long elig = 0;
	for (long long i = (long long)part * NT + tid; i < d.m; i += (long long)nparts * NT) {
		T a;
		if (FROM_PARTIALS) {
			a = __ldcg(d.alpha_part + i);
#pragma unroll 8
			for (int ck = 1; ck < d.nchunk; ++ck) a = a + __ldcg(d.alpha_part + (long long)ck * d.ldb + i);
			d.alpha[i] = a;
		} else {
			a = d.alpha[i];   // already exchanged between the ranks
		}
		if (a > T(0)) {
			++elig;
			const double th = (double)(d.x_b[i] / a);
			if (cand_better(th, i, best_v, best_i)) { best_v = th; best_i = i; }
		}
	}
	block_argmin(best_v, best_i, sh);
	// block-wide eligible count
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) elig += __shfl_xor_sync(0xffffffffu, elig, off);
	__syncthreads();
	if ((tid & 31) == 0) sh.red_c[tid >> 5] = elig;
	__syncthreads();
	if (tid == 0) {
		long long c = 0;
#pragma unroll
		for (int w = 0; w < NWARP; ++w) c += sh.red_c[w];
		d.cand[part].val = best_v;
		d.cand[part].idx = best_i;
		d.cnt[part] = c;
	}
}

// ---------------------------------------------------------------- phase: bookkeeping 1

// row_q = B^-1[q,:] (old), E_q from alpha (v4:331-332, 210-215) and the slice
// partials of  row_q.b  (v4:347)  and  c_b_new.E_q  (v4:354; c_b[q] already
// replaced by c[p], v4:340).
template <typename T>
__device__ void book1_phase(const Dev<T>& d, Smem& sh, long long p, long long q, int part, int nparts) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const T alpha_q = d.alpha[q];
	const T c_p = d.c[p];
	for (long long s = part; s < d.nslice; s += nparts) {
		const long long i = s * SLICE + tid;
		T t1 = T(0), t2 = T(0);
		if (i < d.m) {
			const T rq = d.B[q + i * d.ldb];
			const T eq = (i != q) ? (-d.alpha[i] / alpha_q) : (T)(1.0 / (double)alpha_q - 1.0);
			d.row_q[i] = rq;
			d.E_q[i] = eq;
			T cb = d.c_b[i];
			if (i == q) { d.ctl->c_b_q = (double)cb; cb = c_p; }
			t1 = fma_t(rq, d.b[i], T(0));
			t2 = fma_t(cb, eq, T(0));
		}
		t1 = warp_butterfly_sum(t1);
		t2 = warp_butterfly_sum(t2);
		__syncthreads();
		if (lane == 0) { sh.dsum[0][warp] = (double)t1; sh.dsum[1][warp] = (double)t2; }
		__syncthreads();
		if (tid < 2) {
			T a = T(0);
#pragma unroll
			for (int w = 0; w < NWARP; ++w) a = a + (T)sh.dsum[tid][w];
			d.dpart[(long long)tid * d.nslice + s] = a;
		}
	}
}

// ---------------------------------------------------------------- phase: bookkeeping 2

// x_b += (row_q.b) E_q (v4:348);  y += ((c_b_new.E_q) + (c_p - c_b_q)) row_q (v4:355-356);
// c_b[q] = c[p], b_ixs[q] = p (v4:340-342)
template <typename T>
__device__ void book2_phase(const Dev<T>& d, Smem& sh, long long p, long long q, int part, int nparts) {
	const int tid = threadIdx.x;
	__syncthreads();
	if (tid < 2) {
		T a = T(0);
#pragma unroll 8
		for (int s = 0; s < d.nslice; ++s) a = a + __ldcg(d.dpart + (long long)tid * d.nslice + s);
		if (tid == 1) a += d.c[p] - (T)__ldcg(&d.ctl->c_b_q);
		sh.bc_s[tid] = (double)a;
	}
	__syncthreads();
	const T sx = (T)sh.bc_s[0], sy = (T)sh.bc_s[1];
	for (long long i = (long long)part * NT + tid; i < d.m; i += (long long)nparts * NT) {
		const T eq = d.E_q[i], rq = d.row_q[i];
		d.x_b[i] = fma_t(sx, eq, d.x_b[i]);
		d.y[i] = fma_t(sy, rq, d.y[i]);
		if (i == q) { d.c_b[i] = d.c[p]; d.b_ixs[i] = (int)p; }
	}
}

// z = c_b . x_b in slice order (v4:365); single CTA
template <typename T>
__device__ double objective(const Dev<T>& d, Smem& sh) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	T z = T(0);
	for (long long s = 0; s < d.nslice; ++s) {
		const long long i = s * SLICE + tid;
		T t = i < d.m ? fma_t(d.c_b[i], d.x_b[i], T(0)) : T(0);
		t = warp_butterfly_sum(t);
		__syncthreads();
		if (lane == 0) sh.dsum[0][warp] = (double)t;
		__syncthreads();
		T a = T(0);
#pragma unroll
		for (int w = 0; w < NWARP; ++w) a = a + (T)sh.dsum[0][w];
		z = z + a;
	}
	return (double)z;
}

// ---------------------------------------------------------------- the persistent kernel

// The whole loop of v4:286-359 on the device.  Every CTA takes the same
// branches because every decision is recomputed from the same global data.
template <typename T, int WC>
__global__ void __launch_bounds__(NT, MIN_CTAS) simplex_persistent(Dev<T> d) {
	__shared__ Smem sh;
	Ctl* ctl = d.ctl;
	const int G = gridDim.x, me = blockIdx.x;
	unsigned long long epoch = 0;

	long long it = ctl->iter, pivots = ctl->pivots;
	const long long it_end = ctl->it_end;
	int pending = ctl->pending;
	int status = 0, done = 0;
	long long p = ctl->p, q = ctl->q;
	double min_e = ctl->min_e;

	while (it < it_end) {
		// ---- pricing + entering column (v4:288-302)
		price_phase<T>(d, sh, me, G);
		grid_barrier(ctl, epoch);
		reduce_cands(d.cand, G, min_e, p, sh);
		if (min_e >= -d.eps) { status = 1; done = 1; ++it; break; }

		// ---- pending rank-1 update fused with the FTRAN of column p
		const T* acol = p < d.ns ? d.A + p * d.ld : nullptr;
		if (pending) update_ftran_phase<T, WC, true, true>(d, sh, acol, p - d.ns, me, G);
		else         update_ftran_phase<T, WC, false, true>(d, sh, acol, p - d.ns, me, G);
		pending = 0;
		grid_barrier(ctl, epoch);

		// ---- ratio test (v4:311-325)
		ratio_phase<T, true>(d, sh, me, G);
		grid_barrier(ctl, epoch);
		double th;
		reduce_cands(d.cand, G, th, q, sh);
		const long long elig = reduce_counts(d.cnt, G, sh);
		if (elig == 0) { status = 2; done = 1; ++it; break; }

		// ---- pivot (v4:331-356)
		book1_phase<T>(d, sh, p, q, me, G);
		grid_barrier(ctl, epoch);
		book2_phase<T>(d, sh, p, q, me, G);
		if (me == 0 && threadIdx.x == 0 && pivots < d.trace_cap) d.trace[pivots] = make_int2((int)p, (int)q);
		pending = 1;
		++pivots;
		++it;
		grid_barrier(ctl, epoch);
	}

	if (me == 0) {
		const double z = objective<T>(d, sh);
		if (threadIdx.x == 0) {
			ctl->iter = it; ctl->pivots = pivots; ctl->pending = pending;
			ctl->status = status; ctl->done = done;
			ctl->p = p; ctl->q = q; ctl->min_e = min_e; ctl->z = z;
		}
	}
}

// ---------------------------------------------------------------- sharded (multi-GPU) loop
//
// One process per GPU, the same persistent kernel on every rank.  B^-1 is row-block
// sharded, A column-block sharded, every O(m) vector (y, x_b, c_b, alpha, E_q, row_q,
// b_ixs) is replicated, so all ranks take identical decisions and the per-row / per-column
// sums are exactly the single-GPU ones (results are bit-identical for every GPU count).
// Three exchanges per pivot, all direct peer stores over NVLink into the receivers'
// mailboxes followed by a flag barrier — no NCCL call and no host on the data path:
//   X1  pricing candidate (value, index), 16 B per rank            (replaces the global cub ArgMin)
//   X2  alpha slices of the local row block, m/R values per rank   (after the fused update+FTRAN)
//   X3  row q of B^-1 from its owner, m values                     (replaces cublasScopy, v4:331)
// The entering column a_p is pulled from its owner's A shard with peer loads.

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

// grid barrier whose release/acquire fences are system scope (peer stores must be
// visible to the other GPUs before the flag that announces them)
__device__ __forceinline__ void grid_barrier_sys(Ctl* ctl, unsigned long long& epoch) {
	epoch += gridDim.x;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		if (gridDim.x > 1) {
			atomicAdd(&ctl->bar, 1ULL);
			while (*((volatile unsigned long long*)&ctl->bar) < epoch) { }
		}
		__threadfence_system();
	}
	__syncthreads();
}

// cross-GPU flag barrier, executed by CTA 0 between two grid_barrier_sys calls:
// thread r signals rank r (store into ITS mailbox) and waits for rank r's signal in ours.
template <typename T>
__device__ __forceinline__ void xsync(const Dev<T>& d, unsigned long long xepoch) {
	if (blockIdx.x != 0) return;
	if (threadIdx.x < d.nranks) {
		const int r = threadIdx.x;
		__threadfence_system();
		st_release_sys(&reinterpret_cast<XHdr*>(d.mbox_peer[r])->xflag[d.rank], xepoch);
		const unsigned long long* mine = &reinterpret_cast<XHdr*>(d.mbox_peer[d.rank])->xflag[r];
		const unsigned long long t0 = globaltimer_ns();
		while (ld_acquire_sys(mine) < xepoch) {
			if (globaltimer_ns() - t0 > 20000000000ULL) { d.ctl->bad = 1; break; }   // 20 s: a peer died
		}
		__threadfence_system();
	}
	__syncthreads();
}

// full exchange step: everything stored to peers before it is visible to them after it
template <typename T>
__device__ __forceinline__ void xbarrier(const Dev<T>& d, unsigned long long& epoch, unsigned long long& xepoch) {
	grid_barrier_sys(d.ctl, epoch);
	xsync(d, ++xepoch);
	grid_barrier_sys(d.ctl, epoch);
}

// X1: CTA 0 reduces this rank's per-CTA candidates and stores the winner into every mailbox
template <typename T>
__device__ void push_price_candidate(const Dev<T>& d, Smem& sh) {
	if (blockIdx.x != 0) return;
	double v; long long i;
	reduce_cands(d.cand, gridDim.x, v, i, sh);
	if (threadIdx.x < d.nranks) {
		Cand* dst = &reinterpret_cast<XHdr*>(d.mbox_peer[threadIdx.x])->candx[d.rank];
		dst->val = v;
		dst->idx = i;
	}
}

// pull the entering column from its owner's A shard into the local staging vector
template <typename T>
__device__ void fetch_column(const Dev<T>& d, long long p, int part, int nparts) {
	using M = Mem<T>;
	using V = typename VecT<T>::V;
	constexpr int VN = VecT<T>::N;
	int o = 0;
	while (o + 1 < d.nranks && p >= d.colstart[o + 1]) ++o;
	const T* src = d.A_peer[o] + (p - d.colstart[o]) * d.ld;
	for (long long i = ((long long)part * NT + threadIdx.x) * VN; i < d.ld; i += (long long)nparts * NT * VN)
		*reinterpret_cast<V*>(d.acol + i) = M::ld_nc(src + i);
}

// X2: alpha of the local rows = sum of the chunk partials (left to right), stored into every rank's alpha
template <typename T>
__device__ void push_alpha_phase(const Dev<T>& d, int part, int nparts) {
	const long long nloc = d.ldb;
	for (long long i = (long long)part * NT + threadIdx.x; i < nloc; i += (long long)nparts * NT) {
		T a = __ldcg(d.alpha_part + i);
#pragma unroll 8
		for (int ck = 1; ck < d.nchunk; ++ck) a = a + __ldcg(d.alpha_part + (long long)ck * d.ldb + i);
		for (int r = 0; r < d.nranks; ++r)
			reinterpret_cast<T*>(d.mbox_peer[r] + sizeof(XHdr))[d.row0 + i] = a;
	}
}

// book1, sharded: E_q and the c_b.E_q slice partials on every rank (replicated data);
// X3: the owner of row q gathers it from its B^-1 block and stores it into every rank's row_q.
template <typename T>
__device__ void book1a_sharded(const Dev<T>& d, Smem& sh, long long p, long long q, int part, int nparts) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const T alpha_q = d.alpha[q];
	const T c_p = d.c[p];
	const bool owner = q >= d.row0 && q < d.row0 + d.ldb;
	for (long long s = part; s < d.nslice; s += nparts) {
		const long long i = s * SLICE + tid;
		T t2 = T(0);
		if (i < d.m) {
			const T eq = (i != q) ? (-d.alpha[i] / alpha_q) : (T)(1.0 / (double)alpha_q - 1.0);
			d.E_q[i] = eq;
			T cb = d.c_b[i];
			if (i == q) { d.ctl->c_b_q = (double)cb; cb = c_p; }
			t2 = fma_t(cb, eq, T(0));
			if (owner) {
				const T rq = d.B[(q - d.row0) + i * d.ldb];
				for (int r = 0; r < d.nranks; ++r)
					(reinterpret_cast<T*>(d.mbox_peer[r] + sizeof(XHdr)) + d.ld)[i] = rq;
			}
		}
		t2 = warp_butterfly_sum(t2);
		__syncthreads();
		if (lane == 0) sh.dsum[1][warp] = (double)t2;
		__syncthreads();
		if (tid == 0) {
			T a = T(0);
#pragma unroll
			for (int w = 0; w < NWARP; ++w) a = a + (T)sh.dsum[1][w];
			d.dpart[(long long)d.nslice + s] = a;
		}
	}
}

// row_q.b slice partials once row_q has arrived
template <typename T>
__device__ void book1b_sharded(const Dev<T>& d, Smem& sh, int part, int nparts) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (long long s = part; s < d.nslice; s += nparts) {
		const long long i = s * SLICE + tid;
		T t1 = i < d.m ? fma_t(d.row_q[i], d.b[i], T(0)) : T(0);
		t1 = warp_butterfly_sum(t1);
		__syncthreads();
		if (lane == 0) sh.dsum[0][warp] = (double)t1;
		__syncthreads();
		if (tid == 0) {
			T a = T(0);
#pragma unroll
			for (int w = 0; w < NWARP; ++w) a = a + (T)sh.dsum[0][w];
			d.dpart[s] = a;
		}
	}
}

template <typename T, int WC>
__global__ void __launch_bounds__(NT, MIN_CTAS) simplex_persistent_sharded(Dev<T> d) {
	__shared__ Smem sh;
	Ctl* ctl = d.ctl;
	const int G = gridDim.x, me = blockIdx.x;
	unsigned long long epoch = 0;
	unsigned long long xepoch = ctl->xepoch;

	long long it = ctl->iter, pivots = ctl->pivots;
	const long long it_end = ctl->it_end;
	int pending = ctl->pending;
	int status = 0, done = 0, bad = 0;
	long long p = ctl->p, q = ctl->q;
	double min_e = ctl->min_e;
	const XHdr* mine = reinterpret_cast<const XHdr*>(d.mbox_peer[d.rank]);

	while (it < it_end) {
		// ---- pricing over the local column block, X1
		price_phase<T>(d, sh, me, G);
		grid_barrier_sys(ctl, epoch);
		push_price_candidate<T>(d, sh);
		xbarrier(d, epoch, xepoch);
		if (*(volatile int*)&ctl->bad) { bad = 1; break; }
		min_e = CUDART_INF; p = LLONG_MAX;
		for (int r = 0; r < d.nranks; ++r) {       // same fixed order on every rank and CTA
			const double cv = __ldcg(&mine->candx[r].val);
			const long long ci = __ldcg(&mine->candx[r].idx);
			if (cand_better(cv, ci, min_e, p)) { min_e = cv; p = ci; }
		}
		if (min_e >= -d.eps) { status = 1; done = 1; ++it; break; }

		// ---- entering column from its owner, then the fused update + FTRAN on the local rows
		const bool dense = p < d.ns;
		if (dense) {
			fetch_column<T>(d, p, me, G);
			grid_barrier(ctl, epoch);
		}
		if (pending) update_ftran_phase<T, WC, true, true>(d, sh, dense ? d.acol : nullptr, p - d.ns, me, G);
		else         update_ftran_phase<T, WC, false, true>(d, sh, dense ? d.acol : nullptr, p - d.ns, me, G);
		pending = 0;
		grid_barrier(ctl, epoch);

		// ---- X2: alpha slices to every rank, then the ratio test on the full alpha (replicated)
		push_alpha_phase<T>(d, me, G);
		xbarrier(d, epoch, xepoch);
		if (*(volatile int*)&ctl->bad) { bad = 1; break; }
		ratio_phase<T, false>(d, sh, me, G);
		grid_barrier(ctl, epoch);
		double th;
		reduce_cands(d.cand, G, th, q, sh);
		const long long elig = reduce_counts(d.cnt, G, sh);
		if (elig == 0) { status = 2; done = 1; ++it; break; }

		// ---- pivot: E_q everywhere, X3 row q from its owner, then the replicated O(m) updates
		book1a_sharded<T>(d, sh, p, q, me, G);
		xbarrier(d, epoch, xepoch);
		if (*(volatile int*)&ctl->bad) { bad = 1; break; }
		book1b_sharded<T>(d, sh, me, G);
		grid_barrier(ctl, epoch);
		book2_phase<T>(d, sh, p, q, me, G);
		if (me == 0 && threadIdx.x == 0 && pivots < d.trace_cap) d.trace[pivots] = make_int2((int)p, (int)q);
		pending = 1;
		++pivots;
		++it;
		grid_barrier(ctl, epoch);
	}

	if (me == 0) {
		const double z = objective<T>(d, sh);
		if (threadIdx.x == 0) {
			ctl->iter = it; ctl->pivots = pivots; ctl->pending = pending;
			ctl->status = status; ctl->done = done; ctl->xepoch = xepoch;
			ctl->p = p; ctl->q = q; ctl->min_e = min_e; ctl->z = z;
			if (bad) ctl->bad = 1;
		}
	}
}

// ---------------------------------------------------------------- stand-alone phase kernels
// (unit tests, one-launch-per-phase mode, sharded driver).  p / q are passed by
// the host or read from ctl by the caller.

template <typename T>
__global__ void __launch_bounds__(NT) k_price(Dev<T> d) {
	__shared__ Smem sh;
	price_phase<T>(d, sh, blockIdx.x, gridDim.x);
}

// final argmin over the per-CTA candidates -> ctl->p / ctl->min_e (kind 0) or ctl->q + eligible (kind 1)
template <typename T>
__global__ void __launch_bounds__(NT) k_pick(Dev<T> d, int ncand, int kind) {
	__shared__ Smem sh;
	double v; long long i;
	reduce_cands(d.cand, ncand, v, i, sh);
	if (kind == 0) {
		if (threadIdx.x == 0) { d.ctl->p = i; d.ctl->min_e = v; }
	} else {
		const long long el = reduce_counts(d.cnt, ncand, sh);
		if (threadIdx.x == 0) { d.ctl->q = i; d.ctl->min_e = v; d.cnt[ncand] = el; }
	}
}

template <typename T, int WC, bool UPDATE, bool FTRAN>
__global__ void __launch_bounds__(NT) k_update_ftran(Dev<T> d, long long p) {
	__shared__ Smem sh;
	update_ftran_phase<T, WC, UPDATE, FTRAN>(d, sh, p < d.ns ? d.A + p * d.ld : nullptr, p - d.ns, blockIdx.x, gridDim.x);
}

template <typename T>
__global__ void __launch_bounds__(NT) k_ratio(Dev<T> d) {
	__shared__ Smem sh;
	ratio_phase<T, true>(d, sh, blockIdx.x, gridDim.x);
}

template <typename T>
__global__ void __launch_bounds__(NT) k_book1(Dev<T> d, long long p, long long q) {
	__shared__ Smem sh;
	book1_phase<T>(d, sh, p, q, blockIdx.x, gridDim.x);
}

template <typename T>
__global__ void __launch_bounds__(NT) k_book2(Dev<T> d, long long p, long long q) {
	__shared__ Smem sh;
	book2_phase<T>(d, sh, p, q, blockIdx.x, gridDim.x);
}

template <typename T>
__global__ void __launch_bounds__(NT) k_objective(Dev<T> d) {
	__shared__ Smem sh;
	const double z = objective<T>(d, sh);
	if (threadIdx.x == 0) d.ctl->z = z;
}

// ---------------------------------------------------------------- setup kernels

// slack-basis initial state (v4:272-277): B^-1 = I, c_b = c[n-m..n), x_b = b,
// b_ixs[j] = n-m+j, y = c_b; padding rows zero.
template <typename T>
__global__ void k_reset(Dev<T> d) {
	const long long tot = d.ldb * d.m;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += stride) {
		const long long i = d.row0 + e % d.ldb, j = e / d.ldb;
		d.B[e] = (i == j) ? T(1) : T(0);
	}
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d.ld; i += stride) {
		const bool in = i < d.m;
		const T cb = in ? d.c[d.n - d.m + i] : T(0);
		d.c_b[i] = cb;
		d.y[i] = cb;
		d.x_b[i] = in ? d.b[i] : T(0);
		d.alpha[i] = T(0);
		d.E_q[i] = T(0);
		d.row_q[i] = T(0);
		if (in) d.b_ixs[i] = (int)(d.n - d.m + i);
	}
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double u01(uint64_t seed, uint64_t stream, uint64_t idx) {
	uint64_t z = mix64(seed * 0x9E3779B97F4A7C15ULL + stream * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL);
	z = mix64(z + idx * 0x9E3779B97F4A7C15ULL);
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// dense synthetic LP (same numbers as oracle/lpgen_dense_*): A_s ~ U(0,1),
// b = (ns/2) U(1,2), c_s ~ U(0.5,1.5), slack costs 0
template <typename T>
__global__ void k_generate_dense(T* A, T* b, T* c, long long m, long long n, long long ns, long long ld, uint64_t seed,
		long long col0, long long nsl) {
	const long long stride = (long long)gridDim.x * blockDim.x;
	const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long tot = ld * nsl;           // this rank's column block [col0, col0 + nsl)
	for (long long e = g; e < tot; e += stride) {
		const long long i = e % ld, j = col0 + e / ld;
		A[e] = i < m ? (T)u01(seed, 0, (uint64_t)(i * ns + j)) : T(0);
	}
	for (long long i = g; i < ld; i += stride) b[i] = i < m ? (T)(0.5 * (double)ns * (1.0 + u01(seed, 1, (uint64_t)i))) : T(0);
	for (long long j = g; j < n; j += stride) c[j] = j < ns ? (T)(0.5 + u01(seed, 2, (uint64_t)j)) : T(0);
}

// zero the padding rows [m, ld) of an ld x ncols column-major matrix
template <typename T>
__global__ void k_zero_pad(T* X, long long m, long long ld, long long ncols) {
	const long long pad = ld - m;
	const long long tot = pad * ncols;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += stride) {
		const long long i = m + e % pad, j = e / pad;
		X[i + j * ld] = T(0);
	}
}

} // namespace b200lp
