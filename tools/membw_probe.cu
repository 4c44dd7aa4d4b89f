// Dev tool (GPU box): what can a streaming kernel reach on this B200?  Roofline context for DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bin/membw_probe tools/membw_probe.cu
//   bin/membw_probe [MiB]
// Modes (296 CTAs x 256 threads, 2 CTAs per SM, like the engine):
//   read   TMA bulk loads into a 4-stage ring, data summed                       (pricing-like, read only)
//   rmw    ld.global.v2.f64 / st.global.v2.f64, 8 loads in flight per thread     (the engine's update pass)
//   tma    TMA bulk load into a ring, modify in shared memory, TMA bulk store    (candidate for round 2)
// All in place over one buffer larger than L2.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int NT = 256;
constexpr int STAGE = 16384;           // bytes per stage
constexpr int NSTAGE = 4;

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
	asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_ld(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tma_st(void* dst, const void* src, unsigned bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}

// mode 0: read only through a TMA ring
__global__ void __launch_bounds__(NT, 2) k_read(const double* buf, size_t nchunks, double* sink) {
	extern __shared__ __align__(128) unsigned char ring[];
	__shared__ unsigned long long full[NSTAGE];
	const int tid = threadIdx.x;
	if (tid == 0) { for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	const size_t first = blockIdx.x, step = gridDim.x;
	size_t issued = first;
	auto issue = [&](int s) {
		if (issued < nchunks) { mbar_expect(&full[s], STAGE); tma_ld(ring + s * STAGE, (const char*)buf + issued * STAGE, STAGE, &full[s]); }
		issued += step;
	};
	if (tid == 0) for (int s = 0; s < NSTAGE - 1; ++s) issue(s);
	double acc = 0;
	int st = 0; unsigned ph = 0;
	for (size_t c = first; c < nchunks; c += step) {
		if (tid == 0) issue((st + NSTAGE - 1) % NSTAGE);
		mbar_wait(&full[st], ph);
		const double2* p = reinterpret_cast<const double2*>(ring + st * STAGE);
#pragma unroll
		for (int k = 0; k < STAGE / 16 / NT; ++k) { double2 v = p[tid + k * NT]; acc += v.x + v.y; }
		__syncthreads();
		if (++st == NSTAGE) { st = 0; ph ^= 1; }
	}
	if (acc == 12345.678) sink[0] = acc;
}

// mode 1: register-staged read-modify-write, 8 x 16 B in flight per thread
__global__ void __launch_bounds__(NT, 2) k_rmw(double* buf, size_t nvec) {
	const size_t stride = (size_t)gridDim.x * NT;
	double2* p = reinterpret_cast<double2*>(buf);
	for (size_t i0 = (size_t)blockIdx.x * NT + threadIdx.x; i0 < nvec; i0 += stride * 8) {
		double2 v[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) if (i0 + u * stride < nvec)
			asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(p + i0 + u * stride) : "memory");
#pragma unroll
		for (int u = 0; u < 8; ++u) if (i0 + u * stride < nvec) {
			v[u].x = fma(v[u].x, 1.0000001, 1e-9); v[u].y = fma(v[u].y, 1.0000001, 1e-9);
			asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p + i0 + u * stride), "d"(v[u].x), "d"(v[u].y) : "memory");
		}
	}
}

// mode 2: TMA load ring -> modify in shared memory -> TMA bulk store
__global__ void __launch_bounds__(NT, 2) k_tma(double* buf, size_t nchunks) {
	extern __shared__ __align__(128) unsigned char ring[];
	__shared__ unsigned long long full[NSTAGE];
	const int tid = threadIdx.x;
	if (tid == 0) { for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	const size_t first = blockIdx.x, step = gridDim.x;
	size_t issued = first;
	auto issue = [&](int s) {
		if (issued < nchunks) { mbar_expect(&full[s], STAGE); tma_ld(ring + s * STAGE, (const char*)buf + issued * STAGE, STAGE, &full[s]); }
		issued += step;
	};
	if (tid == 0) for (int s = 0; s < NSTAGE - 2; ++s) issue(s);
	int st = 0; unsigned ph = 0;
	for (size_t c = first; c < nchunks; c += step) {
		if (tid == 0) {
			// the stage about to be refilled was handed to a bulk store two iterations ago: wait until it has been read
			asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
			issue((st + NSTAGE - 2) % NSTAGE);
		}
		mbar_wait(&full[st], ph);
		double2* p = reinterpret_cast<double2*>(ring + st * STAGE);
#pragma unroll
		for (int k = 0; k < STAGE / 16 / NT; ++k) {
			double2 v = p[tid + k * NT];
			v.x = fma(v.x, 1.0000001, 1e-9); v.y = fma(v.y, 1.0000001, 1e-9);
			p[tid + k * NT] = v;
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();
		if (tid == 0) {
			tma_st((char*)buf + c * STAGE, ring + st * STAGE, STAGE);
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
		if (++st == NSTAGE) { st = 0; ph ^= 1; }
	}
	if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
	const size_t mib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 4096;
	const size_t bytes = mib << 20, nchunks = bytes / STAGE, nvec = bytes / 16;
	double *buf, *sink;
	CK(cudaMalloc(&buf, bytes));
	CK(cudaMalloc(&sink, 8));
	CK(cudaMemset(buf, 0, bytes));
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	const int grid = 2 * sms, smem = NSTAGE * STAGE;
	CK(cudaFuncSetAttribute(k_read, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	const char* names[3] = {"read  (TMA ring, read only)", "rmw   (LDG/STG, 8 x 16 B in flight)", "tma   (TMA load + TMA store)"};
	for (int mode = 0; mode < 3; ++mode) {
		float best = 1e30f;
		for (int rep = 0; rep < 6; ++rep) {
			cudaEventRecord(a);
			if (mode == 0) k_read<<<grid, NT, smem>>>(buf, nchunks, sink);
			else if (mode == 1) k_rmw<<<grid, NT>>>(buf, nvec);
			else k_tma<<<grid, NT, smem>>>(buf, nchunks);
			cudaEventRecord(b);
			CK(cudaEventSynchronize(b));
			CK(cudaGetLastError());
			float ms; cudaEventElapsedTime(&ms, a, b);
			if (rep > 0 && ms < best) best = ms;
		}
		const double traffic = (mode == 0 ? 1.0 : 2.0) * (double)bytes;
		printf("%-40s %8.3f ms  %8.1f GB/s  (%zu MiB, grid %d)\n", names[mode], best, traffic / best / 1e6, mib, grid);
	}
	return 0;
}
