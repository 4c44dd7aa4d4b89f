/*
 * b200lp_io — LP file input/output of the B200 dense revised-simplex engine (host code, no GPU needed).
 *
 * Replaces the reference's parser
 *     load_matrix_impl(std::ifstream&, T* a, int m, int n, const char* name)   src/v4_cub_reduction.cu:94-104
 *     header read `file >> m >> n`, m <= n check                                src/v4_cub_reduction.cu:401-405
 * and adds a binary twin of the same content (the text of the m=8192 case is
 * 134 M tokens; `operator>>` per token takes minutes, a parallel from_chars pass seconds, the
 * binary file a single read).
 *
 * Text format (input/sample.txt): `m n`, then A as m rows of n numbers (row-major text), then
 * b (m), then c (n); n counts ALL columns, the slack identity block is the last m; anything
 * after the last number is ignored.  In memory A is column-major (v4:59-60, 98).
 *
 * Binary format (".b200lp"): 64-byte header { char magic[8] = "B200LP1\0"; int32 dtype
 * (B200LP_F32|F64); int32 reserved; int64 m; int64 n; 32 bytes zero }, then A column-major
 * (m*n), b (m), c (n) in the header's dtype, little endian.
 */
#ifndef B200LP_IO_H
#define B200LP_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int32_t dtype;   /* B200LP_F32 (0) or B200LP_F64 (1) */
	int32_t reserved;
	int64_t m, n;
	void* A;         /* column-major m x n */
	void* b;         /* m */
	void* c;         /* n */
} b200lp_problem;

/* Reads a text or binary LP file (detected by the magic) into host memory of `dtype`
 * (a binary file of the other dtype is converted).  `pinned` != 0 allocates A with
 * cudaMallocHost like the reference (v4:408-414) so the upload runs at full PCIe speed;
 * falls back to pageable memory without a CUDA device.  Errors use the reference's
 * messages ("Could not open <path>.", "Either failed to read m and n, or m > n.",
 * "Failed to read (i,j) for A") through b200lp_last_error().  Returns B200LP_OK or
 * B200LP_ERR_ARG. */
int b200lp_read_lp(const char* path, int32_t dtype, int32_t pinned, b200lp_problem* out);

/* `threads` <= 0: all hardware threads.  Shortest round-trip decimal per number. */
int b200lp_write_lp_text(const char* path, const b200lp_problem* p);
int b200lp_write_lp_binary(const char* path, const b200lp_problem* p);

void b200lp_free_problem(b200lp_problem* p);

#ifdef __cplusplus
}
#endif
#endif /* B200LP_IO_H */
