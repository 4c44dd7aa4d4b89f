#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the reference's own v4 solver into oracle/_ref/
# (git-ignored; the binaries travel to the GPU box, the sources never enter the repo).
#
#   _ref/v4_stock.out       src/v4_cub_reduction.cu exactly as shipped, Makefile flags + -arch
#   _ref/libv4ref_f64.so    v4 + the minimal patch list below, real = double
#   _ref/libv4ref_f32.so    same patches, real = float
#   _ref/v4_cli_f{32,64}.out the patched copy as a program with its own main() (V4_EPS / V4_MAX_ITER from the environment)
#   _ref/v4_shim.out        the STOCK source with only solve() (v4:219-380) cut out + integration/v4_b200.inc + -lb200lp:
#                           the reference's unmodified main() driving the B200 engine
#   _ref/v4_shim_env_f{32,64}.out  the same on the patched copy (run-time EPS / MAX_ITER), for LPs beyond 5 iterations
#
# Patch list (every item except P1 is arithmetic-neutral):
#   P1 real = double and cublasS* -> cublasD*                     (v4:12, 289..365)   [f64 only]
#   P2 EPS / MAX_ITER become run-time variables                   (v4:18-19)
#   P3 2-D init kernels launched with ceil(m/16) x ceil(n/16) grids (v4:272, 279; the
#      shipped grids are sized for 256-wide blocks and leave B_inv / D uninitialised
#      beyond 16*ceil(m/256) rows)
#   P4 y_aug[1..] copy length n - m -> m                          (v4:277)
#   P5 device-resident one / zero for the DEVICE-pointer-mode cuBLAS calls
#      (v4:289-290, 307-308, 333 pass host stack addresses)
#   P6 pivot-trace hook after the q read-back (v4:325), iteration count export,
#      "# Iteration" printing silenced (v4:287)
#   P8 z / x_b / b_ixs read back for every status when ref_always_readback is set (v4:363 does it on
#      OptimumFound only); the same cublasDdot + two copies, after the loop, so the loop is untouched
#   P7 64-bit sizes and element offsets (v4:33, 60, 77, 86, 246, 248, 269, 308): the shipped
#      `int` products m*n, (m+1)*n and p*m overflow at m = 32768, n = 65536 (config C4)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
SRC="$REF/src/v4_cub_reduction.cu"
OUT="$HERE/_ref"
if [ ! -f "$SRC" ]; then
	echo "make_ref: $SRC not found (the reference is only mounted in the build container); keeping prebuilt files" >&2
	exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
NVCC="${NVCC:-nvcc}"
CCBIN="${CCBIN:-/usr/bin/g++-13}"
ARCH="-gencode arch=compute_100a,code=sm_100a"

# stock build: the reference Makefile's command line plus the target architecture
$NVCC --std=c++20 $ARCH "$SRC" -o "$OUT/v4_stock.out" -ccbin "$CCBIN" -lcublas

patch_src() { # $1 = S|D  $2 = out file
	sed -E \
		-e '18s/constexpr real EPS/static real EPS/' \
		-e '33s/int size;/long long size;/' \
		-e '60s/constexpr int R2C\(int i, int j, int m\)/constexpr long long R2C(long long i, long long j, long long m)/' \
		-e '77s/int n, const char/size_t n, const char/' \
		-e '86s/int size, cudaMemcpyKind/size_t size, cudaMemcpyKind/' \
		-e '246s/\{d_A, m \* n\}/{d_A, (long long)m * n}/; 246s/\{d_B_inv, m \* m\}/{d_B_inv, (long long)m * m}/' \
		-e '248s/\{d_D, \(m \+ 1\) \* n\}/{d_D, (long long)(m + 1) * n}/' \
		-e '269s/A, m \* n,/A, (long long)m * n,/' \
		-e '308s/d_A \+ p \* m/d_A + (long long)p * m/' \
		-e '19s/constexpr int MAX_ITER/static int MAX_ITER/' \
		-e '243s/$/ real *d_one, *d_zero; cudaMalloc(\&d_one, sizeof(real)); cudaMalloc(\&d_zero, sizeof(real)); cudaMemcpy(d_one, \&one, sizeof(real), cudaMemcpyHostToDevice); cudaMemcpy(d_zero, \&zero, sizeof(real), cudaMemcpyHostToDevice);/' \
		-e '272s/dim3\(blocks_for_m, blocks_for_m\)/dim3((m + BS_2D - 1) \/ BS_2D, (m + BS_2D - 1) \/ BS_2D)/' \
		-e '277s/n - m/m/' \
		-e '279s/dim3\(blocks_for_n, blocks_for_m \+ 1\)/dim3((n + BS_2D - 1) \/ BS_2D, (m + BS_2D - 1) \/ BS_2D)/' \
		-e '287s/print_iteration\(i\);/if (!ref_quiet) print_iteration(i);/' \
		-e '290s/&one/d_one/; 290s/&zero/d_zero/' \
		-e '308s/&one/d_one/; 308s/&zero/d_zero/' \
		-e '333s/&one/d_one/' \
		-e '325s/$/ if (i < ref_trace_cap) { ref_trace[2 * i] = p; ref_trace[2 * i + 1] = q; }/' \
		-e '363s/if \(status == SolveStatus::OptimumFound\)/if (status == SolveStatus::OptimumFound || ref_always_readback)/' \
		-e '360s/$/ ref_iterations = (status == SolveStatus::MaxIter) ? i : i + 1; cudaFree(d_one); cudaFree(d_zero);/' \
		"$SRC" > "$2"
	if [ "$1" = "D" ]; then
		sed -i -E -e '12s/using real = float;/using real = double;/' \
			-e 's/cublasS(gemm|gemv|copy|ger|dot|axpy)\(/cublasD\1(/g' "$2"
	fi
	# every edit must have landed, otherwise the reference moved under us
	grep -q 'static real EPS' "$2" && grep -q 'static int MAX_ITER' "$2" && grep -q 'd_one, d_y_aug' "$2" \
		&& grep -q 'ref_trace\[2 \* i\]' "$2" && grep -q 'ref_iterations =' "$2" \
		&& grep -q 'if (!ref_quiet)' "$2" && grep -q 'd_c_b, m, cudaMemcpyDeviceToDevice' "$2" \
		&& [ "$(grep -c 'BS_2D - 1' "$2")" = "2" ] && [ "$(grep -c '(long long)' "$2")" -ge 4 ] \
 		&& grep -q 'ref_always_readback)' "$2" \
		&& grep -q 'long long R2C' "$2" && grep -q 'long long size;' "$2" || { echo "make_ref: patch did not apply cleanly" >&2; exit 1; }
}

for V in D S; do
	P="$TMP/v4_patched_$V.cu"
	patch_src "$V" "$P"
	if [ "$V" = "D" ]; then NAME=libv4ref_f64.so; else NAME=libv4ref_f32.so; fi
	$NVCC --std=c++20 $ARCH -O2 -shared -Xcompiler -fPIC -DREF_SOURCE="\"$P\"" "$HERE/ref_harness.cu" \
		-o "$OUT/$NAME" -ccbin "$CCBIN" -lcublas
done

# ---- the drop-in, compiled: reference main() + integration/v4_b200.inc + libb200lp.so (no cuBLAS on the link line)
ROOT="$(cd "$HERE/.." && pwd)"
LIBDIR="$ROOT/simplex_method_gpu_b200"
cut_solve() { # $1 = source, $2 = out: v4:219-380 (the old solve()) replaced by its forward declaration
	sed -E -e '219,380d' "$1" | sed -e '218a\
std::pair<real, SolveStatus> solve(real* A, real* b, real* c, real* x_b, int* b_ixs, int m, int n, TimeStruct\& t);' > "$2"
	grep -q 'TimeStruct& t);' "$2" && ! grep -q 'cublasCreate' "$2" && grep -q 'int main(int argc' "$2" \
		|| { echo "make_ref: cutting solve() out did not apply cleanly" >&2; exit 1; }
}
if [ -f "$LIBDIR/libb200lp.so" ]; then
	LINK="-I$ROOT/include -L$LIBDIR -lb200lp -Xlinker -rpath -Xlinker \$ORIGIN/../../simplex_method_gpu_b200"
	cut_solve "$SRC" "$TMP/v4_cut_stock.cu"
	$NVCC --std=c++20 $ARCH -DREF_SOURCE="\"$TMP/v4_cut_stock.cu\"" "$HERE/v4_shim.cu" -o "$OUT/v4_shim.out" -ccbin "$CCBIN" $LINK
	for V in D S; do
		if [ "$V" = "D" ]; then SFX=f64; else SFX=f32; fi
		cut_solve "$TMP/v4_patched_$V.cu" "$TMP/v4_cut_$V.cu"
		$NVCC --std=c++20 $ARCH -DREF_ENV -DREF_SOURCE="\"$TMP/v4_cut_$V.cu\"" "$HERE/v4_shim.cu" -o "$OUT/v4_shim_env_$SFX.out" -ccbin "$CCBIN" $LINK
		$NVCC --std=c++20 $ARCH -O2 -DREF_SOURCE="\"$TMP/v4_patched_$V.cu\"" "$HERE/ref_cli.cu" -o "$OUT/v4_cli_$SFX.out" -ccbin "$CCBIN" -lcublas
	done
else
	echo "make_ref: libb200lp.so not built yet, skipping the drop-in binaries" >&2
fi
ls -la "$OUT"
