// bw_kernels.cu -- HBM-bound kernels of the training step: BatchNorm statistics / apply / backward, ReLU,
// residual join, max / average pooling, softmax cross-entropy, fused Adam, weight re-layout.
//
// Roofline: HBM bandwidth (MEASURED_PEAKS.json hbm_gbs).  Every tensor pass is a flat, fully coalesced
// 128-bit grid-stride stream over fp32 (4 elements per access) or bf16 (8 per access) tensors with fp32 arithmetic;
// per-channel quantities ride in registers because the grid stride is a multiple of the vectors per row, so a thread
// always sees the same channels (NHWC: C is the contiguous axis).
// Reference semantics: resnet.cu:289-342 (BN fwd), 350-426 (BN bwd), 433-494 (max pool), 500-542 (avg pool),
// 545-602 (ReLU, softmax, CE), 605-662 (Adam).
#include "common.cuh"
#include <stdarg.h>
#include <math.h>

namespace rb {

// ------------------------------------------------------------------------------------------- errors
long long g_launches = 0;
static char g_err[512];
static bool g_has_err = false;
void set_error(const char *fmt, ...) {
	if (g_has_err) return;
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	g_has_err = true;
	fprintf(stderr, "[resnet_b200] error: %s\n", g_err);
}
const char *last_error() { return g_has_err ? g_err : ""; }
void clear_error() { g_has_err = false; g_err[0] = 0; }
bool has_error() { return g_has_err; }
bool trace_on() {
	static const bool on = getenv("RESNET_B200_TRACE") != nullptr;
	return on;
}
int pdl_mode() {
	static const int mode = getenv("RESNET_B200_PDL") ? atoi(getenv("RESNET_B200_PDL")) : 2;
	return mode;
}

// ------------------------------------------------------------------------------------------- helpers
constexpr int kThreads = 256;

// Storage types: float (VEC = 4 or 1 elements per access) and bf16 (VEC = 8): one 128-bit access per thread either way; the
// arithmetic is always fp32.  i indexes VECTORS.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
	uint32_t r;
	asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
	return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

template <typename T, int VEC> __device__ __forceinline__ void ldv(const T *p, long long i, float (&v)[VEC]) {
	if constexpr (sizeof(T) == 2) {
		static_assert(sizeof(T) != 2 || VEC == 8, "bf16 tensors are accessed 8 elements at a time");
		const uint4 t = reinterpret_cast<const uint4 *>(p)[i];
		v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
		v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
	} else if constexpr (VEC == 4) {
		float4 t = reinterpret_cast<const float4 *>(p)[i];
		v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
	} else v[0] = p[i];
}
template <typename T, int VEC> __device__ __forceinline__ void stv(T *p, long long i, const float (&v)[VEC]) {
	if constexpr (sizeof(T) == 2)
		reinterpret_cast<uint4 *>(p)[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
	else if constexpr (VEC == 4) reinterpret_cast<float4 *>(p)[i] = make_float4(v[0], v[1], v[2], v[3]);
	else p[i] = v[0];
}
// raw 128-bit loads first, unpacking later: lets a kernel put U loads per stream in flight before any of them is consumed
template <int VEC> struct RawOf { using type = uint4; };
template <> struct RawOf<1> { using type = float; };
template <typename T, int VEC> __device__ __forceinline__ typename RawOf<VEC>::type ldraw(const T *p, long long i) {
	if constexpr (VEC == 1) return p[i];
	else return reinterpret_cast<const uint4 *>(p)[i];
}
template <typename T, int VEC> __device__ __forceinline__ void unpack(const typename RawOf<VEC>::type &r, float (&v)[VEC]) {
	if constexpr (VEC == 1) v[0] = r;
	else if constexpr (sizeof(T) == 2) {
		v[0] = bf16_lo(r.x); v[1] = bf16_hi(r.x); v[2] = bf16_lo(r.y); v[3] = bf16_hi(r.y);
		v[4] = bf16_lo(r.z); v[5] = bf16_hi(r.z); v[6] = bf16_lo(r.w); v[7] = bf16_hi(r.w);
	} else {
		v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
	}
}
// loads per stream kept in flight per thread by the BatchNorm-backward kernels (the bf16 variants carry 8 channels of
// per-channel state in registers, so they afford fewer)
// and the resident blocks per SM they are compiled for: fp32 3 loads x 3 blocks, bf16 4 loads x 2 blocks (8 channels of per-channel
// state per thread): 256 threads x U x 3 streams x 16 B x blocks = 110 / 98 KB in flight per SM, above the ~55 KB that 6.5 TB/s needs
template <int VEC> struct BatchOf { static constexpr int U = (VEC == 8) ? 4 : 3; static constexpr int kBlocks = (VEC == 8) ? 2 : 3; };
static int bn_bwd_blocks_per_sm(int bf16) { return bf16 ? 2 : 3; }

// scalar element access (small kernels)
template <typename T> __device__ __forceinline__ float ld1(const T *p, long long i) {
	if constexpr (sizeof(T) == 2) return __uint_as_float((uint32_t)reinterpret_cast<const uint16_t *>(p)[i] << 16);
	else return p[i];
}
template <typename T> __device__ __forceinline__ void st1(T *p, long long i, float v) {
	if constexpr (sizeof(T) == 2) reinterpret_cast<uint16_t *>(p)[i] = (uint16_t)(pack_bf16x2(v, 0.f) & 0xffffu);
	else p[i] = v;
}
typedef uint16_t bf16_t;  // raw bf16 bits; only sizeof(T) matters to the accessors above

// vector width of a [rows][C] tensor of the given element type, 0 = unsupported
static int vec_of(int C, int bf16) {
	if (bf16) return (C % 8 == 0) ? 8 : 0;
	return (C % 4 == 0) ? 4 : 1;
}
__device__ __forceinline__ float round_tf32(float x) {
	uint32_t r;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
	return __uint_as_float(r);
}

// Launch shape for a flat stream of nvec vectors with V vectors per row: total threads is a multiple of V when
// possible ("fixed column" mode), capped at max_blocks.
static int flat_grid(long long nvec, int V, int max_blocks, bool *fixed) {
	long long want = (nvec + (long long)kThreads * 4 - 1) / ((long long)kThreads * 4);
	int grid = (int)(want < 1 ? 1 : (want > max_blocks ? max_blocks : want));
	*fixed = false;
	if (kThreads % V == 0) *fixed = true;
	else if (V % kThreads == 0) {
		int m = V / kThreads;
		grid = (grid + m - 1) / m * m;
		if (grid > max_blocks) grid = max_blocks / m * m;
		*fixed = grid > 0;
		if (!*fixed) grid = 1;
	}
	return grid;
}
// every streaming kernel is compiled for a fixed number of resident blocks per SM (__launch_bounds__) and its grid is capped at
// whole waves of that: the bf16 BatchNorm-backward kernels at 80 registers once fitted 3 blocks per SM, so a 592-block grid ran as
// one full wave plus a 1-block-per-SM tail and reached 3 TB/s where the fp32 twin reached 6.4
constexpr int kMaxFlatBlocks = kNumSMs * 8;

// ------------------------------------------------------------------------------------------- BN statistics
// partials[blk][0][c] = sum x, partials[blk][1][c] = sum x^2 over the rows this block streamed.
template <typename T, int VEC, bool FIXED, bool BWD>
__global__ void __launch_bounds__(kThreads, BatchOf<VEC>::kBlocks) bn_reduce_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                            const T *__restrict__ mask, const float *__restrict__ means,
                                                            long long nvec, int V, float *__restrict__ partials, const float *__restrict__ mab,
                                                            const uint8_t *__restrict__ mask_bits) {
	extern __shared__ float sm[];  // [2][C]
	const int Cc = V * VEC;
	for (int i = threadIdx.x; i < 2 * Cc; i += kThreads) sm[i] = 0.f;
	__syncthreads();
	pdl_wait();
	const long long TS = (long long)gridDim.x * kThreads;
	const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
	float s[VEC], q[VEC], mu[VEC];
#pragma unroll
	for (int j = 0; j < VEC; j++) s[j] = q[j] = mu[j] = 0.f;
	const int col0 = (int)(g % V) * VEC;
	// ReLU mask of a plain BN+ReLU layer, recomputed from x with the forward's own folded scale/shift (bit-identical sign to the
	// stored activation, which then need not be read at all); layers with a residual join read the stored mask instead
	const bool remask = FIXED && BWD && mab != nullptr;
	float ma[VEC], mb[VEC];
	if (FIXED && BWD) {
#pragma unroll
		for (int j = 0; j < VEC; j++) {
			mu[j] = means[col0 + j];
			ma[j] = remask ? mab[col0 + j] : 0.f;
			mb[j] = remask ? mab[Cc + col0 + j] : 0.f;
		}
	}
	// batches of U vectors per stream: all loads of a batch are issued before the first is consumed (the compiler kept only one
	// or two in flight when load and use sat in the same unrolled body, and the bf16 kernels ran at 4.5 of 6.5 TB/s)
	constexpr int U = BatchOf<VEC>::U;
	using raw_t = typename RawOf<VEC>::type;
	// ReLU mask of a residual join: one bit per element written by the forward's bn_apply (one byte per 128-bit vector), or, when
	// that is not available (single-operator calls), the sign of the stored block output
	const bool rdbits = BWD && !remask && mask_bits != nullptr;
	const bool rdmask = BWD && !remask && !rdbits && mask != nullptr;
	for (long long i0 = g; i0 < nvec; i0 += TS * U) {
		raw_t rx[U], rd[U], rm[U];
		uint32_t rb[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const long long i = i0 + (long long)u * TS;
			if (i < nvec) {
				rx[u] = ldraw<T, VEC>(x, i);
				if constexpr (BWD) {
					rd[u] = ldraw<T, VEC>(dy, i);
					if (rdmask) rm[u] = ldraw<T, VEC>(mask, i);
					if (rdbits) rb[u] = mask_bits[i];
				}
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const long long i = i0 + (long long)u * TS;
			if (i >= nvec) break;
			float a[VEC];
			unpack<T, VEC>(rx[u], a);
			if constexpr (!BWD) {
				if constexpr (FIXED) {
#pragma unroll
					for (int j = 0; j < VEC; j++) { s[j] += a[j]; q[j] += a[j] * a[j]; }
				} else {
					const int c = (int)(i % V) * VEC;
#pragma unroll
					for (int j = 0; j < VEC; j++) { atomicAdd(&sm[c + j], a[j]); atomicAdd(&sm[Cc + c + j], a[j] * a[j]); }
				}
			} else {
				float d[VEC];
				unpack<T, VEC>(rd[u], d);
				if (remask) {
#pragma unroll
					for (int j = 0; j < VEC; j++) d[j] = fmaf(a[j], ma[j], mb[j]) > 0.f ? d[j] : 0.f;
				} else if (rdbits) {
#pragma unroll
					for (int j = 0; j < VEC; j++) d[j] = ((rb[u] >> j) & 1u) ? d[j] : 0.f;
				} else if (rdmask) {
					float mk[VEC];
					unpack<T, VEC>(rm[u], mk);
#pragma unroll
					for (int j = 0; j < VEC; j++) d[j] = mk[j] > 0.f ? d[j] : 0.f;
				}
				if constexpr (FIXED) {
#pragma unroll
					for (int j = 0; j < VEC; j++) { s[j] += d[j]; q[j] += d[j] * (a[j] - mu[j]); }
				} else {
					const int c = (int)(i % V) * VEC;
#pragma unroll
					for (int j = 0; j < VEC; j++) { atomicAdd(&sm[c + j], d[j]); atomicAdd(&sm[Cc + c + j], d[j] * (a[j] - means[c + j])); }
				}
			}
		}
	}
	if constexpr (FIXED) {
		// deterministic in-block combine: every thread parks its sums, then the first thread of each column adds the
		// kThreads / V threads that share its column in a fixed order (no floating-point atomics on the product path)
		__shared__ float red[kThreads][2 * VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][VEC + j] = q[j]; }
		__syncthreads();
		const int tcol = (int)(g % V);  // this thread's column; threads t, t + V, t + 2V ... of the block share it when V <= kThreads
		if (V <= kThreads) {
			if ((int)threadIdx.x < V) {
				float ss[VEC], qq[VEC];
#pragma unroll
				for (int j = 0; j < VEC; j++) ss[j] = qq[j] = 0.f;
				for (int k = threadIdx.x; k < kThreads; k += V)
#pragma unroll
					for (int j = 0; j < VEC; j++) { ss[j] += red[k][j]; qq[j] += red[k][VEC + j]; }
				// (blockIdx.x * kThreads) % V == 0 here, so thread t < V owns column t
#pragma unroll
				for (int j = 0; j < VEC; j++) { sm[tcol * VEC + j] = ss[j]; sm[Cc + tcol * VEC + j] = qq[j]; }
			}
		} else {
			// V is a multiple of kThreads: one thread per column, the block covers kThreads of the V columns (the rest stay 0)
#pragma unroll
			for (int j = 0; j < VEC; j++) { sm[tcol * VEC + j] = s[j]; sm[Cc + tcol * VEC + j] = q[j]; }
		}
	}
	__syncthreads();
	pdl_trigger();  // the fold kernel behind this one may be scheduled now (it waits for our completion before it reads)
	float *out = partials + (size_t)blockIdx.x * 2 * Cc;
	for (int i = threadIdx.x; i < 2 * Cc; i += kThreads) out[i] = sm[i];
}

// Fold of the per-block partials in fp64.  8 channels x 128 slices per block: slice sy sums partial blocks sy, sy+128, ... (32-byte
// sectors, ~5 independent loads per thread), then a fixed shuffle tree inside each warp and a fixed-order sum of the 32 warp
// results (deterministic).  History: one thread per channel over ~1000 partials was latency-bound at 170 us per launch; 32
// channels x 32 slices still spent 22 us in a 19-deep dependent L2 load chain (profiles/r01_ncu_all_kernels_one_step_summary.txt).
constexpr int kFinC = 8, kFinS = 128;
__device__ __forceinline__ bool fold_partials(const float *__restrict__ partials, int nblk, int Cc, double *s_out, double *q_out, float *zero = nullptr) {
	__shared__ double sm[kFinS / 4][2][kFinC];
	const int cx = threadIdx.x, sy = threadIdx.y, c = blockIdx.x * kFinC + cx;
	double s = 0, q = 0;
	if (c < Cc) {
#pragma unroll 4
		for (int b = sy; b < nblk; b += kFinS) {
			s += (double)partials[(size_t)b * 2 * Cc + c];
			q += (double)partials[(size_t)b * 2 * Cc + Cc + c];
		}
		if (zero) {  // every element of the region is read by exactly one thread: that thread clears it
			for (int b = sy; b < nblk; b += kFinS) { zero[(size_t)b * 2 * Cc + c] = 0.f; zero[(size_t)b * 2 * Cc + Cc + c] = 0.f; }
		}
	}
	// a warp holds 4 consecutive slices x 8 channels: lanes l, l^8, l^16, l^24 share a channel
	s += __shfl_xor_sync(0xffffffffu, s, 8);  q += __shfl_xor_sync(0xffffffffu, q, 8);
	s += __shfl_xor_sync(0xffffffffu, s, 16); q += __shfl_xor_sync(0xffffffffu, q, 16);
	if ((sy & 3) == 0) { sm[sy >> 2][0][cx] = s; sm[sy >> 2][1][cx] = q; }
	__syncthreads();
	if (sy != 0 || c >= Cc) return false;
	s = 0; q = 0;
#pragma unroll
	for (int j = 0; j < kFinS / 4; j++) { s += sm[j][0][cx]; q += sm[j][1][cx]; }
	*s_out = s;
	*q_out = q;
	return true;
}

// mean, biased variance, a = gamma*rstd, b = beta - mean*a
__global__ void bn_finalize_kernel(const float *partials, int nblk, double inv_n, int Cc, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, float *__restrict__ means, float *__restrict__ vars,
                                   float *__restrict__ ab, float *zero) {
	double s, q;
	pdl_wait();
	pdl_trigger();
	if (!fold_partials(partials, nblk, Cc, &s, &q, zero)) return;
	const int c = blockIdx.x * kFinC + threadIdx.x;
	const double mean = s * inv_n;
	double var = q * inv_n - mean * mean;
	if (var < 0) var = 0;
	const float meanf = (float)mean, varf = (float)var;
	means[c] = meanf;
	vars[c] = varf;
	const float a = gamma[c] / sqrtf(varf + eps);
	ab[c] = a;
	ab[Cc + c] = beta[c] - meanf * a;
}

static void launch_reduce(bool bwd, const void *x, const void *dy, const void *mask, const float *means, long long rows, int C,
                          float *partials, int max_blocks, int *grid_out, cudaStream_t st, const float *mab, int bf16,
                          const uint8_t *mask_bits = nullptr) {
	const int VEC = vec_of(C, bf16);
	*grid_out = 1;
	if (!VEC) { set_error("BatchNorm over bf16 tensors needs C %% 8 == 0 (C = %d)", C); return; }
	const int V = C / VEC;
	const long long nvec = rows * V;
	bool fixed;
	// one whole wave: the bf16 variants keep 8 channels of coefficients per thread and fit 3 blocks per SM, the fp32 ones 4
	// (and the fold cost grows with the number of partial blocks)
	const int wave = kNumSMs * bn_bwd_blocks_per_sm(bf16);
	int cap = max_blocks < wave ? max_blocks : wave;
	int grid = flat_grid(nvec, V, cap, &fixed);
	if (VEC == 1) fixed = false;
	const size_t smem = 2 * (size_t)C * sizeof(float);
#define RB_RED(T_, VEC_, FIX_, BWD_) \
	launch_k(0, bn_reduce_kernel<T_, VEC_, FIX_, BWD_>, grid, kThreads, smem, st, (const T_ *)x, (const T_ *)dy, (const T_ *)mask, means, nvec, V, partials, mab, mask_bits)
#define RB_RED2(T_, VEC_) \
	do { \
		if (fixed) { if (bwd) RB_RED(T_, VEC_, true, true); else RB_RED(T_, VEC_, true, false); } \
		else { if (bwd) RB_RED(T_, VEC_, false, true); else RB_RED(T_, VEC_, false, false); } \
	} while (0)
	if (bf16) RB_RED2(bf16_t, 8);
	else if (VEC == 4) RB_RED2(float, 4);
	else { if (bwd) RB_RED(float, 1, false, true); else RB_RED(float, 1, false, false); }
#undef RB_RED2
#undef RB_RED
	RB_LAUNCH_CHECK();
	RB_TRACE("bn_reduce_kernel", "%s rows=%lld C=%d grid=%d", bwd ? "bwd" : "fwd", rows, C, grid);
	*grid_out = grid;
}

void bn_finalize(float *partials, int nblk, long long rows, int C, const float *gamma, const float *beta, float eps,
                 float *means, float *vars, float *ab, cudaStream_t st, int zero_after) {
	launch_k(2, bn_finalize_kernel, ceil_div(C, kFinC), dim3(kFinC, kFinS), 0, st, (const float *)partials, nblk, 1.0 / (double)rows, C, gamma, beta, eps, means, vars, ab, zero_after ? partials : (float *)nullptr);
	RB_LAUNCH_CHECK();
}

void bn_stats(const void *x, long long rows, int C, const float *gamma, const float *beta, float eps, float *means, float *vars,
              float *ab, float *partials, int max_blocks, cudaStream_t st, int bf16) {
	int grid;
	launch_reduce(false, x, nullptr, nullptr, nullptr, rows, C, partials, max_blocks, &grid, st, nullptr, bf16);
	bn_finalize(partials, grid, rows, C, gamma, beta, eps, means, vars, ab, st);
}

// ------------------------------------------------------------------------------------------- BN apply (+ residual + ReLU)
// BITS: also store the sign bits of y (one byte per 128-bit vector, bit j = element j > 0): the 1-bit ReLU mask BatchNorm backward
// needs of a residual join's output.  Four neighbouring lanes own four consecutive vectors, so the group leader stores one 32-bit
// word (per-thread byte stores made the join 40 % slower); needs nvec % 32 == 0, which keeps `i < nvec` warp-uniform.
template <typename T, int VEC, bool FIXED, bool BITS>
__global__ void __launch_bounds__(kThreads, 4) bn_apply_kernel(const T *__restrict__ x, const float *__restrict__ ab, long long nvec, int V,
                                                           int relu, const T *__restrict__ res, const float *__restrict__ ab2,
                                                           T *__restrict__ y, int rnd, uint32_t *__restrict__ bits_out) {
	const int Cc = V * VEC;
	const long long TS = (long long)gridDim.x * kThreads;
	const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
	float a[VEC], b[VEC], a2[VEC], b2[VEC];
	pdl_wait();
	pdl_trigger();
	if constexpr (FIXED) {
		const int c0 = (int)(g % V) * VEC;
#pragma unroll
		for (int j = 0; j < VEC; j++) {
			a[j] = ab[c0 + j]; b[j] = ab[Cc + c0 + j];
			a2[j] = ab2 ? ab2[c0 + j] : 1.f; b2[j] = ab2 ? ab2[Cc + c0 + j] : 0.f;
		}
	}
	// batches of U vectors per stream, all loads first (the warp shuffles of the BITS variant otherwise fence every iteration's
	// loads behind the previous iteration's stores: one load pair in flight instead of U)
	constexpr int U = (VEC == 8) ? 2 : 4;
	using raw_t = typename RawOf<VEC>::type;
	for (long long i0 = g; i0 < nvec; i0 += TS * U) {
		raw_t rx[U], rr[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const long long i = i0 + (long long)u * TS;
			if (i < nvec) {
				rx[u] = ldraw<T, VEC>(x, i);
				if (res) rr[u] = ldraw<T, VEC>(res, i);
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const long long i = i0 + (long long)u * TS;
			if (i >= nvec) break;  // warp-uniform in the BITS variant (nvec % 32 == 0)
			if constexpr (!FIXED) {
				const int c0 = (int)(i % V) * VEC;
#pragma unroll
				for (int j = 0; j < VEC; j++) {
					a[j] = ab[c0 + j]; b[j] = ab[Cc + c0 + j];
					a2[j] = ab2 ? ab2[c0 + j] : 1.f; b2[j] = ab2 ? ab2[Cc + c0 + j] : 0.f;
				}
			}
			float v[VEC];
			unpack<T, VEC>(rx[u], v);
#pragma unroll
			for (int j = 0; j < VEC; j++) v[j] = fmaf(v[j], a[j], b[j]);
			if (res) {
				float r[VEC];
				unpack<T, VEC>(rr[u], r);
#pragma unroll
				for (int j = 0; j < VEC; j++) v[j] += fmaf(r[j], a2[j], b2[j]);
			}
#pragma unroll
			for (int j = 0; j < VEC; j++) {
				if (relu) v[j] = fmaxf(v[j], 0.f);
				if (rnd) v[j] = round_tf32(v[j]);
			}
			stv<T, VEC>(y, i, v);
			if constexpr (BITS) {
				uint32_t m = 0;
#pragma unroll
				for (int j = 0; j < VEC; j++) m |= (v[j] > 0.f ? 1u : 0u) << j;
				const uint32_t m1 = __shfl_down_sync(0xffffffffu, m, 1), m2 = __shfl_down_sync(0xffffffffu, m, 2), m3 = __shfl_down_sync(0xffffffffu, m, 3);
				if ((threadIdx.x & 3) == 0) bits_out[i >> 2] = m | (m1 << 8) | (m2 << 16) | (m3 << 24);
			}
		}
	}
}

void bn_apply(const void *x, const float *ab, long long rows, int C, int relu, const void *res, const float *ab2, void *y,
              int rnd, cudaStream_t st, int bf16, uint8_t *bits_out) {
	const int VEC = vec_of(C, bf16);
	if (!VEC) { set_error("BatchNorm over bf16 tensors needs C %% 8 == 0 (C = %d)", C); return; }
	const int V = C / VEC;
	const long long nvec = rows * V;
	bool fixed;
	int grid = flat_grid(nvec, V, kMaxFlatBlocks, &fixed);
	if (bits_out && !(fixed && VEC >= 4 && nvec % 32 == 0)) { set_error("bn_apply: mask bits need the fixed-column vector path and nvec %% 32 == 0 (rows %lld, C %d)", rows, C); return; }
#define RB_APPLY(T_, VEC_, FIX_, BITS_) \
	launch_k(0, bn_apply_kernel<T_, VEC_, FIX_, BITS_>, grid, kThreads, 0, st, (const T_ *)x, ab, nvec, V, relu, (const T_ *)res, ab2, (T_ *)y, bf16 ? 0 : rnd, (uint32_t *)bits_out)
	if (bf16) { if (bits_out) RB_APPLY(bf16_t, 8, true, true); else if (fixed) RB_APPLY(bf16_t, 8, true, false); else RB_APPLY(bf16_t, 8, false, false); }
	else if (VEC == 4 && bits_out) RB_APPLY(float, 4, true, true);
	else if (VEC == 4 && fixed) RB_APPLY(float, 4, true, false);
	else if (VEC == 4) RB_APPLY(float, 4, false, false);
	else RB_APPLY(float, 1, false, false);
#undef RB_APPLY
	RB_LAUNCH_CHECK();
	RB_TRACE("bn_apply_kernel", "rows=%lld C=%d res=%d%s grid=%d", rows, C, res ? (ab2 ? 2 : 1) : 0, bits_out ? " +bits" : "", grid);
}

// ------------------------------------------------------------------------------------------- BN backward
// s1 = sum dy', s2 = sum dy' (x - mean);  dbeta = s1, dgamma = s2 * rstd,
// dx = c1*dy' + c2 + c3*(x - mean) with c1 = gamma*rstd, c2 = -c1*s1/n, c3 = -c1*rstd^2*s2/n
// (algebraically the reference's three-term form, resnet.cu:394-422).
__global__ void bn_bwd_finalize_kernel(const float *__restrict__ partials, int nblk, double inv_n, int Cc, const float *__restrict__ gamma,
                                       const float *__restrict__ means, const float *__restrict__ vars, float eps,
                                       float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ coef) {
	double s1, s2;
	pdl_wait();
	pdl_trigger();
	if (!fold_partials(partials, nblk, Cc, &s1, &s2)) return;
	const int c = blockIdx.x * kFinC + threadIdx.x;
	const float rstd = 1.0f / sqrtf(vars[c] + eps);
	dbeta[c] = (float)s1;
	dgamma[c] = (float)(s2 * (double)rstd);
	const double c1 = (double)gamma[c] * (double)rstd;
	const double c2 = -c1 * s1 * inv_n, c3 = -c1 * (double)rstd * (double)rstd * s2 * inv_n;
	coef[c] = (float)c1;
	coef[Cc + c] = (float)(c2 - c3 * (double)means[c]);  // dx = c1 dy' + c3 x + (c2 - c3 mean)
	coef[2 * Cc + c] = (float)c3;
}

// coef [3][C]: dx = c1 * dy' + c3 * x + k with k = c2 - c3 * mean folded by the finalize kernel
template <typename T, int VEC, bool FIXED>
__global__ void __launch_bounds__(kThreads, BatchOf<VEC>::kBlocks) bn_bwd_dx_kernel(const T *__restrict__ x, const T *dy, const T *__restrict__ mask,
                                                            const float *__restrict__ coef, long long nvec, int V, T *dx, int rnd,
                                                            const float *__restrict__ mab, T *__restrict__ masked_out,
                                                            const uint8_t *__restrict__ mask_bits) {
	const int Cc = V * VEC;
	const long long TS = (long long)gridDim.x * kThreads;
	const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
	float c1[VEC], ck[VEC], c3[VEC], ma[VEC], mb[VEC];
	const bool remask = FIXED && mab != nullptr;
	const bool rdbits = !remask && mask_bits != nullptr;
	const bool rdmask = !remask && !rdbits && mask != nullptr;
	pdl_wait();
	pdl_trigger();
	if constexpr (FIXED) {
		const int c0 = (int)(g % V) * VEC;
#pragma unroll
		for (int j = 0; j < VEC; j++) {
			c1[j] = coef[c0 + j]; ck[j] = coef[Cc + c0 + j]; c3[j] = coef[2 * Cc + c0 + j];
			ma[j] = remask ? mab[c0 + j] : 0.f; mb[j] = remask ? mab[Cc + c0 + j] : 0.f;
		}
	}
	constexpr int U = BatchOf<VEC>::U;
	using raw_t = typename RawOf<VEC>::type;
	for (long long i0 = g; i0 < nvec; i0 += TS * U) {
		raw_t rx[U], rd[U], rm[U];
		uint32_t rb[U];
#pragma unroll
		for (int u = 0; u < U; u++) {  // all loads of the batch first (dx may alias dy: every element is read before its own store)
			const long long i = i0 + (long long)u * TS;
			if (i < nvec) {
				rx[u] = ldraw<T, VEC>(x, i);
				rd[u] = ldraw<T, VEC>(dy, i);
				if (rdmask) rm[u] = ldraw<T, VEC>(mask, i);
				if (rdbits) rb[u] = mask_bits[i];
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const long long i = i0 + (long long)u * TS;
			if (i >= nvec) break;
			if constexpr (!FIXED) {
				const int c0 = (int)(i % V) * VEC;
#pragma unroll
				for (int j = 0; j < VEC; j++) { c1[j] = coef[c0 + j]; ck[j] = coef[Cc + c0 + j]; c3[j] = coef[2 * Cc + c0 + j]; }
			}
			float a[VEC], d[VEC];
			unpack<T, VEC>(rx[u], a);
			unpack<T, VEC>(rd[u], d);
			if (remask) {
#pragma unroll
				for (int j = 0; j < VEC; j++) d[j] = fmaf(a[j], ma[j], mb[j]) > 0.f ? d[j] : 0.f;
			} else if (rdbits) {
#pragma unroll
				for (int j = 0; j < VEC; j++) d[j] = ((rb[u] >> j) & 1u) ? d[j] : 0.f;
			} else if (rdmask) {
				float mk[VEC];
				unpack<T, VEC>(rm[u], mk);
#pragma unroll
				for (int j = 0; j < VEC; j++) d[j] = mk[j] > 0.f ? d[j] : 0.f;
			}
			// the ReLU-masked upstream gradient is also the identity shortcut's gradient (reference: resnet.cu:2003-2004 setVal + addVec):
			// written here, while it is in registers, instead of by a separate 3-tensor relu_bwd pass
			if (masked_out) stv<T, VEC>(masked_out, i, d);
#pragma unroll
			for (int j = 0; j < VEC; j++) {
				float r = fmaf(c1[j], d[j], fmaf(c3[j], a[j], ck[j]));
				d[j] = rnd ? round_tf32(r) : r;
			}
			stv<T, VEC>(dx, i, d);
		}
	}
}

void bn_bwd(const void *x, const void *dy, const void *mask_src, const float *gamma, const float *means, const float *vars, float eps,
            long long rows, int C, float *dgamma, float *dbeta, void *dx, float *partials, int max_blocks, float *coef, int rnd,
            cudaStream_t st, const float *mab, int bf16, void *masked_out, const uint8_t *mask_bits) {
	int grid;
	launch_reduce(true, x, dy, mask_src, means, rows, C, partials, max_blocks, &grid, st, mab, bf16, mask_bits);
	launch_k(2, bn_bwd_finalize_kernel, ceil_div(C, kFinC), dim3(kFinC, kFinS), 0, st, partials, grid, 1.0 / (double)rows, C, gamma, means, vars, eps, dgamma, dbeta, coef);
	RB_LAUNCH_CHECK();
	const int VEC = vec_of(C, bf16);
	if (!VEC) return;
	const int V = C / VEC;
	const long long nvec = rows * V;
	bool fixed;
	int g2 = flat_grid(nvec, V, kNumSMs * bn_bwd_blocks_per_sm(bf16) * 2, &fixed);  // two whole waves
#define RB_DX(T_, VEC_, FIX_) \
	launch_k(0, bn_bwd_dx_kernel<T_, VEC_, FIX_>, g2, kThreads, 0, st, (const T_ *)x, (const T_ *)dy, (const T_ *)mask_src, coef, nvec, V, (T_ *)dx, bf16 ? 0 : rnd, mab, (T_ *)masked_out, mask_bits)
	if (bf16) { if (fixed) RB_DX(bf16_t, 8, true); else RB_DX(bf16_t, 8, false); }
	else if (VEC == 4 && fixed) RB_DX(float, 4, true);
	else if (VEC == 4) RB_DX(float, 4, false);
	else RB_DX(float, 1, false);
#undef RB_DX
	RB_LAUNCH_CHECK();
	RB_TRACE("bn_bwd_dx_kernel", "rows=%lld C=%d mask=%s%s grid=%d", rows, C, mab ? "recomputed" : (mask_bits ? "bits" : (mask_src ? "read" : "none")), masked_out ? "+shortcut" : "", g2);
}

// ------------------------------------------------------------------------------------------- ReLU backward (identity shortcut)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 4) relu_bwd_kernel(const T *__restrict__ y, const T *__restrict__ dy, long long n, T *__restrict__ dx) {
	const long long T_ = (long long)gridDim.x * kThreads;
	const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
	const long long nv = n / VEC;
#pragma unroll 4
	for (long long i = g; i < nv; i += T_) {
		float a[VEC], d[VEC];
		ldv<T, VEC>(y, i, a);
		ldv<T, VEC>(dy, i, d);
#pragma unroll
		for (int j = 0; j < VEC; j++) d[j] = a[j] > 0.f ? d[j] : 0.f;
		stv<T, VEC>(dx, i, d);
	}
	for (long long i = nv * VEC + g; i < n; i += T_) st1<T>(dx, i, ld1<T>(y, i) > 0.f ? ld1<T>(dy, i) : 0.f);
}
void relu_bwd(const void *y, const void *dy, long long n, void *dx, cudaStream_t st, int bf16) {
	const int VEC = bf16 ? 8 : 4;
	int grid = (int)((n / VEC + kThreads * 4 - 1) / (kThreads * 4));
	grid = grid < 1 ? 1 : (grid > kMaxFlatBlocks ? kMaxFlatBlocks : grid);
	if (bf16) relu_bwd_kernel<bf16_t, 8><<<grid, kThreads, 0, st>>>((const bf16_t *)y, (const bf16_t *)dy, n, (bf16_t *)dx);
	else relu_bwd_kernel<float, 4><<<grid, kThreads, 0, st>>>((const float *)y, (const float *)dy, n, (float *)dx);
	RB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------- max pool
// reference resnet.cu:433-471: init -1024, strict '>', row-major window scan, flat input index.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(const T *__restrict__ x, int N, int S, int C, int k, int stride,
                                                              int *__restrict__ inds, T *__restrict__ out) {
	const int So = S / stride, half = k / 2, V = C / VEC;
	const long long total = (long long)N * So * So * V;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
		const int cv = (int)(i % V);
		long long p = i / V;
		const int ow = (int)(p % So); p /= So;
		const int oh = (int)(p % So);
		const int n = (int)(p / So);
		float mv[VEC]; int mi[VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) { mv[j] = -1024.f; mi[j] = -1024; }
		for (int r = -half; r <= half; r++) {
			const int h = stride * oh + r;
			if (h < 0 || h >= S) continue;
			for (int c = -half; c <= half; c++) {
				const int w = stride * ow + c;
				if (w < 0 || w >= S) continue;
				const long long base = (((long long)n * S + h) * S + w) * C + (long long)cv * VEC;
				float v[VEC];
				ldv<T, VEC>(x, base / VEC, v);
#pragma unroll
				for (int j = 0; j < VEC; j++) if (v[j] > mv[j]) { mv[j] = v[j]; mi[j] = (int)(base + j); }
			}
		}
		stv<T, VEC>(out, i, mv);
		if constexpr (VEC >= 4) {
#pragma unroll
			for (int h = 0; h < VEC / 4; h++) reinterpret_cast<int4 *>(inds)[i * (VEC / 4) + h] = make_int4(mi[4 * h], mi[4 * h + 1], mi[4 * h + 2], mi[4 * h + 3]);
		} else inds[i] = mi[0];
	}
}
// k = 3, stride = 2 (the only pooling the network uses, reference: resnet.cu:3248): the nine window loads are issued before the
// first comparison (the generic loop above serialises load -> compare -> branch nine times: 3.1 of 6.5 TB/s), comparisons keep the
// reference's row-major order so ties resolve identically.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 3) maxpool3s2_fwd_kernel(const T *__restrict__ x, int N, int S, int C, int *__restrict__ inds, T *__restrict__ out) {
	using raw_t = typename RawOf<VEC>::type;
	const int So = S / 2, V = C / VEC;
	const long long total = (long long)N * So * So * V;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
		const int cv = (int)(i % V);
		long long p = i / V;
		const int ow = (int)(p % So); p /= So;
		const int oh = (int)(p % So);
		const int n = (int)(p / So);
		raw_t rv[9];
		bool ok[9];
		int base[9];
#pragma unroll
		for (int t = 0; t < 9; t++) {
			const int h = 2 * oh + t / 3 - 1, w = 2 * ow + t % 3 - 1;
			ok[t] = h >= 0 && h < S && w >= 0 && w < S;
			base[t] = (int)((((long long)n * S + h) * S + w) * C + (long long)cv * VEC);  // flat input index (int, as the reference's max_inds)
			if (ok[t]) rv[t] = ldraw<T, VEC>(x, base[t] / VEC);
		}
		float mv[VEC]; int mi[VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) { mv[j] = -1024.f; mi[j] = -1024; }
#pragma unroll
		for (int t = 0; t < 9; t++) {
			if (!ok[t]) continue;
			float v[VEC];
			unpack<T, VEC>(rv[t], v);
#pragma unroll
			for (int j = 0; j < VEC; j++) if (v[j] > mv[j]) { mv[j] = v[j]; mi[j] = base[t] + j; }
		}
		stv<T, VEC>(out, i, mv);
#pragma unroll
		for (int h = 0; h < VEC / 4; h++) reinterpret_cast<int4 *>(inds)[i * (VEC / 4) + h] = make_int4(mi[4 * h], mi[4 * h + 1], mi[4 * h + 2], mi[4 * h + 3]);
	}
}
// backward twin: an input pixel belongs to at most 2 x 2 windows; their argmax vectors and gradients are loaded together, then summed
// in ascending (oh, ow) order like the generic kernel
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 3) maxpool3s2_bwd_kernel(const int *__restrict__ inds, const T *__restrict__ dout, int N, int S, int C, T *__restrict__ din) {
	using raw_t = typename RawOf<VEC>::type;
	const int So = S / 2, V = C / VEC;
	const long long total = (long long)N * S * S * V;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
		const int cv = (int)(i % V);
		long long p = i / V;
		const int w = (int)(p % S); p /= S;
		const int h = (int)(p % S);
		const int n = (int)(p / S);
		// windows oh with 2*oh - 1 <= h <= 2*oh + 1: oh = h/2 for even h; (h-1)/2 and (h+1)/2 for odd h
		const int oh0 = h >> 1, ow0 = w >> 1;
		const bool two_h = (h & 1) && oh0 + 1 < So, two_w = (w & 1) && ow0 + 1 < So;
		raw_t rd[4];
		int4 ri[4][VEC / 4];
		bool ok[4];
#pragma unroll
		for (int t = 0; t < 4; t++) {
			const int oh = oh0 + (t >> 1), ow = ow0 + (t & 1);
			ok[t] = ((t >> 1) == 0 || two_h) && ((t & 1) == 0 || two_w);
			if (ok[t]) {
				const long long o = ((((long long)n * So + oh) * So + ow) * C) / VEC + cv;
				rd[t] = ldraw<T, VEC>(dout, o);
#pragma unroll
				for (int q = 0; q < VEC / 4; q++) ri[t][q] = reinterpret_cast<const int4 *>(inds)[o * (VEC / 4) + q];
			}
		}
		const int me = (int)(i * VEC);
		float acc[VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) acc[j] = 0.f;
#pragma unroll
		for (int t = 0; t < 4; t++) {
			if (!ok[t]) continue;
			float d[VEC];
			unpack<T, VEC>(rd[t], d);
#pragma unroll
			for (int q = 0; q < VEC / 4; q++) {
				if (ri[t][q].x == me + 4 * q) acc[4 * q] += d[4 * q];
				if (ri[t][q].y == me + 4 * q + 1) acc[4 * q + 1] += d[4 * q + 1];
				if (ri[t][q].z == me + 4 * q + 2) acc[4 * q + 2] += d[4 * q + 2];
				if (ri[t][q].w == me + 4 * q + 3) acc[4 * q + 3] += d[4 * q + 3];
			}
		}
		stv<T, VEC>(din, i, acc);
	}
}

// The same sums with each pooled vector read four times instead of nine: a thread owns the 2 x 2 input block (2a..2a+1, 2b..2b+1),
// whose pixels belong to windows (a, b), (a, b+1), (a+1, b), (a+1, b+1) only -- four (gradient, argmax) vector pairs in, four input
// gradient vectors out, summed in the same ascending (oh, ow) order (bit-identical to the per-pixel kernel above, which ran at
// 2.1-2.5 TB/s of DRAM traffic on its L1 / L2 request rate: 2.25 pooled vector pairs per input vector).
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 3) maxpool3s2_bwd_block_kernel(const int *__restrict__ inds, const T *__restrict__ dout, int N, int S, int C, T *__restrict__ din) {
	using raw_t = typename RawOf<VEC>::type;
	const int So = S / 2, V = C / VEC;
	const long long total = (long long)N * So * So * V;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
		const int cv = (int)(i % V);
		long long p = i / V;
		const int b = (int)(p % So); p /= So;
		const int a = (int)(p % So);
		const int n = (int)(p / So);
		raw_t rd[4];
		int4 ri[4][VEC / 4];
		bool ok[4];
#pragma unroll
		for (int t = 0; t < 4; t++) {
			const int oh = a + (t >> 1), ow = b + (t & 1);
			ok[t] = oh < So && ow < So;
			if (ok[t]) {
				const long long o = ((((long long)n * So + oh) * So + ow) * C) / VEC + cv;
				rd[t] = ldraw<T, VEC>(dout, o);
#pragma unroll
				for (int q = 0; q < VEC / 4; q++) ri[t][q] = reinterpret_cast<const int4 *>(inds)[o * (VEC / 4) + q];
			}
		}
		float d[4][VEC];
#pragma unroll
		for (int t = 0; t < 4; t++) {
			if (ok[t]) unpack<T, VEC>(rd[t], d[t]);
		}
#pragma unroll
		for (int px = 0; px < 4; px++) {  // input pixel (2a + px / 2, 2b + px % 2) collects from windows t with (t / 2 <= px / 2) and (t % 2 <= px % 2)
			const int h = 2 * a + (px >> 1), w = 2 * b + (px & 1);
			const long long vi = (((long long)n * S + h) * S + w) * V + cv;
			const int me = (int)(vi * VEC);
			float acc[VEC];
#pragma unroll
			for (int j = 0; j < VEC; j++) acc[j] = 0.f;
#pragma unroll
			for (int t = 0; t < 4; t++) {
				if ((t >> 1) > (px >> 1) || (t & 1) > (px & 1) || !ok[t]) continue;
#pragma unroll
				for (int q = 0; q < VEC / 4; q++) {
					if (ri[t][q].x == me + 4 * q) acc[4 * q] += d[t][4 * q];
					if (ri[t][q].y == me + 4 * q + 1) acc[4 * q + 1] += d[t][4 * q + 1];
					if (ri[t][q].z == me + 4 * q + 2) acc[4 * q + 2] += d[t][4 * q + 2];
					if (ri[t][q].w == me + 4 * q + 3) acc[4 * q + 3] += d[t][4 * q + 3];
				}
			}
			stv<T, VEC>(din, vi, acc);
		}
	}
}

void maxpool_fwd(const void *x, int N, int S, int C, int k, int stride, int *max_inds, void *out, cudaStream_t st, int bf16) {
	const int So = S / stride;
	const int VEC = vec_of(C, bf16);
	if (!VEC) { set_error("max pool over bf16 tensors needs C %% 8 == 0 (C = %d)", C); return; }
	long long total = (long long)N * So * So * (C / VEC);
	int grid = (int)((total + kThreads - 1) / kThreads); grid = grid > kMaxFlatBlocks * 4 ? kMaxFlatBlocks * 4 : grid;
	if (k == 3 && stride == 2 && S % 2 == 0 && VEC >= 4 && (long long)N * S * S * C < (1LL << 31)) {
		grid = grid > kNumSMs * 3 * 4 ? kNumSMs * 3 * 4 : grid;
		if (bf16) maxpool3s2_fwd_kernel<bf16_t, 8><<<grid, kThreads, 0, st>>>((const bf16_t *)x, N, S, C, max_inds, (bf16_t *)out);
		else maxpool3s2_fwd_kernel<float, 4><<<grid, kThreads, 0, st>>>((const float *)x, N, S, C, max_inds, (float *)out);
		RB_LAUNCH_CHECK();
		return;
	}
	if (bf16) maxpool_fwd_kernel<bf16_t, 8><<<grid, kThreads, 0, st>>>((const bf16_t *)x, N, S, C, k, stride, max_inds, (bf16_t *)out);
	else if (VEC == 4) maxpool_fwd_kernel<float, 4><<<grid, kThreads, 0, st>>>((const float *)x, N, S, C, k, stride, max_inds, (float *)out);
	else maxpool_fwd_kernel<float, 1><<<grid, kThreads, 0, st>>>((const float *)x, N, S, C, k, stride, max_inds, (float *)out);
	RB_LAUNCH_CHECK();
}

// Gather form of the reference's scatter (resnet.cu:476-494): every input element sums the gradients of the
// windows whose recorded argmax is that element.  Deterministic, and accumulates where the reference's
// overlapping-window scatter races (its cuDNN variants accumulate too).
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const int *__restrict__ inds, const T *__restrict__ dout, int N, int S, int C,
                                                              int k, int stride, T *__restrict__ din) {
	const int So = S / stride, half = k / 2, V = C / VEC;
	const long long total = (long long)N * S * S * V;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
		const int cv = (int)(i % V);
		long long p = i / V;
		const int w = (int)(p % S); p /= S;
		const int h = (int)(p % S);
		const int n = (int)(p / S);
		// windows oh with stride*oh - half <= h <= stride*oh + half
		int oh_lo = (h - half + stride - 1); oh_lo = oh_lo < 0 ? 0 : oh_lo / stride;
		int oh_hi = (h + half) / stride; oh_hi = oh_hi >= So ? So - 1 : oh_hi;
		int ow_lo = (w - half + stride - 1); ow_lo = ow_lo < 0 ? 0 : ow_lo / stride;
		int ow_hi = (w + half) / stride; ow_hi = ow_hi >= So ? So - 1 : ow_hi;
		const int me = (int)(i * VEC);  // flat index of this thread's first channel in the input tensor
		float acc[VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) acc[j] = 0.f;
		for (int oh = oh_lo; oh <= oh_hi; oh++)
			for (int ow = ow_lo; ow <= ow_hi; ow++) {
				const long long o = ((((long long)n * So + oh) * So + ow) * C) / VEC + cv;
				float d[VEC];
				ldv<T, VEC>(dout, o, d);
				int id[VEC];
				if constexpr (VEC >= 4) {
#pragma unroll
					for (int h = 0; h < VEC / 4; h++) {
						int4 t = reinterpret_cast<const int4 *>(inds)[o * (VEC / 4) + h];
						id[4 * h] = t.x; id[4 * h + 1] = t.y; id[4 * h + 2] = t.z; id[4 * h + 3] = t.w;
					}
				} else id[0] = inds[o];
#pragma unroll
				for (int j = 0; j < VEC; j++) if (id[j] == me + j) acc[j] += d[j];
			}
		stv<T, VEC>(din, i, acc);
	}
}
void maxpool_bwd(const int *max_inds, const void *dout, int N, int S, int C, int k, int stride, void *din, cudaStream_t st, int bf16) {
	const int VEC = vec_of(C, bf16);
	if (!VEC) { set_error("max pool over bf16 tensors needs C %% 8 == 0 (C = %d)", C); return; }
	long long total = (long long)N * S * S * (C / VEC);
	int grid = (int)((total + kThreads - 1) / kThreads); grid = grid > kMaxFlatBlocks * 8 ? kMaxFlatBlocks * 8 : grid;
	if (k == 3 && stride == 2 && S % 2 == 0 && VEC >= 4 && (long long)N * S * S * C < (1LL << 31)) {
		if (!getenv("RESNET_B200_POOL_BWD_PIXEL")) {  // one thread per 2 x 2 input block
			const long long blocks = ((long long)N * (S / 2) * (S / 2) * (C / VEC) + kThreads - 1) / kThreads;
			const int g2 = (int)(blocks > kNumSMs * 3 * 8 ? kNumSMs * 3 * 8 : blocks);
			if (bf16) maxpool3s2_bwd_block_kernel<bf16_t, 8><<<g2, kThreads, 0, st>>>(max_inds, (const bf16_t *)dout, N, S, C, (bf16_t *)din);
			else maxpool3s2_bwd_block_kernel<float, 4><<<g2, kThreads, 0, st>>>(max_inds, (const float *)dout, N, S, C, (float *)din);
			RB_LAUNCH_CHECK();
			return;
		}
		grid = grid > kNumSMs * 3 * 8 ? kNumSMs * 3 * 8 : grid;
		if (bf16) maxpool3s2_bwd_kernel<bf16_t, 8><<<grid, kThreads, 0, st>>>(max_inds, (const bf16_t *)dout, N, S, C, (bf16_t *)din);
		else maxpool3s2_bwd_kernel<float, 4><<<grid, kThreads, 0, st>>>(max_inds, (const float *)dout, N, S, C, (float *)din);
		RB_LAUNCH_CHECK();
		return;
	}
	if (bf16) maxpool_bwd_kernel<bf16_t, 8><<<grid, kThreads, 0, st>>>(max_inds, (const bf16_t *)dout, N, S, C, k, stride, (bf16_t *)din);
	else if (VEC == 4) maxpool_bwd_kernel<float, 4><<<grid, kThreads, 0, st>>>(max_inds, (const float *)dout, N, S, C, k, stride, (float *)din);
	else maxpool_bwd_kernel<float, 1><<<grid, kThreads, 0, st>>>(max_inds, (const float *)dout, N, S, C, k, stride, (float *)din);
	RB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------- fused stem tail
// The stem's BatchNorm + ReLU output (reference: init_conv_activated, N x 112 x 112 x 64 -- the largest tensor of the network, 822 MB
// in fp32 at batch 256) is read by exactly one consumer, the 3x3/2 max pool, and its gradient by exactly one, the stem BatchNorm's
// backward.  Unless the trainer keeps every tensor, neither is materialised:
//   forward   bn_pool_fwd: pooled = maxpool(relu(x * a + b)), argmax indices as before            (X0 in, P0 + max_inds out;
//             the separate kernels moved 2 E + 1.5 E, this one 1 E + 0.5 E, E = bytes of X0 with 4-byte indices counted at fp32 size)
//   backward  pool_bn_bwd: the pool's gradient gather (the 2 x 2-block form of maxpool3s2_bwd_block_kernel) feeds the BatchNorm
//             backward's two passes directly from (dP0, max_inds): reduce reads X0 + 0.5 E, dx reads the same and writes dX0
//             (separate kernels: 1.5 E + 2 E + 3 E; fused: 1.5 E + 2.5 E).
// Per element the arithmetic is the one of bn_apply_kernel / maxpool3s2_fwd_kernel / maxpool3s2_bwd_block_kernel / bn_bwd_dx_kernel
// in the same order (values and indices are bit-identical to the unfused path; only the order of the dgamma / dbeta partial sums differs).
template <typename T> __device__ __forceinline__ float round_as_stored(float v, int rnd) {
	if constexpr (sizeof(T) == 2) return bf16_lo(pack_bf16x2(v, 0.f));
	else return rnd ? round_tf32(v) : v;
}
// (32-bit index arithmetic throughout: the host guarantees N * S * S * C < 2^31, as the int32 max_inds already require; the 64-bit
// divisions of the flat index cost more instructions than the nine compares)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, VEC == 8 ? 2 : 3) bn_pool_fwd_kernel(const T *__restrict__ x, const float *__restrict__ ab, int N, int S, int C, int rnd,
                                                                 int *__restrict__ inds, T *__restrict__ out) {
	using raw_t = typename RawOf<VEC>::type;
	const int So = S / 2, V = C / VEC;
	const int total = N * So * So * V;
	const int cv = (int)(threadIdx.x % V);  // kThreads % V == 0: a thread keeps its channels
	float a[VEC], b[VEC];
#pragma unroll
	for (int j = 0; j < VEC; j++) { a[j] = ab[cv * VEC + j]; b[j] = ab[C + cv * VEC + j]; }
	for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
		int p = i / V;
		const int ow = p % So; p /= So;
		const int oh = p % So;
		const int n = p / So;
		raw_t rv[9];
		bool ok[9];
		const int base0 = ((n * S + 2 * oh - 1) * S + 2 * ow - 1) * C + cv * VEC;  // window corner (may lie outside: only used with ok[t])
#pragma unroll
		for (int t = 0; t < 9; t++) {
			const int h = 2 * oh + t / 3 - 1, w = 2 * ow + t % 3 - 1;
			ok[t] = h >= 0 && h < S && w >= 0 && w < S;
			if (ok[t]) rv[t] = ldraw<T, VEC>(x, (base0 + ((t / 3) * S + t % 3) * C) / VEC);
		}
		float mv[VEC]; int mi[VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) { mv[j] = -1024.f; mi[j] = -1024; }
#pragma unroll
		for (int t = 0; t < 9; t++) {
			if (!ok[t]) continue;
			float v[VEC];
			unpack<T, VEC>(rv[t], v);
			const int bt = base0 + ((t / 3) * S + t % 3) * C;  // flat input index (int, as the reference's max_inds)
#pragma unroll
			for (int j = 0; j < VEC; j++) {
				const float y = round_as_stored<T>(fmaxf(fmaf(v[j], a[j], b[j]), 0.f), rnd);
				if (y > mv[j]) { mv[j] = y; mi[j] = bt + j; }
			}
		}
		stv<T, VEC>(out, i, mv);
#pragma unroll
		for (int h = 0; h < VEC / 4; h++) reinterpret_cast<int4 *>(inds)[(long long)i * (VEC / 4) + h] = make_int4(mi[4 * h], mi[4 * h + 1], mi[4 * h + 2], mi[4 * h + 3]);
	}
}
bool bn_pool_fwd_supported(int N, int S, int C, int k, int stride, int bf16) {
	const int VEC = vec_of(C, bf16);
	return k == 3 && stride == 2 && S % 2 == 0 && VEC >= 4 && kThreads % (C / VEC) == 0 && (long long)N * S * S * C < (1LL << 31);
}
void bn_pool_fwd(const void *x, const float *ab, int N, int S, int C, int rnd, int *max_inds, void *out, cudaStream_t st, int bf16) {
	const int VEC = vec_of(C, bf16);
	long long total = (long long)N * (S / 2) * (S / 2) * (C / VEC);
	int grid = (int)((total + kThreads - 1) / kThreads); grid = grid > kNumSMs * 3 * 4 ? kNumSMs * 3 * 4 : grid;
	if (bf16) bn_pool_fwd_kernel<bf16_t, 8><<<grid, kThreads, 0, st>>>((const bf16_t *)x, ab, N, S, C, 0, max_inds, (bf16_t *)out);
	else bn_pool_fwd_kernel<float, 4><<<grid, kThreads, 0, st>>>((const float *)x, ab, N, S, C, rnd, max_inds, (float *)out);
	RB_LAUNCH_CHECK();
	RB_TRACE("bn_pool_fwd_kernel", "N=%d S=%d C=%d grid=%d", N, S, C, grid);
}

// DX = false: partial sums of dy' and dy' (x - mean) per block; DX = true: dx = c1 dy' + c3 x + k.  dy' = relu'(x a + b) * (pool gradient).
template <typename T, int VEC> __device__ __forceinline__ float elem_of(const typename RawOf<VEC>::type &r, int j) {
	if constexpr (sizeof(T) == 2) {
		const uint32_t w = (j >> 1) == 0 ? r.x : ((j >> 1) == 1 ? r.y : ((j >> 1) == 2 ? r.z : r.w));
		return (j & 1) ? bf16_hi(w) : bf16_lo(w);
	} else return __uint_as_float(j == 0 ? r.x : (j == 1 ? r.y : (j == 2 ? r.z : r.w)));
}
template <typename T, int VEC, bool DX>
__global__ void __launch_bounds__(kThreads, 2) pool_bn_bwd_kernel(const int *__restrict__ inds, const T *__restrict__ dout, const T *__restrict__ x,
                                                                 const float *__restrict__ mab, const float *__restrict__ means,
                                                                 const float *__restrict__ coef, int N, int S, int C, int rnd,
                                                                 float *__restrict__ partials, T *__restrict__ dx) {
	using raw_t = typename RawOf<VEC>::type;
	const int So = S / 2, V = C / VEC;
	const int total = N * So * So * V;
	const int cv = (int)(threadIdx.x % V);
	float ma[VEC], mb[VEC], mu[VEC], c1[VEC], ck[VEC], c3[VEC], s[VEC], q[VEC];
#pragma unroll
	for (int j = 0; j < VEC; j++) {
		const int c = cv * VEC + j;
		ma[j] = mab[c]; mb[j] = mab[C + c];
		if constexpr (DX) { c1[j] = coef[c]; ck[j] = coef[C + c]; c3[j] = coef[2 * C + c]; }
		else { mu[j] = means[c]; s[j] = q[j] = 0.f; }
	}
	for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
		int p = i / V;
		const int b = p % So; p /= So;
		const int a = p % So;
		const int n = p / So;
		raw_t rd[4], rx[4];
		int4 ri[4][VEC / 4];
		bool ok[4];
		const int o0 = ((n * So + a) * So + b) * V + cv;       // pooled vector of window (a, b)
		const int v0 = ((n * S + 2 * a) * S + 2 * b) * V + cv;  // input vector of pixel (2a, 2b)
#pragma unroll
		for (int t = 0; t < 4; t++) {
			ok[t] = a + (t >> 1) < So && b + (t & 1) < So;
			if (ok[t]) {
				const int o = o0 + ((t >> 1) * So + (t & 1)) * V;
				rd[t] = ldraw<T, VEC>(dout, o);
#pragma unroll
				for (int k = 0; k < VEC / 4; k++) ri[t][k] = reinterpret_cast<const int4 *>(inds)[(long long)o * (VEC / 4) + k];
			}
			rx[t] = ldraw<T, VEC>(x, v0 + ((t >> 1) * S + (t & 1)) * V);
		}
#pragma unroll
		for (int px = 0; px < 4; px++) {  // pixel (2a + px / 2, 2b + px % 2) collects from windows t with (t / 2 <= px / 2) and (t % 2 <= px % 2)
			const int vi = v0 + ((px >> 1) * S + (px & 1)) * V;
			const int me = vi * VEC;
			float acc[VEC], xv[VEC];
			unpack<T, VEC>(rx[px], xv);
#pragma unroll
			for (int j = 0; j < VEC; j++) acc[j] = 0.f;
#pragma unroll
			for (int t = 0; t < 4; t++) {
				if ((t >> 1) > (px >> 1) || (t & 1) > (px & 1) || !ok[t]) continue;
#pragma unroll
				for (int k = 0; k < VEC / 4; k++) {
					if (ri[t][k].x == me + 4 * k) acc[4 * k] += elem_of<T, VEC>(rd[t], 4 * k);
					if (ri[t][k].y == me + 4 * k + 1) acc[4 * k + 1] += elem_of<T, VEC>(rd[t], 4 * k + 1);
					if (ri[t][k].z == me + 4 * k + 2) acc[4 * k + 2] += elem_of<T, VEC>(rd[t], 4 * k + 2);
					if (ri[t][k].w == me + 4 * k + 3) acc[4 * k + 3] += elem_of<T, VEC>(rd[t], 4 * k + 3);
				}
			}
#pragma unroll
			for (int j = 0; j < VEC; j++) {
				// the unfused path stores the pool gradient in T before BatchNorm backward reads it
				float g = round_as_stored<T>(acc[j], 0);
				g = fmaf(xv[j], ma[j], mb[j]) > 0.f ? g : 0.f;
				if constexpr (DX) {
					const float r = fmaf(c1[j], g, fmaf(c3[j], xv[j], ck[j]));
					acc[j] = (sizeof(T) == 4 && rnd) ? round_tf32(r) : r;
				} else {
					s[j] += g;
					q[j] += g * (xv[j] - mu[j]);
				}
			}
			if constexpr (DX) stv<T, VEC>(dx, vi, acc);
		}
	}
	if constexpr (!DX) {
		// deterministic in-block combine, as in bn_reduce_kernel: thread t < V adds the kThreads / V threads of its column in a fixed order
		__shared__ float red[kThreads][2 * VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][VEC + j] = q[j]; }
		__syncthreads();
		if ((int)threadIdx.x < V) {
			float ss[VEC], qq[VEC];
#pragma unroll
			for (int j = 0; j < VEC; j++) ss[j] = qq[j] = 0.f;
			for (int k = threadIdx.x; k < kThreads; k += V)
#pragma unroll
				for (int j = 0; j < VEC; j++) { ss[j] += red[k][j]; qq[j] += red[k][VEC + j]; }
			float *o = partials + (size_t)blockIdx.x * 2 * C;
#pragma unroll
			for (int j = 0; j < VEC; j++) { o[cv * VEC + j] = ss[j]; o[C + cv * VEC + j] = qq[j]; }
		}
	}
}
// BatchNorm backward of the stem fed by the max pool's gradient gather; mab = the forward's folded scale / shift [2][C]
void pool_bn_bwd(const int *max_inds, const void *dpool, const void *x, const float *gamma, const float *means, const float *vars, float eps,
                 const float *mab, int N, int S, int C, float *dgamma, float *dbeta, void *dx, float *partials, int max_blocks, float *coef,
                 int rnd, cudaStream_t st, int bf16) {
	const int VEC = vec_of(C, bf16);
	const long long total = (long long)N * (S / 2) * (S / 2) * (C / VEC);
	const long long blocks = (total + kThreads - 1) / kThreads;
	int cap = kNumSMs * 2;  // one whole wave of the reduce pass: the fold cost grows with the number of partial blocks
	if (cap > max_blocks) cap = max_blocks;
	const int g1 = (int)(blocks > cap ? cap : blocks), g2 = (int)(blocks > kNumSMs * 2 * 4 ? kNumSMs * 2 * 4 : blocks);
	const long long rows = (long long)N * S * S;
	if (bf16) pool_bn_bwd_kernel<bf16_t, 8, false><<<g1, kThreads, 0, st>>>(max_inds, (const bf16_t *)dpool, (const bf16_t *)x, mab, means, nullptr, N, S, C, 0, partials, nullptr);
	else pool_bn_bwd_kernel<float, 4, false><<<g1, kThreads, 0, st>>>(max_inds, (const float *)dpool, (const float *)x, mab, means, nullptr, N, S, C, 0, partials, nullptr);
	RB_LAUNCH_CHECK();
	launch_k(2, bn_bwd_finalize_kernel, ceil_div(C, kFinC), dim3(kFinC, kFinS), 0, st, (const float *)partials, g1, 1.0 / (double)rows, C, gamma, means, vars, eps, dgamma, dbeta, coef);
	RB_LAUNCH_CHECK();
	if (bf16) pool_bn_bwd_kernel<bf16_t, 8, true><<<g2, kThreads, 0, st>>>(max_inds, (const bf16_t *)dpool, (const bf16_t *)x, mab, nullptr, coef, N, S, C, 0, nullptr, (bf16_t *)dx);
	else pool_bn_bwd_kernel<float, 4, true><<<g2, kThreads, 0, st>>>(max_inds, (const float *)dpool, (const float *)x, mab, nullptr, coef, N, S, C, rnd, nullptr, (float *)dx);
	RB_LAUNCH_CHECK();
	RB_TRACE("pool_bn_bwd_kernel", "N=%d S=%d C=%d grids=%d,%d", N, S, C, g1, g2);
}

// ------------------------------------------------------------------------------------------- average pool
// activations in T, pooled values and their gradient in fp32 (the FC head stays fp32)
template <typename T>
__global__ void avgpool_fwd_kernel(const T *__restrict__ x, int N, int SS, int C, float *__restrict__ out) {
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (n, c)
	if (i >= (long long)N * C) return;
	const int c = (int)(i % C), n = (int)(i / C);
	float s = 0.f;
	for (int p = 0; p < SS; p++) s += ld1<T>(x, ((long long)n * SS + p) * C + c);
	out[i] = s / (float)SS;
}
void avgpool_fwd(const void *x, int N, int S, int C, float *out, cudaStream_t st, int bf16) {
	if (bf16) avgpool_fwd_kernel<bf16_t><<<ceil_div((long long)N * C, 256), 256, 0, st>>>((const bf16_t *)x, N, S * S, C, out);
	else avgpool_fwd_kernel<float><<<ceil_div((long long)N * C, 256), 256, 0, st>>>((const float *)x, N, S * S, C, out);
	RB_LAUNCH_CHECK();
}
template <typename T>
__global__ void avgpool_bwd_kernel(const float *__restrict__ dp, int N, int SS, int C, T *__restrict__ din) {
	const long long total = (long long)N * SS * C;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int c = (int)(i % C);
		const int n = (int)(i / ((long long)SS * C));
		st1<T>(din, i, dp[(long long)n * C + c] / (float)SS);
	}
}
// 128-bit stores (the scalar kernel above wrote the 103 MB of the last block's output gradient at 1.3 TB/s); same value per element
template <typename T, int VEC>
__global__ void __launch_bounds__(256) avgpool_bwd_vec_kernel(const float *__restrict__ dp, int N, int SS, int C, T *__restrict__ din) {
	const int V = C / VEC, total = N * SS * V;
	const float inv = (float)SS;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
		const int cv = i % V, n = i / (SS * V);
		float v[VEC];
#pragma unroll
		for (int j = 0; j < VEC; j++) v[j] = dp[n * C + cv * VEC + j] / inv;
		stv<T, VEC>(din, i, v);
	}
}
void avgpool_bwd(const float *dpooled, int N, int S, int C, void *din, cudaStream_t st, int bf16) {
	long long total = (long long)N * S * S * C;
	const int VEC = bf16 ? 8 : 4;
	if (C % VEC == 0 && total < (1LL << 31)) {
		long long nv = total / VEC;
		int grid = (int)((nv + 255) / 256); grid = grid > kMaxFlatBlocks ? kMaxFlatBlocks : grid;
		if (bf16) avgpool_bwd_vec_kernel<bf16_t, 8><<<grid, 256, 0, st>>>(dpooled, N, S * S, C, (bf16_t *)din);
		else avgpool_bwd_vec_kernel<float, 4><<<grid, 256, 0, st>>>(dpooled, N, S * S, C, (float *)din);
		RB_LAUNCH_CHECK();
		return;
	}
	int grid = (int)((total + 255) / 256); grid = grid > kMaxFlatBlocks * 4 ? kMaxFlatBlocks * 4 : grid;
	if (bf16) avgpool_bwd_kernel<bf16_t><<<grid, 256, 0, st>>>(dpooled, N, S * S, C, (bf16_t *)din);
	else avgpool_bwd_kernel<float><<<grid, 256, 0, st>>>(dpooled, N, S * S, C, (float *)din);
	RB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------- softmax + cross entropy
// one warp per row; max-subtracted (reference resnet_cudnn.cu:568-587), dlogits = pred - onehot with no 1/N.
__global__ void softmax_ce_kernel(const float *__restrict__ logits, const int *__restrict__ labels, int N, int L, float *__restrict__ pred,
                                  float *__restrict__ dlogits, float *__restrict__ row_loss, int *__restrict__ row_wrong) {
	const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (row >= N) return;
	const float *x = logits + (size_t)row * L;
	float mx = -INFINITY;
	for (int j = lane; j < L; j += 32) mx = fmaxf(mx, x[j]);
	for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	float sum = 0.f;
	for (int j = lane; j < L; j += 32) sum += expf(x[j] - mx);
	for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
	const int lab = labels[row];
	const float pl = expf(x[lab] - mx) / sum;
	int wrong = 0;
	for (int j = lane; j < L; j += 32) {
		const float p = expf(x[j] - mx) / sum;
		pred[(size_t)row * L + j] = p;
		if (dlogits) dlogits[(size_t)row * L + j] = (j == lab) ? p - 1.f : p;
		if (j != lab && p >= pl) wrong = 1;
	}
	wrong = __any_sync(0xffffffffu, wrong);
	if (lane == 0) {
		if (row_loss) row_loss[row] = -logf(pl);
		if (row_wrong) row_wrong[row] = wrong;
	}
}
void softmax_ce(const float *logits, const int *labels, int N, int L, float *pred, float *dlogits, float *row_loss, int *row_wrong,
                cudaStream_t st) {
	softmax_ce_kernel<<<ceil_div((long long)N * 32, 128), 128, 0, st>>>(logits, labels, N, L, pred, dlogits, row_loss, row_wrong);
	RB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------- Adam
// reference resnet.cu:605-662 fused: one pass reads p,g,m,v and writes p,m,v and g=0 (the reference memsets the
// gradients afterwards, resnet.cu:2972-2975).  Non-finite gradient: moments kept; non-finite result: param kept.
__global__ void __launch_bounds__(kThreads) adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                                                       long long n4, float lr, float wd, float b1, float b2, float cb1, float cb2, float eps,
                                                       int *__restrict__ bad) {
	const long long T = (long long)gridDim.x * kThreads;
	int nbad = 0;
	const float ib1 = 1.f - cb1, ib2 = 1.f - cb2;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += T) {
		float4 P = reinterpret_cast<float4 *>(p)[i], G = reinterpret_cast<float4 *>(g)[i];
		float4 M = reinterpret_cast<float4 *>(m)[i], Vv = reinterpret_cast<float4 *>(v)[i];
		float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {G.x, G.y, G.z, G.w}, mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {Vv.x, Vv.y, Vv.z, Vv.w};
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (isfinite(gg[j])) {
				const float gd = gg[j] + wd * pp[j];
				mm[j] = b1 * mm[j] + (1.f - b1) * gd;
				vv[j] = b2 * vv[j] + (1.f - b2) * gd * gd;
			} else nbad++;
			const float ma = mm[j] / ib1, va = vv[j] / ib2;
			const float np_ = pp[j] - (lr * (ma / (sqrtf(va) + eps)) + wd * pp[j]);
			if (isfinite(np_)) pp[j] = np_; else nbad++;
		}
		reinterpret_cast<float4 *>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
		reinterpret_cast<float4 *>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
		reinterpret_cast<float4 *>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
		reinterpret_cast<float4 *>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	if (nbad && bad) atomicAdd(bad, nbad);
}
void adam_step(float *p, float *g, float *m, float *v, long long n, float lr, float wd, float b1, float b2, float cur_b1, float cur_b2,
               float eps, int *bad, cudaStream_t st) {
	if (n % 4) { set_error("adam_step: arena length %lld not a multiple of 4", n); return; }
	long long n4 = n / 4;
	int grid = (int)((n4 + kThreads * 2 - 1) / (kThreads * 2)); grid = grid < 1 ? 1 : (grid > kMaxFlatBlocks ? kMaxFlatBlocks : grid);
	adam_kernel<<<grid, kThreads, 0, st>>>(p, g, m, v, n4, lr, wd, b1, b2, cur_b1, cur_b2, eps, bad);
	RB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------- fp32 SGEMM (FC head)
// 64x64 tile, BK 16, 4x4 per thread.  Used for the 2048x1000 fully-connected layer (0.03% of step FLOPs).
// blockIdx.z = split-K slice: slice z reduces k in [z*kc, (z+1)*kc) into its own [M][N] plane of Cm (summed in a fixed order by
// splitk_sum_kernel): at batch 256 the forward GEMM has only 64 output tiles, one wave of 128 serial K steps (360 us).
__global__ void __launch_bounds__(256) sgemm_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ Cm, int M, int N, int K,
                                                   int ta, int tb, int kc) {
	__shared__ float As[16][65], Bs[16][65];
	const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
	const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
	float acc[4][4] = {};
	const int kbeg = blockIdx.z * kc, kend = min(K, kbeg + kc);  // K stays the leading dimension of the operands
	Cm += (size_t)blockIdx.z * M * N;
	for (int k0 = kbeg; k0 < kend; k0 += 16) {
		for (int e = threadIdx.x; e < 1024; e += 256) {
			int kk, mm;
			if (ta) { mm = e % 64; kk = e / 64; } else { kk = e % 16; mm = e / 16; }
			const int gm = m0 + mm, gk = k0 + kk;
			As[kk][mm] = (gm < M && gk < kend) ? (ta ? A[(size_t)gk * M + gm] : A[(size_t)gm * K + gk]) : 0.f;
			int kb, nn;
			if (tb) { kb = e % 16; nn = e / 16; } else { nn = e % 64; kb = e / 64; }
			const int gn = n0 + nn, gkb = k0 + kb;
			Bs[kb][nn] = (gn < N && gkb < kend) ? (tb ? B[(size_t)gn * K + gkb] : B[(size_t)gkb * N + gn]) : 0.f;
		}
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < 16; kk++) {
			float a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
			if (gm < M && gn < N) Cm[(size_t)gm * N + gn] = acc[i][j];
		}
}
__global__ void splitk_sum_kernel(const float *__restrict__ planes, int nsplit, long long n, float *__restrict__ out) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		float s = 0.f;
		for (int z = 0; z < nsplit; z++) s += planes[(long long)z * n + i];
		out[i] = s;
	}
}
// ws != NULL (>= sgemm_ws_floats(M, N) floats): split K over enough slices to fill the GPU, deterministic two-pass sum
size_t sgemm_ws_floats(int M, int N) { return (size_t)8 * M * N; }
void sgemm(const float *A, const float *B, float *Cm, int M, int N, int K, int ta, int tb, cudaStream_t st, float *ws) {
	const int tiles = ceil_div(N, 64) * ceil_div(M, 64);
	int nsplit = 1;
	if (ws) {
		nsplit = ceil_div(2 * kNumSMs, tiles);
		nsplit = nsplit > 8 ? 8 : nsplit;
		while (nsplit > 1 && K / nsplit < 64) nsplit--;
	}
	const int kc = ceil_div(ceil_div(K, nsplit), 16) * 16;
	nsplit = ceil_div(K, kc);
	dim3 grid(ceil_div(N, 64), ceil_div(M, 64), nsplit);
	sgemm_kernel<<<grid, 256, 0, st>>>(A, B, nsplit > 1 ? ws : Cm, M, N, K, ta, tb, kc);
	RB_LAUNCH_CHECK();
	if (nsplit > 1) {
		const long long n = (long long)M * N;
		splitk_sum_kernel<<<ceil_div(n, 256 * 4), 256, 0, st>>>(ws, nsplit, n, Cm);
		RB_LAUNCH_CHECK();
	}
}

// ------------------------------------------------------------------------------------------- weight re-layout
// One block re-lays 32 co x 32 ci x taps weights through shared memory so that all three streams are coalesced: the source
// [co][ci][tap] is read in runs of 32*taps floats, Wf[co][tap][ci] is written 32 ci at a time and Wd[ci][tap][co] 32 co at a time
// (the element-wise version wrote Wd at a stride of taps*cout elements and ran the 0.4 GB of traffic at 0.8 TB/s).  Jobs whose
// channel counts are not multiples of 32 (the 3-channel stem) take the element-wise path.
constexpr int kPackT = 32, kPackMaxTaps = 9;
__host__ __device__ inline bool pack_tiled(int cout, int cin, int taps) { return !(cout % kPackT) && !(cin % kPackT) && taps <= kPackMaxTaps; }
// number of blocks a job needs: one per 32 x 32 x taps tile, or one per 4096 elements on the element-wise path
int pack_job_blocks(int cout, int cin, int taps) {
	if (pack_tiled(cout, cin, taps)) return (cout / kPackT) * (cin / kPackT);
	return ceil_div((long long)cout * cin * taps, 4096);
}
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackJob *__restrict__ jobs, int njobs, int rnd) {
	__shared__ float sm[kPackT][kPackT * kPackMaxTaps + 1];
	// jobs[].first_block is the prefix sum of pack_job_blocks: find this block's job
	int lo = 0, hi = njobs - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (jobs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
	}
	const PackJob jb = jobs[lo];
	const int tile = blockIdx.x - jb.first_block, taps = jb.taps;
	if (!pack_tiled(jb.cout, jb.cin, taps)) {
		const long long total = (long long)jb.cout * jb.cin * taps, end = min(total, (long long)(tile + 1) * 4096);
		for (long long i = (long long)tile * 4096 + threadIdx.x; i < end; i += 256) {
			const int tap = (int)(i % taps), ci = (int)((i / taps) % jb.cin), co = (int)(i / ((long long)taps * jb.cin));
			float w = jb.src[i];
			if (rnd && sizeof(T) == 4) w = round_tf32(w);
			st1<T>((T *)jb.wf, ((long long)co * taps + tap) * jb.cin + ci, w);
			if (jb.wd) st1<T>((T *)jb.wd, ((long long)ci * taps + tap) * jb.cout + co, w);
		}
		return;
	}
	const int ci_tiles = jb.cin / kPackT, run = kPackT * taps;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int co0 = (tile / ci_tiles) * kPackT, ci0 = (tile % ci_tiles) * kPackT;
	for (int e = threadIdx.x; e < kPackT * run; e += 256) {  // sm[co_l][ci_l * taps + tap]
		const int co_l = e / run, r = e % run;
		float w = jb.src[((long long)(co0 + co_l) * jb.cin + ci0) * taps + r];
		if (rnd && sizeof(T) == 4) w = round_tf32(w);
		sm[co_l][r] = w;
	}
	__syncthreads();
	for (int p = warp; p < kPackT * taps; p += 8) {  // p = (row, tap); lane = the contiguous output index
		const int row = p / taps, tap = p % taps;
		st1<T>((T *)jb.wf, ((long long)(co0 + row) * taps + tap) * jb.cin + ci0 + lane, sm[row][lane * taps + tap]);
		if (jb.wd) st1<T>((T *)jb.wd, ((long long)(ci0 + row) * taps + tap) * jb.cout + co0 + lane, sm[lane][row * taps + tap]);
	}
}
void pack_weights(const PackJob *jobs_dev, int njobs, int total_blocks, int rnd, cudaStream_t st, int bf16) {
	if (njobs <= 0 || total_blocks <= 0) return;
	if (bf16) pack_weights_kernel<bf16_t><<<total_blocks, 256, 0, st>>>(jobs_dev, njobs, rnd);
	else pack_weights_kernel<float><<<total_blocks, 256, 0, st>>>(jobs_dev, njobs, rnd);
	RB_LAUNCH_CHECK();
}

// fp32 <-> bf16 conversion of a flat buffer (test entry points and the single-operator C API)
__global__ void f32_to_bf16_kernel(const float *__restrict__ s, long long n, bf16_t *__restrict__ d) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) st1<bf16_t>(d, i, s[i]);
}
__global__ void bf16_to_f32_kernel(const bf16_t *__restrict__ s, long long n, float *__restrict__ d) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) d[i] = ld1<bf16_t>(s, i);
}
void convert_f32_to_bf16(const float *s, long long n, void *d, cudaStream_t st) {
	int grid = (int)((n + 255) / 256); grid = grid < 1 ? 1 : (grid > kMaxFlatBlocks * 4 ? kMaxFlatBlocks * 4 : grid);
	f32_to_bf16_kernel<<<grid, 256, 0, st>>>(s, n, (bf16_t *)d);
	RB_LAUNCH_CHECK();
}
void convert_bf16_to_f32(const void *s, long long n, float *d, cudaStream_t st) {
	int grid = (int)((n + 255) / 256); grid = grid < 1 ? 1 : (grid > kMaxFlatBlocks * 4 ? kMaxFlatBlocks * 4 : grid);
	bf16_to_f32_kernel<<<grid, 256, 0, st>>>((const bf16_t *)s, n, d);
	RB_LAUNCH_CHECK();
}

// Three shapes of the same fixed-order sum (deterministic), chosen per layer from the ncu launch list of one step:
//  * many splits (>= 48: the 56x56 / 28x28 layers, small dW): 32 outputs x 8 split lanes per block, lane y sums splits y, y+8, ...
//    (independent coalesced loads), shared memory combines the 8 lanes;
//  * few splits, 1x1: one thread per output, reads and writes both contiguous;
//  * few splits, 3x3 (large dW): threads follow the OUTPUT order [co][ci][tap] so the 75 MB of writes are coalesced; the strided
//    reads of partial[tap][co][ci] hit lines the neighbouring threads of the block use completely.
constexpr int kWrX = 32, kWrY = 8;
__global__ void __launch_bounds__(kWrX * kWrY) wgrad_reduce_lanes_kernel(const float *__restrict__ partial, int splits, int cout, int cin, int taps,
                                                                          float *__restrict__ dw) {
	__shared__ float sm[kWrY][kWrX];
	const long long per = (long long)taps * cout * cin;
	const int tx = threadIdx.x, ty = threadIdx.y;
	pdl_wait();
	pdl_trigger();
	for (long long base = (long long)blockIdx.x * kWrX; base < per; base += (long long)gridDim.x * kWrX) {
		const long long i = base + tx;  // indexes partial [tap][co][ci]
		float s = 0.f;
		if (i < per) {
#pragma unroll 4
			for (int sp = ty; sp < splits; sp += kWrY) s += partial[(long long)sp * per + i];
		}
		sm[ty][tx] = s;
		__syncthreads();
		if (ty == 0 && i < per) {
#pragma unroll
			for (int y = 1; y < kWrY; y++) s += sm[y][tx];
			const int ci = (int)(i % cin);
			const int co = (int)((i / cin) % cout);
			const int tap = (int)(i / ((long long)cin * cout));
			dw[((long long)co * cin + ci) * taps + tap] = s;
		}
		__syncthreads();
	}
}
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, int splits, int cout, int cin, int taps, float *__restrict__ dw) {
	const long long per = (long long)taps * cout * cin, cc = (long long)cout * cin;
	for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < per; o += (long long)gridDim.x * blockDim.x) {
		// o indexes dw [co][ci][tap]
		const int tap = (int)(o % taps);
		const long long i = (long long)tap * cc + o / taps;
		float s = 0.f;
#pragma unroll 4
		for (int sp = 0; sp < splits; sp++) s += partial[(long long)sp * per + i];
		dw[o] = s;
	}
}
// The same sums (same order: bit-identical) with 128-bit accesses on both sides.  A block owns kWvPairs consecutive (co, ci) pairs:
// thread (tap, quad) sums the float4 of pairs 4*quad..+3 of its tap plane over the splits (a tap plane's 32 quads are 512 contiguous
// bytes; four loads in flight per thread), the [pair][tap] re-layout goes through shared memory, and the block's kWvPairs * TAPS outputs
// -- contiguous in dW -- leave as float4 rows.  The scalar kernel above ran the 1x1 / large 3x3 reduces at ~2.2 TB/s.
constexpr int kWvPairs = 128;
template <int TAPS, int GROUPS>  // GROUPS independent 128-pair groups per block (TAPS = 1: 8 groups = 256 threads)
__global__ void __launch_bounds__(32 * TAPS * GROUPS) wgrad_reduce_vec_kernel(const float *__restrict__ partial, int splits, long long cc, float *__restrict__ dw) {
	__shared__ float sm[TAPS == 1 ? 1 : kWvPairs * TAPS];
	static_assert(TAPS == 1 || GROUPS == 1, "the shared-memory re-layout is per block");
	const long long per = cc * TAPS;
	const int grp = threadIdx.x / (32 * TAPS), tap = (threadIdx.x / 32) % TAPS, quad = threadIdx.x % 32;
	pdl_wait();
	pdl_trigger();
	for (long long p0 = ((long long)blockIdx.x * GROUPS + grp) * kWvPairs; p0 < cc; p0 += (long long)gridDim.x * GROUPS * kWvPairs) {
		const float4 *src = reinterpret_cast<const float4 *>(partial + (long long)tap * cc + p0) + quad;
		const long long step = per / 4;
		float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
		int sp = 0;
		for (; sp + 4 <= splits; sp += 4) {
			const float4 a = src[(long long)sp * step], b = src[(long long)(sp + 1) * step], c = src[(long long)(sp + 2) * step], d = src[(long long)(sp + 3) * step];
			s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
			s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
			s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
			s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
		}
		for (; sp < splits; sp++) {
			const float4 a = src[(long long)sp * step];
			s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
		}
		if (TAPS == 1) {
			reinterpret_cast<float4 *>(dw + p0)[quad] = s;
		} else {
			__syncthreads();  // the previous round's readers are done
			sm[(4 * quad) * TAPS + tap] = s.x;
			sm[(4 * quad + 1) * TAPS + tap] = s.y;
			sm[(4 * quad + 2) * TAPS + tap] = s.z;
			sm[(4 * quad + 3) * TAPS + tap] = s.w;
			__syncthreads();
			// kWvPairs * TAPS floats = 32 * TAPS float4: one per thread
			reinterpret_cast<float4 *>(dw + p0 * TAPS)[threadIdx.x] = reinterpret_cast<const float4 *>(sm)[threadIdx.x];
		}
	}
}
void wgrad_reduce(const float *partial, int splits, int cout, int cin, int taps, float *dw, cudaStream_t st) {
	long long per = (long long)taps * cout * cin;
	if (splits >= 48) {
		int grid = (int)((per + kWrX - 1) / kWrX); grid = grid > kNumSMs * 32 ? kNumSMs * 32 : grid;
		launch_k(3, wgrad_reduce_lanes_kernel, grid, dim3(kWrX, kWrY), 0, st, partial, splits, cout, cin, taps, dw);
	} else if ((taps == 1 || taps == 9) && ((long long)cout * cin) % (8 * kWvPairs) == 0 && (uintptr_t)partial % 16 == 0 && (uintptr_t)dw % 16 == 0 &&
	           !getenv("RESNET_B200_WGRAD_REDUCE_SCALAR")) {
		const long long cc = (long long)cout * cin;
		if (taps == 1) {
			int grid = (int)(cc / (8 * kWvPairs)); grid = grid > kNumSMs * 8 ? kNumSMs * 8 : grid;
			launch_k(3, wgrad_reduce_vec_kernel<1, 8>, grid, 256, 0, st, partial, splits, cc, dw);
		} else {
			int grid = (int)(cc / kWvPairs); grid = grid > kNumSMs * 7 ? kNumSMs * 7 : grid;
			launch_k(3, wgrad_reduce_vec_kernel<9, 1>, grid, 32 * 9, 0, st, partial, splits, cc, dw);
		}
	} else {
		int grid = (int)((per + 255) / 256); grid = grid > kMaxFlatBlocks * 4 ? kMaxFlatBlocks * 4 : grid;
		wgrad_reduce_kernel<<<grid, 256, 0, st>>>(partial, splits, cout, cin, taps, dw);
	}
	RB_LAUNCH_CHECK();
}

}  // namespace rb
