// common.cuh -- shared declarations of libresnet_b200.so (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace rb {

// First error wins; readable through resnet_b200_last_error().  The reference ignores every CUDA status
// (SURVEY.md 8b "error conventions"); we keep the void signatures but record the failure.
void set_error(const char *fmt, ...);
const char *last_error();
bool has_error();
void clear_error();

#define RB_CUDA(call)                                                                              \
	do {                                                                                           \
		cudaError_t _e = (call);                                                                   \
		if (_e != cudaSuccess) rb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
	} while (0)

extern long long g_launches;  // kernels launched by this library (bench.py gpu_launches)
#define RB_LAUNCH_CHECK()          \
	do {                           \
		rb::g_launches++;          \
		RB_CUDA(cudaGetLastError()); \
	} while (0)

// RESNET_B200_TRACE=1: one "[k] <kernel> <label>" line on stderr per instrumented launch, in launch order; tools/ncu_summary.py
// joins them (one queue per kernel name) with ncu's launch list to label every kernel with its layer / tensor shape
bool trace_on();
#define RB_TRACE(kernel, ...)                     \
	do {                                          \
		if (rb::trace_on()) {                     \
			fprintf(stderr, "[k] %s ", kernel);   \
			fprintf(stderr, __VA_ARGS__);         \
			fprintf(stderr, "\n");                \
		}                                         \
	} while (0)

constexpr int kNumSMs = 148;  // B200

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (round 2).  A step is ~490 (ResNet-50) to ~1100 (ResNet-152) short launches on one stream; between two
// of them the GPU drains, launches the next grid and runs its prologue (for the convolution kernels: barrier init, TMEM allocation,
// tensor-map prefetch).  A kernel launched through launch_k with the programmatic-stream-serialization attribute may have its blocks
// scheduled while the previous kernel of the stream is still running; they run their global-memory-free prologue and then block in
// pdl_wait() (griddepcontrol.wait) until the previous grid has COMPLETED and its writes are visible -- the data dependency stays the
// full one of ordinary stream order, only launch latency and prologue move under the predecessor's tail.  Rules: every kernel that
// can be launched this way calls pdl_wait() in every thread before its first global-memory access (reads AND writes: the predecessor
// may still be reading what this kernel overwrites); pdl_trigger() lets ITS successor be scheduled.  Without the attribute both
// instructions are no-ops and the launch is serialized on both sides.
// Measured (profiles/r02_pdl_ab.txt, same-box A/B): pre-launching the CONVOLUTION kernels under the bandwidth-bound kernel before them
// is worth 0 / 0.3 / 1.6 % of the ResNet-50 TF32 / bf16 / ResNet-152 step; pre-launching the bandwidth-bound kernels costs 3-5 %
// (their blocks sit on the SMs next to a running convolution, or land unevenly behind another streaming kernel); the few-microsecond
// fold kernels gain 1 % on ResNet-152 and nothing on ResNet-50 -- so the default is mask 2: only the convolution kernels carry the
// attribute (RESNET_B200_PDL is a bit mask of kernel classes, see pdl_mode).
int pdl_mode();  // RESNET_B200_PDL: bit mask of the kernel classes launched with the attribute -- 1 streaming BatchNorm kernels, 2 convolutions, 4 the small fold kernels, 8 the split-K reduces of the weight gradients
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline void launch_k(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	// cls: 0 = streaming bandwidth-bound kernel, 1 = convolution kernel (one CTA per SM holding ~200 KB of shared memory), 2 = fold kernel (a few us),
	// 3 = split-K reduce behind a weight-gradient kernel
	cfg.numAttrs = ((pdl_mode() >> cls) & 1) ? 1 : 0;
	RB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// conv geometry shared by the SIMT and tcgen05 paths
struct ConvGeom {
	int N, S, cin, cout, k, stride;  // input spatial S x S, zero pad k/2, output So = S / stride
	__host__ __device__ int So() const { return S / stride; }
	__host__ __device__ int taps() const { return k * k; }
	__host__ __device__ long long in_elems() const { return (long long)N * S * S * cin; }
	__host__ __device__ long long out_elems() const { return (long long)N * So() * So() * cout; }
	__host__ __device__ long long w_elems() const { return (long long)cout * cin * k * k; }
};

// ---------------------------------------------------------------------------------------------
// bandwidth-bound kernels (bw_kernels.cu).  All take a stream.  Activation tensors are NHWC in one of two element types, chosen
// by the trailing `bf16` flag: fp32 (C % 4 == 0 for the vector path) or bf16 (C % 8 == 0); the arithmetic, the per-channel
// vectors (gamma, beta, means, vars, folded coefficients) and every partial sum are fp32 / fp64 in both modes.

// statistics of x[rows][C]: means, biased vars, and the folded scale/shift a = gamma*rstd, b = beta - mean*a
void bn_stats(const void *x, long long rows, int C, const float *gamma, const float *beta, float eps, float *means,
              float *vars, float *ab /* [2][C] */, float *partials, int max_blocks, cudaStream_t st, int bf16 = 0);
// finalize only (statistics partials already produced, e.g. by a conv epilogue): partials [nblk][2][C]
// zero_after: the kernel also clears the partials it folded (the fused-statistics buffer stays all-zero between convolutions)
void bn_finalize(float *partials, int nblk, long long rows, int C, const float *gamma, const float *beta, float eps,
                 float *means, float *vars, float *ab, cudaStream_t st, int zero_after = 0);
// y = act(x*a + b [+ residual]);  residual: res (identity) or res*a2 + b2 (projected, ab2 != NULL)
// bits_out != NULL: also stores the sign bits of y, one byte per 128-bit vector (bit j = element j of the vector is > 0): the
// 1-bit ReLU mask BatchNorm backward needs of a residual join's output (bn_bwd mask_bits)
void bn_apply(const void *x, const float *ab, long long rows, int C, int relu, const void *res, const float *ab2, void *y,
              int round_tf32, cudaStream_t st, int bf16 = 0, uint8_t *bits_out = nullptr);
// BatchNorm backward.  mask_src != NULL: dy is masked where mask_src <= 0 (ReLU).  Produces dgamma, dbeta and
// dx (may alias dy).  coef scratch [4][C].
// mask_ab != NULL ([2][C] folded scale/shift of the forward): the ReLU mask is recomputed as (x*a + b > 0) instead of read
void bn_bwd(const void *x, const void *dy, const void *mask_src, const float *gamma, const float *means, const float *vars,
            float eps, long long rows, int C, float *dgamma, float *dbeta, void *dx, float *partials, int max_blocks,
            float *coef, int round_tf32, cudaStream_t st, const float *mask_ab = nullptr, int bf16 = 0, void *masked_out = nullptr,
            const uint8_t *mask_bits = nullptr);
// mask_bits != NULL: the ReLU mask comes from bn_apply's bits_out instead of the sign of mask_src
// masked_out != NULL: the masked upstream gradient dy' (the identity shortcut's gradient) is also stored there
void relu_bwd(const void *y, const void *dy, long long n, void *dx, cudaStream_t st, int bf16 = 0);
void maxpool_fwd(const void *x, int N, int S, int C, int k, int stride, int *max_inds, void *out, cudaStream_t st, int bf16 = 0);
void maxpool_bwd(const int *max_inds, const void *dout, int N, int S, int C, int k, int stride, void *din, cudaStream_t st, int bf16 = 0);
// fused stem tail (bw_kernels.cu "fused stem tail"): pooled = maxpool3x3/2(relu(x * a + b)) without materialising the activation, and
// the stem BatchNorm's backward fed by the pool's gradient gather without materialising the activation's gradient
bool bn_pool_fwd_supported(int N, int S, int C, int k, int stride, int bf16);
void bn_pool_fwd(const void *x, const float *ab, int N, int S, int C, int round_tf32, int *max_inds, void *out, cudaStream_t st, int bf16 = 0);
void pool_bn_bwd(const int *max_inds, const void *dpool, const void *x, const float *gamma, const float *means, const float *vars, float eps,
                 const float *mab, int N, int S, int C, float *dgamma, float *dbeta, void *dx, float *partials, int max_blocks, float *coef,
                 int round_tf32, cudaStream_t st, int bf16 = 0);
// pooled values and their gradient are fp32 in both modes (the FC head is fp32)
void avgpool_fwd(const void *x, int N, int S, int C, float *out, cudaStream_t st, int bf16 = 0);
void avgpool_bwd(const float *dpooled, int N, int S, int C, void *din, cudaStream_t st, int bf16 = 0);
// pred = softmax(logits); dlogits = pred - onehot(labels) (no 1/N, reference resnet.cu:1806-1811);
// per-row loss = -log pred[label] and wrong flag (ties wrong, reference resnet.cu:3376)
void softmax_ce(const float *logits, const int *labels, int N, int L, float *pred, float *dlogits, float *row_loss,
                int *row_wrong, cudaStream_t st);
// fused Adam over a flat arena (reference resnet.cu:605-662); zeroes g; counts non-finite hits in *bad
void adam_step(float *p, float *g, float *m, float *v, long long n, float lr, float wd, float b1, float b2, float cur_b1,
               float cur_b2, float eps, int *bad, cudaStream_t st);
// C[M][N] = op(A) * op(B), fp32 FMA; ta: A stored [K][M]; tb: B stored [N][K]
// ws (optional, >= sgemm_ws_floats(M, N) floats): lets small-output GEMMs split K (deterministic two-pass sum)
void sgemm(const float *A, const float *B, float *Cm, int M, int N, int K, int ta, int tb, cudaStream_t st, float *ws = nullptr);
size_t sgemm_ws_floats(int M, int N);

// weight re-layout [Cout][Cin][k][k] (fp32 master) -> Wf [Cout][k*k][Cin] and Wd [Cin][k*k][Cout], tf32-rounded fp32 or bf16
// first_block: prefix sum of pack_job_blocks() over the job list (one launch covers all jobs, one block per 32x32xtaps tile)
struct PackJob { const float *src; void *wf; void *wd; int cout, cin, taps; int first_block; };
int pack_job_blocks(int cout, int cin, int taps);
void pack_weights(const PackJob *jobs_dev, int njobs, int total_blocks, int round_tf32, cudaStream_t st, int bf16 = 0);
// dW [Cout][Cin][k][k] = sum_s partial[s][tap][Cout][Cin]  (deterministic split-K reduce + re-layout)
void wgrad_reduce(const float *partial, int splits, int cout, int cin, int taps, float *dw, cudaStream_t st);
void convert_f32_to_bf16(const float *s, long long n, void *d, cudaStream_t st);
void convert_bf16_to_f32(const void *s, long long n, float *d, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// SIMT fp32 implicit-GEMM convolution (simt_conv.cu): exact-fp32 device-side checker and the C1 stem path.
void simt_conv_fprop(const ConvGeom &g, const float *x, const float *wf, float *y, cudaStream_t st);
void simt_conv_dgrad(const ConvGeom &g, const float *dy, const float *wd, float *dx, int accumulate, cudaStream_t st);
// writes dW in the public [Cout][Cin][k][k] layout (zeroes it first, split-K atomics)
void simt_conv_wgrad(const ConvGeom &g, const float *x, const float *dy, float *dw, cudaStream_t st);

}  // namespace rb
