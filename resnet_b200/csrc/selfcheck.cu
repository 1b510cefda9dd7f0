// selfcheck.cu -- in-situ verification of the tensor-core convolutions at the REAL batch size and on the trainer's REAL tensors.
//
// With resnet_b200_selfcheck(1) set before init_trainer, every tcgen05 convolution launch of forward_pass / backwards_pass (fprop,
// dgrad incl. the accumulating residual-join variant, wgrad, and the stem) is re-derived right after it ran by the fp32 SIMT
// implicit GEMM (simt_conv.cu, the device-side restatement of the reference's doConvolution / convolutionDerivInput /
// convolutionDerivWeights, resnet.cu:109-281) from the SAME input buffers, and the largest deviation relative to the tensor's
// largest magnitude is recorded per kernel family.  Rounding noise therefore cannot compound or flip ReLU masks between layers,
// and the plans are exactly the ones the benchmark runs (persistent multi-wave scheduling, resident weight operand, paired co
// tiles, two epilogue groups, >= 48-way split-K) -- tests/test_gpu_network.py::test_batch256_step_selfcheck.
// Test infrastructure inside the product library: costs nothing unless switched on (one flag test per conv launch).
#include "engine.h"
#include "../../include/resnet_b200.h"

namespace rb {

int g_selfcheck = 0;  // resnet_b200_selfcheck(): trainers created while set carry the checker

struct SelfCheck {
	float *xf, *dyf, *ref, *got;       // fp32 scratch: conv input, output gradient, re-derived result, widened copy of the actual result
	float *wf, *wd, *dwref;             // fp32 packed weights, re-derived weight gradient
	PackJob *job;
	float *red;                         // device [2]: max |got - ref|, max |ref| (as non-negative float bits)
	long long max_act, max_w;
	float worst[3];                     // per family: 0 fprop, 1 dgrad, 2 wgrad
	char where[3][96];
	long long checks[3];
};

__global__ void sc_maxdiff_kernel(const float *__restrict__ a, const float *__restrict__ b, long long n, unsigned int *__restrict__ out) {
	float md = 0.f, mr = 0.f;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const float x = a[i], y = b[i];
		float d = fabsf(x - y);
		if (!(d == d)) d = __int_as_float(0x7f800000);  // NaN anywhere is an infinite deviation
		md = fmaxf(md, d);
		mr = fmaxf(mr, fabsf(y));
	}
	for (int o = 16; o; o >>= 1) { md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, o)); mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o)); }
	if ((threadIdx.x & 31) == 0) { atomicMax(out, __float_as_uint(md)); atomicMax(out + 1, __float_as_uint(mr)); }
}

static SelfCheck *sc_of(Engine *e) { return (SelfCheck *)e->selfcheck; }

void selfcheck_init(Engine *e, long long max_act_elems, long long max_w_elems) {
	e->selfcheck = nullptr;
	if (!g_selfcheck) return;
	SelfCheck *s = new SelfCheck();
	memset(s, 0, sizeof(*s));
	s->max_act = max_act_elems; s->max_w = max_w_elems;
	for (float **p : {&s->xf, &s->dyf, &s->ref, &s->got}) RB_CUDA(cudaMalloc(p, (size_t)max_act_elems * sizeof(float)));
	for (float **p : {&s->wf, &s->wd, &s->dwref}) RB_CUDA(cudaMalloc(p, (size_t)max_w_elems * sizeof(float)));
	RB_CUDA(cudaMalloc(&s->job, sizeof(PackJob)));
	RB_CUDA(cudaMalloc(&s->red, 2 * sizeof(float)));
	e->selfcheck = s;
}
void selfcheck_release(Engine *e) {
	SelfCheck *s = sc_of(e);
	if (!s) return;
	for (float *p : {s->xf, s->dyf, s->ref, s->got, s->wf, s->wd, s->dwref, s->red}) cudaFree(p);
	cudaFree(s->job);
	delete s;
	e->selfcheck = nullptr;
}

// fp32 view of an activation tensor of the engine's storage type (bf16 tensors are widened into `scratch`)
static const float *widen(Engine *e, const void *act, long long n, float *scratch) {
	if (!e->bf16) return (const float *)act;
	convert_bf16_to_f32(act, n, scratch, e->stream);
	return scratch;
}
static void record(Engine *e, int fam, const char *what, const ConvGeom &g, const float *got, const float *ref, long long n) {
	SelfCheck *s = sc_of(e);
	RB_CUDA(cudaMemsetAsync(s->red, 0, 2 * sizeof(float), e->stream));
	sc_maxdiff_kernel<<<kNumSMs * 4, 256, 0, e->stream>>>(got, ref, n, (unsigned int *)s->red);
	float h[2] = {0.f, 0.f};
	RB_CUDA(cudaMemcpyAsync(h, s->red, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
	RB_CUDA(cudaStreamSynchronize(e->stream));
	const float rel = h[0] / fmaxf(h[1], 1e-30f);
	s->checks[fam]++;
	if (rel > s->worst[fam] || !(rel == rel)) {
		s->worst[fam] = (rel == rel) ? rel : __builtin_inff();
		snprintf(s->where[fam], sizeof(s->where[fam]), "%s %dx%d/%d %d->%d @%d (max|diff| %.3g, max|ref| %.3g)", what, g.k, g.k, g.stride, g.cin, g.cout, g.S, h[0], h[1]);
	}
}
static void pack_fp32(Engine *e, const ConvGeom &g, const float *w) {
	SelfCheck *s = sc_of(e);
	PackJob job{w, s->wf, s->wd, g.cout, g.cin, g.taps(), 0};
	RB_CUDA(cudaMemcpyAsync(s->job, &job, sizeof(job), cudaMemcpyHostToDevice, e->stream));
	pack_weights(s->job, 1, pack_job_blocks(g.cout, g.cin, g.taps()), 0, e->stream, 0);
}

// y was just produced from x by the tensor-core fprop (x: activation type, or the fp32 batch for the stem)
void selfcheck_fprop(Engine *e, const ConvGeom &g, const float *w, const void *x, bool x_is_fp32_batch, const void *y) {
	SelfCheck *s = sc_of(e);
	if (!s) return;
	pack_fp32(e, g, w);
	const float *xf = x_is_fp32_batch ? (const float *)x : widen(e, x, g.in_elems(), s->xf);
	simt_conv_fprop(g, xf, s->wf, s->ref, e->stream);
	record(e, 0, "fprop", g, widen(e, y, g.out_elems(), s->got), s->ref, g.out_elems());
}
// accumulate: call selfcheck_dgrad_snapshot(dx) BEFORE the tensor-core dgrad ran, so that the re-derivation starts from the same base
void selfcheck_dgrad_snapshot(Engine *e, const ConvGeom &g, const void *dx) {
	SelfCheck *s = sc_of(e);
	if (!s) return;
	if (e->bf16) convert_bf16_to_f32(dx, g.in_elems(), s->ref, e->stream);
	else RB_CUDA(cudaMemcpyAsync(s->ref, dx, (size_t)g.in_elems() * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
}
void selfcheck_dgrad(Engine *e, const ConvGeom &g, const float *w, const void *dy, const void *dx, int accumulate) {
	SelfCheck *s = sc_of(e);
	if (!s) return;
	pack_fp32(e, g, w);
	const float *dyf = widen(e, dy, g.out_elems(), s->dyf);
	simt_conv_dgrad(g, dyf, s->wd, s->ref, accumulate, e->stream);
	record(e, 1, accumulate ? "dgrad+=" : "dgrad", g, widen(e, dx, g.in_elems(), s->got), s->ref, g.in_elems());
}
void selfcheck_wgrad(Engine *e, const ConvGeom &g, const void *x, bool x_is_fp32_batch, const void *dy, const float *dw) {
	SelfCheck *s = sc_of(e);
	if (!s) return;
	const float *xf = x_is_fp32_batch ? (const float *)x : widen(e, x, g.in_elems(), s->xf);
	const float *dyf = widen(e, dy, g.out_elems(), s->dyf);
	simt_conv_wgrad(g, xf, dyf, s->dwref, e->stream);
	record(e, 2, "wgrad", g, dw, s->dwref, g.w_elems());
}

}  // namespace rb

using namespace rb;

extern "C" {

int resnet_b200_selfcheck(int enable) { g_selfcheck = enable ? 1 : 0; return 0; }

int resnet_b200_selfcheck_read(Train_ResNet *t, int family, float *worst_rel, long long *n_checks, char *where, int where_len) {
	Engine *e = engine_of(t);
	SelfCheck *s = e ? (SelfCheck *)e->selfcheck : nullptr;
	if (!s || family < 0 || family > 2) { set_error("selfcheck_read: trainer was not created with resnet_b200_selfcheck(1)"); return 1; }
	if (worst_rel) *worst_rel = s->worst[family];
	if (n_checks) *n_checks = s->checks[family];
	if (where && where_len > 0) snprintf(where, (size_t)where_len, "%s", s->where[family]);
	return 0;
}

}  // extern "C"
