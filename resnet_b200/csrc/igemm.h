// igemm.h -- host API of the tcgen05/TMA implicit-GEMM convolution (igemm.cu).
#pragma once
#include "common.cuh"

namespace rb {

struct TcPlan;  // tensor maps + tile schedule for one (layer, buffers) pair; built once at init, launched every step

// `bf16` selects the element type of the activation / packed-weight tensors of a plan: 0 = fp32 storage with kind::tf32 MMAs,
// 1 = bf16 storage with kind::f16 MMAs.  Accumulation is fp32 in TMEM either way; weight gradients are always fp32.
bool tc_supported(const ConvGeom &g, int bf16);
// y[N][So][So][Cout] = conv(x[N][S][S][Cin], Wf[Cout][tap][Cin])
TcPlan *tc_make_fprop(const ConvGeom &g, const void *x, const void *wf, void *y, int bf16);
// dx[N][S][S][Cin] (+)= conv^T(dy[N][So][So][Cout], Wd[Cin][tap][Cout])
TcPlan *tc_make_dgrad(const ConvGeom &g, const void *dy, const void *wd, void *dx, int accumulate, int bf16);
// dw[Cout][Cin][k][k] = sum_pixels dy (x) x   (split-K partials in `workspace`, deterministic reduce)
TcPlan *tc_make_wgrad(const ConvGeom &g, const void *x, const void *dy, float *dw, float *workspace, size_t ws_bytes, int bf16);
size_t tc_wgrad_workspace_bytes(const ConvGeom &g, int bf16);
// stem 7x7/2, Cin = 3 on the same kernels: zero-bordered NHWC4 copy of the (fp32) batch + packed [Cout][7][taps][4] weights
bool tc_stem_supported(int S, int k, int cin, int cout, int stride, int bf16);
size_t stem_xp_bytes(int N, int S, int bf16);
size_t stem_wfs_bytes(int cout, int bf16);
void stem_pad_input(const float *x, int N, int S, void *xp, int round_tf32, int bf16, cudaStream_t st);
void stem_pack_weights(const float *w, int cout, void *wfs, int round_tf32, int bf16, cudaStream_t st);
TcPlan *tc_make_stem_fprop(int N, int S, int cout, const void *xp, const void *wfs, void *y, int bf16);
size_t tc_stem_wgrad_workspace_bytes(int N, int S, int cout, int bf16);
TcPlan *tc_make_stem_wgrad(int N, int S, int cout, const void *xp, const void *dy, float *dw, float *workspace, size_t ws_bytes, int bf16);
// fused BatchNorm statistics in the fprop epilogue: returns the number of partial rows (0 = not available for this plan)
// prezeroed: the caller keeps `partials` all-zero between uses (bn_finalize with zero_after clears exactly what it folded), so tc_run
// issues no memset in front of the kernel -- a memset node between two kernels also breaks the programmatic launch chain
int tc_attach_stats(TcPlan *pl, float *partials, int prezeroed = 0);
size_t tc_stats_floats(int cout);
void tc_run(TcPlan *pl, cudaStream_t st);
void tc_free(TcPlan *pl);
void tc_describe(const TcPlan *pl, char *buf, size_t n);

}  // namespace rb
