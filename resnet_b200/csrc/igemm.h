// igemm.h -- host API of the tcgen05/TMA implicit-GEMM convolution (igemm.cu).
#pragma once
#include "common.cuh"

namespace rb {

struct TcPlan;  // tensor maps + tile schedule for one (layer, buffers) pair; built once at init, launched every step

bool tc_supported(const ConvGeom &g);
// y[N][So][So][Cout] = conv(x[N][S][S][Cin], Wf[Cout][tap][Cin])
TcPlan *tc_make_fprop(const ConvGeom &g, const float *x, const float *wf, float *y);
// dx[N][S][S][Cin] (+)= conv^T(dy[N][So][So][Cout], Wd[Cin][tap][Cout])
TcPlan *tc_make_dgrad(const ConvGeom &g, const float *dy, const float *wd, float *dx, int accumulate);
// dw[Cout][Cin][k][k] = sum_pixels dy (x) x   (split-K partials in `workspace`, deterministic reduce)
TcPlan *tc_make_wgrad(const ConvGeom &g, const float *x, const float *dy, float *dw, float *workspace, size_t ws_bytes);
size_t tc_wgrad_workspace_bytes(const ConvGeom &g);
// stem 7x7/2, Cin = 3 on the same kernels: zero-bordered NHWC4 copy of the batch + packed [Cout][7][8][4] weights
bool tc_stem_supported(int S, int k, int cin, int cout, int stride);
size_t stem_xp_elems(int N, int S);
void stem_pad_input(const float *x, int N, int S, float *xp, int round_tf32, cudaStream_t st);
void stem_pack_weights(const float *w, int cout, float *wfs, int round_tf32, cudaStream_t st);
TcPlan *tc_make_stem_fprop(int N, int S, int cout, const float *xp, const float *wfs, float *y);
size_t tc_stem_wgrad_workspace_bytes(int N, int S, int cout);
TcPlan *tc_make_stem_wgrad(int N, int S, int cout, const float *xp, const float *dy, float *dw, float *workspace, size_t ws_bytes);
// fused BatchNorm statistics in the fprop epilogue: returns the number of partial rows (0 = not available for this plan)
int tc_attach_stats(TcPlan *pl, float *partials);
size_t tc_stats_floats(int cout);
void tc_run(TcPlan *pl, cudaStream_t st);
void tc_free(TcPlan *pl);
void tc_describe(const TcPlan *pl, char *buf, size_t n);

}  // namespace rb
