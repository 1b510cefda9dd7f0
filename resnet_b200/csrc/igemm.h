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
void tc_run(TcPlan *pl, cudaStream_t st);
void tc_free(TcPlan *pl);
void tc_describe(const TcPlan *pl, char *buf, size_t n);

}  // namespace rb
