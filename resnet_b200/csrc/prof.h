// prof.h -- optional per-kernel-family timing with CUDA events on the launching stream (bench.py's roofline leg).
// Off by default; when on, every instrumented launch is bracketed by an event pair (pooled, created lazily).
#pragma once
#include "common.cuh"

namespace rb {
// PROF_IGEMM_1X1: fprop / dgrad of the 1x1 convolutions, kept apart from the 3x3 / stem launches of PROF_IGEMM_KMAJOR because they are bound by HBM,
// not by the tensor pipe (work = FLOPs, work2 = algorithmic HBM bytes: input + output tensor once)
enum ProfFamily { PROF_IGEMM_KMAJOR = 0, PROF_IGEMM_WGRAD = 1, PROF_BN_ELTWISE = 2, PROF_STEM_SIMT = 3, PROF_OTHER = 4, PROF_IGEMM_1X1 = 5, PROF_NFAM = 6 };
void prof_enable(bool on);
bool prof_enabled();
void prof_begin(cudaStream_t st, int family, double work, double work2 = 0);  // work: algorithmic FLOPs (tensor) or bytes (HBM)
void prof_end(cudaStream_t st);
// sums over all records since the last reset (synchronises); returns 0 on success
int prof_read(int family, double *ms, long long *launches, double *work, double *work2 = nullptr);
void prof_reset();
struct ProfScope {
	cudaStream_t st;
	bool on;
	ProfScope(cudaStream_t s, int family, double work, double work2 = 0) : st(s), on(prof_enabled()) { if (on) prof_begin(st, family, work, work2); }
	~ProfScope() { if (on) prof_end(st); }
};
}  // namespace rb
