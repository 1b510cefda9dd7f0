// simt_conv.cu -- fp32 FMA implicit-GEMM convolution on CUDA cores.
//
// NOT the product convolution: the product path is the tcgen05/TMA implicit GEMM in igemm.cu.  This file is
// (a) the exact-fp32 device-side checker the parity tests compare the tensor-core kernels against at full
// size (the host oracle needs minutes there), selected with RESNET_B200_CONV=simt, and (b) the fp32 path of
// BASELINE config 1 (stem fwd+bwd in fp32).
// Semantics: reference resnet.cu:109-156 (fprop), 166-219 (dgrad), 227-281 (wgrad); zero pad k/2.
//
// GEMM views (64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread):
//   fprop : M = N*So*So pixels, Ncol = Cout,      K = taps*Cin   A = gathered x,  B = Wf[Cout][tap][Cin]
//   dgrad : M = N*S*S pixels,   Ncol = Cin,       K = taps*Cout  A = gathered dy, B = Wd[Cin][tap][Cout]
//   wgrad : M = Cout,           Ncol = taps*Cin,  K = N*So*So    A = dy^T,        B = gathered x (split-K, atomics)
#include "common.cuh"

namespace rb {

enum { MODE_FPROP = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

template <int MODE>
__global__ void __launch_bounds__(256) simt_conv_kernel(ConvGeom g, const float *__restrict__ src, const float *__restrict__ w_or_x,
                                                       float *__restrict__ out, int accumulate, long long k_per_split) {
	__shared__ float As[16][65], Bs[16][65];
	const int S = g.S, So = g.S / g.stride, k = g.k, half = g.k / 2, st = g.stride, cin = g.cin, cout = g.cout, taps = g.k * g.k;
	long long M, Ncol, K;
	if (MODE == MODE_FPROP) { M = (long long)g.N * So * So; Ncol = cout; K = (long long)taps * cin; }
	else if (MODE == MODE_DGRAD) { M = (long long)g.N * S * S; Ncol = cin; K = (long long)taps * cout; }
	else { M = cout; Ncol = (long long)taps * cin; K = (long long)g.N * So * So; }
	const long long m0 = (long long)blockIdx.y * 64, n0 = (long long)blockIdx.x * 64;
	long long kbeg = 0, kend = K;
	if (MODE == MODE_WGRAD) { kbeg = (long long)blockIdx.z * k_per_split; kend = kbeg + k_per_split < K ? kbeg + k_per_split : K; }
	const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
	float acc[4][4] = {};

	for (long long k0 = kbeg; k0 < kend; k0 += 16) {
		for (int e = threadIdx.x; e < 1024; e += 256) {
			// ---- A tile element (mm, kk)
			{
				int mm, kk;
				if (MODE == MODE_WGRAD) { mm = e % 64; kk = e / 64; } else { kk = e % 16; mm = e / 16; }
				const long long gm = m0 + mm, gk = k0 + kk;
				float v = 0.f;
				if (gm < M && gk < kend) {
					if (MODE == MODE_FPROP) {
						const int ci = (int)(gk % cin), tap = (int)(gk / cin), kh = tap / k, kw = tap % k;
						const int ow = (int)(gm % So), oh = (int)((gm / So) % So), n = (int)(gm / ((long long)So * So));
						const int ih = st * oh + kh - half, iw = st * ow + kw - half;
						if (ih >= 0 && ih < S && iw >= 0 && iw < S) v = src[(((long long)n * S + ih) * S + iw) * cin + ci];
					} else if (MODE == MODE_DGRAD) {
						const int co = (int)(gk % cout), tap = (int)(gk / cout), kh = tap / k, kw = tap % k;
						const int x = (int)(gm % S), y = (int)((gm / S) % S), n = (int)(gm / ((long long)S * S));
						const int th = y + half - kh, tw = x + half - kw;
						if (th >= 0 && tw >= 0 && th % st == 0 && tw % st == 0) {
							const int oh = th / st, ow = tw / st;
							if (oh < So && ow < So) v = src[(((long long)n * So + oh) * So + ow) * cout + co];
						}
					} else {
						v = src[gk * cout + gm];  // dy[pixel][co]
					}
				}
				As[kk][mm] = v;
			}
			// ---- B tile element (kk, nn)
			{
				int kk, nn;
				if (MODE == MODE_WGRAD) { nn = e % 64; kk = e / 64; } else { kk = e % 16; nn = e / 16; }
				const long long gn = n0 + nn, gk = k0 + kk;
				float v = 0.f;
				if (gn < Ncol && gk < kend) {
					if (MODE == MODE_WGRAD) {
						const int ci = (int)(gn % cin), tap = (int)(gn / cin), kh = tap / k, kw = tap % k;
						const int ow = (int)(gk % So), oh = (int)((gk / So) % So), n = (int)(gk / ((long long)So * So));
						const int ih = st * oh + kh - half, iw = st * ow + kw - half;
						if (ih >= 0 && ih < S && iw >= 0 && iw < S) v = w_or_x[(((long long)n * S + ih) * S + iw) * cin + ci];
					} else {
						v = w_or_x[gn * K + gk];  // Wf[co][K] or Wd[ci][K]
					}
				}
				Bs[kk][nn] = v;
			}
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
			const long long gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
			if (gm >= M || gn >= Ncol) continue;
			if (MODE == MODE_WGRAD) {
				const int ci = (int)(gn % cin), tap = (int)(gn / cin);
				atomicAdd(&out[((long long)gm * cin + ci) * taps + tap], acc[i][j]);  // public [co][ci][kh][kw]
			} else {
				float *o = &out[gm * Ncol + gn];
				*o = accumulate ? *o + acc[i][j] : acc[i][j];
			}
		}
}

void simt_conv_fprop(const ConvGeom &g, const float *x, const float *wf, float *y, cudaStream_t st) {
	const long long M = (long long)g.N * g.So() * g.So();
	dim3 grid(ceil_div(g.cout, 64), ceil_div(M, 64));
	simt_conv_kernel<MODE_FPROP><<<grid, 256, 0, st>>>(g, x, wf, y, 0, 0);
	RB_LAUNCH_CHECK();
}

void simt_conv_dgrad(const ConvGeom &g, const float *dy, const float *wd, float *dx, int accumulate, cudaStream_t st) {
	const long long M = (long long)g.N * g.S * g.S;
	dim3 grid(ceil_div(g.cin, 64), ceil_div(M, 64));
	simt_conv_kernel<MODE_DGRAD><<<grid, 256, 0, st>>>(g, dy, wd, dx, accumulate, 0);
	RB_LAUNCH_CHECK();
}

void simt_conv_wgrad(const ConvGeom &g, const float *x, const float *dy, float *dw, cudaStream_t st) {
	const long long K = (long long)g.N * g.So() * g.So();
	const int tiles = ceil_div(g.cout, 64) * ceil_div((long long)g.taps() * g.cin, 64);
	int splits = (kNumSMs * 4 + tiles - 1) / tiles;
	if (splits > 1024) splits = 1024;
	long long per = (K + splits - 1) / splits;
	per = (per + 15) / 16 * 16;
	splits = (int)((K + per - 1) / per);
	RB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * g.w_elems(), st));
	dim3 grid(ceil_div((long long)g.taps() * g.cin, 64), ceil_div(g.cout, 64), splits);
	simt_conv_kernel<MODE_WGRAD><<<grid, 256, 0, st>>>(g, dy, x, dw, 0, per);
	RB_LAUNCH_CHECK();
}

}  // namespace rb
