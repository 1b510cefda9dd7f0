/*
 * oracle/ops.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Host-core C restatement of the arithmetic of the reference's hand-written CUDA kernels
 * (als244/ResNet, resnet.cu).  Every function cites the reference kernel it follows.  It is
 * the checker for the parity tests (tests/), for __graft_entry__.smoke() and for bench.py's
 * cpu_baseline / reference legs; nothing in resnet_b200/ may link or call it.
 *
 * Conventions (all as in resnet.cu): fp32, activations NHWC, conv weights
 * [Cout][Cin][kh][kw], zero padding k/2, out_spatial = in_spatial / stride.
 * Each output is accumulated in the same sequential order as the reference kernel's loop
 * nest (cited per function) so that fp32 results agree with the reference to rounding of
 * identical operation sequences; loops are only re-nested across *independent* outputs
 * (vectorised over channels, OpenMP over pixels).
 *
 * Pinning status: pinned against the reference's own compiled kernels (oracle/_ref, built
 * from /root/reference/resnet.cu by oracle/Makefile) on a B200 -- see tests/golden/ and
 * oracle/gen_golden.py.  The reference ships no golden vectors of its own.
 *
 * Build: make -C oracle   (gcc -O3 -march=native -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX4(n, h, w, c, H, W, C) ((((size_t)(n) * (H) + (h)) * (W) + (w)) * (C) + (c))

int oracle_num_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* reference: resnet.cu:109-156 doConvolution.  Per output: sum over (kh, kw, ci) in that order. */
void oracle_conv_fwd(const float *in, const float *w, int S, int k, int cin, int cout, int stride, int N, float *out) {
	const int So = S / stride, half = k / 2;
	/* re-lay weights as [kh][kw][ci][co] so the independent co axis is the vector axis */
	float *wt = (float *)malloc(sizeof(float) * (size_t)k * k * cin * cout);
	for (int co = 0; co < cout; co++)
		for (int ci = 0; ci < cin; ci++)
			for (int r = 0; r < k; r++)
				for (int c = 0; c < k; c++)
					wt[(((size_t)r * k + c) * cin + ci) * cout + co] = w[(((size_t)co * cin + ci) * k + r) * k + c];
#pragma omp parallel for collapse(2) schedule(static)
	for (int n = 0; n < N; n++) {
		for (int oh = 0; oh < So; oh++) {
			float *acc = (float *)malloc(sizeof(float) * cout);
			for (int ow = 0; ow < So; ow++) {
				for (int co = 0; co < cout; co++) acc[co] = 0.0f;
				for (int r = 0; r < k; r++) {
					const int ih = stride * oh + r - half;
					for (int c = 0; c < k; c++) {
						const int iw = stride * ow + c - half;
						if (ih < 0 || ih >= S || iw < 0 || iw >= S) continue; /* adds w*0 in the reference */
						const float *xp = in + IDX4(n, ih, iw, 0, S, S, cin);
						const float *wp = wt + ((size_t)r * k + c) * cin * cout;
						for (int ci = 0; ci < cin; ci++) {
							const float x = xp[ci];
							const float *wr = wp + (size_t)ci * cout;
							for (int co = 0; co < cout; co++) acc[co] += wr[co] * x;
						}
					}
				}
				memcpy(out + IDX4(n, oh, ow, 0, So, So, cout), acc, sizeof(float) * cout);
			}
			free(acc);
		}
	}
	free(wt);
}

/* reference: resnet.cu:166-219 convolutionDerivInput.  Per input element: sum over
 * (co, row_offset, col_offset) in that order; to_add accumulates into the existing buffer. */
void oracle_conv_dgrad(const float *w, const float *dout, int S, int k, int cin, int cout, int stride, int N, int to_add,
                       float *din) {
	const int So = S / stride, half = k / 2;
	float *wt = (float *)malloc(sizeof(float) * (size_t)k * k * cin * cout); /* [co][kh][kw][ci] */
	for (int co = 0; co < cout; co++)
		for (int ci = 0; ci < cin; ci++)
			for (int r = 0; r < k; r++)
				for (int c = 0; c < k; c++)
					wt[(((size_t)co * k + r) * k + c) * cin + ci] = w[(((size_t)co * cin + ci) * k + r) * k + c];
#pragma omp parallel for collapse(2) schedule(static)
	for (int n = 0; n < N; n++) {
		for (int h = 0; h < S; h++) {
			float *acc = (float *)malloc(sizeof(float) * cin);
			for (int x = 0; x < S; x++) {
				for (int ci = 0; ci < cin; ci++) acc[ci] = 0.0f;
				const int oh0 = h / stride, ow0 = x / stride;
				for (int co = 0; co < cout; co++) {
					for (int ro = -half; ro <= half; ro++) {
						const int oh = oh0 + ro;
						const int kr = h - oh * stride + half;
						if (kr < 0 || kr >= k || oh < 0 || oh >= So) continue;
						for (int cf = -half; cf <= half; cf++) {
							const int ow = ow0 + cf;
							const int kc = x - ow * stride + half;
							if (kc < 0 || kc >= k || ow < 0 || ow >= So) continue;
							const float d = dout[IDX4(n, oh, ow, co, So, So, cout)];
							const float *wr = wt + (((size_t)co * k + kr) * k + kc) * cin;
							for (int ci = 0; ci < cin; ci++) acc[ci] += wr[ci] * d;
						}
					}
				}
				float *dp = din + IDX4(n, h, x, 0, S, S, cin);
				if (to_add) for (int ci = 0; ci < cin; ci++) dp[ci] += acc[ci];
				else memcpy(dp, acc, sizeof(float) * cin);
			}
			free(acc);
		}
	}
	free(wt);
}

/* reference: resnet.cu:227-281 convolutionDerivWeights.  Per weight: sum over
 * (sample, out_row, out_col) in that order; overwrites weight_deriv. */
void oracle_conv_wgrad(const float *in, const float *dout, int S, int k, int cin, int cout, int stride, int N, float *dw) {
	const int So = S / stride, half = k / 2;
#pragma omp parallel for collapse(2) schedule(dynamic)
	for (int r = 0; r < k; r++) {
		for (int c = 0; c < k; c++) {
			float *acc = (float *)calloc((size_t)cout * cin, sizeof(float)); /* [co][ci] */
			for (int n = 0; n < N; n++) {
				for (int oh = 0; oh < So; oh++) {
					const int ih = stride * oh + r - half;
					if (ih < 0 || ih >= S) continue;
					for (int ow = 0; ow < So; ow++) {
						const int iw = stride * ow + c - half;
						if (iw < 0 || iw >= S) continue;
						const float *xp = in + IDX4(n, ih, iw, 0, S, S, cin);
						const float *dp = dout + IDX4(n, oh, ow, 0, So, So, cout);
						for (int co = 0; co < cout; co++) {
							const float d = dp[co];
							float *ar = acc + (size_t)co * cin;
							for (int ci = 0; ci < cin; ci++) ar[ci] += xp[ci] * d;
						}
					}
				}
			}
			for (int co = 0; co < cout; co++)
				for (int ci = 0; ci < cin; ci++) dw[(((size_t)co * cin + ci) * k + r) * k + c] = acc[(size_t)co * cin + ci];
			free(acc);
		}
	}
}

/* reference: resnet.cu:289-342 doBatchNormAndActivate.  Per channel: mean, biased variance
 * (two-pass), x_hat, gamma*x_hat+beta, optional ReLU.  Rows are (n, h, w) in that order.
 * xhat / normalized may be NULL (the B200 path does not keep them). */
void oracle_bn_fwd(const float *x, const float *gamma, const float *beta, int S, int C, int N, float eps, float *means,
                   float *vars, float *xhat, float *normalized, float *activated, int relu) {
	const size_t rows = (size_t)N * S * S;
	const float cnt = (float)(N * S * S);
#pragma omp parallel for schedule(static)
	for (int c0 = 0; c0 < C; c0 += 16) {
		const int c1 = c0 + 16 < C ? c0 + 16 : C;
		float sum[16] = {0}, vs[16] = {0}, mean[16], var[16];
		for (size_t r = 0; r < rows; r++)
			for (int c = c0; c < c1; c++) sum[c - c0] += x[r * C + c];
		for (int c = c0; c < c1; c++) { mean[c - c0] = sum[c - c0] / cnt; means[c] = mean[c - c0]; }
		for (size_t r = 0; r < rows; r++)
			for (int c = c0; c < c1; c++) { const float d = x[r * C + c] - mean[c - c0]; vs[c - c0] += d * d; }
		for (int c = c0; c < c1; c++) { var[c - c0] = vs[c - c0] / cnt; vars[c] = var[c - c0]; }
		for (size_t r = 0; r < rows; r++)
			for (int c = c0; c < c1; c++) {
				const float xh = (x[r * C + c] - mean[c - c0]) / sqrtf(var[c - c0] + eps);
				const float nv = gamma[c] * xh + beta[c];
				if (xhat) xhat[r * C + c] = xh;
				if (normalized) normalized[r * C + c] = nv;
				activated[r * C + c] = relu ? fmaxf(nv, 0.0f) : nv;
			}
	}
}

/* reference: resnet.cu:350-426 activationAndBatchNormDeriv.  mask_src is the forward
 * "activated" tensor (ReLU mask = activated <= 0 -> zero), used when relu != 0.
 * dx uses the reference's three-term form (resnet.cu:422). */
void oracle_bn_bwd(const float *x, const float *gamma, int S, int C, int N, float eps, const float *means,
                   const float *vars, const float *mask_src, const float *dy, float *dgamma, float *dbeta, float *dx,
                   int relu) {
	const size_t rows = (size_t)N * S * S;
	const float n_samples = (float)(N * S * S);
#pragma omp parallel for schedule(static)
	for (int c = 0; c < C; c++) {
		const float g = gamma[c], mean = means[c], var = vars[c];
		const float rstd = 1.0f / sqrtf(var + eps);
		float dG = 0.0f, dB = 0.0f;
		for (size_t r = 0; r < rows; r++) {
			const size_t i = r * C + c;
			if (relu && mask_src[i] <= 0.0f) continue;
			const float xh = (x[i] - mean) / sqrtf(var + eps);
			dG += dy[i] * xh;
			dB += dy[i];
		}
		dgamma[c] = dG;
		dbeta[c] = dB;
		const float three_halfs = -0.5f * powf(var + eps, -1.5f);
		const float neg_rstd = -1.0f * rstd;
		float dVar = 0.0f, dMean = 0.0f, pvar = 0.0f;
		for (size_t r = 0; r < rows; r++) {
			const size_t i = r * C + c;
			const float dxh = (relu && mask_src[i] <= 0.0f) ? 0.0f : dy[i] * g;
			dVar += dxh * (x[i] - mean) * three_halfs;
			dMean += dxh * neg_rstd;
			pvar += -2.0f * (x[i] - mean);
		}
		dMean += dVar * pvar / n_samples;
		for (size_t r = 0; r < rows; r++) {
			const size_t i = r * C + c;
			const float dxh = (relu && mask_src[i] <= 0.0f) ? 0.0f : dy[i] * g;
			dx[i] = dxh * rstd + dVar * (2.0f * (x[i] - mean)) / n_samples + dMean / n_samples;
		}
	}
}

/* reference: resnet.cu:433-471 doMaxPool.  init -1024, strict '>', row-major window scan,
 * argmax recorded as flat index into the input tensor. */
void oracle_maxpool_fwd(const float *in, int k, int stride, int S, int C, int N, int *max_inds, float *out) {
	const int So = S / stride, half = k / 2;
#pragma omp parallel for collapse(2) schedule(static)
	for (int n = 0; n < N; n++)
		for (int oh = 0; oh < So; oh++)
			for (int ow = 0; ow < So; ow++)
				for (int c = 0; c < C; c++) {
					float mv = -1024.0f;
					int mi = -1024;
					for (int ro = -half; ro <= half; ro++)
						for (int cf = -half; cf <= half; cf++) {
							const int h = stride * oh + ro, w = stride * ow + cf;
							if (h < 0 || h >= S || w < 0 || w >= S) continue;
							const size_t ii = IDX4(n, h, w, c, S, S, C);
							if (in[ii] > mv) { mv = in[ii]; mi = (int)ii; }
						}
					const size_t oi = IDX4(n, oh, ow, c, So, So, C);
					max_inds[oi] = mi;
					out[oi] = mv;
				}
}

/* reference: resnet.cu:476-494 maxPoolDeriv, with the race fixed: the reference overwrites
 * (dx[max_ind] = dy) although 3x3/2 windows overlap, so the result at a pixel chosen by two
 * windows is whichever thread wrote last; its cuDNN variants accumulate, and so do we
 * (SURVEY.md appendix B-7).  din must be zeroed by the caller (reference: resnet.cu:2186). */
void oracle_maxpool_bwd(const int *max_inds, const float *dout, size_t n_out, float *din) {
	for (size_t i = 0; i < n_out; i++) din[max_inds[i]] += dout[i];
}

/* reference: resnet.cu:500-517 doFilterAvgPool */
void oracle_avgpool_fwd(const float *in, int S, int C, int N, float *out) {
	for (int n = 0; n < N; n++)
		for (int c = 0; c < C; c++) {
			float s = 0.0f;
			for (int h = 0; h < S; h++)
				for (int w = 0; w < S; w++) s += in[IDX4(n, h, w, c, S, S, C)];
			out[(size_t)n * C + c] = s / (float)(S * S);
		}
}

/* reference: resnet.cu:522-542 filterAvgPoolDeriv */
void oracle_avgpool_bwd(const float *dpooled, int S, int C, int N, float *din) {
	for (int n = 0; n < N; n++)
		for (int h = 0; h < S; h++)
			for (int w = 0; w < S; w++)
				for (int c = 0; c < C; c++) din[IDX4(n, h, w, c, S, S, C)] = dpooled[(size_t)n * C + c] / (float)(S * S);
}

/* reference: resnet.cu:70-85 matMul (row-major, sequential k).  ta / tb select the transposed
 * operand forms used by resnet.cu:1482-1509 (transpose into a temp, then matMul). */
void oracle_matmul(const float *A, const float *B, int m, int k, int n, int ta, int tb, float *out) {
#pragma omp parallel for schedule(static)
	for (int i = 0; i < m; i++) {
		for (int j = 0; j < n; j++) {
			float v = 0.0f;
			for (int z = 0; z < k; z++) {
				const float a = ta ? A[(size_t)z * m + i] : A[(size_t)i * k + z];
				const float b = tb ? B[(size_t)j * k + z] : B[(size_t)z * n + j];
				v += a * b;
			}
			out[(size_t)i * n + j] = v;
		}
	}
}

/* reference: resnet_cudnn.cu:568-587 softMax (max-subtracted; identical in exact arithmetic
 * to the unstabilised resnet.cu:569-580) */
void oracle_softmax(const float *X, int N, int L, float *out) {
	for (int i = 0; i < N; i++) {
		float mx = X[(size_t)i * L];
		for (int j = 0; j < L; j++) if (X[(size_t)i * L + j] > mx) mx = X[(size_t)i * L + j];
		float sum = 0.0f;
		for (int j = 0; j < L; j++) sum += expf(X[(size_t)i * L + j] - mx);
		for (int j = 0; j < L; j++) out[(size_t)i * L + j] = expf(X[(size_t)i * L + j] - mx) / sum;
	}
}

/* reference: resnet.cu:1800-1804 + 597-602: d = pred; d[i][label[i]] -= 1; no 1/N (1806-1811) */
void oracle_ce_deriv(const float *pred, const int *labels, int N, int L, float *d) {
	memcpy(d, pred, sizeof(float) * (size_t)N * L);
	for (int i = 0; i < N; i++) d[(size_t)i * L + labels[i]] -= 1.0f;
}

/* reference: resnet.cu:3363-3383 host loss / accuracy: loss = sum -logf(p[label]);
 * a sample is wrong when any other class has p >= p[label] (ties are wrong). */
void oracle_loss_acc(const float *pred, const int *labels, int N, int L, float *loss_sum, int *n_wrong) {
	float ls = 0.0f;
	int nw = 0;
	for (int s = 0; s < N; s++) {
		const float pc = pred[(size_t)s * L + labels[s]];
		ls += -1.0f * logf(pc);
		for (int c = 0; c < L; c++)
			if (c != labels[s] && pred[(size_t)s * L + c] >= pc) { nw++; break; }
	}
	*loss_sum = ls;
	*n_wrong = nw;
}

/* reference: resnet.cu:605-662 updateMeans / updateVars / updateParams (Adam with the
 * reference's weight-decay form and eps outside the sqrt).  cur_*_decay are the running
 * products beta^t *after* this step's multiply (resnet.cu:2920-2921). */
void oracle_adam(float *p, const float *g, float *m, float *v, size_t n, float lr, float wd, float b1, float b2,
                 float cur_b1, float cur_b2, float eps) {
	for (size_t i = 0; i < n; i++) {
		if (!(isnan(g[i]) || isinf(g[i]))) {
			const float gd = g[i] + wd * p[i];
			m[i] = b1 * m[i] + (1 - b1) * gd;
			v[i] = b2 * v[i] + (1 - b2) * gd * gd;
		}
		const float ma = m[i] / (1 - cur_b1);
		const float va = v[i] / (1 - cur_b2);
		const float old = p[i];
		const float np_ = old - (lr * (ma / (sqrtf(va) + eps)) + wd * old);
		p[i] = (isnan(np_) || isinf(np_)) ? old : np_;
	}
}

/* residual join, reference: resnet.cu:1717 addVec + 1723 doActivation */
void oracle_add_relu(const float *a, const float *b, size_t n, float *sum, float *act) {
	for (size_t i = 0; i < n; i++) {
		const float s = a[i] + b[i];
		if (sum) sum[i] = s;
		act[i] = fmaxf(0.0f, s);
	}
}

/* reference: resnet.cu:553-564 doActivationDeriv (mask on pre-activation input > 0) */
void oracle_relu_bwd(const float *pre, const float *up, size_t n, float *out) {
	for (size_t i = 0; i < n; i++) out[i] = pre[i] > 0.0f ? up[i] : 0.0f;
}
