// l2_handoff.cu -- does a streaming consumer that walks a tensor in the OPPOSITE direction of its producer find the producer's tail in L2?
// Kernel P writes Y (and reads X) front to back; kernel C then reads Y (a) front to back, (b) back to front.  Same grid-stride shape as
// the BatchNorm kernels of bw_kernels.cu (148 x 8 blocks of 256 threads, 128-bit accesses, 4 loads in flight).  Times kernel C with CUDA events.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
template <int NIN>
__global__ void __launch_bounds__(256, 4) stream_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, float4 *__restrict__ y, long long n, int rev, int write) {
	const long long TS = (long long)gridDim.x * 256, g = (long long)blockIdx.x * 256 + threadIdx.x;
	float4 acc = make_float4(0, 0, 0, 0);
	for (long long i0 = g; i0 < n; i0 += TS * 4) {
		float4 ra[4], rb[4];
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const long long j = i0 + u * TS;
			if (j < n) { const long long i = rev ? n - 1 - j : j; ra[u] = a[i]; if (NIN == 2) rb[u] = b[i]; }
		}
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const long long j = i0 + u * TS;
			if (j < n) {
				const long long i = rev ? n - 1 - j : j;
				float4 v = ra[u];
				if (NIN == 2) { v.x += rb[u].x; v.y += rb[u].y; v.z += rb[u].z; v.w += rb[u].w; }
				if (write) y[i] = v; else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
			}
		}
	}
	if (!write && acc.x + acc.y + acc.z + acc.w == 1234.5f) y[0] = acc;
}
int main() {
	const int grid = 148 * 8;
	const size_t maxb = (size_t)1 << 30;
	float4 *X, *Y, *Z, *W;
	CK(cudaMalloc(&X, maxb)); CK(cudaMalloc(&Y, maxb)); CK(cudaMalloc(&Z, maxb)); CK(cudaMalloc(&W, maxb));
	CK(cudaMemset(X, 0, maxb)); CK(cudaMemset(Y, 0, maxb)); CK(cudaMemset(Z, 0, maxb)); CK(cudaMemset(W, 0, maxb));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	printf("# producer P: Y = X (front to back).  consumer C: Z = Y.  time of C in us and GB/s over its algorithmic bytes (2 x size)\n");
	printf("%8s | %10s %8s | %10s %8s | %s\n", "MB", "C fwd us", "GB/s", "C rev us", "GB/s", "gain");
	for (size_t mb : {25, 50, 100, 200, 400, 800}) {
		const long long n = (long long)mb * 1000000 / 16;
		float t[2];
		for (int rev = 0; rev < 2; rev++) {
			float best = 1e9;
			for (int it = 0; it < 5; it++) {
				stream_kernel<1><<<grid, 256>>>(W, nullptr, X, n, 0, 1);  // unrelated traffic first
				stream_kernel<1><<<grid, 256>>>(X, nullptr, Y, n, 0, 1);  // P
				CK(cudaEventRecord(e0));
				stream_kernel<1><<<grid, 256>>>(Y, nullptr, Z, n, rev, 1);  // C
				CK(cudaEventRecord(e1));
				CK(cudaEventSynchronize(e1));
				float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
				if (ms < best) best = ms;
			}
			t[rev] = best;
		}
		printf("%8zu | %10.1f %8.0f | %10.1f %8.0f | %.2fx\n", mb, t[0] * 1e3, 2.0 * mb * 1e6 / (t[0] * 1e-3) / 1e9, t[1] * 1e3, 2.0 * mb * 1e6 / (t[1] * 1e-3) / 1e9, t[0] / t[1]);
	}
	printf("# reduce -> dx shape: R reads (A, B) front to back without writing; D then reads (A, B) and writes Z.  time of D\n");
	printf("%8s | %10s %8s | %10s %8s | %s\n", "MB each", "D fwd us", "GB/s", "D rev us", "GB/s", "gain");
	for (size_t mb : {25, 50, 100, 200, 400, 800}) {
		const long long n = (long long)mb * 1000000 / 16;
		float t[2];
		for (int rev = 0; rev < 2; rev++) {
			float best = 1e9;
			for (int it = 0; it < 5; it++) {
				stream_kernel<1><<<grid, 256>>>(W, nullptr, Z, n, 0, 1);
				stream_kernel<2><<<grid, 256>>>(X, Y, Z, n, 0, 0);  // R
				CK(cudaEventRecord(e0));
				stream_kernel<2><<<grid, 256>>>(X, Y, Z, n, rev, 1);  // D
				CK(cudaEventRecord(e1));
				CK(cudaEventSynchronize(e1));
				float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
				if (ms < best) best = ms;
			}
			t[rev] = best;
		}
		printf("%8zu | %10.1f %8.0f | %10.1f %8.0f | %.2fx\n", mb, t[0] * 1e3, 3.0 * mb * 1e6 / (t[0] * 1e-3) / 1e9, t[1] * 1e3, 3.0 * mb * 1e6 / (t[1] * 1e-3) / 1e9, t[0] / t[1]);
	}
	return 0;
}
