// mma_rate2.cu -- companion of mma_rate.cu for CTA PAIRS: clocks per tcgen05.mma.cta_group::2 (M = 256 over two SMs, SS mode) issued by one
// elected lane of the leader CTA, operands zero-filled in the shared memory of both CTAs (A: each CTA's own 128 rows; B: each CTA holds
// N / 2 of the N rows), accumulators in the tensor memory of both SMs.  Question: the ~45-clock gap between consecutive MMAs of one
// issuing thread (profiles/r02_mma_rate.txt) -- is it paid per instruction (then a pair instruction halves it per SM, and the 64- /
// 128-column layers would gain up to 2x / 1.4x from 2-CTA tiles) or per SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I resnet_b200/csrc tools/mma_rate2.cu -o tools/mma_rate2.bin
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "ptx.cuh"

using namespace rb::ptx;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {
	asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
	asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_ss(int bf16, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	if (bf16)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc),
		             "l"(bdesc), "r"(idesc), "r"(accumulate)
		             : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc),
		             "l"(bdesc), "r"(idesc), "r"(accumulate)
		             : "memory");
}
// arrives (once all previously issued MMAs of this thread completed) on the barrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit2(uint64_t *bar, uint16_t mask) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct Cfg { int N, bf16, n_mma; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(Cfg c, unsigned long long *clk) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	__shared__ uint64_t bar;
	__shared__ uint32_t tmem_slot;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t rank = cluster_ctarank();
	for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(base)[i] = make_uint4(0, 0, 0, 0);
	if (warp == 1) {
		if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
		__syncwarp();
		tmem_alloc2(&tmem_slot, 512);
		tmem_relinquish2();
	}
	fence_proxy_async();
	tc_fence_before();
	cluster_sync_all();
	tc_fence_after();
	const uint32_t tmem_base = tmem_slot;
	long long t0 = 0;
	if (warp == 1) {
		if (rank == 0) {  // leader CTA: warp-uniform loop, elect.sync around the instructions
			const uint32_t idesc = c.bf16 ? make_idesc_bf16(256, c.N, 0, 0) : make_idesc_tf32(256, c.N, 0, 0);
			const uint64_t adesc = make_smem_desc(smem_u32(base), 16, 1024);
			const uint64_t bdesc = make_smem_desc(smem_u32(base + 96 * 1024), 16, 1024);
			t0 = clock64();
			for (int i = 0; i < c.n_mma; i += 4) {
				if (elect_one()) {
#pragma unroll
					for (int k = 0; k < 4; k++) mma2_ss(c.bf16, tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
				}
				__syncwarp();
			}
			if (elect_one()) mma_commit2(&bar, (uint16_t)3);
			__syncwarp();
		}
		mbar_wait(&bar, 0);  // both CTAs: the pair's MMAs have completed
		if (rank == 0 && lane == 0) clk[blockIdx.x / 2] = (unsigned long long)(clock64() - t0);
	}
	tc_fence_before();
	cluster_sync_all();
	if (warp == 1) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

int main() {
	const int SM = 148, smem = 161 * 1024 + 1024;
	cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	unsigned long long *d;
	cudaMalloc(&d, SM * sizeof(unsigned long long));
	printf("%-5s %4s %4s %12s %14s %10s\n", "kind", "M", "N", "clk/mma(pair)", "clk/mma per SM", "pipe floor");
	for (int bf = 1; bf >= 0; bf--)
		for (int N : {256, 128, 64, 32}) {
			Cfg c{N, bf, 8192};
			rate2_kernel<<<SM, 128, smem>>>(c, d);  // warm-up
			cudaMemset(d, 0, SM * sizeof(unsigned long long));
			rate2_kernel<<<SM, 128, smem>>>(c, d);
			cudaError_t e = cudaDeviceSynchronize();
			if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
			std::vector<unsigned long long> h(SM / 2);
			cudaMemcpy(h.data(), d, (SM / 2) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
			unsigned long long mx = 0;
			for (auto v : h) mx = v > mx ? v : mx;
			const double per = (double)mx / c.n_mma;
			// one pair instruction = M 256 x N x K(32 B): each SM's pipe works 128 x N / 256 clocks on it
			printf("%-5s %4d %4d %12.1f %14.1f %10.1f\n", bf ? "bf16" : "tf32", 256, N, per, per / 2.0, 128.0 * N / 256.0);
			fflush(stdout);
		}
	return 0;
}
