// mma_rate.cu -- how many clocks does ONE tcgen05.mma (SS mode, cta_group::1) cost as a function of its shape and of what the
// previous instruction was?  One CTA per SM, operands zero-filled in shared memory (no TMA, no epilogue), one thread issues a long
// train of MMAs and times it with clock64().  Answers the question the conv kernels raised: the layers with 64 / 128 output
// channels run at 22 / 45 % tensor-pipe although their TMA feed was cut 6x (haloed patch) -- is a narrow MMA bound by its operand
// reads, by the dependency on its own accumulator, or by the issue path?
// Round 2: the issuing loops run warp-uniform with elect.sync around the tcgen05 instructions (Cfg::lane0 = 0, the product's form since
// round 2).  Cfg::lane0 = 1 reproduces round 1's `if (lane == 0)` form, which nvcc compiles into an ELECT / R2UR.BROADCAST / BRA.U.ANY
// waterfall loop around EVERY UTCHMMA -- the "115-clock issue floor" of profiles/r01_mma_rate.txt was that loop, not the hardware.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I resnet_b200/csrc tools/mma_rate.cu -o tools/mma_rate.bin
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "ptx.cuh"

using namespace rb::ptx;
__device__ __forceinline__ void mma_ss_rt(int bf16, uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
	if (bf16) mma_f16_ss(d, a, b, idesc, acc);
	else mma_tf32_ss(d, a, b, idesc, acc);
}

struct Cfg {
	int M, N, bf16;      // instruction shape (K = 32 bytes) and kind
	int nacc;            // accumulators used round-robin (columns nacc * N <= 512)
	int a_shift;         // bytes added to the A start address (128 = one swizzle row: the haloed-patch addressing)
	int a_tiles;         // distinct 16 KB A tiles visited round-robin (one per group of 4 MMAs)
	int b_tiles;         // distinct B tiles visited round-robin
	int commit_every;    // tcgen05.commit to a (never waited) mbarrier every this many MMAs (0 = only at the end)
	int n_mma;
	const char *note;
	int issuers = 1;     // warps issuing concurrently (one thread each, own accumulator): is the ~115-clock floor per issuing thread or per SM?
	int same_acc = 0;    // the concurrent issuers accumulate into ONE accumulator; operands are all ones and every element of D is checked
	int lane0 = 0;       // 1 = round 1's divergent `if (lane == 0)` issue loop (waterfalled UTCHMMA); 0 = warp-uniform loop + elect.sync
};

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, unsigned long long *clk) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bars[2];
	__shared__ uint64_t ibars[4];
	__shared__ unsigned long long iclk[4];
	__shared__ uint32_t tmem_slot;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t fillw = c.same_acc ? (c.bf16 ? 0x3F803F80u : 0x3F800000u) : 0u;
	for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(base)[i] = make_uint4(fillw, fillw, fillw, fillw);
	if (warp == 1) {
		if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); for (int i = 0; i < 4; i++) mbar_init(&ibars[i], 1); fence_barrier_init(); }
		__syncwarp();
		tmem_alloc(&tmem_slot, 512);
		tmem_relinquish();
	}
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = tmem_slot;
	if (c.issuers > 1) {
		const uint32_t idesc = c.bf16 ? make_idesc_bf16(c.M, c.N, 0, 0) : make_idesc_tf32(c.M, c.N, 0, 0);
		if (c.same_acc) {  // D = A * B once (overwrite), completed before anyone accumulates
			if (warp == 0 && lane == 0) {
				mma_ss_rt(c.bf16, tmem_base, make_smem_desc(smem_u32(base), 16, 1024), make_smem_desc(smem_u32(base + 96 * 1024), 16, 1024), idesc, 0u);
				mma_commit(&bars[1]);
				mbar_wait(&bars[1], 0);
			}
			tc_fence_before();
			__syncthreads();
			tc_fence_after();
		}
		if (warp < c.issuers && !c.lane0) {  // warp-uniform loops + elect.sync
			const uint64_t adesc = make_smem_desc(smem_u32(base) + (uint32_t)warp * 24576u, 16, 1024);
			const uint64_t bdesc = make_smem_desc(smem_u32(base + 96 * 1024), 16, 1024);
			const uint32_t d = tmem_base + (c.same_acc ? 0u : (uint32_t)(warp * c.N));
			const long long t0 = clock64();
			for (int i = 0; i < c.n_mma; i += 4) {
				if (elect_one()) {
#pragma unroll
					for (int k = 0; k < 4; k++) mma_ss_rt(c.bf16, d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
				}
				__syncwarp();
			}
			if (elect_one()) mma_commit(&ibars[warp]);
			__syncwarp();
			mbar_wait(&ibars[warp], 0);
			if (lane == 0) iclk[warp] = (unsigned long long)(clock64() - t0);
		} else
		if (warp < c.issuers && lane == 0) {
			const uint64_t adesc = make_smem_desc(smem_u32(base) + (uint32_t)warp * 24576u, 16, 1024);
			const uint64_t bdesc = make_smem_desc(smem_u32(base + 96 * 1024), 16, 1024);
			const uint32_t d = tmem_base + (c.same_acc ? 0u : (uint32_t)(warp * c.N));
			const long long t0 = clock64();
			for (int i = 0; i < c.n_mma; i += 4) {
#pragma unroll
				for (int k = 0; k < 4; k++) mma_ss_rt(c.bf16, d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
			}
			mma_commit(&ibars[warp]);
			mbar_wait(&ibars[warp], 0);
			iclk[warp] = (unsigned long long)(clock64() - t0);
		}
		tc_fence_before();
		__syncthreads();
		tc_fence_after();
		if (threadIdx.x == 0) {
			unsigned long long m = 0;
			for (int i = 0; i < c.issuers; i++) m = iclk[i] > m ? iclk[i] : m;
			clk[blockIdx.x] = m / (unsigned long long)c.issuers;  // clocks per MMA of the whole SM = this / n_mma
		}
		if (c.same_acc) {  // every element of D must be (1 + issuers * n_mma) * K ones-products
			const float expect = (float)(1 + c.issuers * c.n_mma) * (c.bf16 ? 16.f : 8.f);
			int bad = 0;
			for (int col = 0; col < c.N; col += 32) {
				float v[32];
				tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, v);
#pragma unroll
				for (int j = 0; j < 32; j++) bad += (v[j] != expect);
			}
			if (bad) atomicAdd(&clk[148], (unsigned long long)bad);
		}
	} else
	if (warp == 1 && !c.lane0) {
		// warp-uniform issue loop: all 32 lanes run it, elect.sync picks the lane that executes the tcgen05 instructions
		const uint32_t idesc = c.bf16 ? make_idesc_bf16(c.M, c.N, 0, 0) : make_idesc_tf32(c.M, c.N, 0, 0);
		const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 96 * 1024);
		const long long t0 = clock64();
		int acc = 0, at = 0, bt = 0;
		for (int i = 0; i < c.n_mma; i += 4) {
			if (elect_one()) {
				const uint64_t adesc = make_smem_desc(a0 + (uint32_t)at * 24576u + (uint32_t)c.a_shift, 16, 1024);
				const uint64_t bdesc = make_smem_desc(b0 + (uint32_t)bt * 32768u, 16, 1024);
				int a2 = acc;
#pragma unroll
				for (int k = 0; k < 4; k++) {
					mma_ss_rt(c.bf16, tmem_base + (uint32_t)(a2 * c.N), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
					if (c.nacc > 1) { if (++a2 == c.nacc) a2 = 0; }
				}
				if (c.commit_every && ((i + 4) % c.commit_every) == 0) mma_commit(&bars[1]);
			}
			__syncwarp();
			if (c.nacc > 1) acc = (acc + 4) % c.nacc;
			if (++at == c.a_tiles) at = 0;
			if (++bt == c.b_tiles) bt = 0;
		}
		if (elect_one()) mma_commit(&bars[0]);
		__syncwarp();
		mbar_wait(&bars[0], 0);
		const long long t1 = clock64();
		if (lane == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
	} else
	if (warp == 1 && lane == 0) {
		const uint32_t idesc = c.bf16 ? make_idesc_bf16(c.M, c.N, 0, 0) : make_idesc_tf32(c.M, c.N, 0, 0);
		const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 96 * 1024);  // A tiles: 24 KB pitch (4 fit); B tiles: 32 KB pitch (3 fit)
		const long long t0 = clock64();
		int acc = 0, at = 0, bt = 0;
		for (int i = 0; i < c.n_mma; i += 4) {
			const uint64_t adesc = make_smem_desc(a0 + (uint32_t)at * 24576u + (uint32_t)c.a_shift, 16, 1024);
			const uint64_t bdesc = make_smem_desc(b0 + (uint32_t)bt * 32768u, 16, 1024);
#pragma unroll
			for (int k = 0; k < 4; k++) {
				mma_ss_rt(c.bf16, tmem_base + (uint32_t)(acc * c.N), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
				if (c.nacc > 1) { if (++acc == c.nacc) acc = 0; }
			}
			if (++at == c.a_tiles) at = 0;
			if (++bt == c.b_tiles) bt = 0;
			if (c.commit_every && ((i + 4) % c.commit_every) == 0) mma_commit(&bars[1]);
		}
		mma_commit(&bars[0]);
		mbar_wait(&bars[0], 0);
		const long long t1 = clock64();
		clk[blockIdx.x] = (unsigned long long)(t1 - t0);
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
	const int SM = 148, smem = 201 * 1024 + 1024;
	cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	unsigned long long *d;
	cudaMalloc(&d, (SM + 1) * sizeof(unsigned long long));  // [148] = mismatching accumulator elements of the same_acc check
	std::vector<Cfg> cfgs;
	const int n = 8192;
	for (int bf = 1; bf >= 0; bf--) {
		for (int N : {256, 128, 64, 32, 16}) cfgs.push_back({128, N, bf, 1, 0, 1, 1, 0, n, "one accumulator, same tiles"});
		for (int N : {128, 64}) for (int na : {2, 4}) cfgs.push_back({128, N, bf, na, 0, 1, 1, 0, n, "accumulators round-robin per MMA"});
		for (int N : {256, 64}) cfgs.push_back({128, N, bf, 1, 0, 4, 3, 0, n, "4 A tiles / 3 B tiles round-robin"});
		for (int N : {256, 64}) for (int sh : {128, 1280, 2176}) cfgs.push_back({128, N, bf, 1, sh, 1, 1, 0, n, "A start moved by whole rows"});
		for (int N : {256, 64}) cfgs.push_back({128, N, bf, 1, 0, 4, 3, 4, n, "commit every 4 MMAs"});
		for (int N : {256, 128, 64}) cfgs.push_back({64, N, bf, 1, 0, 1, 1, 0, n, "M = 64"});
		for (int N : {128, 64}) for (int is : {2, 4}) { Cfg c{128, N, bf, 1, 0, 1, 1, 0, n, "issuing warps in parallel, own accumulators"}; c.issuers = is; cfgs.push_back(c); }
		{ Cfg c{128, 256, bf, 1, 0, 1, 1, 0, n, "issuing warps in parallel, own accumulators"}; c.issuers = 2; cfgs.push_back(c); }
		for (int N : {256, 128, 64}) for (int is : {2, 4}) { Cfg c{128, N, bf, 1, 0, 1, 1, 0, n, "issuing warps in parallel, ONE accumulator, result checked"}; c.issuers = is; c.same_acc = 1; cfgs.push_back(c); }
	}
	{   // round 1's divergent form for comparison (same binary, same box)
		std::vector<Cfg> legacy;
		for (int bf = 1; bf >= 0; bf--)
			for (int N : {256, 128, 64}) { Cfg c{128, N, bf, 1, 0, 1, 1, 0, n, "LEGACY if (lane == 0) loop (waterfalled UTCHMMA)"}; c.lane0 = 1; legacy.push_back(c); }
		cfgs.insert(cfgs.end(), legacy.begin(), legacy.end());
	}
	printf("%-5s %4s %4s %4s %6s %3s %3s %6s %10s %10s  %s\n", "kind", "M", "N", "nacc", "ashift", "At", "Bt", "commit", "clk/mma", "floor", "note");
	for (const Cfg &c : cfgs) {
		rate_kernel<<<SM, 128, smem>>>(c, d);  // warm-up
		cudaMemset(d, 0, (SM + 1) * sizeof(unsigned long long));
		rate_kernel<<<SM, 128, smem>>>(c, d);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
		std::vector<unsigned long long> h(SM + 1);
		cudaMemcpy(h.data(), d, (SM + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
		const unsigned long long nbad = h[SM];
		h.resize(SM);
		unsigned long long mx = 0, mn = ~0ull;
		for (auto v : h) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
		const double floor_clk = (c.M < 128 ? 128.0 : c.M) * c.N / 256.0;
		printf("%-5s %4d %4d %4d %6d %3d %3d %6d %10.1f %10.1f  %s%s (min SM %.1f)\n", c.bf16 ? "bf16" : "tf32", c.M, c.N, c.nacc, c.a_shift, c.a_tiles, c.b_tiles,
		       c.commit_every, (double)mx / c.n_mma, floor_clk, c.note, c.issuers == 2 ? " x2" : (c.issuers == 4 ? " x4" : ""), (double)mn / c.n_mma);
		if (c.same_acc) printf("      ^ accumulator check: %llu wrong elements of %d\n", nbad, SM * 128 * c.N);
		fflush(stdout);
	}
	return 0;
}
