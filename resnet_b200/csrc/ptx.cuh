// ptx.cuh -- hand-written sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA, TMEM).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace rb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
	uint32_t pred = 0;
	asm volatile(
	    "{\n\t.reg .pred P;\n\t"
	    "elect.sync _|P, 0xffffffff;\n\t"
	    "selp.b32 %0, 1, 0, P;\n\t}"
	    : "=r"(pred));
	return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n\t.reg .pred P1;\n\t"
	    "WAIT_LOOP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
	    "@P1 bra DONE;\n\t"
	    "bra WAIT_LOOP;\n\t"
	    "DONE:\n\t}" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) {
	asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
	    "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
	    : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
	asm volatile(
	    "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
	        smem_u32(smem)),
	    "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	    : "memory");
}

__device__ __forceinline__ void tma_load_5d(void *smem, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
	asm volatile(
	    "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
	        smem_u32(smem)),
	    "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
	    : "memory");
}

// shared -> global tile store / fp32 reduce-add through the tensor map (out-of-bounds rows are clipped by the TMA unit)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *m, const void *smem, int c0, int c1, int c2, int c3) {
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
	             "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	             : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap *m, const void *smem, int c0, int c1, int c2, int c3) {
	asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
	             "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	             : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_barrier_sync(uint32_t id, uint32_t nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate; issued by ONE thread for the CTA.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// same, bf16 inputs (kind::f16; the a/b formats in the instruction descriptor select bf16)
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
template <bool BF16> __device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	if constexpr (BF16) mma_f16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
	else mma_tf32_ss(d_tmem, adesc, bdesc, idesc, accumulate);
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (taddr.lane + i), columns taddr.col..+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
	uint32_t r[32];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
	      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
	      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
	      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr)
	    : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// two back-to-back 32-column loads (64 consecutive columns) behind ONE wait
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, float (&v)[64]) {
	uint32_t r[64];
#pragma unroll
	for (int h = 0; h < 2; h++) {
		uint32_t *q = r + 32 * h;
		asm volatile(
		    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		    : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
		      "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]),
		      "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]),
		      "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
		    : "r"(taddr + (uint32_t)(32 * h))
		    : "memory");
	}
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 64; i++) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 format): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout type [61,64) (2 = 128-byte swizzle, 1 = 128-byte swizzle with 32-byte atoms).
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
	d |= (uint64_t)1 << 46;
	d |= (uint64_t)(layout_type & 7) << 61;
	return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate: c_format=F32 [4,6), a/b_format=TF32 [7,10)/[10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
// fmt: 2 = TF32 (kind::tf32), 1 = BF16, 0 = F16 (kind::f16).
__host__ __device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, uint32_t fmt) {
	uint32_t d = 0;
	d |= 1u << 4;
	d |= fmt << 7;
	d |= fmt << 10;
	d |= (uint32_t)(a_mn_major & 1) << 15;
	d |= (uint32_t)(b_mn_major & 1) << 16;
	d |= (uint32_t)(N >> 3) << 17;
	d |= (uint32_t)(M >> 4) << 24;
	return d;
}
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) { return make_idesc(M, N, a_mn_major, b_mn_major, 2u); }
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) { return make_idesc(M, N, a_mn_major, b_mn_major, 1u); }

}  // namespace ptx
}  // namespace rb
