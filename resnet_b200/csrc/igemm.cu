// igemm.cu -- tcgen05 / TMA implicit-GEMM convolution for sm_100a: fprop, dgrad (K-major operands) and wgrad
// (MN-major operands); fp32 tensors with kind::tf32 MMAs or bf16 tensors with kind::f16 MMAs, fp32 accumulation in TMEM.
//
// Replaces the reference's doConvolution / convolutionDerivInput / convolutionDerivWeights
// (reference: resnet.cu:109-156, 166-219, 227-281 and their launchers 1386-1429).
//
// Data layout: activations NHWC (fp32 or bf16); packed weights Wf[Cout][tap][Cin] (fprop) and
// Wd[Cin][tap][Cout] (dgrad), both K-major for the GEMM that reads them.  A "K chunk" is one 128-byte swizzle row:
// 32 tf32 or 64 bf16 elements; the shared-memory tiles are byte-identical in the two modes.
//
// fprop / dgrad kernel (igemm_kmajor_kernel<BF16, MINB>): D[128 pixels x BN] += A[128 x chunk] * B[BN x chunk]^T per pipeline stage.
//   * A tile = a (bw x bh x bn) box of output pixels; for filter tap (kh, kw) the SAME box shifted by the tap
//     offset is fetched by ONE 4-D TMA tiled load {chunk, bw, bh, bn}; out-of-bounds coordinates are zero-filled
//     by the TMA unit, which is exactly the convolution's zero padding (im2col never exists in memory).
//     Stride-2 layers read through four "parity" tensor maps (even/odd rows x even/odd cols of the input), so
//     every tap is again a dense shifted box.  Stride-2 dgrad is decomposed into the four output parities
//     (1 + 2 + 2 + 4 taps): no multiply by structural zeros.
//   * B tile = BN weight rows x chunk via a 2-D TMA load.  Both tiles land in 128-byte-swizzled K-major layout,
//     the canonical operand layout of tcgen05.mma (UMMA) descriptors.
//   * warp 0 (+ 6): TMA producer; warp 1: TMEM allocator + tcgen05.mma issuer (one elected lane, 4 MMAs per stage, K = 8 tf32 / 16 bf16);
//     warps 2-5 (and 6-9 on short-K layers): epilogue -- tcgen05.ld TMEM -> registers -> swizzled smem staging tile -> TMA
//     tile store / reduce-add, plus the fused BatchNorm statistics -- overlapped with the next tile's mainloop through a
//     double-buffered TMEM accumulator (2 x BN <= 512 columns).  Persistent: one CTA per SM; the narrow-N, long-K layers (BN <= 64:
//     the 64-channel 3x3 convolutions and the stem) run TWO independent CTAs per SM -- the <., 2> instantiation, half the shared
//     memory and 2 x BN TMEM columns each -- because one issuing thread cannot keep the tensor pipe busy with N = 64 MMAs.
//
// wgrad kernel (igemm_mnmajor_kernel<BF16>): dW[tap][128 co x BN ci] += dY[px x 128 co]^T * X_tap[px x BN ci], px = 32 (tf32) /
//   64 (bf16) pixels per stage.  The reduction (GEMM K) runs over pixels, so both operands are MN-major straight out of NHWC
//   memory: a TMA box of px pixels x 128 bytes of channels is one swizzled MN-major atom column.  Split-K over pixel ranges
//   into a workspace, reduced deterministically (and re-laid to [Cout][Cin][kh][kw]) by wgrad_reduce.
#include "common.cuh"
#include "ptx.cuh"
#include "igemm.h"
#include <algorithm>
#include <mutex>
#include <vector>

namespace rb {
using namespace ptx;

// ------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
	static EncodeTiledFn fn = nullptr;
	if (!fn) {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
		if (e != cudaSuccess || !p) { set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e)); return nullptr; }
		fn = (EncodeTiledFn)p;
	}
	return fn;
}

// element type of a tensor map: fp32 (tf32 MMAs) or bf16
// L2 promotion of the tensor maps: how much L2 fetches from HBM around every 128-byte request of a TMA box.  256 bytes: the next K
// chunk of the same pixel row comes along (0.5 % on the whole conv set, 3-7 % on the 1x1 dgrads with 4 KB rows;
// profiles/r02_conv_bench_l2_promotion_f32.txt).  RESNET_B200_L2_PROMO = 0 | 64 | 128 | 256 overrides.
static CUtensorMapL2promotion map_promo() {
	if (const char *e = getenv("RESNET_B200_L2_PROMO")) {
		const int v = atoi(e);
		if (v == 128) return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
		if (v == 64) return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
		if (v == 0) return CU_TENSOR_MAP_L2_PROMOTION_NONE;
	}
	return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}
static inline CUtensorMapDataType map_dtype(int bf16) { return bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
static inline size_t esize(int bf16) { return bf16 ? 2 : 4; }
static inline const void *eptr(const void *base, long long elems, int bf16) { return (const char *)base + elems * (long long)esize(bf16); }

// rank-4 map, dims[0] innermost (channels), 128B swizzle, zero OOB fill. strides in ELEMENTS for dims 1..3.
static bool make_map4(CUtensorMap *m, const void *base, const long long dims[4], const long long strides_elems[3], const int box[4],
                      int bf16, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
	EncodeTiledFn enc = get_encode();
	if (!enc) return false;
	cuuint64_t gd[4], gs[3];
	cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
	for (int i = 0; i < 4; i++) { gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
	for (int i = 0; i < 3; i++) gs[i] = (cuuint64_t)strides_elems[i] * esize(bf16);
	CUresult r = enc(m, map_dtype(bf16), 4, (void *)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
	                 map_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("cuTensorMapEncodeTiled(4d) failed: %d dims=(%lld,%lld,%lld,%lld) box=(%d,%d,%d,%d)", (int)r, dims[0], dims[1], dims[2],
		          dims[3], box[0], box[1], box[2], box[3]);
		return false;
	}
	return true;
}
// rank-5 view of the same NHWC tensors for the wgrad operands: channels are split into {128 bytes inner (32 fp32 / 64 bf16), outer}
// and the outer block index is made the SLOWEST box dimension, so ONE TMA op lands `cblk` consecutive [pixels][128 B] boxes (the
// MN-major operand layout) instead of one op per channel block -- the single producer thread was issue-bound at 12-36 ops per stage.
static bool make_map5_cblk(CUtensorMap *m, const void *base, const long long dims4[4], const long long strides_elems[3], const int box4[4],
                           int cblk_box, int bf16, CUtensorMapSwizzle swz) {
	EncodeTiledFn enc = get_encode();
	if (!enc) return false;
	const long long C = dims4[0];
	const cuuint64_t es_ = esize(bf16), cb = 128 / es_;
	cuuint64_t gd[5] = {cb, (cuuint64_t)dims4[1], (cuuint64_t)dims4[2], (cuuint64_t)dims4[3], (cuuint64_t)((C + cb - 1) / cb)};
	cuuint64_t gs[4] = {(cuuint64_t)strides_elems[0] * es_, (cuuint64_t)strides_elems[1] * es_, (cuuint64_t)strides_elems[2] * es_, 128};
	cuuint32_t bx[5] = {(cuuint32_t)cb, (cuuint32_t)box4[1], (cuuint32_t)box4[2], (cuuint32_t)box4[3], (cuuint32_t)cblk_box}, es[5] = {1, 1, 1, 1, 1};
	if (C < (long long)cb) gd[0] = (cuuint64_t)C;
	CUresult r = enc(m, map_dtype(bf16), 5, (void *)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
	                 map_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("cuTensorMapEncodeTiled(5d) failed: %d dims=(%lld,%lld,%lld,%lld) box=(%d,%d,%d,%d,%d)", (int)r, dims4[0], dims4[1], dims4[2],
		          dims4[3], box4[0], box4[1], box4[2], box4[3], cblk_box);
		return false;
	}
	return true;
}
static bool make_map2(CUtensorMap *m, const void *base, long long cols, long long rows, long long row_stride_elems, int box_cols, int box_rows,
                      int bf16) {
	EncodeTiledFn enc = get_encode();
	if (!enc) return false;
	cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gs[1] = {(cuuint64_t)row_stride_elems * esize(bf16)};
	cuuint32_t bx[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
	CUresult r = enc(m, map_dtype(bf16), 2, (void *)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                 CU_TENSOR_MAP_SWIZZLE_128B, map_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d) failed: %d cols=%lld rows=%lld box=(%d,%d)", (int)r, cols, rows, box_cols, box_rows); return false; }
	return true;
}

// ------------------------------------------------------------------------------------------ kernel parameters
struct TapDesc { int dx, dy, amap, bcol; };
struct GroupDesc { int ntaps, oh_off, ow_off, omap; TapDesc taps[9]; };  // omap: output tensor map of the group (stride-2 dgrad parity)

struct alignas(64) IgemmParams {
	CUtensorMap amap[4];
	CUtensorMap bmap;
	CUtensorMap omap[4];  // output tile store / reduce-add maps, box {32, bw, bh, bn}
	GroupDesc groups[4];
	int ngroups;
	int bw, bh, bn, tiles_w, tiles_h, tiles_b, m_tiles;
	int Wm, Hm, Nn;
	int kchunks, kelems;  // K chunks per tap and elements per chunk (one 128-byte swizzle row: 32 tf32 / 64 bf16)
	int BN, n_tiles, Ncol;
	int stages;
	int nstaging;  // epilogue staging tiles PER GROUP (2..4): up to nstaging - 2 TMA stores stay in flight behind the chunk being staged
	int nprod;       // TMA-producer warps (1, 2 or 4; divides `stages`): warp 0 plus warps 6.. of the idle second epilogue group
	int epi_groups;  // 1 or 2 groups of four epilogue warps; with 2, the 128-byte column chunks of a CTA alternate between them
	uint32_t a_bytes, b_bytes, a_tx_bytes;  // smem slot sizes; bytes one A box actually transfers (bw*bh*bn rows)
	// resident_b: the CTA's whole weight operand (all taps x K chunks of its one N tile, resb_bytes) is loaded ONCE into shared memory
	// ahead of the pipeline stages, which then carry activation tiles only.  Used when it fits and every tile of a CTA has the same
	// N tile (gridDim.x % n_tiles == 0): the 64-channel 3x3 layers and the stem re-fetched their few KB of weights for every tile
	// and sat at the ~10-12 TB/s the TMA / L2 path delivers chip-wide.
	int resident_b;
	uint32_t resb_bytes;
	int pdl_early;  // 1: let the next kernel of the stream be scheduled as soon as this one runs, 0: when its CTAs are done (see tc_run)
	// TMEM columns this CTA allocates: 512 for the one-CTA-per-SM plans; 2 * BN (a power of two >= 32) when two CTAs share an SM
	// (finish_kmajor `two`): the narrow-N layers are bound by what ONE issuing thread can push into the tensor pipe (~95 clocks per
	// N = 64 MMA for 32 clocks of pipe time, profiles/r02_mma_rate.txt: two threads issuing side by side reach 62), and a second,
	// fully independent CTA on the SM is a second issuing thread with its own ring, accumulators and epilogue
	uint32_t tmem_cols;
	int debug;  // profiling aid (RESNET_B200_DEBUG_SKIP): bit 0 = issue no MMAs (feed only), bit 1 = epilogue drains TMEM but stores nothing
	float *out;
	int OH, OW, os, accumulate;
	int tma_store;  // epilogue: 1 = swizzled smem staging + TMA tile store (reduce-add when accumulate), 0 = per-thread row stores
	float *stats;   // optional BatchNorm partial sums [gridDim.x * 4][2][Ncol] (sum y, sum y^2), one row per epilogue warp
};

struct alignas(64) WgradParams {
	CUtensorMap amap;
	CUtensorMap bmap[4];
	TapDesc taps[9];
	int ntaps;
	int bw, bh, bn, tiles_w, tiles_h, tiles_b, k_boxes;
	int splits, boxes_per_split;
	int tpt;  // filter taps per tile (they share the dY tile of each stage); tpt * BN <= 256 TMEM columns
	int co_tiles, ci_tiles, BN, cin, cout;
	int a_blocks, cb;  // 128-byte channel blocks per 128-channel A tile (4 tf32 / 2 bf16), channels per block (32 / 64)
	// m_pair = 2: one work item covers TWO 128-row co tiles that share the X tile of every stage (A = 256 co x px, two M = 128 MMAs
	// per K step against the same B descriptor, accumulators in all 512 TMEM columns, single-buffered: a work item runs for hundreds
	// of stages, so its one epilogue need not overlap).  Cuts the bytes the TMA path must deliver per MAC by a third on the layers
	// that are bound by it (every wgrad with Cout >= 256 sat at the ~42 B/clk/SM the TMA / L2 path delivers chip-wide).
	int m_pair, co_items;  // co_items = co_tiles / m_pair
	int nprod;         // TMA-producer warps: 1, or 2 when the ring depth is even (a slot is always refilled by the same warp)
	int split_major;   // work-item order, see wgrad_tile()
	int merge_taps;    // the taps of a group sit back to back in shared memory AND in TMEM: one MMA of N = ntaps * BN covers them all
	uint32_t kadv;     // descriptor start-address advance per MMA (16-byte units): 8 tf32 / 16 bf16 pixel rows
	int stages;
	int pdl_early;
	uint32_t a_bytes, b_bytes, lbo, sbo, layout_type;
	float *partial;
};

// Block sizes: TMA warp, MMA warp, 4 epilogue warps (wgrad) / up to two groups of 4 epilogue warps (fprop, dgrad).
// (Round 1 carried an experimental mode with 2-4 MMA-issuing warps per CTA.  With the issue loops warp-uniform it measured 1-25 % SLOWER
// on every layer, profiles/r02_conv_bench_issuers_*.txt -- the kernels are bound by their operand feed and epilogue, not by MMA issue --
// and it gave up bitwise reproducibility, so it was removed.)
constexpr int kIgemmThreads = 224;  // wgrad: warp 0 TMA, 1 MMA, 2-5 epilogue, 6 second TMA producer
constexpr int kKmajorThreads = 320;
constexpr int kTmemCols = 512;
constexpr uint32_t kABytes = 128 * 32 * 4;  // 128 rows x 32 tf32 = 16 KB

// 1024-byte aligned start of the dynamic shared memory (the 128-byte swizzle wants 1024).  Plain pointer arithmetic on the __shared__
// array -- not an integer round trip -- so that the compiler keeps the address space of everything derived from it: the epilogue's
// staging stores and the statistics loads are then STS / LDS with 32-bit addresses instead of generic ST / LD with 64-bit address math.
__device__ __forceinline__ uint8_t *align1024(uint8_t *p) {
	return p + ((1024u - (ptx::smem_u32(p) & 1023u)) & 1023u);
}

// ------------------------------------------------------------------------------------------ fprop / dgrad
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
	uint32_t r;
	asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
	return r;
}

// BF16 = false: fp32 tensors, kind::tf32 MMAs, 32 output columns per staged 128-byte row;
// BF16 = true:  bf16 tensors, kind::f16 MMAs, 64 output columns per staged row.  The shared-memory tiles are byte-identical
// in both modes (128 rows x 128 bytes, 4 MMAs per stage each advancing 32 bytes along K).
// MINB = 2: the same kernel compiled for two resident CTAs per SM (<= 102 registers per thread)
template <bool BF16, int MINB>
__global__ void __launch_bounds__(kKmajorThreads, MINB) igemm_kmajor_kernel(const __grid_constant__ IgemmParams p) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];  // (no static shared memory in this kernel: the dynamic window starts 1024-aligned; align1024 stays as a guard)
	uint8_t *base = align1024(smem_raw);
	if (MINB == 2 && base != smem_raw) __trap();  // the two-CTA plans have no spare kilobyte for a misaligned window (see finish_kmajor)
	const uint32_t stage_bytes = p.resident_b ? p.a_bytes : p.a_bytes + p.b_bytes;
	uint8_t *resb = base;                                           // resident weight slots [tap * kchunks + kc] (resb_bytes, 0 when unused)
	uint8_t *stage0 = base + p.resb_bytes;
	uint8_t *staging = stage0 + (size_t)p.stages * stage_bytes;  // epi_groups x nstaging x 16 KB epilogue tiles (128 rows x 128 B, 128B-swizzled)
	uint64_t *full = reinterpret_cast<uint64_t *>(staging + p.epi_groups * p.nstaging * kABytes);
	uint64_t *empty = full + p.stages;
	uint64_t *tfull = empty + p.stages;
	uint64_t *tempty = tfull + 2;
	uint64_t *bfull = tempty + 2;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bfull + 1);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (warp == 0 && lane == 0) {
		for (int i = 0; i < 4; i++) prefetch_tmap(&p.amap[i]);
		prefetch_tmap(&p.bmap);
		for (int i = 0; i < 4; i++) prefetch_tmap(&p.omap[i]);
	}
	if (warp == 1) {
		if (lane == 0) {
			for (int i = 0; i < p.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
			for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * p.epi_groups); }
			mbar_init(bfull, 1);
			fence_barrier_init();
		}
		__syncwarp();
		tmem_alloc(tmem_slot, p.tmem_cols);
		tmem_relinquish();
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;
	// everything above touched shared memory, TMEM and the kernel parameters only; global memory from here on (common.cuh launch_k)
	pdl_wait();
	if (p.pdl_early) pdl_trigger();

	const int total_tiles = p.ngroups * p.m_tiles * p.n_tiles;

	// The producer and the MMA warps run their loops with all 32 lanes (warp-uniform control flow and addresses) and elect one lane
	// only around the TMA / tcgen05 instructions themselves.  Round 1 had these loops under `if (lane == 0)`: nvcc then treats the
	// whole region as divergent and wraps EVERY UTCHMMA / UTMALDG / UTCBAR in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop,
	// which is what made a tcgen05.mma cost ~115 clocks to issue (profiles/r02_mma_rate.txt).
	// Several producer warps: one warp needs ~480 clocks per pipeline stage (mbarrier try_wait on the empty slot, elect, expect_tx,
	// two TMA issues, tap / tile index arithmetic: profiles/r02_ncu_conv_layers.md), more than the 380 / 445 clocks a stage of four
	// N = 64 / 128 MMAs takes -- the 64- and 128-channel layers were bound by their TMA-issuing warp.  Producer `pi` of `nprod` owns the
	// ring slots s = pi (mod nprod) (nprod divides the ring depth, so a slot is always refilled by the same warp); the extra producers
	// are the warps of the second epilogue group (6..8), idle in the long-K plans that need them.
	const int pi = (warp == 0) ? 0 : ((warp >= 6 && warp - 5 < p.nprod && p.epi_groups == 1) ? warp - 5 : -1);
	if (pi >= 0) {
		int stage = 0, turn = 0;  // turn = (stage counter) mod nprod
		uint32_t phase = 0;
		if (pi == 0 && p.resident_b && (int)blockIdx.x < total_tiles) {  // the weights of this CTA's N tile, once
			const GroupDesc &g = p.groups[0];
			const int nt = (int)blockIdx.x % p.n_tiles;
			if (elect_one()) {
				mbar_expect_tx(bfull, p.resb_bytes);
				for (int t = 0; t < g.ntaps; t++)
					for (int kc = 0; kc < p.kchunks; kc++)
						tma_load_2d(resb + (size_t)(t * p.kchunks + kc) * p.b_bytes, &p.bmap, bfull, g.taps[t].bcol + kc * p.kelems, nt * p.BN);
			}
			__syncwarp();
		}
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			const int nt = tile % p.n_tiles;
			int r = tile / p.n_tiles;
			const int mt = r % p.m_tiles;
			const GroupDesc &g = p.groups[r / p.m_tiles];
			const int ow0 = (mt % p.tiles_w) * p.bw, oh0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh, n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
			for (int t = 0; t < g.ntaps; t++) {
				const TapDesc tp = g.taps[t];
				for (int kc = 0; kc < p.kchunks; kc++) {
					if (turn == pi) {
						mbar_wait(&empty[stage], phase ^ 1);
						if (elect_one()) {
							uint8_t *sa = stage0 + (size_t)stage * stage_bytes;
							mbar_expect_tx(&full[stage], p.resident_b ? p.a_tx_bytes : p.a_tx_bytes + p.b_bytes);
							tma_load_4d(sa, &p.amap[tp.amap], &full[stage], kc * p.kelems, ow0 + tp.dx, oh0 + tp.dy, n0);
							if (!p.resident_b) tma_load_2d(sa + p.a_bytes, &p.bmap, &full[stage], tp.bcol + kc * p.kelems, nt * p.BN);
						}
						__syncwarp();
					}
					if (++turn == p.nprod) turn = 0;
					if (++stage == p.stages) { stage = 0; phase ^= 1; }
				}
			}
		}
	} else if (warp == 1) {
		const uint32_t idesc = BF16 ? make_idesc_bf16(128, p.BN, 0, 0) : make_idesc_tf32(128, p.BN, 0, 0);
		int stage = 0, acc = 0;
		uint32_t phase = 0, accphase = 0;
		if (p.resident_b && (int)blockIdx.x < total_tiles) mbar_wait(bfull, 0);
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			const GroupDesc &g = p.groups[(tile / p.n_tiles) / p.m_tiles];
			const int iters = g.ntaps * p.kchunks;
			mbar_wait(&tempty[acc], accphase ^ 1);  // the epilogue has drained this accumulator
			tc_fence_after();
			const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
			for (int it = 0; it < iters; it++) {
				mbar_wait(&full[stage], phase);
				tc_fence_after();
				if (elect_one()) {
					const uint32_t a_addr = smem_u32(stage0 + (size_t)stage * stage_bytes);
					const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
					const uint64_t bdesc = make_smem_desc(p.resident_b ? smem_u32(resb + (size_t)it * p.b_bytes) : a_addr + p.a_bytes, 16, 1024);
					if (!(p.debug & 1)) {
#pragma unroll
						for (int k = 0; k < 4; k++) mma_ss<BF16>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((it | k) != 0));
					}
					mma_commit(&empty[stage]);
					if (it == iters - 1) mma_commit(&tfull[acc]);
				}
				__syncwarp();
				if (++stage == p.stages) { stage = 0; phase ^= 1; }
			}
			acc ^= 1;
			if (acc == 0) accphase ^= 1;
		}
	} else if ((warp - 2) / 4 < p.epi_groups) {
		// Epilogue.  A warp may only touch TMEM lanes 32 * (warp % 4) .. +31, so warps 2-5 and 6-9 each cover the four lane quarters.
		// With two groups the 128-byte column chunks (counted over the CTA's whole tile sequence) alternate between them: the
		// short-K 1x1 layers spend their time here (TMEM load -> convert -> swizzled st.shared -> barrier -> TMA store -> statistics
		// is a serial ~0.8 us chain per chunk for four warps), and a second group doubles the chunks in flight.
		const int eg = (warp - 2) / 4;
		const int q = warp & 3;
		const int row = q * 32 + lane;
		const int wq = row % p.bw, hq = (row / p.bw) % p.bh, nq = row / (p.bw * p.bh);
		const bool store_warp = ((warp - 2) % 4 == 0);  // its elected lane issues the group's TMA stores (elect.sync names the same lane every time: bulk groups are per thread)
		uint8_t *const gstaging = staging + (size_t)eg * p.nstaging * kABytes;
		int acc = 0;
		uint32_t accphase = 0, sbuf = 0, chunk_no = 0;
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			const int nt = tile % p.n_tiles;
			int r = tile / p.n_tiles;
			const int mt = r % p.m_tiles;
			const GroupDesc &g = p.groups[r / p.m_tiles];
			const int ow0 = (mt % p.tiles_w) * p.bw, oh0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh, n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
			const bool row_valid = (nq < p.bn) && (ow0 + wq < p.Wm) && (oh0 + hq < p.Hm) && (n0 + nq < p.Nn);
			mbar_wait(&tfull[acc], accphase);
			tc_fence_after();
			const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
			if (p.debug & 2) {
				// feed-only profiling: nothing to drain
			} else if (p.tma_store) {
				// TMEM -> registers -> swizzled smem tile -> one TMA tile store (or reduce-add for the residual join) per 128-byte
				// column chunk: fully coalesced lines, rows outside the tensor are clipped by the TMA unit
				constexpr int CW = BF16 ? 64 : 32;  // output columns per staged 128-byte row
				for (int c = 0; c < p.BN / CW; c++) {
					if (p.epi_groups == 2 && ((chunk_no++) & 1u) != (uint32_t)eg) continue;  // the other group's chunk
					float v[CW];
					if constexpr (BF16) tmem_ld_32x64(taddr + (uint32_t)(c * CW), v);
					else tmem_ld_32x32(taddr + (uint32_t)(c * CW), v);
					if (p.nstaging == 1) {
						// a single staging tile (the two-CTA plans, which trade it for a ring slot): the previous chunk's store must have
						// read the tile, and everybody must know it, before anyone overwrites it -- one more barrier per chunk, which the
						// long main loops of these layers hide
						if (store_warp) tma_wait_group_read<0>();
						named_barrier_sync(1 + eg, 128);
					}
					uint8_t *buf = gstaging + sbuf * kABytes;
					sbuf = (sbuf + 1 == (uint32_t)p.nstaging ? 0 : sbuf + 1);
					uint8_t *rowp = buf + row * 128;
					if (p.stats && !row_valid) {  // rows the TMA store clips must not pollute the fused statistics
#pragma unroll
						for (int j = 0; j < CW; j++) v[j] = 0.f;
					}
#pragma unroll
					for (int j = 0; j < 8; j++) {
						if constexpr (BF16)
							*reinterpret_cast<uint4 *>(rowp + ((j ^ (row & 7)) << 4)) =
							    make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
							               pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
						else
							*reinterpret_cast<float4 *>(rowp + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
					}
					fence_proxy_async();
					// the store that last read the NEXT buffer in the ring is done before anyone rewrites it; nstaging - 2 younger
					// stores may still be in flight (the write-bound 1x1 layers were serialised on the store round trip with 2 tiles)
					if (store_warp) { if (p.nstaging == 4) tma_wait_group_read<2>(); else if (p.nstaging == 3) tma_wait_group_read<1>(); else if (p.nstaging == 2) tma_wait_group_read<0>(); }  // (lanes without bulk groups fall through)
					named_barrier_sync(1 + eg, 128);
					if (store_warp) {
						if (elect_one()) {
							if (p.accumulate) tma_reduce_add_4d(&p.omap[g.omap], buf, nt * p.BN + c * CW, ow0, oh0, n0);
							else tma_store_4d(&p.omap[g.omap], buf, nt * p.BN + c * CW, ow0, oh0, n0);
							tma_commit_group();
						}
						__syncwarp();
					}
					if (p.stats) {
						// fused BatchNorm statistics: lane j sums 32-bit word j (one fp32 column / two bf16 columns) of this warp's 32 rows
						// out of the staged tile (one conflict-free LDS per row) and adds them to the warp's private row of partial sums
						// in global memory with fire-and-forget reductions (always the same thread for a given address, in program
						// order: deterministic; no load latency on the epilogue's critical path)
						float cs = 0.f, cq = 0.f, cs1 = 0.f, cq1 = 0.f;
#pragma unroll 8
						for (int rr = 0; rr < 32; rr++) {
							const int r2 = q * 32 + rr;
							const uint32_t w = *reinterpret_cast<const uint32_t *>(buf + r2 * 128 + ((((lane >> 2) ^ (r2 & 7)) << 4) | ((lane & 3) << 2)));
							if constexpr (BF16) {
								const float y0 = __uint_as_float(w << 16), y1 = __uint_as_float(w & 0xffff0000u);
								cs += y0; cq = fmaf(y0, y0, cq);
								cs1 += y1; cq1 = fmaf(y1, y1, cq1);
							} else {
								const float y = __uint_as_float(w);
								cs += y;
								cq = fmaf(y, y, cq);
							}
						}
						float *sp = p.stats + ((size_t)((blockIdx.x * p.epi_groups + eg) * 4 + q) * 2) * p.Ncol + (size_t)nt * p.BN + c * CW + (BF16 ? 2 * lane : lane);
						atomicAdd(sp, cs);
						atomicAdd(sp + p.Ncol, cq);
						if constexpr (BF16) {
							atomicAdd(sp + 1, cs1);
							atomicAdd(sp + p.Ncol + 1, cq1);
						}
					}
				}
			} else if constexpr (!BF16) {
				const int ow = ow0 + wq, oh = oh0 + hq, n = n0 + nq;
				const bool valid = (nq < p.bn) && (ow < p.Wm) && (oh < p.Hm) && (n < p.Nn);
				float *dst = p.out + (((size_t)n * p.OH + (size_t)(oh * p.os + g.oh_off)) * p.OW + (size_t)(ow * p.os + g.ow_off)) * p.Ncol + (size_t)nt * p.BN;
				for (int c = 0; c < p.BN / 32; c++) {
					float v[32];
					tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
					if (valid) {
						float4 *d4 = reinterpret_cast<float4 *>(dst + c * 32);
						if (p.accumulate) {
#pragma unroll
							for (int j = 0; j < 8; j++) {
								float4 o = d4[j];
								d4[j] = make_float4(o.x + v[4 * j], o.y + v[4 * j + 1], o.z + v[4 * j + 2], o.w + v[4 * j + 3]);
							}
						} else {
#pragma unroll
							for (int j = 0; j < 8; j++) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
						}
					}
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&tempty[acc]);
			acc ^= 1;
			if (acc == 0) accphase ^= 1;
		}
		if (store_warp) tma_wait_group0();  // shared memory must outlive the last store's reads; global writes complete before exit
	}
	pdl_trigger();  // this thread's role loop is finished: the stream's next kernel may be scheduled once every CTA got here
	tc_fence_before();
	__syncthreads();
	if (warp == 1) {
		tc_fence_after();
		tmem_dealloc(tmem_base, p.tmem_cols);
	}
}

// ------------------------------------------------------------------------------------------ wgrad
// work item -> (pixel-range split, output tile r = (tap group * co_items + cot) * ci_tiles + cit).  split_major: consecutive items -- the
// ones the persistent grid runs at the same time -- are the DIFFERENT output tiles of the SAME pixel range, so a range of dY / X is
// fetched from HBM once and its re-reads by the other (co, ci, tap-group) tiles hit L2.  Round 1 ran split-fastest: one output tile's
// splits side by side, every other tile re-reading the whole tensors from HBM later (1.5-1.9x the algorithmic DRAM bytes).
__device__ __forceinline__ void wgrad_tile(const WgradParams &p, int tile, int n_groups, int &split, int &r) {
	if (p.split_major) {
		const int T = n_groups * p.co_items * p.ci_tiles;
		split = tile / T;
		r = tile - split * T;
		// split_major = 2: pixel ranges from the END of the tensors first -- dY was written front to back by the kernel just before this
		// one, so its tail is what the 126 MB L2 still holds (the partial sums land in the same workspace planes: results unchanged)
		if (p.split_major == 2) split = p.splits - 1 - split;
	} else {
		split = tile % p.splits;
		r = tile / p.splits;
	}
}
template <bool BF16>
__global__ void __launch_bounds__(kIgemmThreads, 1) igemm_mnmajor_kernel(const __grid_constant__ WgradParams p) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t *base = align1024(smem_raw);
	const uint32_t stage_bytes = p.a_bytes + (uint32_t)p.tpt * p.b_bytes;
	uint64_t *full = reinterpret_cast<uint64_t *>(base + (size_t)p.stages * stage_bytes);
	uint64_t *empty = full + p.stages;
	uint64_t *tfull = empty + p.stages;
	uint64_t *tempty = tfull + 2;
	uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (warp == 0 && lane == 0) {
		prefetch_tmap(&p.amap);
		for (int i = 0; i < 4; i++) prefetch_tmap(&p.bmap[i]);
	}
	if (warp == 1) {
		if (lane == 0) {
			for (int i = 0; i < p.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
			for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
			fence_barrier_init();
		}
		__syncwarp();
		tmem_alloc(tmem_slot, kTmemCols);
		tmem_relinquish();
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;
	// everything above touched shared memory, TMEM and the kernel parameters only; global memory from here on (common.cuh launch_k)
	pdl_wait();
	if (p.pdl_early) pdl_trigger();

	// tile = ((tap_group * co_tiles + cot) * ci_tiles + cit) * splits + split; a tap group shares one dY (A) tile per
	// stage between up to `tpt` filter taps, each with its own BN-column accumulator (tpt * BN <= 256 TMEM columns)
	const int n_groups = (p.ntaps + p.tpt - 1) / p.tpt;
	const int total_tiles = n_groups * p.co_items * p.ci_tiles * p.splits;
	const int nb_boxes = p.BN / p.cb;  // [px][128 B of channels] boxes per B tile
	const uint32_t grp_cols = (uint32_t)(p.tpt * p.BN);          // accumulator columns of one 128-row co tile
	const uint32_t acc_cols = grp_cols * (uint32_t)p.m_pair;     // ... of one work item
	const int nbuf = (2 * acc_cols <= (uint32_t)kTmemCols) ? 2 : 1;  // TMEM accumulator buffers

	// producer warps: warp 0 and, with nprod = 2, warp 6; producer pi owns the ring slots s = pi (mod nprod) (see igemm_kmajor_kernel)
	const int pi = (warp == 0) ? 0 : ((warp == 6 && p.nprod == 2) ? 1 : -1);
	if (pi >= 0) {  // all lanes run the loops, one elected lane issues
		int stage = 0;
		uint32_t phase = 0;
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			int split, r;
			wgrad_tile(p, tile, n_groups, split, r);
			const int cit = r % p.ci_tiles; r /= p.ci_tiles;
			const int cot = r % p.co_items;
			const int tap0 = (r / p.co_items) * p.tpt;
			const int ntg = min(p.tpt, p.ntaps - tap0);
			const int kb0 = split * p.boxes_per_split;
			const int kb1 = min(kb0 + p.boxes_per_split, p.k_boxes);
			for (int kb = kb0; kb < kb1; kb++) {
				if ((stage & (p.nprod - 1)) == pi) {
					const int ow0 = (kb % p.tiles_w) * p.bw, oh0 = ((kb / p.tiles_w) % p.tiles_h) * p.bh, n0 = (kb / (p.tiles_w * p.tiles_h)) * p.bn;
					mbar_wait(&empty[stage], phase ^ 1);
					if (elect_one()) {
						uint8_t *sa = base + (size_t)stage * stage_bytes;
						mbar_expect_tx(&full[stage], p.a_bytes + (uint32_t)ntg * p.b_bytes);
						tma_load_5d(sa, &p.amap, &full[stage], 0, ow0, oh0, n0, cot * p.a_blocks * p.m_pair);  // all [px][128 B of co] boxes in one op
						for (int t = 0; t < ntg; t++) {
							const TapDesc tp = p.taps[tap0 + t];
							tma_load_5d(sa + p.a_bytes + (size_t)t * p.b_bytes, &p.bmap[tp.amap], &full[stage], 0, ow0 + tp.dx, oh0 + tp.dy, n0, cit * nb_boxes);
						}
					}
					__syncwarp();
				}
				if (++stage == p.stages) { stage = 0; phase ^= 1; }
			}
		}
	} else if (warp == 1) {
		const uint32_t idesc = BF16 ? make_idesc_bf16(128, p.BN, 1, 1) : make_idesc_tf32(128, p.BN, 1, 1);
		int stage = 0, acc = 0;
		uint32_t phase = 0, accphase = 0;
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			int split, r;
			wgrad_tile(p, tile, n_groups, split, r);
			const int tap0 = (r / p.ci_tiles / p.co_items) * p.tpt;
			const int ntg = min(p.tpt, p.ntaps - tap0);
			const int kb0 = split * p.boxes_per_split;
			const int kb1 = min(kb0 + p.boxes_per_split, p.k_boxes);
			mbar_wait(&tempty[acc], accphase ^ 1);  // the epilogue has drained this accumulator
			tc_fence_after();
			const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
			for (int kb = kb0; kb < kb1; kb++) {
				mbar_wait(&full[stage], phase);
				tc_fence_after();
				if (elect_one()) {
					const uint32_t a_addr = smem_u32(base + (size_t)stage * stage_bytes);
					for (int m = 0; m < p.m_pair; m++) {  // the 128-row co tiles of the item: same B tiles, own A rows and accumulators
						const uint64_t adesc = make_smem_desc(a_addr + (uint32_t)m * kABytes, p.lbo, p.sbo, p.layout_type);
						const uint32_t d_m = d_tmem + (uint32_t)m * grp_cols;
						if (p.merge_taps) {
							// the group's B tiles are contiguous [tap][channel block][px][128 B] with one LBO pitch, and its accumulators are
							// contiguous TMEM columns: issue the whole group as ONE N = ntg * BN MMA per K step instead of ntg narrow ones
							const uint32_t idesc_g = BF16 ? make_idesc_bf16(128, ntg * p.BN, 1, 1) : make_idesc_tf32(128, ntg * p.BN, 1, 1);
							const uint64_t bdesc = make_smem_desc(a_addr + p.a_bytes, p.lbo, p.sbo, p.layout_type);
#pragma unroll
							for (int k = 0; k < 4; k++)
								mma_ss<BF16>(d_m, adesc + (uint64_t)(k * p.kadv), bdesc + (uint64_t)(k * p.kadv), idesc_g, (uint32_t)((kb > kb0) || (k != 0)));
						} else {
							for (int t = 0; t < ntg; t++) {
								const uint64_t bdesc = make_smem_desc(a_addr + p.a_bytes + (uint32_t)t * p.b_bytes, p.lbo, p.sbo, p.layout_type);
#pragma unroll
								for (int k = 0; k < 4; k++)  // 8 (tf32, K=8) or 16 (bf16, K=16) pixel rows of 128 B per MMA
									mma_ss<BF16>(d_m + (uint32_t)(t * p.BN), adesc + (uint64_t)(k * p.kadv), bdesc + (uint64_t)(k * p.kadv), idesc,
									             (uint32_t)((kb > kb0) || (k != 0)));
							}
						}
					}
					mma_commit(&empty[stage]);
					if (kb == kb1 - 1) mma_commit(&tfull[acc]);
				}
				__syncwarp();
				if (++stage == p.stages) { stage = 0; phase ^= 1; }
			}
			if (++acc == nbuf) { acc = 0; accphase ^= 1; }
		}
	} else if (warp >= 2 && warp < 6) {
		const int q = warp & 3;
		const int row = q * 32 + lane;
		int acc = 0;
		uint32_t accphase = 0;
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			int split, r;
			wgrad_tile(p, tile, n_groups, split, r);
			const int cit = r % p.ci_tiles; r /= p.ci_tiles;
			const int cot = r % p.co_items;
			const int tap0 = (r / p.co_items) * p.tpt;
			const int ntg = min(p.tpt, p.ntaps - tap0);
			mbar_wait(&tfull[acc], accphase);
			tc_fence_after();
			for (int m = 0; m < p.m_pair; m++) {
				const int co = (cot * p.m_pair + m) * 128 + row;
				const bool valid = co < p.cout;
				for (int t = 0; t < ntg; t++) {
					float *dst = p.partial + (((size_t)split * p.ntaps + tap0 + t) * p.cout + co) * p.cin + (size_t)cit * p.BN;
					const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)m * grp_cols + (uint32_t)(t * p.BN);
					for (int c = 0; c < p.BN / 32; c++) {
						float v[32];
						tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
						if (valid) {
							float4 *d4 = reinterpret_cast<float4 *>(dst + c * 32);
#pragma unroll
							for (int j = 0; j < 8; j++) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
						}
					}
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&tempty[acc]);
			if (++acc == nbuf) { acc = 0; accphase ^= 1; }
		}
	}
	pdl_trigger();  // this thread's role loop is finished: the stream's next kernel may be scheduled once every CTA got here
	tc_fence_before();
	__syncthreads();
	if (warp == 1) {
		tc_fence_after();
		tmem_dealloc(tmem_base, kTmemCols);
	}
}

// ------------------------------------------------------------------------------------------ host side
struct TcPlan {
	int kind;  // 0 = kmajor (fprop / dgrad), 1 = wgrad, 2 = stem wgrad
	int bf16;  // element type of the activation / packed-weight tensors: 0 = fp32 (tf32 MMAs), 1 = bf16
	IgemmParams ip;
	WgradParams wp;
	int grid;
	size_t smem;
	int two;  // fprop / dgrad: two CTAs per SM (igemm_kmajor_kernel<., 2>)
	// wgrad epilogue
	float *dw;
	int cout, cin, taps;
	// fused BatchNorm statistics (fprop): rows of partial sums and their byte size
	int stats_rows;
	size_t stats_bytes;
	int stats_prezeroed;
	double flops;  // algorithmic FLOPs of one launch: 2 * N * Ho * Wo * Cout * Cin * k^2
	char what[40];
};

static const size_t kMaxDynSmem = 227 * 1024;
static inline int kelems_of(int bf16) { return bf16 ? 64 : 32; }  // elements per 128-byte swizzle row

// choose a pixel box (bw, bh, bn) with product <= cap (exact == cap if exact) maximising coverage of (W, H, N)
static void choose_box(int W, int H, int N, int cap, bool exact, int *bw, int *bh, int *bn) {
	if (const char *e = getenv("RESNET_B200_BOX")) {  // bring-up / tuning aid: "bw,bh,bn" for the 128-pixel boxes of fprop / dgrad
		int a, b, c;
		if (!exact && sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a * b * c <= cap && a >= 1 && b >= 1 && c >= 1) { *bw = a; *bh = b; *bn = c; return; }
	}
	double best = -1;
	*bw = 1; *bh = 1; *bn = cap;
	for (int w = 1; w <= cap; w++) {
		for (int h = 1; w * h <= cap; h++) {
			int n = cap / (w * h);
			if (exact && w * h * n != cap) continue;
			if (n < 1) continue;
			// boxes may overhang (TMA zero-fills), but avoid boxes larger than the tensor when another choice exists
			double over = (w > W ? 0.5 : 1.0) * (h > H ? 0.5 : 1.0) * (n > N ? 0.75 : 1.0);
			long long tiles = (long long)ceil_div(W, w) * ceil_div(H, h) * ceil_div(N, n);
			double util = (double)W * H * N / ((double)tiles * cap) * over;
			util += 1e-6 * (w * h) + 1e-8 * w;  // tie-break: larger spatial patch (tap halo reuse), then wider rows
			if (util > best) { best = util; *bw = w; *bh = h; *bn = n; }
		}
	}
}

static int pick_bn(int ncol, int bf16) {
	int cap = 256;
	if (const char *e = getenv("RESNET_B200_MAX_BN")) { int v = atoi(e); if (v == 128 || v == 64) cap = v; }  // tuning aid: narrower N tiles
	for (int bn : {256, 128, 64, 32})
		if (bn <= cap && ncol % bn == 0 && bn >= kelems_of(bf16)) return bn;
	return 0;
}

bool tc_supported(const ConvGeom &g, int bf16) {
	const int q = kelems_of(bf16);
	if (g.cin % q || g.cout % q) return false;
	if (!(g.k == 1 || g.k == 3)) return false;
	if (g.stride == 2 && (g.k != 3 || (g.S % 2))) return false;
	if (g.stride != 1 && g.stride != 2) return false;
	return true;
}

// parity decomposition of "stride*o + k - pad" for stride 2, k 3, pad 1:  k=0 -> (odd, o-1), k=1 -> (even, o), k=2 -> (odd, o)
static void s2_tap(int kk, int *parity, int *d) {
	if (kk == 0) { *parity = 1; *d = -1; }
	else if (kk == 1) { *parity = 0; *d = 0; }
	else { *parity = 1; *d = 0; }
}

// maps over an NHWC tensor [N][S][S][C] as seen by a conv of stride `stride`: 1 map (stride 1) or 4 parity maps
static bool make_input_maps(CUtensorMap *maps, const void *x, int N, int S, int C, int stride, const int box[4], bool flat, int bf16,
                            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int cblk = 0) {
	auto mk = [&](CUtensorMap *m, const void *base, const long long dims[4], const long long str[3]) {
		return cblk > 0 ? make_map5_cblk(m, base, dims, str, box, cblk, bf16, swz) : make_map4(m, base, dims, str, box, bf16, swz);
	};
	if (flat) {  // 1x1: pixels are a flat list
		long long P = (long long)N * S * S;
		long long dims[4] = {C, P, 1, 1}, str[3] = {C, P * C, P * C};
		return mk(&maps[0], x, dims, str);
	}
	if (stride == 1) {
		long long dims[4] = {C, S, S, N}, str[3] = {C, (long long)S * C, (long long)S * S * C};
		return mk(&maps[0], x, dims, str);
	}
	for (int ph = 0; ph < 2; ph++)
		for (int pw = 0; pw < 2; pw++) {
			long long dims[4] = {C, S / 2, S / 2, N}, str[3] = {2LL * C, 2LL * S * C, (long long)S * S * C};
			if (!mk(&maps[ph * 2 + pw], eptr(x, ((long long)ph * S + pw) * C, bf16), dims, str)) return false;
		}
	return true;
}

// RESNET_B200_TMA_STORE=0 falls back to per-thread row stores in the epilogue (bring-up aid, fp32 only)
static int tma_store_enabled(int bf16) {
	const char *e = getenv("RESNET_B200_TMA_STORE");
	return (e && !bf16) ? atoi(e) != 0 : 1;
}

static int env_two_nstaging() {
	if (const char *e = getenv("RESNET_B200_TWO_NSTAGING")) { int v = atoi(e); if (v >= 1 && v <= 2) return v; }
	return 1;
}
static void finish_kmajor(TcPlan *pl) {
	IgemmParams &p = pl->ip;
	int max_stages_override = 0;  // RESNET_B200_STAGES: profiling aid
	p.a_bytes = kABytes;
	p.a_tx_bytes = (uint32_t)(p.bw * p.bh * p.bn) * 128;
	p.b_bytes = (uint32_t)p.BN * 128;
	const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
	// short mainloops (1x1 layers with few K chunks) are bound by the epilogue's stores: give them a deeper store ring
	int max_iters = 0;
	for (int gi = 0; gi < p.ngroups; gi++) max_iters = std::max(max_iters, p.groups[gi].ntaps * p.kchunks);
	p.epi_groups = (max_iters <= 8 && p.tma_store) ? 2 : 1;
	p.nstaging = 2;
	if (const char *e = getenv("RESNET_B200_EPI_GROUPS")) { int v = atoi(e); if (v >= 1 && v <= 2 && p.tma_store) p.epi_groups = v; }
	if (const char *e = getenv("RESNET_B200_NSTAGING")) { int v = atoi(e); if (v >= 2 && v <= 4) p.nstaging = v; }
	const int total = p.ngroups * p.m_tiles * p.n_tiles;
	// Two CTAs per SM for the narrow-N, long-K layers (the 64- / 128-channel 3x3 convolutions): see IgemmParams::tmem_cols.  Each CTA
	// gets half of the shared memory (fewer ring slots each, the same number per SM) and 2 * BN TMEM columns.
	// RESNET_B200_TWO_CTA = largest BN that takes this form (0 = never).  Default: 128 for the 3x3 layers, 64 for the 1x1 layers (at
	// N = 128 those are HBM-bound and measured 0-7 % slower with two CTAs; profiles/r02_two_cta.txt).
	const bool one_tap = p.ngroups == 1 && p.groups[0].ntaps == 1;
	int two_max_bn = one_tap ? 64 : 128;
	if (const char *e = getenv("RESNET_B200_TWO_CTA")) two_max_bn = atoi(e);
	int two_min_iters = 7;  // long main loops only: the short-K 1x1 layers are bound by HBM and their epilogue, measured below
	if (const char *e = getenv("RESNET_B200_TWO_CTA_MINK")) two_min_iters = atoi(e);
	const bool two_any_size = getenv("RESNET_B200_TWO_CTA_FORCE") != nullptr;  // unit tests: also on problems with a handful of tiles
	pl->two = p.BN <= two_max_bn && p.BN >= 32 && p.tma_store && max_iters >= two_min_iters && (total >= 4 * kNumSMs || two_any_size);
	if (pl->two) {
		p.epi_groups = 1;  // (the stem: 7 stages per tile would take two epilogue groups; two CTAs bring two groups per SM anyway)
		p.nstaging = env_two_nstaging();  // one staging tile: the 16 KB buy each CTA a ring slot (4 instead of 3 at BN = 64)
	}
	const size_t staging_bytes = (size_t)p.epi_groups * p.nstaging * kABytes;
	const size_t smem_budget = pl->two ? (kMaxDynSmem + 1024) / 2 - 1024 : kMaxDynSmem;  // 228 KB per SM, 1 KB reserved per CTA
	p.tmem_cols = pl->two ? (uint32_t)(2 * p.BN) : (uint32_t)kTmemCols;
	pl->grid = pl->two ? (total < 2 * kNumSMs ? total : 2 * kNumSMs) : (total < kNumSMs ? total : kNumSMs);
	// weights resident in shared memory when they fit next to >= 4 activation stages and every tile of a CTA shares one N tile
	const size_t resb = (size_t)p.groups[0].ntaps * p.kchunks * p.b_bytes;
	int want_res = 1;
	if (const char *e = getenv("RESNET_B200_RESIDENT_B")) want_res = atoi(e);
	// (RESNET_B200_RESIDENT_B=2 also takes it when a CTA has a single tile: how the unit tests reach this path on small problems)
	p.resident_b = want_res && p.ngroups == 1 && pl->grid % p.n_tiles == 0 && resb <= 80 * 1024 && (total >= 2 * pl->grid || want_res == 2) &&
	               smem_budget >= 2048 + staging_bytes + resb + 4 * (size_t)p.a_bytes;
	p.resb_bytes = p.resident_b ? (uint32_t)resb : 0;
	p.debug = 0;
	if (const char *e = getenv("RESNET_B200_DEBUG_SKIP")) p.debug = atoi(e);
	if (const char *e = getenv("RESNET_B200_STAGES")) { int v = atoi(e); if (v >= 1) max_stages_override = v; }
	const uint32_t pipe_stage = p.resident_b ? p.a_bytes : stage_bytes;
	// bytes kept for the barriers (< 256) and, in the one-CTA plans, for rounding the dynamic window's base up to 1024; the two-CTA plans
	// count every byte (4 x 24 KB + 16 KB + 256 B fill the half SM exactly) and rely on the 1024-aligned base the kernel declares and checks
	const size_t reserve = pl->two ? 1024 : 2048, base_pad = pl->two ? 0 : 1024;
	int stages = (int)((smem_budget - reserve - staging_bytes - p.resb_bytes) / pipe_stage);
	p.stages = stages > 8 ? 8 : stages;
	if (max_stages_override > 0 && max_stages_override < p.stages) p.stages = max_stages_override;
	// A ring slot must always be refilled by the SAME producer warp: a warp that skipped a pass of a slot could find the slot's `empty`
	// barrier one phase behind and mbarrier.try_wait.parity would report "free" (phase aliasing).  So nprod divides the ring depth.
	p.nprod = p.epi_groups == 1 ? 2 : 1;
	if (const char *e = getenv("RESNET_B200_PRODUCERS")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) p.nprod = p.epi_groups == 1 ? v : 1; }
	while (p.nprod > 1 && p.stages < 2 * p.nprod) p.nprod /= 2;
	p.stages = p.stages / p.nprod * p.nprod;
	pl->smem = (size_t)p.resb_bytes + (size_t)p.stages * pipe_stage + staging_bytes + base_pad + 256;
	pl->kind = 0;
}

TcPlan *tc_make_fprop(const ConvGeom &g, const void *x, const void *wf, void *y, int bf16) {
	if (!tc_supported(g, bf16)) { set_error("tc_make_fprop: unsupported geometry"); return nullptr; }
	TcPlan *pl = new TcPlan();
	memset(pl, 0, sizeof(*pl));
	pl->bf16 = bf16;
	IgemmParams &p = pl->ip;
	const int So = g.So(), ke = kelems_of(bf16);
	const bool flat = (g.k == 1);
	if (flat) { p.Wm = (int)((long long)g.N * So * So); p.Hm = 1; p.Nn = 1; p.bw = 128; p.bh = 1; p.bn = 1; }
	else { p.Wm = So; p.Hm = So; p.Nn = g.N; choose_box(So, So, g.N, 128, false, &p.bw, &p.bh, &p.bn); }
	p.tiles_w = ceil_div(p.Wm, p.bw); p.tiles_h = ceil_div(p.Hm, p.bh); p.tiles_b = ceil_div(p.Nn, p.bn);
	p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
	p.Ncol = g.cout; p.BN = pick_bn(g.cout, bf16); p.n_tiles = g.cout / p.BN;
	p.kelems = ke;
	p.kchunks = g.cin / ke;
	const int box[4] = {ke, p.bw, p.bh, p.bn};
	bool ok = make_input_maps(p.amap, x, g.N, g.S, g.cin, g.stride, box, flat, bf16);
	for (int i = 1; i < 4; i++) if (g.stride == 1) p.amap[i] = p.amap[0];
	ok = ok && make_map2(&p.bmap, wf, (long long)g.taps() * g.cin, g.cout, (long long)g.taps() * g.cin, ke, p.BN, bf16);
	p.ngroups = 1;
	GroupDesc &gr = p.groups[0];
	gr.oh_off = gr.ow_off = 0;
	gr.ntaps = g.taps();
	for (int kh = 0; kh < g.k; kh++)
		for (int kw = 0; kw < g.k; kw++) {
			TapDesc &t = gr.taps[kh * g.k + kw];
			t.bcol = (kh * g.k + kw) * g.cin;
			if (g.k == 1) { t.dx = t.dy = 0; t.amap = 0; }
			else if (g.stride == 1) { t.dx = kw - 1; t.dy = kh - 1; t.amap = 0; }
			else { int ph, pw; s2_tap(kh, &ph, &t.dy); s2_tap(kw, &pw, &t.dx); t.amap = ph * 2 + pw; }
		}
	p.out = (float *)y;
	if (flat) { p.OH = 1; p.OW = p.Wm; } else { p.OH = So; p.OW = So; }
	p.os = 1; p.accumulate = 0;
	ok = ok && make_input_maps(p.omap, y, g.N, So, g.cout, 1, box, flat, bf16);  // output tile store map, same pixel box as the A tile
	for (int i = 1; i < 4; i++) p.omap[i] = p.omap[0];
	gr.omap = 0;
	p.tma_store = tma_store_enabled(bf16);
	finish_kmajor(pl);
	pl->flops = 2.0 * g.N * So * So * (double)g.cout * g.cin * g.taps();
	snprintf(pl->what, sizeof(pl->what), "fprop %dx%d/%d %d->%d @%d", g.k, g.k, g.stride, g.cin, g.cout, g.S);
	if (!ok) { delete pl; return nullptr; }
	return pl;
}

TcPlan *tc_make_dgrad(const ConvGeom &g, const void *dy, const void *wd, void *dx, int accumulate, int bf16) {
	if (!tc_supported(g, bf16)) { set_error("tc_make_dgrad: unsupported geometry"); return nullptr; }
	TcPlan *pl = new TcPlan();
	memset(pl, 0, sizeof(*pl));
	pl->bf16 = bf16;
	IgemmParams &p = pl->ip;
	const int So = g.So(), ke = kelems_of(bf16);
	const bool flat = (g.k == 1);
	// GEMM-M space: input pixels (stride 1) or one output-parity class of them (stride 2) == the dy grid
	if (flat) { p.Wm = (int)((long long)g.N * So * So); p.Hm = 1; p.Nn = 1; p.bw = 128; p.bh = 1; p.bn = 1; }
	else { p.Wm = So; p.Hm = So; p.Nn = g.N; choose_box(So, So, g.N, 128, false, &p.bw, &p.bh, &p.bn); }
	p.tiles_w = ceil_div(p.Wm, p.bw); p.tiles_h = ceil_div(p.Hm, p.bh); p.tiles_b = ceil_div(p.Nn, p.bn);
	p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
	p.Ncol = g.cin; p.BN = pick_bn(g.cin, bf16); p.n_tiles = g.cin / p.BN;
	p.kelems = ke;
	p.kchunks = g.cout / ke;
	const int box[4] = {ke, p.bw, p.bh, p.bn};
	bool ok = make_input_maps(p.amap, dy, g.N, So, g.cout, 1, box, flat, bf16);
	for (int i = 1; i < 4; i++) p.amap[i] = p.amap[0];
	ok = ok && make_map2(&p.bmap, wd, (long long)g.taps() * g.cout, g.cin, (long long)g.taps() * g.cout, ke, p.BN, bf16);
	p.out = (float *)dx;
	p.accumulate = accumulate;
	if (g.k == 1) {
		p.ngroups = 1;
		p.groups[0].ntaps = 1; p.groups[0].oh_off = p.groups[0].ow_off = 0;
		p.groups[0].taps[0] = TapDesc{0, 0, 0, 0};
		p.OH = 1; p.OW = p.Wm; p.os = 1;
		ok = ok && make_input_maps(p.omap, dx, g.N, g.S, g.cin, 1, box, true, bf16);
	} else if (g.stride == 1) {
		p.ngroups = 1;
		GroupDesc &gr = p.groups[0];
		gr.ntaps = 9; gr.oh_off = gr.ow_off = 0;
		for (int kh = 0; kh < 3; kh++)
			for (int kw = 0; kw < 3; kw++) gr.taps[kh * 3 + kw] = TapDesc{1 - kw, 1 - kh, 0, (kh * 3 + kw) * g.cout};
		p.OH = g.S; p.OW = g.S; p.os = 1;
		ok = ok && make_input_maps(p.omap, dx, g.N, g.S, g.cin, 1, box, false, bf16);
	} else {
		// dx[2h'+ph] gathers dy[h' + d] * W[kh]:  ph = 0 -> (kh 1, d 0);  ph = 1 -> (kh 0, d +1), (kh 2, d 0)
		p.ngroups = 4;
		const int order[4][2] = {{1, 1}, {1, 0}, {0, 1}, {0, 0}};  // heaviest first
		for (int gi = 0; gi < 4; gi++) {
			const int ph = order[gi][0], pw = order[gi][1];
			GroupDesc &gr = p.groups[gi];
			gr.oh_off = ph; gr.ow_off = pw; gr.ntaps = 0;
			gr.omap = ph * 2 + pw;  // the parity view of dx this phase writes
			int khs[2], dhs[2], nh, kws[2], dws[2], nw;
			if (ph == 0) { nh = 1; khs[0] = 1; dhs[0] = 0; } else { nh = 2; khs[0] = 0; dhs[0] = 1; khs[1] = 2; dhs[1] = 0; }
			if (pw == 0) { nw = 1; kws[0] = 1; dws[0] = 0; } else { nw = 2; kws[0] = 0; dws[0] = 1; kws[1] = 2; dws[1] = 0; }
			for (int a = 0; a < nh; a++)
				for (int b = 0; b < nw; b++) gr.taps[gr.ntaps++] = TapDesc{dws[b], dhs[a], 0, (khs[a] * 3 + kws[b]) * g.cout};
		}
		p.OH = g.S; p.OW = g.S; p.os = 2;
		ok = ok && make_input_maps(p.omap, dx, g.N, g.S, g.cin, 2, box, false, bf16);  // four parity views of dx
	}
	if (g.k == 1 || g.stride == 1) for (int i = 1; i < 4; i++) p.omap[i] = p.omap[0];
	p.tma_store = tma_store_enabled(bf16);
	finish_kmajor(pl);
	pl->flops = 2.0 * g.N * So * So * (double)g.cout * g.cin * g.taps();
	snprintf(pl->what, sizeof(pl->what), "dgrad %dx%d/%d %d->%d @%d", g.k, g.k, g.stride, g.cin, g.cout, g.S);
	if (!ok) { delete pl; return nullptr; }
	return pl;
}

// shape decisions of a wgrad launch, shared by the workspace query and the plan builder.  One pipeline stage reduces over a box
// of `px` pixels = 4 MMAs: 32 pixels (tf32, K = 8) or 64 pixels (bf16, K = 16).
struct WgradShape { int bw, bh, bn, tiles_w, tiles_h, tiles_b, k_boxes, BN, ci_tiles, co_tiles, ntaps, tpt, splits, boxes_per_split, m_pair, co_items; };
static WgradShape wgrad_shape(int Wm, int Hm, int Nn, bool flat, int cin_cols, int cout, int ntaps, int bf16) {
	WgradShape w;
	const int px = bf16 ? 64 : 32;
	if (flat) { w.bw = px; w.bh = 1; w.bn = 1; }
	else choose_box(Wm, Hm, Nn, px, true, &w.bw, &w.bh, &w.bn);
	w.tiles_w = ceil_div(Wm, w.bw); w.tiles_h = ceil_div(Hm, w.bh); w.tiles_b = ceil_div(Nn, w.bn);
	w.k_boxes = w.tiles_w * w.tiles_h * w.tiles_b;
	w.BN = pick_bn(cin_cols, bf16);
	w.ci_tiles = cin_cols / w.BN;
	w.co_tiles = ceil_div(cout, 128);
	w.ntaps = ntaps;
	w.tpt = 256 / w.BN;  // double-buffered accumulators: 2 * tpt * BN <= 512 TMEM columns
	if (w.tpt > ntaps) w.tpt = ntaps;
	if (w.tpt < 1) w.tpt = 1;
	// two co tiles per work item (WgradParams::m_pair) when there is an even number of them AND the items stay long: pairing halves
	// the tile count, so the split count doubles and an item gets half the stages, while its epilogue (512 TMEM columns, not
	// overlapped with the next item) doubles.  Measured (profiles/r01_conv_probe_wgrad_mpair.txt): -15..20 % on the three large
	// projections (196 stages per item), +10..40 % on the 14x14 / 7x7 layers (11-25 stages per item).  RESNET_B200_WGRAD_MPAIR=0 / 2
	// forces it off / on.
	int force = -1;
	if (const char *e = getenv("RESNET_B200_WGRAD_MPAIR")) force = atoi(e);
	for (int pair = (w.co_tiles % 2 == 0 && force != 0) ? 2 : 1; pair >= 1; pair--) {
		w.m_pair = pair;
		w.co_items = w.co_tiles / pair;
		const int tiles = ceil_div(ntaps, w.tpt) * w.co_items * w.ci_tiles;
		// two whole waves of the persistent grid: floor, not ceil -- 3 tiles x 99 splits = 297 work items ran as three rounds with the
		// last one 1 % full (wave efficiency 0.67-0.81 on 13 of the 22 layer shapes)
		int splits = (2 * kNumSMs) / tiles;
		if (splits > w.k_boxes) splits = w.k_boxes;
		if (splits < 1) splits = 1;
		w.boxes_per_split = ceil_div(w.k_boxes, splits);
		w.splits = ceil_div(w.k_boxes, w.boxes_per_split);
		if (pair == 1 || force == 2 || w.boxes_per_split >= 64) break;
	}
	return w;
}

// operand layout of the MN-major tiles: tf32 exists only as "128B swizzle, 32B atom" (4-row atoms, descriptor layout type 1);
// bf16 uses the plain 128B swizzle (8-row atoms, layout type 2)
static CUtensorMapSwizzle wgrad_layout(WgradParams &p, int bf16) {
	const int px = bf16 ? 64 : 32;
	p.cb = kelems_of(bf16);
	p.a_blocks = 128 / p.cb;
	p.lbo = (uint32_t)px * 128;      // distance between 128-byte channel blocks (one TMA box of px pixel rows)
	p.sbo = bf16 ? 1024 : 512;        // pitch of the swizzle atoms along the pixel (K) axis
	p.layout_type = bf16 ? 2 : 1;
	p.kadv = bf16 ? 128 : 64;         // 16 / 8 pixel rows of 128 B per MMA, in 16-byte units
	p.a_bytes = (uint32_t)p.m_pair * 128 * 128;  // m_pair x (128 co x px pixels): 4 x [32 px][32 fp32] or 2 x [64 px][64 bf16] each
	p.b_bytes = (uint32_t)p.BN * 128;
	p.merge_taps = 1;
	if (const char *e = getenv("RESNET_B200_WGRAD_MERGE")) p.merge_taps = atoi(e) != 0;
	p.split_major = 2;  // from the end: 5 same-box A/B pairs, 40.21 40.21 40.55 40.50 40.63 -> 40.24 39.92 40.39 40.49 40.33 ms (c2), identical results
	if (const char *e = getenv("RESNET_B200_WGRAD_SPLIT_MAJOR")) p.split_major = atoi(e);  // A/B aid: 0 split-fastest, 1 split-major, 2 split-major from the end
	return bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
}

size_t tc_wgrad_workspace_bytes(const ConvGeom &g, int bf16) {
	const int So = g.So();
	const bool flat = (g.k == 1);
	WgradShape w = flat ? wgrad_shape((int)((long long)g.N * So * So), 1, 1, true, g.cin, g.cout, 1, bf16)
	                    : wgrad_shape(So, So, g.N, false, g.cin, g.cout, g.taps(), bf16);
	return (size_t)w.splits * g.taps() * g.cout * g.cin * sizeof(float);
}

TcPlan *tc_make_wgrad(const ConvGeom &g, const void *x, const void *dy, float *dw, float *workspace, size_t ws_bytes, int bf16) {
	if (!tc_supported(g, bf16)) { set_error("tc_make_wgrad: unsupported geometry"); return nullptr; }
	TcPlan *pl = new TcPlan();
	memset(pl, 0, sizeof(*pl));
	pl->bf16 = bf16;
	WgradParams &p = pl->wp;
	const int So = g.So();
	const bool flat = (g.k == 1);
	WgradShape w = flat ? wgrad_shape((int)((long long)g.N * So * So), 1, 1, true, g.cin, g.cout, 1, bf16)
	                    : wgrad_shape(So, So, g.N, false, g.cin, g.cout, g.taps(), bf16);
	p.bw = w.bw; p.bh = w.bh; p.bn = w.bn; p.tiles_w = w.tiles_w; p.tiles_h = w.tiles_h; p.tiles_b = w.tiles_b; p.k_boxes = w.k_boxes;
	p.cin = g.cin; p.cout = g.cout; p.BN = w.BN; p.ci_tiles = w.ci_tiles; p.co_tiles = w.co_tiles; p.ntaps = w.ntaps; p.tpt = w.tpt;
	p.splits = w.splits; p.boxes_per_split = w.boxes_per_split; p.m_pair = w.m_pair; p.co_items = w.co_items;
	const int tiles = ceil_div(p.ntaps, p.tpt) * p.co_items * p.ci_tiles;
	if ((size_t)p.splits * p.ntaps * g.cout * g.cin * sizeof(float) > ws_bytes) { set_error("tc_make_wgrad: workspace too small"); delete pl; return nullptr; }
	CUtensorMapSwizzle swz = wgrad_layout(p, bf16);
	const int box[4] = {p.cb, p.bw, p.bh, p.bn};
	if (const char *e = getenv("RESNET_B200_WGRAD_DESC")) {  // bring-up aid: "lbo,sbo,layout_type,tma_swizzle_enum"
		unsigned a, b, c, d;
		if (sscanf(e, "%u,%u,%u,%u", &a, &b, &c, &d) == 4) { p.lbo = a; p.sbo = b; p.layout_type = c; swz = (CUtensorMapSwizzle)d; }
	}
	bool ok = make_input_maps(&p.amap, dy, g.N, So, g.cout, 1, box, flat, bf16, swz, p.a_blocks * p.m_pair);
	ok = ok && make_input_maps(p.bmap, x, g.N, g.S, g.cin, g.stride, box, flat, bf16, swz, p.BN / p.cb);
	if (g.stride == 1) for (int i = 1; i < 4; i++) p.bmap[i] = p.bmap[0];
	for (int kh = 0; kh < g.k; kh++)
		for (int kw = 0; kw < g.k; kw++) {
			TapDesc &t = p.taps[kh * g.k + kw];
			t.bcol = 0;
			if (g.k == 1) { t.dx = t.dy = 0; t.amap = 0; }
			else if (g.stride == 1) { t.dx = kw - 1; t.dy = kh - 1; t.amap = 0; }
			else { int ph, pw; s2_tap(kh, &ph, &t.dy); s2_tap(kw, &pw, &t.dx); t.amap = ph * 2 + pw; }
		}
	const uint32_t stage_bytes = p.a_bytes + (uint32_t)p.tpt * p.b_bytes;
	int stages = (int)((kMaxDynSmem - 2048) / stage_bytes);
	p.stages = stages > 8 ? 8 : stages;
	if (const char *e = getenv("RESNET_B200_WGRAD_STAGES")) { int v = atoi(e); if (v >= 2 && v < p.stages) p.stages = v; }  // tuning aid
	p.nprod = (p.stages % 2 == 0 && p.stages >= 4) ? 2 : 1;
	if (const char *e = getenv("RESNET_B200_PRODUCERS")) { if (atoi(e) == 1) p.nprod = 1; }
	p.partial = workspace;
	pl->smem = (size_t)p.stages * stage_bytes + 1024 + 256;
	const int total = tiles * p.splits;
	pl->grid = total < kNumSMs ? total : kNumSMs;
	pl->kind = 1;
	pl->dw = dw; pl->cout = g.cout; pl->cin = g.cin; pl->taps = g.taps();
	pl->flops = 2.0 * g.N * So * So * (double)g.cout * g.cin * g.taps();
	snprintf(pl->what, sizeof(pl->what), "wgrad %dx%d/%d %d->%d @%d", g.k, g.k, g.stride, g.cin, g.cout, g.S);
	if (!ok) { delete pl; return nullptr; }
	return pl;
}

// ------------------------------------------------------------------------------------------ stem (7x7 / 2, Cin = 3)
// reference: resnet.cu:1547 (fprop) and 2243 (wgrad only).  TMA needs 16-byte pixel strides, so the batch is first copied
// into a zero-bordered NHWC4 buffer xp[N][S][S+PW][4] (3 pixels of left border = the convolution's padding).  One filter ROW
// (kh) of one output pixel is then one contiguous 128-byte run -- 8 taps x 4 channels of fp32 or 16 taps x 4 channels of bf16;
// the taps past the 7th and the 4th channel meet zero weights -- starting 2 pixels after the previous output pixel's: an
// overlapping tensor map {128 B, ow (stride 2 pixels), row, n} turns every filter row into one K chunk of the same implicit GEMM
// (7 chunks), on the same kernels as every other layer.  Input rows 2*oh + kh - 3 are reached through even / odd row maps
// (dy below), out-of-range rows are zero-filled.
constexpr int kStemK = 7, kStemPadL = 3;
static inline int stem_padw(int bf16) { return bf16 ? 16 : 8; }     // extra (zero) pixels per padded row
static inline int stem_row_elems(int bf16) { return bf16 ? 64 : 32; }  // elements of one filter row chunk: taps x 4 channels

template <bool BF16>
__global__ void stem_pad_input_kernel(const float *__restrict__ x, int N, int S, void *__restrict__ xp, int Wp, int rnd) {
	const long long total = (long long)N * S * Wp;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int wp = (int)(i % Wp);
		const long long row = i / Wp;  // n * S + h
		const int w = wp - kStemPadL;
		float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
		if (w >= 0 && w < S) {
			const float *s = x + (row * S + w) * 3;
			v.x = s[0]; v.y = s[1]; v.z = s[2];
			if (rnd && !BF16) {
				uint32_t r;
				asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v.x)); v.x = __uint_as_float(r);
				asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v.y)); v.y = __uint_as_float(r);
				asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v.z)); v.z = __uint_as_float(r);
			}
		}
		if constexpr (BF16) reinterpret_cast<uint2 *>(xp)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, 0.f));
		else reinterpret_cast<float4 *>(xp)[i] = v;
	}
}
// w [Cout][3][7][7] -> wfs [Cout][7][RE/4][4] (zero for kw >= 7 and c = 3); RE = elements of one filter-row chunk
template <bool BF16>
__global__ void stem_pack_weights_kernel(const float *__restrict__ w, int cout, void *__restrict__ wfs, int RE, int rnd) {
	const int total = cout * kStemK * RE;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
		const int c = i % 4, kw = (i / 4) % (RE / 4), kh = (i / RE) % kStemK, co = i / (RE * kStemK);
		float v = 0.f;
		if (c < 3 && kw < kStemK) v = w[((co * 3 + c) * kStemK + kh) * kStemK + kw];
		if constexpr (BF16) reinterpret_cast<uint16_t *>(wfs)[i] = (uint16_t)(pack_bf16x2(v, 0.f) & 0xffffu);
		else {
			if (rnd) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); v = __uint_as_float(r); }
			reinterpret_cast<float *>(wfs)[i] = v;
		}
	}
}
// dw [Cout][3][7][7] = sum_s partial[s][kh][Cout][kw*4 + c].  The stem's one wgrad tile is split ~300 ways, so a block sums 32
// outputs over 8 split lanes (lane y takes splits y, y + 8, ...: independent loads) and combines the lanes in a fixed order
// (deterministic); one thread per output walked the splits as one serial chain of dependent-latency loads (145 us for 9408 outputs).
constexpr int kSwrX = 32, kSwrY = 8;
__global__ void __launch_bounds__(kSwrX * kSwrY) stem_wgrad_reduce_kernel(const float *__restrict__ partial, int splits, int cout, int RE, float *__restrict__ dw) {
	__shared__ float sm[kSwrY][kSwrX];
	const int total = cout * 3 * kStemK * kStemK;
	const long long per = (long long)kStemK * cout * RE;
	const int i = blockIdx.x * kSwrX + threadIdx.x, ty = threadIdx.y;
	float s = 0.f;
	if (i < total) {
		const int kw = i % kStemK, kh = (i / kStemK) % kStemK, c = (i / (kStemK * kStemK)) % 3, co = i / (3 * kStemK * kStemK);
		const long long src = ((long long)kh * cout + co) * RE + kw * 4 + c;
#pragma unroll 4
		for (int sp = ty; sp < splits; sp += kSwrY) s += partial[sp * per + src];
	}
	sm[ty][threadIdx.x] = s;
	__syncthreads();
	if (ty == 0 && i < total) {
#pragma unroll
		for (int y = 1; y < kSwrY; y++) s += sm[y][threadIdx.x];
		dw[i] = s;
	}
}

void stem_pad_input(const float *x, int N, int S, void *xp, int rnd, int bf16, cudaStream_t st) {
	const int Wp = S + stem_padw(bf16);
	long long total = (long long)N * S * Wp;
	int grid = (int)((total + 255) / 256);
	grid = grid > kNumSMs * 16 ? kNumSMs * 16 : grid;
	if (bf16) stem_pad_input_kernel<true><<<grid, 256, 0, st>>>(x, N, S, xp, Wp, rnd);
	else stem_pad_input_kernel<false><<<grid, 256, 0, st>>>(x, N, S, xp, Wp, rnd);
	RB_LAUNCH_CHECK();
}
void stem_pack_weights(const float *w, int cout, void *wfs, int rnd, int bf16, cudaStream_t st) {
	const int RE = stem_row_elems(bf16), grid = ceil_div(cout * kStemK * RE, 256);
	if (bf16) stem_pack_weights_kernel<true><<<grid, 256, 0, st>>>(w, cout, wfs, RE, rnd);
	else stem_pack_weights_kernel<false><<<grid, 256, 0, st>>>(w, cout, wfs, RE, rnd);
	RB_LAUNCH_CHECK();
}
size_t stem_xp_bytes(int N, int S, int bf16) { return (size_t)N * S * (S + stem_padw(bf16)) * 4 * esize(bf16); }
size_t stem_wfs_bytes(int cout, int bf16) { return (size_t)cout * kStemK * stem_row_elems(bf16) * esize(bf16); }

// row parity / offset of input row 2*oh + kh - 3 in the even/odd row maps
static void stem_row(int kh, int *parity, int *dy) {
	static const int par[7] = {1, 0, 1, 0, 1, 0, 1}, off[7] = {-2, -1, -1, 0, 0, 1, 1};
	*parity = par[kh]; *dy = off[kh];
}
static bool make_stem_maps(CUtensorMap *maps, const void *xp, int N, int S, const int box[4], int bf16, CUtensorMapSwizzle swz, int cblk = 0) {
	const int So = S / 2, Wp = S + stem_padw(bf16);
	for (int ph = 0; ph < 2; ph++) {
		long long dims[4] = {stem_row_elems(bf16), So, So, N}, str[3] = {8, 2LL * Wp * 4, (long long)S * Wp * 4};
		const void *base = eptr(xp, (long long)ph * Wp * 4, bf16);
		if (!(cblk > 0 ? make_map5_cblk(&maps[ph], base, dims, str, box, cblk, bf16, swz) : make_map4(&maps[ph], base, dims, str, box, bf16, swz))) return false;
	}
	return true;
}
bool tc_stem_supported(int S, int k, int cin, int cout, int stride, int bf16) {
	return k == kStemK && cin == 3 && stride == 2 && S % 2 == 0 && cout % kelems_of(bf16) == 0 && cout <= 128;
}

TcPlan *tc_make_stem_fprop(int N, int S, int cout, const void *xp, const void *wfs, void *y, int bf16) {
	TcPlan *pl = new TcPlan();
	memset(pl, 0, sizeof(*pl));
	pl->bf16 = bf16;
	IgemmParams &p = pl->ip;
	const int So = S / 2, RE = stem_row_elems(bf16);
	p.Wm = So; p.Hm = So; p.Nn = N;
	choose_box(So, So, N, 128, false, &p.bw, &p.bh, &p.bn);
	p.tiles_w = ceil_div(p.Wm, p.bw); p.tiles_h = ceil_div(p.Hm, p.bh); p.tiles_b = ceil_div(p.Nn, p.bn);
	p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
	p.Ncol = cout; p.BN = pick_bn(cout, bf16); p.n_tiles = cout / p.BN;
	p.kchunks = 1; p.kelems = RE;
	const int box[4] = {RE, p.bw, p.bh, p.bn};
	bool ok = make_stem_maps(p.amap, xp, N, S, box, bf16, CU_TENSOR_MAP_SWIZZLE_128B);
	p.amap[2] = p.amap[0]; p.amap[3] = p.amap[1];
	ok = ok && make_map2(&p.bmap, wfs, kStemK * RE, cout, kStemK * RE, RE, p.BN, bf16);
	p.ngroups = 1;
	GroupDesc &gr = p.groups[0];
	gr.ntaps = kStemK; gr.oh_off = gr.ow_off = 0;
	for (int kh = 0; kh < kStemK; kh++) {
		int par, dy;
		stem_row(kh, &par, &dy);
		gr.taps[kh] = TapDesc{0, dy, par, kh * RE};
	}
	p.out = (float *)y; p.OH = So; p.OW = So; p.os = 1; p.accumulate = 0;
	ok = ok && make_input_maps(p.omap, y, N, So, cout, 1, box, false, bf16);
	for (int i = 1; i < 4; i++) p.omap[i] = p.omap[0];
	gr.omap = 0;
	p.tma_store = tma_store_enabled(bf16);
	finish_kmajor(pl);
	pl->flops = 2.0 * N * So * So * (double)cout * 3 * kStemK * kStemK;
	snprintf(pl->what, sizeof(pl->what), "fprop 7x7/2 3->%d @%d", cout, S);
	if (!ok) { delete pl; return nullptr; }
	return pl;
}

size_t tc_stem_wgrad_workspace_bytes(int N, int S, int cout, int bf16) {
	const int RE = stem_row_elems(bf16);
	WgradShape w = wgrad_shape(S / 2, S / 2, N, false, RE, cout, kStemK, bf16);
	return (size_t)w.splits * kStemK * cout * RE * sizeof(float);
}

TcPlan *tc_make_stem_wgrad(int N, int S, int cout, const void *xp, const void *dy, float *dw, float *workspace, size_t ws_bytes, int bf16) {
	TcPlan *pl = new TcPlan();
	memset(pl, 0, sizeof(*pl));
	pl->bf16 = bf16;
	WgradParams &p = pl->wp;
	const int So = S / 2, RE = stem_row_elems(bf16);
	WgradShape w = wgrad_shape(So, So, N, false, RE, cout, kStemK, bf16);
	p.bw = w.bw; p.bh = w.bh; p.bn = w.bn; p.tiles_w = w.tiles_w; p.tiles_h = w.tiles_h; p.tiles_b = w.tiles_b; p.k_boxes = w.k_boxes;
	p.cin = RE; p.cout = cout; p.BN = w.BN; p.ci_tiles = w.ci_tiles; p.co_tiles = w.co_tiles; p.ntaps = w.ntaps; p.tpt = w.tpt;
	p.splits = w.splits; p.boxes_per_split = w.boxes_per_split; p.m_pair = w.m_pair; p.co_items = w.co_items;
	if ((size_t)p.splits * kStemK * cout * RE * sizeof(float) > ws_bytes) { set_error("tc_make_stem_wgrad: workspace too small"); delete pl; return nullptr; }
	const CUtensorMapSwizzle swz = wgrad_layout(p, bf16);
	const int box[4] = {p.cb, p.bw, p.bh, p.bn};
	bool ok = make_input_maps(&p.amap, dy, N, So, cout, 1, box, false, bf16, swz, p.a_blocks * p.m_pair);
	ok = ok && make_stem_maps(p.bmap, xp, N, S, box, bf16, swz, 1);
	p.bmap[2] = p.bmap[0]; p.bmap[3] = p.bmap[1];
	for (int kh = 0; kh < kStemK; kh++) {
		int par, dyy;
		stem_row(kh, &par, &dyy);
		p.taps[kh] = TapDesc{0, dyy, par, 0};
	}
	const uint32_t stage_bytes = p.a_bytes + (uint32_t)p.tpt * p.b_bytes;
	int stages = (int)((kMaxDynSmem - 2048) / stage_bytes);
	p.stages = stages > 8 ? 8 : stages;
	if (const char *e = getenv("RESNET_B200_WGRAD_STAGES")) { int v = atoi(e); if (v >= 2 && v < p.stages) p.stages = v; }  // tuning aid
	p.nprod = (p.stages % 2 == 0 && p.stages >= 4) ? 2 : 1;
	if (const char *e = getenv("RESNET_B200_PRODUCERS")) { if (atoi(e) == 1) p.nprod = 1; }
	p.partial = workspace;
	pl->smem = (size_t)p.stages * stage_bytes + 1024 + 256;
	const int total = ceil_div(p.ntaps, p.tpt) * p.co_items * p.ci_tiles * p.splits;
	pl->grid = total < kNumSMs ? total : kNumSMs;
	pl->kind = 2;
	pl->dw = dw; pl->cout = cout; pl->cin = 3; pl->taps = kStemK * kStemK;
	pl->flops = 2.0 * N * So * So * (double)cout * 3 * kStemK * kStemK;
	snprintf(pl->what, sizeof(pl->what), "wgrad 7x7/2 3->%d @%d", cout, S);
	if (!ok) { delete pl; return nullptr; }
	return pl;
}

void tc_run(TcPlan *pl, cudaStream_t st) {
	if (!pl) { set_error("tc_run: null plan"); return; }
	if (trace_on()) {  // one line per launch, in launch order (tools/ncu_summary.py joins it with ncu's list)
		char buf[256];
		tc_describe(pl, buf, sizeof(buf));
		fprintf(stderr, "[tc_run] %s flops=%.6g\n", buf, pl->flops);
	}
	// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once on every device that launches (one process
	// may hold trainers on several GPUs)
	static std::mutex attr_mu;
	static bool attr_done[64] = {false};
	int dev = 0;
	RB_CUDA(cudaGetDevice(&dev));
	std::lock_guard<std::mutex> attr_lk(attr_mu);
	bool &attr_set = attr_done[dev & 63];
	if (!attr_set) {
		const void *kernels[] = {(const void *)igemm_kmajor_kernel<false, 1>, (const void *)igemm_kmajor_kernel<true, 1>,
		                         (const void *)igemm_kmajor_kernel<false, 2>, (const void *)igemm_kmajor_kernel<true, 2>,
		                         (const void *)igemm_mnmajor_kernel<false>, (const void *)igemm_mnmajor_kernel<true>};
		for (const void *k : kernels) RB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
		attr_set = true;
	}
	// When may the kernel after this one be scheduled?  Its blocks would sit next to our CTA for our whole run time, blocked in
	// griddepcontrol.wait -- measured: BatchNorm kernels pre-launched under a convolution cost the step 3-5 % (profiles/r02_pdl_ab.txt),
	// so by default a convolution releases its successor only when its own CTAs are done.  RESNET_B200_PDL_CONV_EARLY=1: at its start.
	static const int conv_early = getenv("RESNET_B200_PDL_CONV_EARLY") ? atoi(getenv("RESNET_B200_PDL_CONV_EARLY")) : 0;
	pl->ip.pdl_early = pl->wp.pdl_early = conv_early;
	if (pl->kind == 0) {
		if (pl->ip.stats && !pl->stats_prezeroed) RB_CUDA(cudaMemsetAsync(pl->ip.stats, 0, pl->stats_bytes, st));
		if (pl->two) {
			if (pl->bf16) launch_k(1, igemm_kmajor_kernel<true, 2>, pl->grid, kKmajorThreads, pl->smem, st, pl->ip);
			else launch_k(1, igemm_kmajor_kernel<false, 2>, pl->grid, kKmajorThreads, pl->smem, st, pl->ip);
		} else {
			if (pl->bf16) launch_k(1, igemm_kmajor_kernel<true, 1>, pl->grid, kKmajorThreads, pl->smem, st, pl->ip);
			else launch_k(1, igemm_kmajor_kernel<false, 1>, pl->grid, kKmajorThreads, pl->smem, st, pl->ip);
		}
		RB_LAUNCH_CHECK();
	} else {
		if (pl->bf16) launch_k(1, igemm_mnmajor_kernel<true>, pl->grid, kIgemmThreads, pl->smem, st, pl->wp);
		else launch_k(1, igemm_mnmajor_kernel<false>, pl->grid, kIgemmThreads, pl->smem, st, pl->wp);
		RB_LAUNCH_CHECK();
		if (pl->kind == 1) wgrad_reduce(pl->wp.partial, pl->wp.splits, pl->cout, pl->cin, pl->taps, pl->dw, st);
		else {
			stem_wgrad_reduce_kernel<<<ceil_div(pl->cout * 3 * kStemK * kStemK, kSwrX), dim3(kSwrX, kSwrY), 0, st>>>(pl->wp.partial, pl->wp.splits, pl->cout, pl->wp.cin, pl->dw);
			RB_LAUNCH_CHECK();
		}
	}
}

void tc_free(TcPlan *pl) { delete pl; }

// Attach BatchNorm statistics to an fprop plan: the epilogue then also produces [rows][2][Cout] partial sums of the
// conv output (sum, sum of squares) in `partials` (>= tc_stats_floats(cout) floats); returns the row count for bn_finalize.
int tc_attach_stats(TcPlan *pl, float *partials, int prezeroed) {
	if (!pl || pl->kind != 0 || pl->ip.ngroups != 1 || !pl->ip.tma_store || pl->ip.accumulate) return 0;
	pl->ip.stats = partials;
	pl->stats_prezeroed = prezeroed;
	pl->stats_rows = pl->grid * 4 * pl->ip.epi_groups;
	pl->stats_bytes = (size_t)pl->stats_rows * 2 * pl->ip.Ncol * sizeof(float);
	return pl->stats_rows;
}
size_t tc_stats_floats(int cout) { return (size_t)kNumSMs * 8 * 2 * cout; }

void tc_describe(const TcPlan *pl, char *buf, size_t n) {
	if (!pl) { snprintf(buf, n, "null"); return; }
	if (pl->kind == 0) {
		const IgemmParams &p = pl->ip;
		snprintf(buf, n, "%s | kmajor %s box=(%d,%d,%d) m_tiles=%d n_tiles=%d BN=%d groups=%d kchunks=%d stages=%d epi=%d prod=%d grid=%d smem=%zu resB=%u ctas/SM=%d", pl->what,
		         pl->bf16 ? "bf16" : "tf32", p.bw, p.bh, p.bn, p.m_tiles, p.n_tiles, p.BN, p.ngroups, p.kchunks, p.stages, p.epi_groups, p.nprod, pl->grid, pl->smem, p.resb_bytes,
		         pl->two ? 2 : 1);
	} else {
		const WgradParams &p = pl->wp;
		snprintf(buf, n, "%s | wgrad %s box=(%d,%d,%d) k_boxes=%d splits=%d co_tiles=%d m_pair=%d ci_tiles=%d BN=%d taps=%d tpt=%d stages=%d prod=%d grid=%d smem=%zu",
		         pl->what, pl->bf16 ? "bf16" : "tf32", p.bw, p.bh, p.bn, p.k_boxes, p.splits, p.co_tiles, p.m_pair, p.ci_tiles, p.BN, p.ntaps, p.tpt, p.stages, p.nprod, pl->grid, pl->smem);
	}
}

}  // namespace rb
