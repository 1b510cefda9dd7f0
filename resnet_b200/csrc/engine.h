// engine.h -- private state behind the public resnet.h structs (side tables keyed by the public pointers, so
// the public struct prefix stays byte-identical to the reference's).
#pragma once
#include "common.cuh"
#include "igemm.h"
#include "../../include/resnet.h"
#include <vector>

namespace rb {

// One contiguous fp32 arena per Params tree; locations[i] are 256-byte aligned views in the reference's order
// (reference: resnet.cu:839-943), so Adam / zeroing / allreduce are single passes over [base, base + total).
struct ParamStore {
	Params *tree;
	float *base;
	long long total;              // floats, including alignment padding (padding stays zero)
	std::vector<long long> offs;  // per location
};
ParamStore *param_store_of(const Params *p);

struct BnRef {
	int C;
	long long rows;
	float *gamma, *beta, *dgamma, *dbeta;
	float *means, *vars;
	float *ab;  // [2][C] folded scale / shift of the current batch
	Cache_BatchNorm *cache;
};

struct ConvRef {
	ConvGeom g;
	int loc;
	float *w, *dw;    // public [Cout][Cin][k][k]
	float *wf, *wd;   // packed
	bool use_tc;
	TcPlan *fprop, *dgrad, *wgrad;
	int stats_rows;  // > 0: the fprop epilogue leaves this many rows of BatchNorm partial sums (fused statistics)
};

struct BlockRef {
	bool has_proj;
	ConvRef reduce, spatial, expand, proj;
	BnRef bn_r, bn_s, bn_e, bn_p;
	const float *x_in;  // block input (previous block's output_activated / init_convblock_input)
	float *Xr, *Yr, *Xs, *Ys, *Xe, *Xp, *OA;
	// gradient buffers for this block (may alias scratch)
	float *dOA, *dBI, *dXe, *dXp, *dYs, *dXs, *dYr, *dXr;
	long long n_in, n_red_in, n_red_out, n_exp_out;
	uint8_t *oa_bits;  // sign bits of OA, one byte per 128-bit vector (NULL: backward reads OA itself)
};

struct Engine {
	Train_ResNet *trainer;
	cudaStream_t stream;
	int conv_mode;    // 0 = tcgen05 (default), 1 = simt fp32
	int round_tf32;   // producers round conv inputs to tf32 (tensor-core mode only)
	int keep_all;     // materialise every reference buffer (debug / parity)
	int bf16, esz;    // activation / packed-weight storage: 0 = fp32 (tf32 MMAs), 1 = bf16; bytes per element
	int N;
	// stem
	ConvRef stem;
	BnRef bn0;
	float *X0, *Y0, *P0;  // init_conv_applied, init_conv_activated, init_convblock_input
	bool fuse_stem_tail;  // Y0 and its gradient are not materialised: BatchNorm + ReLU + max pool in one kernel, forward and backward
	int *max_inds;
	float *dP0, *dY0, *dX0;
	// stem on the tensor cores: zero-bordered NHWC4 copy of the batch, packed [Cout][7][8][4] weights, plans
	bool stem_tc;
	float *stem_xp, *stem_wfs;
	TcPlan *stem_fprop, *stem_wgrad;
	std::vector<BlockRef> blocks;
	// head
	float *pooled, *logits, *pred, *dlogits, *dpooled, *row_loss;
	int *row_wrong;
	float *pred_host;  // pinned
	// on-device loss / accuracy bookkeeping: running sums over the steps since the last reset, so a host loop
	// needs neither pred_cpu nor a synchronisation per step
	int pred_copy;              // 1 (default, the reference's contract): forward_pass returns with pred_cpu valid; 0: no copy, no sync
	double *epoch_acc;          // device [2]: sum of -log p[label], number of wrong predictions
	long long epoch_images;     // images forwarded since the last reset
	// workspaces
	float *bn_partials;
	float *stats_partials;  // partial sums of the fused conv-epilogue statistics; all-zero between uses (bn_finalize clears what it folds)
	int bn_max_blocks;
	float *bn_coef;
	float *fc_ws;  // split-K planes of the fully-connected GEMMs
	float *wgrad_ws;
	size_t wgrad_ws_bytes;
	PackJob *pack_jobs_dev;
	int n_pack_jobs, pack_max_elems;
	float *ones, *zeros, *tmp_ab, *tmp_mv;  // keep-all helpers
	int *bad_dev, *bad_host;
	std::vector<void *> allocs;  // per-tensor fallback allocations (empty when the arena is in use)
	void *arena;                 // model.cu Arena: one contiguous activation arena (virtual range + physical granules)
	// double-buffered host -> device staging of the next batch (resnet_b200_prefetch_batch / commit_batch), created lazily
	cudaStream_t copy_stream;
	float *stage_img;
	int *stage_lab;
	cudaEvent_t ev_staged, ev_consumed;
	// Weight gradients on a side stream (RESNET_B200_ASYNC_WGRAD, default on): a layer's wgrad only feeds update_parameters, so it is
	// forked off right after the BatchNorm backward that produces its dX operand and runs next to the HBM-bound BatchNorm kernels of the
	// following layers (a persistent wgrad CTA takes the shared memory of an SM but few threads / registers, so BatchNorm blocks share
	// the SM with it; two convolution kernels never co-reside).  ev_rd[k]: the last wgrad reading role buffer k (0 dXe, 1 dXs, 2 dXr,
	// 3 dXp) has finished -- awaited by the main stream before the buffer's next writer.  NULL wstream = synchronous.
	cudaStream_t wstream;
	cudaEvent_t ev_fork, ev_join, ev_rd[4];
	bool ev_rd_live[4];
	// data-parallel hook (dp.cu)
	void *dp;
	// in-situ SIMT re-derivation of every tensor-core convolution (selfcheck.cu), NULL unless resnet_b200_selfcheck(1)
	void *selfcheck;
	// step timing
	cudaEvent_t ev0, ev1;
};
Engine *engine_of(const Train_ResNet *t);
extern int g_selfcheck;     // selfcheck.cu: trainers created while set verify every conv launch in place
extern int g_default_bf16;  // storage type of the next init_trainer (resnet_b200_set_dtype)

// dp.cu
void dp_rank_world(const Engine *e, int *rank, int *world);  // (0, 1) without data parallelism
void dp_release(Engine *e);                                   // destroys the communicator, stream and events
// loader.cu
void loader_release(Batch *bb);
// selfcheck.cu
void selfcheck_init(Engine *e, long long max_act_elems, long long max_w_elems);
void selfcheck_release(Engine *e);
void selfcheck_fprop(Engine *e, const ConvGeom &g, const float *w, const void *x, bool x_is_fp32_batch, const void *y);
void selfcheck_dgrad_snapshot(Engine *e, const ConvGeom &g, const void *dx);
void selfcheck_dgrad(Engine *e, const ConvGeom &g, const float *w, const void *dy, const void *dx, int accumulate);
void selfcheck_wgrad(Engine *e, const ConvGeom &g, const void *x, bool x_is_fp32_batch, const void *dy, const float *dw);
// model.cu: drops the side-table entries of a trainer (resnet_b200_destroy_trainer); releases the activation arena
void arena_close(void *arena);
void engine_forget(const Train_ResNet *t);
void dp_block_done(Engine *e, int block);  // block's gradients are enqueued: issue the buckets that became complete
void dp_allreduce_grads(Engine *e);       // end of backward: flush remaining buckets, compute stream waits

}  // namespace rb
