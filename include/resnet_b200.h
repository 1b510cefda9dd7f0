/*
 * resnet_b200.h -- extra C-ABI entry points of libresnet_b200.so, beyond the reference's own surface (resnet.h).
 *
 * (1) The reference's launch wrappers, one per kernel family, on caller-supplied DEVICE buffers.  These are what
 *     the reference's embedded self-tests call (reference: resnet.cu:3109-3218 testConvolution -> resnet.cu:1386
 *     prepareAndDoConvolution, etc.); the parity tests in tests/ go through them.
 * (2) Small runtime services a host driver needs because the library owns its stream: device malloc/copy helpers
 *     for FFI callers without a CUDA binding, synchronisation, error string, step timers.
 * (3) Data-parallel control (new; the reference is single-GPU).
 *
 * Plain pointers and sizes only; every function returns 0 on success (or the requested value) and records the
 * first failure for resnet_b200_last_error().
 */
#ifndef RESNET_B200_EXTRA_H
#define RESNET_B200_EXTRA_H

#include "resnet.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime services ---------------------------------------------------------------------- */
const char * resnet_b200_last_error(void);              /* "" when no error was recorded */
void resnet_b200_clear_error(void);
int resnet_b200_set_device(int device);                 /* cudaSetDevice */
void * resnet_b200_malloc(size_t bytes);                /* device memory */
void resnet_b200_free(void * dev_ptr);
void * resnet_b200_malloc_host(size_t bytes);           /* pinned host memory */
void resnet_b200_free_host(void * host_ptr);
int resnet_b200_memcpy_h2d(void * dev_dst, const void * host_src, size_t bytes);
int resnet_b200_memcpy_d2h(void * host_dst, const void * dev_src, size_t bytes);
int resnet_b200_memcpy_d2d(void * dev_dst, const void * dev_src, size_t bytes);
int resnet_b200_memset(void * dev_dst, int value, size_t bytes);
int resnet_b200_sync(void);                             /* cudaDeviceSynchronize */
/* curandGenerator_t (CURAND_RNG_PSEUDO_DEFAULT) seeded like the reference's main (resnet.cu:3264-3267) */
void * resnet_b200_rng_create(unsigned long long seed);
void resnet_b200_rng_destroy(void * gen);

/* ---- storage / MMA type ------------------------------------------------------------------- */
/* BASELINE config 2 keeps the reference's fp32 tensors (convolutions on kind::tf32 tensor-core MMAs); configs 3-5 store
 * activations, activation gradients and packed weights as bf16 (kind::f16 MMAs, fp32 accumulation).  Master weights,
 * gradients of parameters, Adam state, BatchNorm statistics, pooled features, logits and pred stay fp32 in both modes, so
 * locations[] / param_derivs / pred_cpu keep the reference's types; the named activation buffers of resnet.h then hold
 * bf16 bits behind their float* names (SURVEY.md 8b "in fast mode buffers may be bf16").
 * set_dtype: storage of trainers created AFTER the call: -1 = follow $RESNET_B200_DTYPE ("bf16" or unset), 0 = fp32, 1 = bf16. */
int resnet_b200_set_dtype(int bf16);
int resnet_b200_trainer_dtype(Train_ResNet * trainer);  /* 0 = fp32/TF32, 1 = bf16, -1 = unknown trainer */
/* element type of the ACTIVATION tensors the single-operator entry points below read and write (0 = fp32, 1 = bf16) */
int resnet_b200_set_op_dtype(int bf16);
/* device-side conversion of n elements: to_bf16 = 1: fp32 -> bf16 (round to nearest even), 0: bf16 -> fp32 */
int resnet_b200_convert(const void * dev_src, void * dev_dst, long long n, int to_bf16);

/* ---- trainer services ---------------------------------------------------------------------- */
/* asynchronous copies on the trainer's stream: host (pinned) -> cur_batch->images / correct_classes */
int resnet_b200_stage_batch(Train_ResNet * trainer, const float * images_host, const int * labels_host);
int resnet_b200_stage_batch_device(Train_ResNet * trainer, const float * images_dev, const int * labels_dev);
/* overlapped variant: prefetch_batch starts the host (pinned) -> device copy of the NEXT batch on a copy stream while the
 * current step computes; commit_batch (before forward_pass) orders the compute stream after it and moves it into cur_batch */
int resnet_b200_prefetch_batch(Train_ResNet * trainer, const float * images_host, const int * labels_host);
int resnet_b200_commit_batch(Train_ResNet * trainer);
int resnet_b200_trainer_sync(Train_ResNet * trainer);   /* waits for the trainer's stream */
/* CUDA-event bracket on the trainer's stream: begin / end-and-return-milliseconds */
int resnet_b200_timer_begin(Train_ResNet * trainer);
float resnet_b200_timer_end_ms(Train_ResNet * trainer);
/* on-device loss / accuracy of the last forward_pass (same definitions as reference resnet.cu:3363-3383) */
int resnet_b200_loss_accuracy(Train_ResNet * trainer, float * loss_sum, int * n_wrong);
/* epoch bookkeeping without a per-step synchronisation (the reference's loop reads N x 1000 floats of pred_cpu every step to
 * compute these on the host, resnet.cu:3363-3412): forward_pass adds the batch's loss sum and wrong count to running device
 * sums; epoch_stats reads (and optionally resets) them.  set_pred_copy(0) makes forward_pass return without copying pred to
 * pred_cpu and without synchronising; fetch_pred brings pred_cpu up to date on demand. */
int resnet_b200_set_pred_copy(Train_ResNet * trainer, int on);
int resnet_b200_fetch_pred(Train_ResNet * trainer);
int resnet_b200_epoch_stats(Train_ResNet * trainer, double * loss_sum, long long * n_wrong, long long * n_images, int reset);
/* number of kernels this library launched since process start (bench.py's gpu_launches) */
long long resnet_b200_launch_count(void);
/* per-kernel-family timing with CUDA events on the launching stream (bench.py's roofline leg): enable (resets the
 * records), run steps, then read family f: 0 = tcgen05 fprop+dgrad (3x3 and stem), 1 = tcgen05 wgrad (+ split-K reduce),
 * 2 = BatchNorm / elementwise, 3 = SIMT conv (stem).  work = algorithmic FLOPs (0, 1, 3) or HBM bytes (2). */
void resnet_b200_profile(int enable);
int resnet_b200_profile_read(int family, double * ms, long long * launches, double * work);
/* family 5 = fprop / dgrad of the 1x1 convolutions (HBM-bound; family 0 then holds the 3x3 / stem launches only): work = FLOPs,
 * work2 = algorithmic HBM bytes (input + output tensor once) */
int resnet_b200_profile_read2(int family, double * ms, long long * launches, double * work, double * work2);
/* 1 when conv layers of this trainer run on the tcgen05 path, 0 when on the fp32 SIMT path */
int resnet_b200_uses_tensor_cores(Train_ResNet * trainer);
void resnet_b200_destroy_trainer(Train_ResNet * trainer);

/* ---- single-operator entry points (device pointers, fp32, NHWC, weights [Cout][Cin][k][k]) -- */
/* impl: 0 = tcgen05/TMA implicit GEMM (product path), 1 = fp32 SIMT */
/* reference: resnet.cu:1386 prepareAndDoConvolution */
int resnet_b200_conv_forward(int in_spatial_dim, int kern_dim, int in_filters, int out_filters, int stride, int batch_size,
                             const float * input, const float * weights, float * output, int impl);
/* reference: resnet.cu:1399 prepreAndDoConvolutionDeriv (input_deriv may be NULL == toComputeInputDeriv false) */
int resnet_b200_conv_backward(int in_spatial_dim, int kern_dim, int in_filters, int out_filters, int stride, int batch_size, int to_add,
                              const float * input, const float * weights, const float * out_deriv, float * input_deriv,
                              float * weight_deriv, int impl);
/* measurement aid (tools/conv_bench.py): mean milliseconds per launch of ONE convolution pass of this geometry on synthetic device
 * data, CUDA events around `iters` launches after `warmup`; pass 0 = fprop (with_stats: fused BatchNorm statistics), 1 = dgrad,
 * 2 = dgrad accumulating, 3 = wgrad + reduce; activations in the op dtype; < 0 on error; desc receives the launch plan */
float resnet_b200_conv_bench(int in_spatial_dim, int kern_dim, int in_filters, int out_filters, int stride, int batch_size, int pass,
                             int with_stats, int warmup, int iters, char * desc, int desc_len);
/* reference: resnet.cu:1431 prepareAndDoBatchNormAndActivate (normalized_temp / normalized outputs may be NULL) */
int resnet_b200_batchnorm_forward(int spatial_dim, int filters, int batch_size, float eps, const float * input, const float * gamma,
                                  const float * beta, float * means, float * vars, float * activated, int to_activate,
                                  const float * residual, int round_tf32);
/* reference: resnet.cu:1455 prepareAndDoActivationAndBatchNormDeriv */
int resnet_b200_batchnorm_backward(int spatial_dim, int filters, int batch_size, float eps, const float * input, const float * gamma,
                                   const float * means, const float * vars, const float * activated, const float * out_layer_deriv,
                                   float * gamma_deriv, float * beta_deriv, float * input_deriv, int to_activate_deriv);
/* reference: resnet.cu:433 doMaxPool / 476 maxPoolDeriv */
int resnet_b200_maxpool_forward(const float * input, int kern_dim, int stride, int in_spatial_dim, int filters, int batch_size,
                                int * max_inds, float * out);
int resnet_b200_maxpool_backward(const int * max_inds, const float * out_deriv, int kern_dim, int in_spatial_dim, int stride, int filters,
                                 int batch_size, float * input_deriv);
/* reference: resnet.cu:500 doFilterAvgPool / 522 filterAvgPoolDeriv */
int resnet_b200_avgpool_forward(const float * input, int spatial_dim, int filters, int batch_size, float * out);
int resnet_b200_avgpool_backward(const float * pooled_deriv, int filters, int batch_size, int spatial_dim, float * out);
/* reference: resnet.cu:70 matMul, 1482/1496 prepareAndDoMatMul{Left,Right}Transpose */
int resnet_b200_matmul(const float * A, const float * B, int m, int k, int n, int transpose_a, int transpose_b, float * out);
/* reference: resnet.cu:569 softMax + 597 crossEntropyDeriv (output_deriv may be NULL) */
int resnet_b200_softmax_ce(const float * logits, const int * labels, int batch_size, int output_len, float * pred, float * output_deriv);
/* reference: resnet.cu:605-662 updateMeans/updateVars/updateParams, fused; n must be a multiple of 4 */
int resnet_b200_adam(float * params, float * grads, float * means, float * vars, long long n, float learning_rate, float weight_decay,
                     float base_mean_decay, float base_var_decay, float cur_mean_decay, float cur_var_decay, float eps);

/* ---- in-situ verification (selfcheck.cu) ----------------------------------------------------- */
/* selfcheck(1): trainers created afterwards re-derive EVERY tensor-core convolution launch of forward_pass / backwards_pass (fprop,
 * dgrad, accumulating dgrad, wgrad, stem) with the fp32 SIMT restatement of the reference's kernels (resnet.cu:109-281) from the same
 * input buffers, at the real batch size.  selfcheck_read: family 0 = fprop, 1 = dgrad, 2 = wgrad: the worst max|diff| / max|ref| seen
 * so far, the number of launches checked and the layer that produced the worst value.  A test aid: slow (fp32 CUDA-core convolutions). */
int resnet_b200_selfcheck(int enable);
int resnet_b200_selfcheck_read(Train_ResNet * trainer, int family, float * worst_rel, long long * n_checks, char * where, int where_len);

/* ---- data parallel (new) -------------------------------------------------------------------- */
/* 128-byte NCCL unique id, created on rank 0 and passed to every rank by the launcher (torch.distributed) */
int resnet_b200_dp_unique_id(void * out_id_128_bytes);
/* joins `trainer` to a world of `world_size` replicas: gradients are all-reduced (sum) in buckets on a side
 * stream, overlapped with backwards_pass; update_parameters waits per bucket.  bucket_bytes <= 0: default. */
int resnet_b200_dp_init(Train_ResNet * trainer, const void * id_128_bytes, int rank, int world_size, long long bucket_bytes);
int resnet_b200_dp_world_size(Train_ResNet * trainer);
/* host-only (no GPU): the (shard, batch) pairs that n_calls consecutive load_new_batch calls deliver on rank `rank` of `world_size`,
 * from a fresh cursor -- the single-GPU traversal of the reference (resnet.cu:1260-1295) taken with stride world_size */
int resnet_b200_loader_plan(int rank, int world_size, int batch_size, int shard_n_images, int n_calls, int * out_shard, int * out_batch);

#ifdef __cplusplus
}
#endif

#endif /* RESNET_B200_EXTRA_H */
