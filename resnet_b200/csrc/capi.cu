// capi.cu -- C-ABI services and single-operator entry points declared in include/resnet_b200.h.
#include "engine.h"
#include "prof.h"
#include "../../include/resnet_b200.h"
#include <curand.h>

using namespace rb;

namespace {
struct Tmp {
	std::vector<void *> ptrs;
	template <typename T> T *get(long long n) {
		void *p = nullptr;
		RB_CUDA(cudaMalloc(&p, (size_t)(n > 0 ? n : 1) * sizeof(T)));
		ptrs.push_back(p);
		return (T *)p;
	}
	~Tmp() { for (void *p : ptrs) cudaFree(p); }
};
int status() { return has_error() ? 1 : 0; }
int g_op_bf16 = 0;  // element type of the activation tensors the single-operator entry points read and write

// packs [Cout][Cin][k][k] into Wf / Wd on the default stream
void pack_one(const float *w, void *wf, void *wd, int cout, int cin, int taps, int rnd, Tmp &tmp) {
	PackJob job{w, wf, wd, cout, cin, taps, 0};
	PackJob *jd = tmp.get<PackJob>(1);
	RB_CUDA(cudaMemcpy(jd, &job, sizeof(job), cudaMemcpyHostToDevice));
	pack_weights(jd, 1, pack_job_blocks(cout, cin, taps), rnd, 0, g_op_bf16);
}
bool simt_in_bf16(int impl) {
	if (impl != 0 && g_op_bf16) { set_error("the SIMT fp32 convolution has no bf16 variant (impl must be 0 when the op dtype is bf16)"); return true; }
	return false;
}
}  // namespace

extern "C" {

const char *resnet_b200_last_error(void) { return last_error(); }
void resnet_b200_clear_error(void) { clear_error(); }
// storage type of trainers created from now on: -1 = follow $RESNET_B200_DTYPE (default), 0 = fp32 tensors / TF32 MMAs, 1 = bf16
int resnet_b200_set_dtype(int bf16) { g_default_bf16 = bf16 < 0 ? -1 : (bf16 ? 1 : 0); return 0; }
int resnet_b200_trainer_dtype(Train_ResNet *t) {
	Engine *e = engine_of(t);
	return e ? e->bf16 : -1;
}
// element type of the ACTIVATION tensors passed to the single-operator entry points below (weights, statistics, pooled values,
// logits and all gradients of parameters stay fp32): 0 = fp32, 1 = bf16
int resnet_b200_set_op_dtype(int bf16) { g_op_bf16 = bf16 ? 1 : 0; return 0; }
int resnet_b200_convert(const void *src, void *dst, long long n, int to_bf16) {
	if (to_bf16) convert_f32_to_bf16((const float *)src, n, dst, 0);
	else convert_bf16_to_f32(src, n, (float *)dst, 0);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}
int resnet_b200_set_device(int device) { RB_CUDA(cudaSetDevice(device)); return status(); }
void *resnet_b200_malloc(size_t bytes) { void *p = nullptr; RB_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); return p; }
void resnet_b200_free(void *p) { if (p) RB_CUDA(cudaFree(p)); }
void *resnet_b200_malloc_host(size_t bytes) { void *p = nullptr; RB_CUDA(cudaMallocHost(&p, bytes ? bytes : 1)); return p; }
void resnet_b200_free_host(void *p) { if (p) RB_CUDA(cudaFreeHost(p)); }
int resnet_b200_memcpy_h2d(void *d, const void *h, size_t n) { RB_CUDA(cudaMemcpy(d, h, n, cudaMemcpyHostToDevice)); return status(); }
int resnet_b200_memcpy_d2h(void *h, const void *d, size_t n) { RB_CUDA(cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost)); return status(); }
int resnet_b200_memcpy_d2d(void *d, const void *s, size_t n) { RB_CUDA(cudaMemcpy(d, s, n, cudaMemcpyDeviceToDevice)); return status(); }
int resnet_b200_memset(void *d, int v, size_t n) { RB_CUDA(cudaMemset(d, v, n)); return status(); }
int resnet_b200_sync(void) { RB_CUDA(cudaDeviceSynchronize()); return status(); }

void *resnet_b200_rng_create(unsigned long long seed) {
	curandGenerator_t *g = (curandGenerator_t *)malloc(sizeof(curandGenerator_t));
	if (curandCreateGenerator(g, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) { set_error("curandCreateGenerator failed"); free(g); return NULL; }
	curandSetPseudoRandomGeneratorSeed(*g, seed);
	return g;
}
void resnet_b200_rng_destroy(void *gen) {
	if (!gen) return;
	curandDestroyGenerator(*(curandGenerator_t *)gen);
	free(gen);
}

int resnet_b200_stage_batch(Train_ResNet *t, const float *images_host, const int *labels_host) {
	Engine *e = engine_of(t);
	if (!e) { set_error("stage_batch: unknown trainer"); return 1; }
	Batch *b = t->cur_batch;
	RB_CUDA(cudaMemcpyAsync(b->images, images_host, (size_t)t->batch_size * b->image_size * sizeof(float), cudaMemcpyHostToDevice, e->stream));
	RB_CUDA(cudaMemcpyAsync(b->correct_classes, labels_host, (size_t)t->batch_size * sizeof(int), cudaMemcpyHostToDevice, e->stream));
	return status();
}
int resnet_b200_stage_batch_device(Train_ResNet *t, const float *images_dev, const int *labels_dev) {
	Engine *e = engine_of(t);
	if (!e) { set_error("stage_batch_device: unknown trainer"); return 1; }
	Batch *b = t->cur_batch;
	RB_CUDA(cudaMemcpyAsync(b->images, images_dev, (size_t)t->batch_size * b->image_size * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
	RB_CUDA(cudaMemcpyAsync(b->correct_classes, labels_dev, (size_t)t->batch_size * sizeof(int), cudaMemcpyDeviceToDevice, e->stream));
	return status();
}
// Overlapped input path: prefetch_batch enqueues the host -> device copy of the NEXT batch on a copy stream (into a staging
// buffer) while the current step computes; commit_batch makes the compute stream wait for it and moves it into cur_batch with a
// device-to-device copy (0.05 ms for 154 MB).  The host buffers must be pinned and stay valid until commit_batch returns.
int resnet_b200_prefetch_batch(Train_ResNet *t, const float *images_host, const int *labels_host) {
	Engine *e = engine_of(t);
	if (!e) { set_error("prefetch_batch: unknown trainer"); return 1; }
	Batch *b = t->cur_batch;
	const size_t ib = (size_t)t->batch_size * b->image_size * sizeof(float), lb = (size_t)t->batch_size * sizeof(int);
	if (!e->copy_stream) {
		RB_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
		RB_CUDA(cudaMalloc(&e->stage_img, ib));
		RB_CUDA(cudaMalloc(&e->stage_lab, lb));
		RB_CUDA(cudaEventCreateWithFlags(&e->ev_staged, cudaEventDisableTiming));
		RB_CUDA(cudaEventCreateWithFlags(&e->ev_consumed, cudaEventDisableTiming));
		RB_CUDA(cudaEventRecord(e->ev_consumed, e->stream));
	}
	RB_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_consumed, 0));  // the previous commit has read the staging buffers
	RB_CUDA(cudaMemcpyAsync(e->stage_img, images_host, ib, cudaMemcpyHostToDevice, e->copy_stream));
	RB_CUDA(cudaMemcpyAsync(e->stage_lab, labels_host, lb, cudaMemcpyHostToDevice, e->copy_stream));
	RB_CUDA(cudaEventRecord(e->ev_staged, e->copy_stream));
	return status();
}
int resnet_b200_commit_batch(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e || !e->copy_stream) { set_error("commit_batch: no batch was prefetched"); return 1; }
	Batch *b = t->cur_batch;
	const size_t ib = (size_t)t->batch_size * b->image_size * sizeof(float), lb = (size_t)t->batch_size * sizeof(int);
	RB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_staged, 0));
	RB_CUDA(cudaMemcpyAsync(b->images, e->stage_img, ib, cudaMemcpyDeviceToDevice, e->stream));
	RB_CUDA(cudaMemcpyAsync(b->correct_classes, e->stage_lab, lb, cudaMemcpyDeviceToDevice, e->stream));
	RB_CUDA(cudaEventRecord(e->ev_consumed, e->stream));
	return status();
}
int resnet_b200_trainer_sync(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) { set_error("trainer_sync: unknown trainer"); return 1; }
	RB_CUDA(cudaStreamSynchronize(e->stream));
	return status();
}
int resnet_b200_timer_begin(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) return 1;
	RB_CUDA(cudaEventRecord(e->ev0, e->stream));
	return status();
}
float resnet_b200_timer_end_ms(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) return -1.f;
	float ms = -1.f;
	RB_CUDA(cudaEventRecord(e->ev1, e->stream));
	RB_CUDA(cudaEventSynchronize(e->ev1));
	RB_CUDA(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
	return ms;
}
int resnet_b200_loss_accuracy(Train_ResNet *t, float *loss_sum, int *n_wrong) {
	Engine *e = engine_of(t);
	if (!e) return 1;
	std::vector<float> l(e->N);
	std::vector<int> w(e->N);
	RB_CUDA(cudaMemcpyAsync(l.data(), e->row_loss, e->N * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	RB_CUDA(cudaMemcpyAsync(w.data(), e->row_wrong, e->N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
	RB_CUDA(cudaStreamSynchronize(e->stream));
	float ls = 0.f;
	int nw = 0;
	for (int i = 0; i < e->N; i++) { ls += l[i]; nw += w[i]; }
	*loss_sum = ls;
	*n_wrong = nw;
	return status();
}
int resnet_b200_set_pred_copy(Train_ResNet *t, int on) {
	Engine *e = engine_of(t);
	if (!e) { set_error("set_pred_copy: unknown trainer"); return 1; }
	e->pred_copy = on ? 1 : 0;
	return status();
}
int resnet_b200_fetch_pred(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) { set_error("fetch_pred: unknown trainer"); return 1; }
	RB_CUDA(cudaMemcpyAsync(e->pred_host, e->pred, (size_t)e->N * t->model->dims->output * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
	RB_CUDA(cudaStreamSynchronize(e->stream));
	return status();
}
int resnet_b200_epoch_stats(Train_ResNet *t, double *loss_sum, long long *n_wrong, long long *n_images, int reset) {
	Engine *e = engine_of(t);
	if (!e) { set_error("epoch_stats: unknown trainer"); return 1; }
	double acc[2] = {0, 0};
	RB_CUDA(cudaMemcpyAsync(acc, e->epoch_acc, sizeof(acc), cudaMemcpyDeviceToHost, e->stream));
	if (reset) RB_CUDA(cudaMemsetAsync(e->epoch_acc, 0, sizeof(acc), e->stream));
	RB_CUDA(cudaStreamSynchronize(e->stream));
	if (loss_sum) *loss_sum = acc[0];
	if (n_wrong) *n_wrong = (long long)(acc[1] + 0.5);
	if (n_images) *n_images = e->epoch_images;
	if (reset) e->epoch_images = 0;
	return status();
}
long long resnet_b200_launch_count(void) { return g_launches; }
void resnet_b200_profile(int enable) { prof_reset(); prof_enable(enable != 0); }
int resnet_b200_profile_read(int family, double *ms, long long *launches, double *work) { return prof_read(family, ms, launches, work); }
int resnet_b200_profile_read2(int family, double *ms, long long *launches, double *work, double *work2) { return prof_read(family, ms, launches, work, work2); }
int resnet_b200_uses_tensor_cores(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) return 0;
	int n = 0;
	for (auto &b : e->blocks) n += b.reduce.use_tc + b.spatial.use_tc + b.expand.use_tc;
	return n > 0;
}
// Frees everything init_trainer / init_resnet / init_general_batch created for this trainer on the device and in the side tables
// (the reference has no destroy functions; nothing is ever freed there).  The trainer, its model and its batch are unusable afterwards;
// a second call, or any entry point on the stale pointer, finds no engine and does nothing.
void resnet_b200_destroy_trainer(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) return;
	cudaStreamSynchronize(e->stream);
	if (t->cur_batch) loader_release(t->cur_batch);  // stops the prefetch thread before its streams / buffers go away
	dp_release(e);
	selfcheck_release(e);
	for (auto &b : e->blocks)
		for (ConvRef *c : {&b.reduce, &b.spatial, &b.expand, &b.proj}) {
			if (c->fprop) tc_free(c->fprop);
			if (c->dgrad) tc_free(c->dgrad);
			if (c->wgrad) tc_free(c->wgrad);
		}
	if (e->stem_fprop) tc_free(e->stem_fprop);
	if (e->stem_wgrad) tc_free(e->stem_wgrad);
	for (void *p : e->allocs) cudaFree(p);
	arena_close(e->arena);
	for (Params *P : {t->model->params, t->backprop_buffer->param_derivs, t->backprop_buffer->prev_means, t->backprop_buffer->prev_vars}) {
		ParamStore *ps = param_store_of(P);
		if (ps) { cudaFree(ps->base); delete ps; }
	}
	if (e->copy_stream) {
		cudaStreamSynchronize(e->copy_stream);
		cudaFree(e->stage_img); cudaFree(e->stage_lab);
		cudaEventDestroy(e->ev_staged); cudaEventDestroy(e->ev_consumed);
		cudaStreamDestroy(e->copy_stream);
	}
	cudaFreeHost(e->pred_host);
	cudaFreeHost(e->bad_host);
	if (e->wstream) {
		cudaStreamSynchronize(e->wstream);
		cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_join);
		for (int k = 0; k < 4; k++) cudaEventDestroy(e->ev_rd[k]);
		cudaStreamDestroy(e->wstream);
	}
	cudaEventDestroy(e->ev0);
	cudaEventDestroy(e->ev1);
	cudaStreamDestroy(e->stream);
	if (Batch *b = t->cur_batch) {
		cudaFree(b->images); cudaFree(b->correct_classes);
		cudaFreeHost(b->images_float_cpu); cudaFreeHost(b->correct_classes_cpu);
		b->images = nullptr; b->correct_classes = nullptr; b->images_float_cpu = nullptr; b->correct_classes_cpu = nullptr;
	}
	engine_forget(t);
	delete e;
}

// ---------------------------------------------------------------------------------------------- single operators
int resnet_b200_conv_forward(int S, int k, int cin, int cout, int stride, int N, const float *input, const float *weights, float *output, int impl) {
	Tmp tmp;
	const int bf = g_op_bf16;
	if (simt_in_bf16(impl)) return status();
	ConvGeom g{N, S, cin, cout, k, stride};
	void *wf = tmp.get<char>(g.w_elems() * 4);
	pack_one(weights, wf, nullptr, cout, cin, k * k, 0, tmp);
	if (impl == 0 && tc_stem_supported(S, k, cin, cout, stride, bf)) {
		void *xp = tmp.get<char>((long long)stem_xp_bytes(N, S, bf)), *wfs = tmp.get<char>((long long)stem_wfs_bytes(cout, bf));
		stem_pad_input(input, N, S, xp, 0, bf, 0);  // the stem's input is the fp32 batch in both modes
		stem_pack_weights(weights, cout, wfs, 0, bf, 0);
		TcPlan *pl = tc_make_stem_fprop(N, S, cout, xp, wfs, output, bf);
		if (pl) { tc_run(pl, 0); RB_CUDA(cudaDeviceSynchronize()); tc_free(pl); }
	} else if (impl == 0) {
		TcPlan *pl = tc_make_fprop(g, input, wf, output, bf);
		if (pl && getenv("RESNET_B200_OP_STATS")) tc_attach_stats(pl, tmp.get<float>((long long)tc_stats_floats(cout)));  // tools/conv_probe.py: epilogue with the fused statistics
		if (pl) { tc_run(pl, 0); RB_CUDA(cudaDeviceSynchronize()); tc_free(pl); }
	} else {
		simt_conv_fprop(g, input, (const float *)wf, output, 0);
	}
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}

int resnet_b200_conv_backward(int S, int k, int cin, int cout, int stride, int N, int to_add, const float *input, const float *weights,
                              const float *out_deriv, float *input_deriv, float *weight_deriv, int impl) {
	Tmp tmp;
	const int bf = g_op_bf16;
	if (simt_in_bf16(impl)) return status();
	ConvGeom g{N, S, cin, cout, k, stride};
	void *wf = tmp.get<char>(g.w_elems() * 4), *wd = tmp.get<char>(g.w_elems() * 4);
	pack_one(weights, wf, wd, cout, cin, k * k, 0, tmp);
	if (impl == 0 && tc_stem_supported(S, k, cin, cout, stride, bf)) {
		// stem: weight gradient on the tensor cores; the (never needed, reference: resnet.cu:2243-2245) input gradient stays SIMT
		if (input_deriv && !bf) simt_conv_dgrad(g, out_deriv, (const float *)wd, input_deriv, to_add, 0);
		if (input_deriv && bf) set_error("the stem has no input gradient in bf16 mode (reference: resnet.cu:2243-2245 never computes it)");
		void *xp = tmp.get<char>((long long)stem_xp_bytes(N, S, bf));
		stem_pad_input(input, N, S, xp, 0, bf, 0);
		size_t ws = tc_stem_wgrad_workspace_bytes(N, S, cout, bf);
		float *wsp = (float *)tmp.get<char>((long long)ws);
		TcPlan *pl = tc_make_stem_wgrad(N, S, cout, xp, out_deriv, weight_deriv, wsp, ws, bf);
		if (pl) { tc_run(pl, 0); RB_CUDA(cudaDeviceSynchronize()); tc_free(pl); }
	} else if (impl == 0) {
		if (input_deriv) {
			TcPlan *pl = tc_make_dgrad(g, out_deriv, wd, input_deriv, to_add, bf);
			if (pl) { tc_run(pl, 0); RB_CUDA(cudaDeviceSynchronize()); tc_free(pl); }
		}
		size_t ws = tc_wgrad_workspace_bytes(g, bf);
		float *wsp = (float *)tmp.get<char>((long long)ws);
		TcPlan *pl = tc_make_wgrad(g, input, out_deriv, weight_deriv, wsp, ws, bf);
		if (pl) { tc_run(pl, 0); RB_CUDA(cudaDeviceSynchronize()); tc_free(pl); }
	} else {
		if (input_deriv) simt_conv_dgrad(g, out_deriv, (const float *)wd, input_deriv, to_add, 0);
		simt_conv_wgrad(g, input, out_deriv, weight_deriv, 0);
	}
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}

// Times one convolution pass of the given geometry on synthetic device data: plan built once, `warmup` + `iters` launches on the
// default stream bracketed by CUDA events; returns the mean milliseconds per launch (< 0 on error).  pass: 0 = fprop (with_stats: the
// epilogue also produces the fused BatchNorm statistics), 1 = dgrad, 2 = dgrad accumulating into dx, 3 = wgrad (+ its reduce).
// Activations in the op dtype (resnet_b200_set_op_dtype).  tools/conv_bench.py walks the network's layer shapes with it.
float resnet_b200_conv_bench(int S, int k, int cin, int cout, int stride, int N, int pass, int with_stats, int warmup, int iters, char *desc, int desc_len) {
	Tmp tmp;
	const int bf = g_op_bf16;
	const size_t es = bf ? 2 : 4;
	ConvGeom g{N, S, cin, cout, k, stride};
	const bool stem = tc_stem_supported(S, k, cin, cout, stride, bf);
	if (!stem && !tc_supported(g, bf)) { set_error("conv_bench: unsupported geometry"); return -1.f; }
	if (stem && (pass == 1 || pass == 2)) { set_error("conv_bench: the stem has no input gradient"); return -1.f; }
	curandGenerator_t gen;
	if (curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) { set_error("conv_bench: curandCreateGenerator failed"); return -1.f; }
	curandSetPseudoRandomGeneratorSeed(gen, 99);
	auto rand_act = [&](long long n) -> void * {  // N(0, 1) values in the op dtype
		n = (n + 1) / 2 * 2;
		float *f = tmp.get<float>(n);
		curandGenerateNormal(gen, f, (size_t)n, 0.f, 1.f);
		if (!bf) return f;
		void *h = tmp.get<char>(n * 2);
		convert_f32_to_bf16(f, n, h, 0);
		return h;
	};
	float *w = tmp.get<float>(g.w_elems() + 1);
	curandGenerateNormal(gen, w, (size_t)((g.w_elems() + 1) / 2 * 2), 0.f, 0.05f);
	void *wf = tmp.get<char>(g.w_elems() * 4), *wd = tmp.get<char>(g.w_elems() * 4);
	pack_one(w, wf, wd, cout, cin, k * k, bf ? 0 : 1, tmp);
	void *x = stem ? nullptr : rand_act(g.in_elems());
	void *dy = (pass == 0) ? tmp.get<char>(g.out_elems() * (long long)es) : rand_act(g.out_elems());
	TcPlan *pl = nullptr;
	if (stem) {
		float *img = tmp.get<float>(g.in_elems() + 1);
		curandGenerateNormal(gen, img, (size_t)((g.in_elems() + 1) / 2 * 2), 0.f, 60.f);
		void *xp = tmp.get<char>((long long)stem_xp_bytes(N, S, bf)), *wfs = tmp.get<char>((long long)stem_wfs_bytes(cout, bf));
		stem_pad_input(img, N, S, xp, 1, bf, 0);
		stem_pack_weights(w, cout, wfs, 1, bf, 0);
		if (pass == 0) pl = tc_make_stem_fprop(N, S, cout, xp, wfs, dy, bf);
		else {
			const size_t ws = tc_stem_wgrad_workspace_bytes(N, S, cout, bf);
			pl = tc_make_stem_wgrad(N, S, cout, xp, dy, tmp.get<float>(g.w_elems()), (float *)tmp.get<char>((long long)ws), ws, bf);
		}
	} else if (pass == 0) pl = tc_make_fprop(g, x, wf, dy, bf);
	else if (pass == 1 || pass == 2) pl = tc_make_dgrad(g, dy, wd, tmp.get<char>(g.in_elems() * (long long)es), pass == 2, bf);
	else {
		const size_t ws = tc_wgrad_workspace_bytes(g, bf);
		pl = tc_make_wgrad(g, x, dy, tmp.get<float>(g.w_elems()), (float *)tmp.get<char>((long long)ws), ws, bf);
	}
	curandDestroyGenerator(gen);
	if (!pl) return -1.f;
	if (pass == 0 && with_stats) tc_attach_stats(pl, tmp.get<float>((long long)tc_stats_floats(cout)));
	if (desc && desc_len > 0) tc_describe(pl, desc, (size_t)desc_len);
	cudaEvent_t e0, e1;
	RB_CUDA(cudaEventCreate(&e0));
	RB_CUDA(cudaEventCreate(&e1));
	for (int i = 0; i < warmup; i++) tc_run(pl, 0);
	RB_CUDA(cudaEventRecord(e0, 0));
	for (int i = 0; i < iters; i++) tc_run(pl, 0);
	RB_CUDA(cudaEventRecord(e1, 0));
	RB_CUDA(cudaEventSynchronize(e1));
	float ms = -1.f;
	RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	tc_free(pl);
	RB_CUDA(cudaDeviceSynchronize());
	return has_error() ? -1.f : ms / (float)(iters > 0 ? iters : 1);
}

int resnet_b200_batchnorm_forward(int S, int C, int N, float eps, const float *input, const float *gamma, const float *beta, float *means,
                                  float *vars, float *activated, int to_activate, const float *residual, int rnd) {
	Tmp tmp;
	const long long rows = (long long)N * S * S;
	const int maxb = kNumSMs * 8;
	float *partials = tmp.get<float>((long long)maxb * 2 * C), *ab = tmp.get<float>(2LL * C);
	bn_stats(input, rows, C, gamma, beta, eps, means, vars, ab, partials, maxb, 0, g_op_bf16);
	bn_apply(input, ab, rows, C, to_activate, residual, nullptr, activated, rnd, 0, g_op_bf16);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}

int resnet_b200_batchnorm_backward(int S, int C, int N, float eps, const float *input, const float *gamma, const float *means, const float *vars,
                                   const float *activated, const float *out_layer_deriv, float *gamma_deriv, float *beta_deriv,
                                   float *input_deriv, int to_activate_deriv) {
	Tmp tmp;
	const long long rows = (long long)N * S * S;
	const int maxb = kNumSMs * 8;
	float *partials = tmp.get<float>((long long)maxb * 2 * C), *coef = tmp.get<float>(4LL * C);
	bn_bwd(input, out_layer_deriv, to_activate_deriv ? activated : nullptr, gamma, means, vars, eps, rows, C, gamma_deriv, beta_deriv, input_deriv,
	       partials, maxb, coef, 0, 0, nullptr, g_op_bf16);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}

int resnet_b200_maxpool_forward(const float *input, int k, int stride, int S, int C, int N, int *max_inds, float *out) {
	maxpool_fwd(input, N, S, C, k, stride, max_inds, out, 0, g_op_bf16);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}
int resnet_b200_maxpool_backward(const int *max_inds, const float *out_deriv, int k, int S, int stride, int C, int N, float *input_deriv) {
	maxpool_bwd(max_inds, out_deriv, N, S, C, k, stride, input_deriv, 0, g_op_bf16);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}
int resnet_b200_avgpool_forward(const float *input, int S, int C, int N, float *out) {
	avgpool_fwd(input, N, S, C, out, 0, g_op_bf16);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}
int resnet_b200_avgpool_backward(const float *pooled_deriv, int C, int N, int S, float *out) {
	avgpool_bwd(pooled_deriv, N, S, C, out, 0, g_op_bf16);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}
int resnet_b200_matmul(const float *A, const float *B, int m, int k, int n, int ta, int tb, float *out) {
	sgemm(A, B, out, m, n, k, ta, tb, 0);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}
int resnet_b200_softmax_ce(const float *logits, const int *labels, int N, int L, float *pred, float *output_deriv) {
	softmax_ce(logits, labels, N, L, pred, output_deriv, nullptr, nullptr, 0);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}
int resnet_b200_adam(float *p, float *g, float *m, float *v, long long n, float lr, float wd, float b1, float b2, float cb1, float cb2, float eps) {
	adam_step(p, g, m, v, n, lr, wd, b1, b2, cb1, cb2, eps, nullptr, 0);
	RB_CUDA(cudaDeviceSynchronize());
	return status();
}

}  // extern "C"
