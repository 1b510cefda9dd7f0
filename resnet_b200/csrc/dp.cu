// dp.cu -- batch-sharded data parallelism over the GPUs of one box (new; the reference is single-GPU).
//
// One process per GPU, each with a full Train_ResNet replica.  Gradients live in ONE contiguous fp32 arena in
// the reference's locations[] order (reference: resnet.cu:839-943), so a bucket is a byte range.  backwards_pass
// produces them in reverse location order (FC first, stem last; reference: resnet.cu:1901 loop + 2952 reverse
// update order); as soon as a block's wgrads are enqueued we record an event and issue ncclAllReduce(SUM) for
// every bucket that is now complete on a side stream, so the NVLink traffic of deep layers overlaps the
// backward compute of shallow ones.  update_parameters' Adam waits for the last bucket.  The loss gradient is a
// batch SUM (no 1/N, reference: resnet.cu:1806-1811), so allreduce-SUM reproduces single-GPU large-batch
// gradients exactly; BatchNorm statistics stay per-GPU.
//
// NCCL is bound with dlopen so that single-GPU users do not need libnccl at load time.
#include "engine.h"
#include "../../include/resnet_b200.h"
#include <dlfcn.h>

namespace rb {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };

struct NcclApi {
	void *lib;
	int (*GetUniqueId)(ncclUniqueId *);
	int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
	int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
	int (*CommDestroy)(ncclComm_t);
	const char *(*GetErrorString)(int);
};
static NcclApi *nccl() {
	static NcclApi api;
	static bool tried = false;
	if (tried) return api.lib ? &api : nullptr;
	tried = true;
	for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
		api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
		if (api.lib) break;
	}
	if (!api.lib) { set_error("dlopen(libnccl.so.2) failed: %s", dlerror()); return nullptr; }
	api.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(api.lib, "ncclGetUniqueId");
	api.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
	api.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
	api.Broadcast = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclBroadcast");
	api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
	api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
	if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Broadcast) { set_error("libnccl: missing symbols"); api.lib = nullptr; return nullptr; }
	return &api;
}

struct Bucket { long long off, len; int first_block; };  // ready once block `first_block` (or the stem: -1) is done
struct DpState {
	ncclComm_t comm;
	int rank, world;
	cudaStream_t comm_stream;
	cudaEvent_t ready, ready_w, done;
	std::vector<Bucket> buckets;  // in issue order (deepest first)
	size_t next;
	float *grad_base;
};

// bucket plan: walk locations from the last (FC) to the first (stem), cutting at block boundaries once a bucket
// holds >= bucket_bytes.  Pure function of (offsets, block starts): unit-tested on the host via dp_plan_buckets.
static std::vector<Bucket> plan_buckets(const std::vector<long long> &offs, long long total, const std::vector<int> &block_first_loc,
                                        long long bucket_floats) {
	std::vector<Bucket> out;
	long long hi = total;
	const int nb = (int)block_first_loc.size();
	for (int b = nb - 1; b >= 0; b--) {
		const long long lo = offs[block_first_loc[b]];
		if (hi - lo >= bucket_floats || b == 0) {
			if (b == 0) {  // the last bucket also carries the stem (locations 0..2), ready only after the stem wgrad
				if (hi - lo >= bucket_floats) { out.push_back({lo, hi - lo, 0}); hi = lo; }
				out.push_back({0, hi, -1});
			} else { out.push_back({lo, hi - lo, b}); hi = lo; }
		}
	}
	return out;
}

static void issue_ready(Engine *e, int finished_block) {
	DpState *s = (DpState *)e->dp;
	NcclApi *api = nccl();
	bool recorded = false;
	while (s->next < s->buckets.size()) {
		const Bucket &b = s->buckets[s->next];
		const bool ready = (b.first_block >= 0) ? (finished_block >= 0 ? finished_block <= b.first_block : true) : (finished_block < 0);
		if (!ready) break;
		if (!recorded) {
			RB_CUDA(cudaEventRecord(s->ready, e->stream));
			RB_CUDA(cudaStreamWaitEvent(s->comm_stream, s->ready, 0));
			if (e->wstream) {  // the block's weight gradients are produced on the side stream (Engine::wstream)
				RB_CUDA(cudaEventRecord(s->ready_w, e->wstream));
				RB_CUDA(cudaStreamWaitEvent(s->comm_stream, s->ready_w, 0));
			}
			recorded = true;
		}
		int r = api->AllReduce(s->grad_base + b.off, s->grad_base + b.off, (size_t)b.len, ncclFloat32, ncclSum, s->comm, s->comm_stream);
		if (r != ncclSuccess) set_error("ncclAllReduce failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
		s->next++;
	}
}

void dp_rank_world(const Engine *e, int *rank, int *world) {
	const DpState *s = (const DpState *)e->dp;
	*rank = s ? s->rank : 0;
	*world = s ? s->world : 1;
}

void dp_release(Engine *e) {
	DpState *s = (DpState *)e->dp;
	if (!s) return;
	cudaStreamSynchronize(s->comm_stream);
	NcclApi *api = nccl();
	if (api && api->CommDestroy) api->CommDestroy(s->comm);
	cudaEventDestroy(s->ready);
	cudaEventDestroy(s->ready_w);
	cudaEventDestroy(s->done);
	cudaStreamDestroy(s->comm_stream);
	delete s;
	e->dp = nullptr;
}

void dp_block_done(Engine *e, int block) {
	if (!e->dp) return;
	issue_ready(e, block);
}

// called at the end of backwards_pass: flush the remaining buckets and make the compute stream wait for all of them
void dp_allreduce_grads(Engine *e) {
	if (!e->dp) return;
	DpState *s = (DpState *)e->dp;
	issue_ready(e, -1);
	RB_CUDA(cudaEventRecord(s->done, s->comm_stream));
	RB_CUDA(cudaStreamWaitEvent(e->stream, s->done, 0));
	s->next = 0;
}

}  // namespace rb

using namespace rb;

extern "C" {

int resnet_b200_dp_unique_id(void *out) {
	NcclApi *api = nccl();
	if (!api) return 1;
	ncclUniqueId id;
	int r = api->GetUniqueId(&id);
	if (r != ncclSuccess) { set_error("ncclGetUniqueId failed (%d)", r); return 1; }
	memcpy(out, &id, sizeof(id));
	return 0;
}

int resnet_b200_dp_init(Train_ResNet *t, const void *id_bytes, int rank, int world, long long bucket_bytes) {
	Engine *e = engine_of(t);
	if (!e) { set_error("dp_init: unknown trainer"); return 1; }
	if (world <= 1) return 0;
	NcclApi *api = nccl();
	if (!api) return 1;
	DpState *s = new DpState();
	s->rank = rank; s->world = world; s->next = 0;
	ncclUniqueId id;
	memcpy(&id, id_bytes, sizeof(id));
	int r = api->CommInitRank(&s->comm, world, id, rank);
	if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?"); delete s; return 1; }
	RB_CUDA(cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
	RB_CUDA(cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
	RB_CUDA(cudaEventCreateWithFlags(&s->ready_w, cudaEventDisableTiming));
	RB_CUDA(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
	ParamStore *g = param_store_of(t->backprop_buffer->param_derivs);
	s->grad_base = g->base;
	std::vector<int> first;
	for (auto &b : e->blocks) first.push_back(b.reduce.loc);
	if (bucket_bytes <= 0) bucket_bytes = 32LL << 20;
	s->buckets = plan_buckets(g->offs, g->total, first, bucket_bytes / 4);
	// Replicas must start identical: rank 0's parameters and Adam moments (whatever seed, checkpoint restore or set_params put there)
	// and its optimizer clocks replace every other rank's.  Without this a rank that was seeded or restored differently diverges silently.
	RB_CUDA(cudaStreamSynchronize(e->stream));
	ParamStore *trees[3] = {param_store_of(t->model->params), param_store_of(t->backprop_buffer->prev_means), param_store_of(t->backprop_buffer->prev_vars)};
	for (ParamStore *ps : trees) {
		if (!ps) { set_error("dp_init: trainer has no parameter arena"); continue; }
		r = api->Broadcast(ps->base, ps->base, (size_t)ps->total, ncclFloat32, 0, s->comm, s->comm_stream);
		if (r != ncclSuccess) set_error("ncclBroadcast failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
	}
	float *clk = nullptr;
	RB_CUDA(cudaMalloc(&clk, 2 * sizeof(float)));
	const float hclk[2] = {t->cur_mean_decay, t->cur_var_decay};
	RB_CUDA(cudaMemcpyAsync(clk, hclk, sizeof(hclk), cudaMemcpyHostToDevice, s->comm_stream));
	r = api->Broadcast(clk, clk, 2, ncclFloat32, 0, s->comm, s->comm_stream);
	if (r != ncclSuccess) set_error("ncclBroadcast failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
	float rclk[2] = {hclk[0], hclk[1]};
	RB_CUDA(cudaMemcpyAsync(rclk, clk, sizeof(rclk), cudaMemcpyDeviceToHost, s->comm_stream));
	RB_CUDA(cudaStreamSynchronize(s->comm_stream));
	RB_CUDA(cudaFree(clk));
	t->cur_mean_decay = rclk[0];
	t->cur_var_decay = rclk[1];
	e->dp = s;
	return has_error() ? 1 : 0;
}

int resnet_b200_dp_world_size(Train_ResNet *t) {
	Engine *e = engine_of(t);
	return (e && e->dp) ? ((DpState *)e->dp)->world : 1;
}

// host-only view of the bucket plan for tests: fills off/len/first_block (each up to max entries), returns the count
int resnet_b200_dp_plan(const long long *offs, int n_locs, long long total, const int *block_first_loc, int n_blocks, long long bucket_floats,
                        long long *out_off, long long *out_len, int *out_first_block, int max_out) {
	std::vector<long long> o(offs, offs + n_locs);
	std::vector<int> f(block_first_loc, block_first_loc + n_blocks);
	std::vector<Bucket> b = plan_buckets(o, total, f, bucket_floats);
	for (size_t i = 0; i < b.size() && (int)i < max_out; i++) { out_off[i] = b[i].off; out_len[i] = b[i].len; out_first_block[i] = b[i].first_block; }
	return (int)b.size();
}

}  // extern "C"
