// model.cu -- the reference's entry points (resnet.h) on the B200 path: allocation (init_*), forward_pass,
// backwards_pass, update_parameters, load_new_batch.
//
// reference: resnet.cu:666-1231 (init_*), 1235-1381 (loader, class metadata), 1526-1775 (forward_pass),
// 1777-2248 (backwards_pass; block wiring per resnet_clean.cu:2459-2958), 2910-2987 (update_parameters).
//
// Differences that are deliberate: arena allocation instead of ~1400 cudaMallocs; nothing is
// allocated inside the step; fused kernels (BN apply + residual + ReLU; Adam m/v/p + grad zeroing); recomputable
// buffers (x-hat, pre-ReLU sums) exist only in keep-all mode; one stream instead of the legacy default stream.
#include "engine.h"
#include "prof.h"
#include <curand.h>
#include <map>
#include <mutex>

namespace rb {

static std::map<const Params *, ParamStore *> g_param_stores;
static std::map<const Train_ResNet *, Engine *> g_engines;
static std::mutex g_mu;

ParamStore *param_store_of(const Params *p) {
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_param_stores.find(p);
	return it == g_param_stores.end() ? nullptr : it->second;
}
Engine *engine_of(const Train_ResNet *t) {
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_engines.find(t);
	return it == g_engines.end() ? nullptr : it->second;
}

void engine_forget(const Train_ResNet *t) {
	std::lock_guard<std::mutex> lk(g_mu);
	g_engines.erase(t);
	if (t->model) g_param_stores.erase(t->model->params);
	if (t->backprop_buffer)
		for (const Params *P : {t->backprop_buffer->param_derivs, t->backprop_buffer->prev_means, t->backprop_buffer->prev_vars}) g_param_stores.erase(P);
}

int g_default_bf16 = -1;  // resnet_b200_set_dtype(): -1 = follow $RESNET_B200_DTYPE, 0 = fp32/tf32, 1 = bf16

static int env_int(const char *name, int dflt) {
	const char *v = getenv(name);
	return v ? atoi(v) : dflt;
}

static long long align_up(long long n, long long a) { return (n + a - 1) / a * a; }

__global__ void fill_kernel(float *p, long long n, float v) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
static void fill(float *p, long long n, float v, cudaStream_t st) {
	if (n <= 0) return;
	int grid = (int)((n + 255) / 256);
	fill_kernel<<<grid > 2048 ? 2048 : grid, 256, 0, st>>>(p, n, v);
	RB_LAUNCH_CHECK();
}

// d = pred - onehot (reference: resnet.cu:1800-1804 memcpy + crossEntropyDeriv; no 1/N, 1806-1811)
__global__ void ce_deriv_kernel(const float *__restrict__ pred, const int *__restrict__ labels, int N, int L, float *__restrict__ d) {
	const long long total = (long long)N * L;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int row = (int)(i / L), col = (int)(i % L);
		d[i] = pred[i] - (labels[row] == col ? 1.f : 0.f);
	}
}

// epoch_acc[0] += sum_i row_loss[i], epoch_acc[1] += sum_i row_wrong[i]: one block, fixed-order tree (deterministic), fp64 running sums
__global__ void epoch_accumulate_kernel(const float *__restrict__ row_loss, const int *__restrict__ row_wrong, int N, double *__restrict__ acc) {
	__shared__ double sl[256], sw[256];
	double l = 0, w = 0;
	for (int i = threadIdx.x; i < N; i += 256) { l += (double)row_loss[i]; w += (double)row_wrong[i]; }
	sl[threadIdx.x] = l; sw[threadIdx.x] = w;
	__syncthreads();
	for (int o = 128; o; o >>= 1) {
		if ((int)threadIdx.x < o) { sl[threadIdx.x] += sl[threadIdx.x + o]; sw[threadIdx.x] += sw[threadIdx.x + o]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) { acc[0] += sl[0]; acc[1] += sw[0]; }
}

// ------------------------------------------------------------------------------------------------ parameters
// Builds a Params tree over one arena.  gen != NULL: weights ~ N(0, 2/(fan_in+fan_out)), FC ~ N(0, 1e-4), gamma 1,
// beta 0, drawn with curandGenerateNormal in the reference's order (reference: resnet.cu:730,741,754,790,835,938)
// so the same generator seed reproduces the reference's initial bytes.  gen == NULL: all zeros (is_zero).
static Params *make_params(Dims *d, curandGenerator_t *gen) {
	const int nb = d->n_conv_blocks;
	Params *P = (Params *)calloc(1, sizeof(Params));
	P->n_locations = 16 + 9 * nb;  // reference: resnet.cu:819 (upper bound; trailing entries unused)
	P->locations = (float **)calloc(P->n_locations, sizeof(float *));
	P->sizes = (int *)calloc(P->n_locations, sizeof(int));
	P->conv_blocks = (ConvBlock **)calloc(nb, sizeof(ConvBlock *));

	struct Spec { long long size; int kind; float var; };  // kind 0 weight, 1 gamma, 2 beta
	std::vector<Spec> specs;
	auto add_conv = [&](int cout, int cin, int k) {
		specs.push_back({(long long)cout * cin * k * k, 0, 2.0f / (float)(k * k * (cin + cout))});
		specs.push_back({cout, 1, 0.f});
		specs.push_back({cout, 2, 0.f});
	};
	add_conv(d->init_conv_filters, 3, d->init_kernel_dim);
	int incoming = d->init_conv_filters, spatial = d->input / 4, reduced = d->init_conv_filters, expanded = 4 * d->init_conv_filters;
	struct BlkDims { int incoming, spatial, reduced, expanded, stride; bool proj; };
	std::vector<BlkDims> bd;
	for (int i = 0; i < nb; i++) {
		int stride = 1;
		if (d->is_block_spatial_reduction[i] == 1) { stride = 2; reduced *= 2; expanded *= 2; }
		const bool proj = incoming != expanded;
		bd.push_back({incoming, spatial, reduced, expanded, stride, proj});
		add_conv(reduced, incoming, 1);
		add_conv(reduced, reduced, 3);
		add_conv(expanded, reduced, 1);
		if (proj) add_conv(expanded, incoming, stride == 2 ? 3 : 1);
		if (stride == 2) spatial /= 2;
		incoming = expanded;
	}
	specs.push_back({(long long)expanded * d->output, 0, 0.0001f});
	const int nloc = (int)specs.size();
	if (nloc > P->n_locations) { set_error("make_params: location overflow"); return nullptr; }
	P->n_locations = nloc;  // equals 16 + 9*nb for ResNet-50's four projections (reference: resnet.cu:819)

	ParamStore *ps = new ParamStore();
	ps->tree = P;
	long long off = 0;
	for (auto &s : specs) { ps->offs.push_back(off); off += align_up(s.size, 64); }
	ps->total = off;
	RB_CUDA(cudaMalloc(&ps->base, ps->total * sizeof(float)));
	RB_CUDA(cudaMemset(ps->base, 0, ps->total * sizeof(float)));
	for (int i = 0; i < nloc; i++) {
		P->locations[i] = ps->base + ps->offs[i];
		P->sizes[i] = (int)specs[i].size;
		if (gen) {
			if (specs[i].kind == 0) {
				curandStatus_t cs = curandGenerateNormal(*gen, P->locations[i], (size_t)specs[i].size, 0.f, sqrtf(specs[i].var));
				if (cs != CURAND_STATUS_SUCCESS) set_error("curandGenerateNormal failed (%d) at location %d", (int)cs, i);
			} else if (specs[i].kind == 1) fill(P->locations[i], specs[i].size, 1.0f, 0);
		}
	}
	RB_CUDA(cudaDeviceSynchronize());

	auto mk_bn = [&](int loc, int sp, int depth) {
		BatchNorm *b = (BatchNorm *)calloc(1, sizeof(BatchNorm));
		b->spatial_dim = sp; b->depth = depth; b->gamma = P->locations[loc + 1]; b->beta = P->locations[loc + 2];
		return b;
	};
	int li = 0;
	P->init_conv_layer = P->locations[0];
	P->norm_init_conv = mk_bn(0, d->input / d->init_conv_stride, d->init_conv_filters);
	li = 3;
	for (int i = 0; i < nb; i++) {
		const BlkDims &b = bd[i];
		ConvBlock *cb = (ConvBlock *)calloc(1, sizeof(ConvBlock));
		cb->incoming_filters = b.incoming; cb->incoming_spatial_dim = b.spatial; cb->reduced_depth = b.reduced;
		cb->expanded_depth = b.expanded; cb->stride = b.stride;
		const int sp_out = b.spatial / b.stride;
		cb->depth_reduction = P->locations[li]; cb->norm_depth_reduction = mk_bn(li, b.spatial, b.reduced);
		cb->spatial = P->locations[li + 3]; cb->norm_spatial = mk_bn(li + 3, sp_out, b.reduced);
		cb->depth_expansion = P->locations[li + 6]; cb->norm_expansion = mk_bn(li + 6, sp_out, b.expanded);
		if (b.proj) { cb->projection = P->locations[li + 9]; cb->norm_projection = mk_bn(li + 9, sp_out, b.expanded); li += 12; }
		else { cb->projection = NULL; cb->norm_projection = NULL; li += 9; }
		P->conv_blocks[i] = cb;
	}
	P->fully_connected = P->locations[li];
	{
		std::lock_guard<std::mutex> lk(g_mu);
		g_param_stores[P] = ps;
	}
	return P;
}

// ------------------------------------------------------------------------------------------------ engine
// One activation arena per trainer (the reference cudaMalloc()s ~1400 buffers one by one): a range of virtual
// addresses is reserved up front (CUDA virtual memory management), tensors are bump-allocated from it with 256-byte alignment, and
// physical memory is mapped behind the bump pointer in 512 MB granules -- the arena is contiguous, exactly as large as the trainer
// needs, and nothing is allocated after init_trainer.  Falls back to one cudaMalloc per tensor if the driver refuses the VMM calls.
struct Arena {
	CUdeviceptr base = 0;
	size_t reserved = 0, mapped = 0, used = 0;
	std::vector<CUmemGenericAllocationHandle> handles;
	std::vector<size_t> sizes;
	CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
	CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
	CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
	CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
	CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
	CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
	CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
	CUresult (*granularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
	CUmemAllocationProp prop;
	size_t gran = 0;
	int device = 0;
	bool ok = false;
};
static bool arena_open(Arena *a, size_t reserve_bytes) {
	auto sym = [](const char *name) -> void * {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
		return p;
	};
	a->reserve = (decltype(a->reserve))sym("cuMemAddressReserve");
	a->create = (decltype(a->create))sym("cuMemCreate");
	a->map = (decltype(a->map))sym("cuMemMap");
	a->set_access = (decltype(a->set_access))sym("cuMemSetAccess");
	a->unmap = (decltype(a->unmap))sym("cuMemUnmap");
	a->release = (decltype(a->release))sym("cuMemRelease");
	a->addr_free = (decltype(a->addr_free))sym("cuMemAddressFree");
	a->granularity = (decltype(a->granularity))sym("cuMemGetAllocationGranularity");
	if (!a->reserve || !a->create || !a->map || !a->set_access || !a->unmap || !a->release || !a->addr_free || !a->granularity) return false;
	if (cudaGetDevice(&a->device) != cudaSuccess) return false;
	memset(&a->prop, 0, sizeof(a->prop));
	a->prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	a->prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	a->prop.location.id = a->device;
	size_t g = 0;
	if (a->granularity(&g, &a->prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || g == 0) return false;
	const size_t chunk = (size_t)512 << 20;
	a->gran = (chunk + g - 1) / g * g;
	a->reserved = (reserve_bytes + a->gran - 1) / a->gran * a->gran;
	if (a->reserve(&a->base, a->reserved, 0, 0, 0) != CUDA_SUCCESS) return false;
	a->ok = true;
	return true;
}
// maps physical memory so that [base, base + upto) is backed
static bool arena_back(Arena *a, size_t upto) {
	while (a->mapped < upto) {
		if (a->mapped + a->gran > a->reserved) { set_error("activation arena: reserved range of %zu bytes exhausted", a->reserved); return false; }
		CUmemGenericAllocationHandle h;
		CUresult r = a->create(&h, a->gran, &a->prop, 0);
		if (r != CUDA_SUCCESS) { set_error("activation arena: cuMemCreate(%zu) failed (%d) after %zu bytes", a->gran, (int)r, a->mapped); return false; }
		r = a->map(a->base + a->mapped, a->gran, 0, h, 0);
		if (r != CUDA_SUCCESS) { a->release(h); set_error("activation arena: cuMemMap failed (%d)", (int)r); return false; }
		CUmemAccessDesc acc;
		memset(&acc, 0, sizeof(acc));
		acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
		acc.location.id = a->device;
		acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
		r = a->set_access(a->base + a->mapped, a->gran, &acc, 1);
		if (r != CUDA_SUCCESS) { set_error("activation arena: cuMemSetAccess failed (%d)", (int)r); return false; }
		a->handles.push_back(h);
		a->sizes.push_back(a->gran);
		a->mapped += a->gran;
	}
	return true;
}
void arena_close(void *arena) {
	Arena *a = (Arena *)arena;
	if (!a) return;
	if (a->ok) {
		size_t off = 0;
		for (size_t i = 0; i < a->handles.size(); i++) { a->unmap(a->base + off, a->sizes[i]); a->release(a->handles[i]); off += a->sizes[i]; }
		a->addr_free(a->base, a->reserved);
	}
	delete a;
}

struct Bump {
	Engine *e;
	// activation tensor of n elements in the engine's storage type (fp32 or bf16); the public structs keep float* names
	float *act(long long n) { return (float *)get<char>((n <= 0 ? 1 : n) * (long long)e->esz); }
	template <typename T> T *get(long long n) {
		if (n <= 0) n = 1;
		const size_t bytes = (size_t)align_up(n * (long long)sizeof(T), 256);
		Arena *a = (Arena *)e->arena;
		if (a && a->ok) {
			if (arena_back(a, a->used + bytes)) {
				T *p = (T *)(uintptr_t)(a->base + a->used);
				a->used += bytes;
				return p;
			}
			return nullptr;
		}
		void *p = nullptr;  // fallback: one allocation per tensor
		RB_CUDA(cudaMalloc(&p, bytes));
		e->allocs.push_back(p);
		return (T *)p;
	}
};

static Cache_BatchNorm *mk_cache(Bump &B, long long input_size, int C, bool keep_all, bool with_stats) {
	Cache_BatchNorm *c = (Cache_BatchNorm *)calloc(1, sizeof(Cache_BatchNorm));
	c->input_size = (int)input_size; c->feature_size = C;
	if (with_stats) { c->means = B.get<float>(C); c->vars = B.get<float>(C); }
	if (keep_all) { c->normalized_temp = B.act(input_size); c->normalized = B.act(input_size); }
	return c;
}

static void setup_conv(Engine *e, Bump &B, ConvRef &c, int N, int S, int cin, int cout, int k, int stride, int loc, Params *P, Params *G,
                       std::vector<PackJob> &jobs) {
	c.g = ConvGeom{N, S, cin, cout, k, stride};
	c.loc = loc;
	c.w = P->locations[loc];
	c.dw = G->locations[loc];
	c.wf = B.act(c.g.w_elems());
	c.wd = B.act(c.g.w_elems());
	c.use_tc = (e->conv_mode == 0) && tc_supported(c.g, e->bf16);
	if (e->bf16 && !c.use_tc && k != 7) set_error("bf16 mode: conv %dx%d/%d %d->%d has no tensor-core plan (channels must be multiples of 64)", k, k, stride, cin, cout);
	c.fprop = c.dgrad = c.wgrad = nullptr;
	c.stats_rows = 0;
	jobs.push_back(PackJob{c.w, c.wf, c.wd, cout, cin, k * k, 0});
	if (c.use_tc) {
		size_t ws = tc_wgrad_workspace_bytes(c.g, e->bf16);
		if (ws > e->wgrad_ws_bytes) e->wgrad_ws_bytes = ws;
	}
}

static BnRef mk_bnref(Bump &B, BatchNorm *p, BatchNorm *gp, Cache_BatchNorm *cache, long long rows) {
	BnRef r;
	r.C = p->depth; r.rows = rows; r.gamma = p->gamma; r.beta = p->beta; r.dgamma = gp->gamma; r.dbeta = gp->beta;
	r.means = cache->means; r.vars = cache->vars; r.cache = cache;
	r.ab = B.get<float>(2 * (long long)p->depth);
	return r;
}

static Engine *build_engine(Train_ResNet *t) {
	Engine *e = new Engine();
	e->trainer = t;
	e->N = t->batch_size;
	const char *cm = getenv("RESNET_B200_CONV");
	e->conv_mode = (cm && !strcmp(cm, "simt")) ? 1 : 0;
	// storage / MMA type: fp32 tensors + kind::tf32 (BASELINE config 2) or bf16 tensors + kind::f16 (configs 3-5); fp32 master
	// weights, gradients, optimizer state, BatchNorm statistics and the FC head in both
	{
		const char *dt = getenv("RESNET_B200_DTYPE");
		e->bf16 = g_default_bf16 >= 0 ? g_default_bf16 : ((dt && !strcmp(dt, "bf16")) ? 1 : 0);
		if (e->bf16 && e->conv_mode != 0) { set_error("bf16 storage needs the tensor-core convolution path (RESNET_B200_CONV=simt is fp32 only)"); e->bf16 = 0; }
		e->esz = e->bf16 ? 2 : 4;
	}
	e->round_tf32 = (e->conv_mode == 0 && !e->bf16) ? env_int("RESNET_B200_TF32_ROUND", 1) : 0;
	e->keep_all = env_int("RESNET_B200_KEEP_ALL", 0);
	e->dp = nullptr;
	e->copy_stream = nullptr; e->stage_img = nullptr; e->stage_lab = nullptr;
	e->wgrad_ws_bytes = 0;
	// a BLOCKING stream: it synchronises with the legacy default stream, so a host driver's plain cudaMemcpy / kernels on
	// stream 0 (how the reference's own code touches these buffers) stay ordered with our launches
	RB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamDefault));
	RB_CUDA(cudaEventCreate(&e->ev0));
	RB_CUDA(cudaEventCreate(&e->ev1));
	e->wstream = nullptr;
	if (e->conv_mode == 0 && env_int("RESNET_B200_ASYNC_WGRAD", 1) && !g_selfcheck) {
		RB_CUDA(cudaStreamCreateWithFlags(&e->wstream, cudaStreamNonBlocking));
		RB_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
		RB_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
		for (int k = 0; k < 4; k++) { RB_CUDA(cudaEventCreateWithFlags(&e->ev_rd[k], cudaEventDisableTiming)); e->ev_rd_live[k] = false; }
	}
	e->arena = nullptr;
	if (env_int("RESNET_B200_ARENA", 1)) {
		Arena *a = new Arena();
		size_t free_b = 0, total_b = 0;
		RB_CUDA(cudaMemGetInfo(&free_b, &total_b));  // (creates the context before the driver entry points are used)
		if (arena_open(a, total_b ? total_b : ((size_t)192 << 30))) e->arena = a;  // address space only: reserving the device's size costs nothing
		else delete a;
	}
	Bump B{e};
	Dims *d = t->model->dims;
	Params *P = t->model->params;
	const int N = e->N, nb = d->n_conv_blocks;
	const bool ka = e->keep_all != 0;

	// ---- public structs: forward buffer
	Forward_Buffer *fb = (Forward_Buffer *)calloc(1, sizeof(Forward_Buffer));
	Backprop_Buffer *bb = (Backprop_Buffer *)calloc(1, sizeof(Backprop_Buffer));
	t->forward_buffer = fb;
	t->backprop_buffer = bb;
	bb->param_derivs = make_params(d, nullptr);
	bb->prev_means = make_params(d, nullptr);
	bb->prev_vars = make_params(d, nullptr);
	Params *G = bb->param_derivs;
	Activations *A = (Activations *)calloc(1, sizeof(Activations));
	Activations *DA = (Activations *)calloc(1, sizeof(Activations));
	fb->activations = A;
	bb->activation_derivs = DA;
	A->n_conv_blocks = DA->n_conv_blocks = nb;
	A->activation_conv_blocks = (Activation_ConvBlock **)calloc(nb, sizeof(Activation_ConvBlock *));
	DA->activation_conv_blocks = (Activation_ConvBlock **)calloc(nb, sizeof(Activation_ConvBlock *));

	std::vector<PackJob> jobs;
	const int S0 = d->input, S1 = d->input / d->init_conv_stride, S2 = S1 / d->init_maxpool_stride, F = d->init_conv_filters;
	const long long n_x0 = (long long)N * S1 * S1 * F, n_p0 = (long long)N * S2 * S2 * F;
	// max_inds holds flat int32 indices into init_conv_activated (reference: resnet.cu:459-468) and Cache_BatchNorm::input_size is an
	// int: the largest tensor of the network must stay below 2^31 elements (batch 2674 at 224 x 224 x 64 filters)
	if (n_x0 >= (1LL << 31)) set_error("batch %d: init_conv_activated has %lld elements, above the int32 index range of max_inds", N, n_x0);
	setup_conv(e, B, e->stem, N, S0, 3, F, d->init_kernel_dim, d->init_conv_stride, 0, P, G, jobs);
	e->stem_tc = (e->conv_mode == 0) && (e->bf16 || env_int("RESNET_B200_STEM_TC", 1)) && tc_stem_supported(S0, d->init_kernel_dim, 3, F, d->init_conv_stride, e->bf16);
	if (e->bf16 && !e->stem_tc) set_error("bf16 mode: the stem must be 7x7/2 with a multiple of 64 filters (<= 128)");
	e->stem_xp = e->stem_wfs = nullptr;
	e->stem_fprop = e->stem_wgrad = nullptr;
	if (e->stem_tc) {
		e->stem_xp = (float *)B.get<char>((long long)stem_xp_bytes(N, S0, e->bf16));
		e->stem_wfs = (float *)B.get<char>((long long)stem_wfs_bytes(F, e->bf16));
		size_t ws = tc_stem_wgrad_workspace_bytes(N, S0, F, e->bf16);
		if (ws > e->wgrad_ws_bytes) e->wgrad_ws_bytes = ws;
	}
	e->X0 = A->init_conv_applied = B.act(n_x0);
	A->norm_init_conv = mk_cache(B, n_x0, F, ka, true);
	// the stem's activated tensor has one reader, the max pool: unless every tensor is kept it is never written (bw_kernels.cu "fused stem tail")
	e->fuse_stem_tail = !ka && env_int("RESNET_B200_FUSE_STEM_TAIL", 1) && F % 4 == 0 &&
	                    bn_pool_fwd_supported(N, S1, F, d->init_maxpool_dim, d->init_maxpool_stride, e->bf16);
	e->Y0 = A->init_conv_activated = e->fuse_stem_tail ? nullptr : B.act(n_x0);
	e->max_inds = A->max_inds = B.get<int>(n_p0);
	e->P0 = A->init_convblock_input = B.act(n_p0);
	e->bn0 = mk_bnref(B, P->norm_init_conv, G->norm_init_conv, A->norm_init_conv, (long long)N * S1 * S1);

	// ---- blocks
	e->blocks.resize(nb);
	long long max_exp_out = n_p0, max_red_in = 1, max_red_out = 1;
	const float *x_in = e->P0;
	for (int i = 0; i < nb; i++) {
		ConvBlock *cb = P->conv_blocks[i], *gcb = G->conv_blocks[i];
		BlockRef &b = e->blocks[i];
		Activation_ConvBlock *ab = (Activation_ConvBlock *)calloc(1, sizeof(Activation_ConvBlock));
		A->activation_conv_blocks[i] = ab;
		ab->incoming_filters = cb->incoming_filters; ab->incoming_spatial_dim = cb->incoming_spatial_dim;
		ab->reduced_depth = cb->reduced_depth; ab->expanded_depth = cb->expanded_depth; ab->stride = cb->stride;
		const int Sin = cb->incoming_spatial_dim, Sout = Sin / cb->stride;
		b.has_proj = cb->projection != NULL;
		b.n_in = (long long)N * Sin * Sin * cb->incoming_filters;
		b.n_red_in = (long long)N * Sin * Sin * cb->reduced_depth;
		b.n_red_out = (long long)N * Sout * Sout * cb->reduced_depth;
		b.n_exp_out = (long long)N * Sout * Sout * cb->expanded_depth;
		max_exp_out = std::max(max_exp_out, std::max(b.n_exp_out, b.n_in));
		max_red_in = std::max(max_red_in, b.n_red_in);
		max_red_out = std::max(max_red_out, b.n_red_out);
		int li = 0;
		for (int l = 0; l < P->n_locations; l++) if (P->locations[l] == cb->depth_reduction) { li = l; break; }
		setup_conv(e, B, b.reduce, N, Sin, cb->incoming_filters, cb->reduced_depth, 1, 1, li, P, G, jobs);
		setup_conv(e, B, b.spatial, N, Sin, cb->reduced_depth, cb->reduced_depth, 3, cb->stride, li + 3, P, G, jobs);
		setup_conv(e, B, b.expand, N, Sout, cb->reduced_depth, cb->expanded_depth, 1, 1, li + 6, P, G, jobs);
		if (b.has_proj) setup_conv(e, B, b.proj, N, Sin, cb->incoming_filters, cb->expanded_depth, cb->stride == 2 ? 3 : 1, cb->stride, li + 9, P, G, jobs);
		b.x_in = x_in;
		b.Xr = ab->post_reduced = B.act(b.n_red_in);
		ab->norm_post_reduced = mk_cache(B, b.n_red_in, cb->reduced_depth, ka, true);
		b.Yr = ab->post_reduced_activated = B.act(b.n_red_in);
		b.Xs = ab->post_spatial = B.act(b.n_red_out);
		ab->norm_post_spatial = mk_cache(B, b.n_red_out, cb->reduced_depth, ka, true);
		b.Ys = ab->post_spatial_activated = B.act(b.n_red_out);
		b.Xe = ab->post_expanded = B.act(b.n_exp_out);
		ab->norm_post_expanded = mk_cache(B, b.n_exp_out, cb->expanded_depth, ka, true);
		ab->post_expanded_norm_vals = ka ? B.act(b.n_exp_out) : NULL;
		if (b.has_proj) {
			b.Xp = ab->transformed_residual = B.act(b.n_exp_out);
			ab->norm_post_projection = mk_cache(B, b.n_exp_out, cb->expanded_depth, ka, true);
			ab->post_projection_norm_vals = ka ? B.act(b.n_exp_out) : NULL;
		} else { b.Xp = NULL; }
		ab->output = ka ? B.act(b.n_exp_out) : NULL;
		b.OA = ab->output_activated = B.act(b.n_exp_out);
		// 1-bit ReLU mask of the block output for the two BatchNorm backwards under the residual join (they read OA only for its sign)
		// (needs whole warps of vectors per row group: expanded_depth / vector width a multiple of 32, true for every ResNet width)
		const int vecw = e->bf16 ? 8 : 4;
		const bool bits_ok = cb->expanded_depth % (32 * vecw) == 0 && (256 % (cb->expanded_depth / vecw) == 0 || (cb->expanded_depth / vecw) % 256 == 0);
		b.oa_bits = (bits_ok && env_int("RESNET_B200_BITMASK", 1)) ? B.get<uint8_t>(b.n_exp_out / vecw) : nullptr;
		b.bn_r = mk_bnref(B, cb->norm_depth_reduction, gcb->norm_depth_reduction, ab->norm_post_reduced, (long long)N * Sin * Sin);
		b.bn_s = mk_bnref(B, cb->norm_spatial, gcb->norm_spatial, ab->norm_post_spatial, (long long)N * Sout * Sout);
		b.bn_e = mk_bnref(B, cb->norm_expansion, gcb->norm_expansion, ab->norm_post_expanded, (long long)N * Sout * Sout);
		if (b.has_proj) b.bn_p = mk_bnref(B, cb->norm_projection, gcb->norm_projection, ab->norm_post_projection, (long long)N * Sout * Sout);
		x_in = b.OA;
	}
	// ---- head
	e->pooled = A->final_conv_output_pooled = B.get<float>((long long)N * d->final_depth);
	e->logits = A->linear_output = B.get<float>((long long)N * d->output);
	e->pred = fb->pred = B.get<float>((long long)N * d->output);
	RB_CUDA(cudaMallocHost(&e->pred_host, (size_t)N * d->output * sizeof(float)));
	fb->pred_cpu = e->pred_host;
	e->dlogits = bb->output_layer_deriv = B.get<float>((long long)N * d->output);
	e->dpooled = DA->final_conv_output_pooled = B.get<float>((long long)N * d->final_depth);
	DA->linear_output = e->dlogits;
	e->row_loss = B.get<float>(N);
	e->row_wrong = B.get<int>(N);
	e->fc_ws = B.get<float>((long long)sgemm_ws_floats(N, std::max(d->output, d->final_depth)));
	e->pred_copy = 1;
	e->epoch_acc = B.get<double>(2);
	RB_CUDA(cudaMemset(e->epoch_acc, 0, 2 * sizeof(double)));
	e->epoch_images = 0;

	// ---- gradient buffers: per-role scratch (default) or a full mirror (keep-all)
	float *pp[2] = {nullptr, nullptr}, *T1 = nullptr, *T1p = nullptr, *T2 = nullptr, *T3 = nullptr;
	if (!ka) {
		pp[0] = B.act(max_exp_out); pp[1] = B.act(max_exp_out);
		T1 = B.act(max_exp_out); T2 = B.act(max_red_out); T3 = B.act(max_red_in);
		// the projection branch's dX gets its own buffer when wgrads run on the side stream (it shared T1 with the expansion branch,
		// whose BatchNorm backward overwrites it right after the projection's convolutions were launched)
		T1p = e->wstream ? B.act(max_exp_out) : T1;
	}
	e->dP0 = DA->init_convblock_input = ka ? B.act(n_p0) : pp[1];  // block 0 writes its input gradient into pp[(0+1)&1]
	e->dY0 = DA->init_conv_activated = e->fuse_stem_tail ? nullptr : B.act(n_x0);
	e->dX0 = DA->init_conv_applied = (ka || e->fuse_stem_tail) ? B.act(n_x0) : e->dY0;
	DA->norm_init_conv = mk_cache(B, n_x0, F, false, false);
	DA->max_inds = NULL;
	for (int i = 0; i < nb; i++) {
		BlockRef &b = e->blocks[i];
		Activation_ConvBlock *ab = A->activation_conv_blocks[i];
		Activation_ConvBlock *db = (Activation_ConvBlock *)calloc(1, sizeof(Activation_ConvBlock));
		*db = *ab;
		DA->activation_conv_blocks[i] = db;
		db->norm_post_reduced = db->norm_post_spatial = db->norm_post_expanded = db->norm_post_projection = NULL;
		db->post_expanded_norm_vals = db->post_projection_norm_vals = db->output = NULL;
		if (ka) {
			b.dOA = B.act(b.n_exp_out);
			b.dXe = B.act(b.n_exp_out);
			b.dXp = b.has_proj ? B.act(b.n_exp_out) : NULL;
			b.dYs = B.act(b.n_red_out); b.dXs = B.act(b.n_red_out);
			b.dYr = B.act(b.n_red_in); b.dXr = B.act(b.n_red_in);
			db->output = B.act(b.n_exp_out);
		} else {
			b.dOA = pp[i & 1];
			b.dXe = T1; b.dXp = b.has_proj ? T1p : NULL;
			b.dYs = b.dXs = T2;
			b.dYr = b.dXr = T3;
		}
		db->output_activated = b.dOA;
		db->post_expanded = b.dXe; db->transformed_residual = b.dXp;
		db->post_spatial_activated = b.dYs; db->post_spatial = b.dXs;
		db->post_reduced_activated = b.dYr; db->post_reduced = b.dXr;
	}
	for (int i = 0; i < nb; i++) e->blocks[i].dBI = (i == 0) ? e->dP0 : e->blocks[i - 1].dOA;

	// ---- workspaces
	e->bn_max_blocks = kNumSMs * 8;
	int maxC = d->final_depth > F ? d->final_depth : F;
	e->bn_partials = B.get<float>((long long)e->bn_max_blocks * 2 * maxC);
	e->bn_coef = B.get<float>(4LL * maxC);
	e->stats_partials = B.get<float>((long long)tc_stats_floats(maxC));
	RB_CUDA(cudaMemsetAsync(e->stats_partials, 0, tc_stats_floats(maxC) * sizeof(float), e->stream));
	e->ones = B.get<float>(maxC); e->zeros = B.get<float>(maxC); e->tmp_ab = B.get<float>(2LL * maxC); e->tmp_mv = B.get<float>(2LL * maxC);
	fill(e->ones, maxC, 1.f, e->stream);
	fill(e->zeros, maxC, 0.f, e->stream);
	e->wgrad_ws = e->wgrad_ws_bytes ? (float *)B.get<char>((long long)e->wgrad_ws_bytes) : nullptr;
	e->bad_dev = B.get<int>(1);
	RB_CUDA(cudaMemsetAsync(e->bad_dev, 0, sizeof(int), e->stream));
	RB_CUDA(cudaMallocHost(&e->bad_host, sizeof(int)));
	*e->bad_host = 0;
	e->n_pack_jobs = (int)jobs.size();
	e->pack_max_elems = 0;  // total blocks of the one pack launch
	for (auto &j : jobs) { j.first_block = e->pack_max_elems; e->pack_max_elems += pack_job_blocks(j.cout, j.cin, j.taps); }
	e->pack_jobs_dev = (PackJob *)B.get<char>((long long)(jobs.size() * sizeof(PackJob)));
	RB_CUDA(cudaMemcpyAsync(e->pack_jobs_dev, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, e->stream));
	RB_CUDA(cudaStreamSynchronize(e->stream));

	// ---- tensor-core plans (tensor maps bind the arena addresses, so they are built once)
	const int fused_stats = env_int("RESNET_B200_FUSED_STATS", 1);
	auto plan = [&](ConvRef &c, const float *in, float *out, const float *dout, float *din, int din_accumulate) {
		if (!c.use_tc) return;
		c.fprop = tc_make_fprop(c.g, in, c.wf, out, e->bf16);
		c.stats_rows = fused_stats ? tc_attach_stats(c.fprop, e->stats_partials, 1) : 0;
		if (din) c.dgrad = tc_make_dgrad(c.g, dout, c.wd, din, din_accumulate, e->bf16);
		c.wgrad = tc_make_wgrad(c.g, in, dout, c.dw, e->wgrad_ws, e->wgrad_ws_bytes, e->bf16);
	};
	for (int i = 0; i < nb; i++) {
		BlockRef &b = e->blocks[i];
		plan(b.reduce, b.x_in, b.Xr, b.dXr, b.dBI, 1);  // joins the shortcut gradient (reference: resnet.cu:2157 toAdd=true)
		plan(b.spatial, b.Yr, b.Xs, b.dXs, b.dYr, 0);
		plan(b.expand, b.Ys, b.Xe, b.dXe, b.dYs, 0);
		if (b.has_proj) plan(b.proj, b.x_in, b.Xp, b.dXp, b.dBI, 0);
	}
	if (e->stem_tc) {
		const bool had_error = has_error();
		e->stem_fprop = tc_make_stem_fprop(N, S0, F, e->stem_xp, e->stem_wfs, e->X0, e->bf16);
		e->stem.stats_rows = fused_stats ? tc_attach_stats(e->stem_fprop, e->stats_partials, 1) : 0;
		e->stem_wgrad = tc_make_stem_wgrad(N, S0, F, e->stem_xp, e->dX0, e->stem.dw, e->wgrad_ws, e->wgrad_ws_bytes, e->bf16);
		if (!e->stem_fprop || !e->stem_wgrad) {
			// the overlapping-row tensor map was refused by the driver: keep the fp32 SIMT stem (slower, still correct)
			fprintf(stderr, "[resnet_b200] stem tensor maps unavailable (%s); stem stays on the SIMT path\n", last_error());
			if (!had_error && !e->bf16) clear_error();  // bf16 has no SIMT stem: the error stands
			e->stem_tc = false;
			e->stem.stats_rows = 0;
		}
	}
	{
		long long max_act = std::max(n_x0, (long long)N * S0 * S0 * 3), max_w = e->stem.g.w_elems();
		for (auto &b : e->blocks) {
			max_act = std::max(max_act, std::max(b.n_in, std::max(b.n_exp_out, b.n_red_in)));
			for (ConvRef *c : {&b.reduce, &b.spatial, &b.expand, &b.proj}) if (c->g.cout) max_w = std::max(max_w, c->g.w_elems());
		}
		selfcheck_init(e, max_act, max_w);
	}
	{
		std::lock_guard<std::mutex> lk(g_mu);
		g_engines[t] = e;
	}
	return e;
}

static void conv_fwd(Engine *e, ConvRef &c, const float *in, float *out);

static void stem_forward(Engine *e, const float *images) {
	if (!e->stem_tc) { conv_fwd(e, e->stem, images, e->X0); return; }
	const ConvGeom &g = e->stem.g;
	stem_pack_weights(e->stem.w, g.cout, e->stem_wfs, e->round_tf32, e->bf16, e->stream);
	stem_pad_input(images, g.N, g.S, e->stem_xp, e->round_tf32, e->bf16, e->stream);
	{
		ProfScope ps(e->stream, PROF_IGEMM_KMAJOR, 2.0 * g.N * g.So() * g.So() * (double)g.cout * g.cin * g.k * g.k);
		tc_run(e->stem_fprop, e->stream);
	}
	if (e->selfcheck) selfcheck_fprop(e, g, e->stem.w, images, true, e->X0);
}
static void stem_backward(Engine *e, const float *images) {
	if (!e->stem_tc) {  // no input gradient (reference: resnet.cu:2243-2245)
		ConvRef &c = e->stem;
		ProfScope ps(e->stream, PROF_STEM_SIMT, 2.0 * c.g.N * c.g.So() * c.g.So() * (double)c.g.cout * c.g.cin * c.g.k * c.g.k);
		simt_conv_wgrad(c.g, images, e->dX0, c.dw, e->stream);
		return;
	}
	const ConvGeom &g = e->stem.g;
	if (e->wstream && !prof_enabled()) {
		// same stream as the other weight gradients: they share the split-K workspace
		RB_CUDA(cudaEventRecord(e->ev_fork, e->stream));
		RB_CUDA(cudaStreamWaitEvent(e->wstream, e->ev_fork, 0));
		tc_run(e->stem_wgrad, e->wstream);
	} else {
		ProfScope ps(e->stream, PROF_IGEMM_WGRAD, 2.0 * g.N * g.So() * g.So() * (double)g.cout * g.cin * g.k * g.k);
		tc_run(e->stem_wgrad, e->stream);  // reads the padded copy made by this step's forward_pass
	}
	if (e->selfcheck) selfcheck_wgrad(e, g, images, true, e->dX0, e->stem.dw);
}

// ------------------------------------------------------------------------------------------------ layer helpers
// algorithmic FLOPs of one conv pass: 2 * N * Ho * Wo * Cout * Cin * k^2
static double conv_flops(const ConvGeom &g) { return 2.0 * g.N * g.So() * g.So() * (double)g.cout * g.cin * g.k * g.k; }
// fprop / dgrad of a 1x1 convolution move at least their input and output tensor once: the HBM roofline that bounds them
static double conv_io_bytes(const Engine *e, const ConvGeom &g) { return (double)e->esz * ((double)g.in_elems() + (double)g.out_elems()); }
static int kmajor_family(const ConvGeom &g) { return g.k == 1 ? PROF_IGEMM_1X1 : PROF_IGEMM_KMAJOR; }

static void conv_fwd(Engine *e, ConvRef &c, const float *in, float *out) {
	{
		ProfScope ps(e->stream, c.use_tc ? kmajor_family(c.g) : PROF_STEM_SIMT, conv_flops(c.g), conv_io_bytes(e, c.g));
		if (c.use_tc) tc_run(c.fprop, e->stream);
		else simt_conv_fprop(c.g, in, c.wf, out, e->stream);
	}
	if (e->selfcheck && c.use_tc) selfcheck_fprop(e, c.g, c.w, in, false, out);
}
// the main stream is about to overwrite role buffer k (0 dXe, 1 dXs, 2 dXr, 3 dXp): wait for the side-stream wgrad that still reads it
static void wait_role_readers(Engine *e, int k) {
	if (e->wstream && e->ev_rd_live[k]) RB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_rd[k], 0));
}
// role: which role buffer `dout` is (see Engine::ev_rd), -1 = none (keep-all mode allocates every gradient tensor separately)
static void conv_bwd(Engine *e, ConvRef &c, const float *in, const float *dout, float *din, int accumulate, int role = -1) {
	if (c.use_tc) {
		// instrumented passes (bench.py's per-family timing) run the wgrad in line, so that every kernel is timed alone
		const bool side = e->wstream && !prof_enabled();
		if (side) {  // dout and in are complete on the main stream here: fork the wgrad before the dgrad is enqueued
			RB_CUDA(cudaEventRecord(e->ev_fork, e->stream));
			RB_CUDA(cudaStreamWaitEvent(e->wstream, e->ev_fork, 0));
			tc_run(c.wgrad, e->wstream);
			if (role >= 0) { RB_CUDA(cudaEventRecord(e->ev_rd[role], e->wstream)); e->ev_rd_live[role] = true; }
		}
		if (din) {
			if (e->selfcheck && accumulate) selfcheck_dgrad_snapshot(e, c.g, din);
			{ ProfScope ps(e->stream, kmajor_family(c.g), conv_flops(c.g), conv_io_bytes(e, c.g)); tc_run(c.dgrad, e->stream); }
			if (e->selfcheck) selfcheck_dgrad(e, c.g, c.w, dout, din, accumulate);
		}
		if (!side) {
			ProfScope ps(e->stream, PROF_IGEMM_WGRAD, conv_flops(c.g));
			tc_run(c.wgrad, e->stream);
		}
		if (e->selfcheck) selfcheck_wgrad(e, c.g, in, false, dout, c.dw);
	} else {
		ProfScope ps(e->stream, PROF_STEM_SIMT, conv_flops(c.g) * (din ? 2 : 1));
		if (din) simt_conv_dgrad(c.g, dout, c.wd, din, accumulate, e->stream);
		simt_conv_wgrad(c.g, in, dout, c.dw, e->stream);
	}
}
// algorithmic HBM bytes of the BatchNorm / elementwise kernels, E = rows * C elements of 4 bytes:
// statistics 1E; apply 2E (+1E residual); backward reduce 2E (+1E mask) and dx 3E (+1E mask)
static double bn_bytes(const Engine *e, const BnRef &bn, double passes) { return passes * (double)e->esz * (double)bn.rows * bn.C; }

// stats_rows > 0: the producing conv's epilogue already left [stats_rows][2][C] partial sums in e->stats_partials (fused statistics);
// the fold clears them again, so the next convolution needs no memset in front of it
static void bn_forward(Engine *e, BnRef &bn, const float *x, float eps, int stats_rows = 0) {
	ProfScope ps(e->stream, PROF_BN_ELTWISE, stats_rows ? 0.0 : bn_bytes(e, bn, 1));
	if (stats_rows) bn_finalize(e->stats_partials, stats_rows, bn.rows, bn.C, bn.gamma, bn.beta, eps, bn.means, bn.vars, bn.ab, e->stream, 1);
	else bn_stats(x, bn.rows, bn.C, bn.gamma, bn.beta, eps, bn.means, bn.vars, bn.ab, e->bn_partials, e->bn_max_blocks, e->stream, e->bf16);
	if (e->keep_all && bn.cache->normalized) {
		bn_apply(x, bn.ab, bn.rows, bn.C, 0, nullptr, nullptr, bn.cache->normalized, 0, e->stream, e->bf16);
		bn_stats(x, bn.rows, bn.C, e->ones, e->zeros, eps, e->tmp_mv, e->tmp_mv + bn.C, e->tmp_ab, e->bn_partials, e->bn_max_blocks, e->stream, e->bf16);
		bn_apply(x, e->tmp_ab, bn.rows, bn.C, 0, nullptr, nullptr, bn.cache->normalized_temp, 0, e->stream, e->bf16);
	}
}
static void bn_act(Engine *e, BnRef &bn, const float *x, int relu, const float *res, const float *ab2, float *y, int rnd, uint8_t *bits_out = nullptr) {
	ProfScope ps(e->stream, PROF_BN_ELTWISE, bn_bytes(e, bn, res ? 3 : 2));
	bn_apply(x, bn.ab, bn.rows, bn.C, relu, res, ab2, y, rnd, e->stream, e->bf16, bits_out);
}
static void relu_backward(Engine *e, const float *y, const float *dy, long long n, float *dx) {
	ProfScope ps(e->stream, PROF_BN_ELTWISE, 3.0 * e->esz * (double)n);
	relu_bwd(y, dy, n, dx, e->stream, e->bf16);
}
// remask: plain BN+ReLU layer, the mask is recomputed from x (the stored activation is not read); otherwise `mask` (the block's
// output after the residual join) is read
// masked_out: also store the masked upstream gradient there (the identity shortcut's gradient, +1 E of writes)
// mask_bits: the 1-bit mask bn_apply left for this tensor (then `mask` itself is not read)
static void bn_backward(Engine *e, BnRef &bn, const float *x, const float *dy, const float *mask, float *dx, float eps, bool remask = false,
                        float *masked_out = nullptr, const uint8_t *mask_bits = nullptr) {
	const bool re = remask && (bn.C % 4 == 0) && env_int("RESNET_B200_REMASK", 1);
	ProfScope ps(e->stream, PROF_BN_ELTWISE, bn_bytes(e, bn, (re || mask_bits ? 5 : (mask ? 7 : 5)) + (masked_out ? 1 : 0)));
	bn_bwd(x, dy, mask, bn.gamma, bn.means, bn.vars, eps, bn.rows, bn.C, bn.dgamma, bn.dbeta, dx, e->bn_partials, e->bn_max_blocks, e->bn_coef,
	       e->round_tf32, e->stream, re ? bn.ab : nullptr, e->bf16, masked_out, mask_bits);
}

}  // namespace rb

using namespace rb;

// ================================================================================================ C ABI
extern "C" {

Dims *init_dimensions(int input, int init_kernel_dim, int init_conv_filters, int init_conv_stride, int init_maxpool_dim, int init_maxpool_stride,
                      int n_conv_blocks, int *is_block_spatial_reduction, int final_depth, int output) {
	Dims *d = (Dims *)malloc(sizeof(Dims));
	d->input = input; d->init_kernel_dim = init_kernel_dim; d->init_conv_filters = init_conv_filters; d->init_conv_stride = init_conv_stride;
	d->init_maxpool_dim = init_maxpool_dim; d->init_maxpool_stride = init_maxpool_stride; d->n_conv_blocks = n_conv_blocks;
	d->is_block_spatial_reduction = is_block_spatial_reduction; d->final_depth = final_depth; d->output = output;
	return d;
}

ResNet *init_resnet(Dims *dims, void *gen) {
	ResNet *m = (ResNet *)malloc(sizeof(ResNet));
	m->dims = dims;
	m->params = make_params(dims, (curandGenerator_t *)gen);
	return m;
}

Batch *init_general_batch(int n_images, int image_size, int image_dim, int shard_n_images) {
	Batch *b = (Batch *)calloc(1, sizeof(Batch));
	b->n_images = n_images; b->image_size = image_size; b->image_dim = image_dim;
	RB_CUDA(cudaMallocHost(&b->images_float_cpu, (size_t)n_images * image_size * sizeof(float)));
	RB_CUDA(cudaMalloc(&b->images, (size_t)n_images * image_size * sizeof(float)));
	RB_CUDA(cudaMallocHost(&b->correct_classes_cpu, n_images * sizeof(int)));
	RB_CUDA(cudaMalloc(&b->correct_classes, n_images * sizeof(int)));
	b->cur_shard_id = -1; b->cur_batch_in_shard = -1; b->shard_n_images = shard_n_images;
	// host staging of one shard (reference: resnet.cu:1227-1228); allocated lazily by load_new_batch
	b->full_shard_images = NULL; b->full_shard_correct_classes = NULL;
	return b;
}

Train_ResNet *init_trainer(ResNet *model, Batch *cur_batch, int batch_size, float learning_rate, float weight_decay, float mean_decay,
                           float var_decay, float eps, int n_epochs, const char *dump_dir) {
	Train_ResNet *t = (Train_ResNet *)calloc(1, sizeof(Train_ResNet));
	t->model = model; t->cur_batch = cur_batch; t->batch_size = batch_size;
	t->learning_rate = learning_rate; t->weight_decay = weight_decay; t->base_mean_decay = mean_decay; t->base_var_decay = var_decay;
	t->cur_mean_decay = 1; t->cur_var_decay = 1; t->eps = eps; t->n_epochs = n_epochs; t->cur_dump_id = -1; t->cur_epoch = 0;
	t->loss_per_epoch = (float *)calloc(n_epochs > 0 ? n_epochs : 1, sizeof(float));
	t->accuracy_per_epoch = (float *)calloc(n_epochs > 0 ? n_epochs : 1, sizeof(float));
	t->init_loaded = 0; t->dump_dir = dump_dir;
	build_engine(t);
	return t;
}

// ---- forward (reference: resnet.cu:1526-1775)
void forward_pass(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) { set_error("forward_pass: unknown trainer"); return; }
	cudaStream_t st = e->stream;
	Dims *d = t->model->dims;
	const float eps = t->eps;
	const int rnd = e->round_tf32;
	// weights may have been written through locations[] since the last step (update, checkpoint restore): re-pack
	// (tried in round 2: the re-pack on the side stream next to the stem convolution -- 0.0 / 0.0 / 0.3 % on c2 / c4 / c5, dropped)
	pack_weights(e->pack_jobs_dev, e->n_pack_jobs, e->pack_max_elems, rnd, st, e->bf16);

	stem_forward(e, t->cur_batch->images);
	bn_forward(e, e->bn0, e->X0, eps, e->stem.stats_rows);
	const int S1 = d->input / d->init_conv_stride;
	if (e->fuse_stem_tail) {
		// X0 in, pooled tensor + argmax out: init_conv_activated is never written
		ProfScope ps(st, PROF_BN_ELTWISE, (double)e->bn0.rows * e->bn0.C * (e->esz + 0.25 * (e->esz + 4)));
		bn_pool_fwd(e->X0, e->bn0.ab, e->N, S1, d->init_conv_filters, rnd, e->max_inds, e->P0, st, e->bf16);
	} else {
		bn_act(e, e->bn0, e->X0, 1, nullptr, nullptr, e->Y0, rnd);
		maxpool_fwd(e->Y0, e->N, S1, d->init_conv_filters, d->init_maxpool_dim, d->init_maxpool_stride, e->max_inds, e->P0, st, e->bf16);
	}

	for (size_t i = 0; i < e->blocks.size(); i++) {
		BlockRef &b = e->blocks[i];
		Activation_ConvBlock *ab = t->forward_buffer->activations->activation_conv_blocks[i];
		conv_fwd(e, b.reduce, b.x_in, b.Xr);
		bn_forward(e, b.bn_r, b.Xr, eps, b.reduce.stats_rows);
		bn_act(e, b.bn_r, b.Xr, 1, nullptr, nullptr, b.Yr, rnd);
		conv_fwd(e, b.spatial, b.Yr, b.Xs);
		bn_forward(e, b.bn_s, b.Xs, eps, b.spatial.stats_rows);
		bn_act(e, b.bn_s, b.Xs, 1, nullptr, nullptr, b.Ys, rnd);
		conv_fwd(e, b.expand, b.Ys, b.Xe);
		bn_forward(e, b.bn_e, b.Xe, eps, b.expand.stats_rows);
		if (b.has_proj) {
			conv_fwd(e, b.proj, b.x_in, b.Xp);
			bn_forward(e, b.bn_p, b.Xp, eps, b.proj.stats_rows);
		}
		if (e->keep_all) {
			bn_apply(b.Xe, b.bn_e.ab, b.bn_e.rows, b.bn_e.C, 0, nullptr, nullptr, ab->post_expanded_norm_vals, 0, st, e->bf16);
			if (b.has_proj) bn_apply(b.Xp, b.bn_p.ab, b.bn_p.rows, b.bn_p.C, 0, nullptr, nullptr, ab->post_projection_norm_vals, 0, st, e->bf16);
			bn_apply(b.Xe, b.bn_e.ab, b.bn_e.rows, b.bn_e.C, 0, b.has_proj ? b.Xp : b.x_in, b.has_proj ? b.bn_p.ab : nullptr, ab->output, 0, st, e->bf16);
		}
		// output_activated = relu(bn(expanded) + shortcut)   (reference: resnet.cu:1670-1723, four kernels there)
		bn_act(e, b.bn_e, b.Xe, 1, b.has_proj ? b.Xp : b.x_in, b.has_proj ? b.bn_p.ab : nullptr, b.OA, rnd, b.oa_bits);
	}
	BlockRef &last = e->blocks.back();
	const int Sl = last.expand.g.S;
	avgpool_fwd(last.OA, e->N, Sl, d->final_depth, e->pooled, st, e->bf16);
	sgemm(e->pooled, t->model->params->fully_connected, e->logits, e->N, d->output, d->final_depth, 0, 0, st, e->fc_ws);
	softmax_ce(e->logits, t->cur_batch->correct_classes, e->N, d->output, e->pred, nullptr, e->row_loss, e->row_wrong, st);
	epoch_accumulate_kernel<<<1, 256, 0, st>>>(e->row_loss, e->row_wrong, e->N, e->epoch_acc);
	RB_LAUNCH_CHECK();
	e->epoch_images += e->N;
	if (e->pred_copy) {
		RB_CUDA(cudaMemcpyAsync(e->pred_host, e->pred, (size_t)e->N * d->output * sizeof(float), cudaMemcpyDeviceToHost, st));
		RB_CUDA(cudaStreamSynchronize(st));  // pred_cpu is valid on return, as in the reference (resnet.cu:1774)
	}
}

// ---- backward (reference: resnet.cu:1777-2248; spatial-BN call per resnet_clean.cu:2778)
void backwards_pass(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) { set_error("backwards_pass: unknown trainer"); return; }
	cudaStream_t st = e->stream;
	Dims *d = t->model->dims;
	const float eps = t->eps;
	Params *G = t->backprop_buffer->param_derivs;
	const int N = e->N;
	{
		long long total = (long long)N * d->output;
		int grid = (int)((total + 255) / 256);
		ce_deriv_kernel<<<grid, 256, 0, st>>>(e->pred, t->cur_batch->correct_classes, N, d->output, e->dlogits);
		RB_LAUNCH_CHECK();
	}
	// dW_fc = pooled^T . dlogits ; dpooled = dlogits . W_fc^T   (reference: resnet.cu:1823, 1830)
	sgemm(e->pooled, e->dlogits, G->fully_connected, d->final_depth, d->output, N, 1, 0, st);
	sgemm(e->dlogits, t->model->params->fully_connected, e->dpooled, N, d->final_depth, d->output, 0, 1, st, e->fc_ws);
	BlockRef &last = e->blocks.back();
	avgpool_bwd(e->dpooled, N, last.expand.g.S, d->final_depth, last.dOA, st, e->bf16);

	for (int i = (int)e->blocks.size() - 1; i >= 0; i--) {
		BlockRef &b = e->blocks[i];
		if (e->keep_all) {
			Activation_ConvBlock *db = t->backprop_buffer->activation_derivs->activation_conv_blocks[i];
			relu_bwd(b.OA, b.dOA, b.n_exp_out, db->output, st, e->bf16);  // d(output), reference: resnet.cu:1934
		}
		// shortcut branch first (its scratch is reused by the expanded branch)
		// identity shortcut: its gradient relu'(OA) * dOA (reference: resnet.cu:2003-2004) is stored by the expansion BatchNorm's
		// backward, which has it in registers, unless RESNET_B200_FUSE_SHORTCUT=0 asks for the separate pass
		const bool fuse_short = !b.has_proj && env_int("RESNET_B200_FUSE_SHORTCUT", 1);
		// (wait_role_readers: in the default buffer plan dXe / dXs / dXr / dXp are one scratch tensor each, shared by all blocks; a
		// wgrad of the previous block may still be reading it on the side stream)
		if (b.has_proj) {
			wait_role_readers(e, 3);
			bn_backward(e, b.bn_p, b.Xp, b.dOA, b.OA, b.dXp, eps, false, nullptr, b.oa_bits);
			conv_bwd(e, b.proj, b.x_in, b.dXp, b.dBI, 0, 3);
		} else if (!fuse_short) {
			relu_backward(e, b.OA, b.dOA, b.n_exp_out, b.dBI);
		}
		wait_role_readers(e, 0);
		bn_backward(e, b.bn_e, b.Xe, b.dOA, b.OA, b.dXe, eps, false, fuse_short ? b.dBI : nullptr, b.oa_bits);
		wait_role_readers(e, 1);  // the expansion's dgrad writes dYs, which shares its tensor with dXs
		conv_bwd(e, b.expand, b.Ys, b.dXe, b.dYs, 0, 0);
		bn_backward(e, b.bn_s, b.Xs, b.dYs, b.Ys, b.dXs, eps, true);
		wait_role_readers(e, 2);  // ... and the 3x3's dgrad writes dYr = dXr's tensor
		conv_bwd(e, b.spatial, b.Yr, b.dXs, b.dYr, 0, 1);
		bn_backward(e, b.bn_r, b.Xr, b.dYr, b.Yr, b.dXr, eps, true);
		conv_bwd(e, b.reduce, b.x_in, b.dXr, b.dBI, 1, 2);
		dp_block_done(e, i);
	}
	const int S1 = d->input / d->init_conv_stride;
	if (e->fuse_stem_tail) {
		// two passes over (X0, dP0, max_inds); the second writes dX0: the pool's input gradient is never written
		ProfScope ps(st, PROF_BN_ELTWISE, (double)e->bn0.rows * e->bn0.C * (3.0 * e->esz + 2 * 0.25 * (e->esz + 4)));
		pool_bn_bwd(e->max_inds, e->dP0, e->X0, e->bn0.gamma, e->bn0.means, e->bn0.vars, eps, e->bn0.ab, N, S1, d->init_conv_filters, e->bn0.dgamma,
		            e->bn0.dbeta, e->dX0, e->bn_partials, e->bn_max_blocks, e->bn_coef, e->stem_tc ? e->round_tf32 : 0, st, e->bf16);
	} else {
		maxpool_bwd(e->max_inds, e->dP0, N, S1, d->init_conv_filters, d->init_maxpool_dim, d->init_maxpool_stride, e->dY0, st, e->bf16);
		ProfScope ps(st, PROF_BN_ELTWISE, bn_bytes(e, e->bn0, 5));
		bn_bwd(e->X0, e->dY0, e->Y0, e->bn0.gamma, e->bn0.means, e->bn0.vars, eps, e->bn0.rows, e->bn0.C, e->bn0.dgamma, e->bn0.dbeta, e->dX0,
		       e->bn_partials, e->bn_max_blocks, e->bn_coef, e->stem_tc ? e->round_tf32 : 0, st, e->bn0.ab, e->bf16);
	}
	stem_backward(e, t->cur_batch->images);
	if (e->wstream) {  // join: every weight gradient is complete before anything later on the main stream (allreduce, Adam, a host read)
		RB_CUDA(cudaEventRecord(e->ev_join, e->wstream));
		RB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
	}
	dp_allreduce_grads(e);
}

// ---- Adam (reference: resnet.cu:2910-2987)
void update_parameters(Train_ResNet *t) {
	Engine *e = engine_of(t);
	if (!e) { set_error("update_parameters: unknown trainer"); return; }
	cudaStream_t st = e->stream;
	const float cur_mean_decay = t->cur_mean_decay * t->base_mean_decay;
	const float cur_var_decay = t->cur_var_decay * t->base_var_decay;
	ParamStore *p = param_store_of(t->model->params), *g = param_store_of(t->backprop_buffer->param_derivs);
	ParamStore *m = param_store_of(t->backprop_buffer->prev_means), *v = param_store_of(t->backprop_buffer->prev_vars);
	// Checkpoint cadence of the reference: a full dump before the update whenever cur_dump_id % 1000 == 0 (resnet.cu:2941-2944;
	// cur_dump_id counts load_new_batch calls from 0, and stays -1 -- no dump -- for a host that stages batches itself).
	// RESNET_B200_DUMP_EVERY overrides the period, 0 switches the periodic dump off.
	const int dump_every = env_int("RESNET_B200_DUMP_EVERY", 1000);
	if (dump_every > 0 && t->cur_dump_id >= 0 && t->cur_dump_id % dump_every == 0) {
		printf("DUMPING TRAINER...!\n\n");
		dump_trainer(t->cur_dump_id, t, t->dump_dir);
	}
	if (*e->bad_host) {
		// the reference scans all four arenas on the host every step, dumps to id 99999999 and exit(1)s on the first NaN/Inf
		// (resnet.cu:2879-2900); here the previous update's Adam kernel counted them on the device
		const int bad = *e->bad_host;
		fprintf(stderr, "ERROR: nan or inf found in %d parameter/gradient entries during the previous update\n", bad);
		if (env_int("RESNET_B200_DUMP_ON_NAN", 1)) {
			printf("Dumping data to id=99999999...\n");
			dump_trainer(99999999, t, t->dump_dir);
		}
		set_error("non-finite values in update_parameters (%d entries)", bad);
		if (env_int("RESNET_B200_EXIT_ON_NAN", 0)) exit(1);
		// reported once: the counter restarts, so later updates only complain about NEW non-finite values
		*e->bad_host = 0;
		RB_CUDA(cudaMemsetAsync(e->bad_dev, 0, sizeof(int), st));
	}
	adam_step(p->base, g->base, m->base, v->base, p->total, t->learning_rate, t->weight_decay, t->base_mean_decay, t->base_var_decay,
	          cur_mean_decay, cur_var_decay, t->eps, e->bad_dev, st);
	RB_CUDA(cudaMemcpyAsync(e->bad_host, e->bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
	// reference resets the batch buffers before the next load (resnet.cu:2981-2982)
	Batch *b = t->cur_batch;
	RB_CUDA(cudaMemsetAsync(b->images, 0, (size_t)t->batch_size * b->image_size * sizeof(float), st));
	RB_CUDA(cudaMemsetAsync(b->correct_classes, 0, (size_t)t->batch_size * sizeof(int), st));
	t->cur_mean_decay = cur_mean_decay;
	t->cur_var_decay = cur_var_decay;
}

// ---- class metadata (reference: resnet.cu:1331-1381)
static int read_lines(const char *filename, char **text, int *ints, int n) {
	FILE *fp = fopen(filename, "r");
	if (!fp) return -1;
	char *line = NULL;
	size_t len = 0;
	int cnt = 0;
	while (cnt < n && getline(&line, &len, fp) != -1) {
		if (text) text[cnt] = strdup(line);
		if (ints) ints[cnt] = atoi(line);
		cnt++;
	}
	free(line);
	fclose(fp);
	return cnt;
}
Class_Metadata *populate_class_info(char *label_filename, char *synset_filename, char *class_size_filename, int n_classes) {
	Class_Metadata *c = (Class_Metadata *)malloc(sizeof(Class_Metadata));
	c->labels = (char **)calloc(n_classes, sizeof(char *));
	c->synsets = (char **)calloc(n_classes, sizeof(char *));
	c->counts = (int *)calloc(n_classes, sizeof(int));
	c->n_classes = n_classes;
	if (read_lines(label_filename, c->labels, NULL, n_classes) < 0 || read_lines(synset_filename, c->synsets, NULL, n_classes) < 0 ||
	    read_lines(class_size_filename, NULL, c->counts, n_classes) < 0) {
		fprintf(stderr, "populate_class_info: cannot open metadata file\n");
		exit(EXIT_FAILURE);  // reference: resnet.cu:1341-1342
	}
	return c;
}

}  // extern "C"
