// loader.cu -- load_new_batch (reference: resnet.cu:1235-1325) with the same shard files and traversal, made asynchronous.
//
// Format (reference: build_training_shards.c:120-176): `<dir>/%03d.images` = shard_n_images x image_size raw
// little-endian fp32 (NHWC, mean-subtracted), `<dir>/%03d.labels` = shard_n_images x int32.  Traversal: batches of a shard
// in order, next shard when cur_batch_in_shard * batch >= shard_n_images; cur_dump_id++ per call; init_loaded re-opens the
// checkpointed shard (reference: resnet.cu:1266-1295).  <dir> = $RESNET_B200_SHARD_DIR or the reference's hard-coded path.
//
// The reference fread()s a whole 19.7 GB shard into pageable memory, memcpy()s the batch into pinned memory and does a
// blocking cudaMemcpy.  Here a background thread pread()s only the NEXT batch straight into a second pinned buffer and
// enqueues its H2D copy on a copy stream while the current step computes; load_new_batch then just waits for that event,
// does a device-to-device copy on the compute stream and swaps the pinned host buffers (images_float_cpu /
// correct_classes_cpu keep their meaning: host copy of the current batch).
#include "engine.h"
#include <condition_variable>
#include <fcntl.h>
#include <map>
#include <mutex>
#include <thread>
#include <unistd.h>

namespace rb {

struct Prefetcher {
	std::thread th;
	std::mutex mu;
	std::condition_variable cv;
	bool stop = false, busy = false;
	int req_shard = -1, req_batch = -1;      // what the worker should fetch
	int have_shard = -1, have_batch = -1;    // what the slot holds (valid when ok)
	bool ok = false;
	float *img_pinned = nullptr, *img_dev = nullptr;
	int *lab_pinned = nullptr, *lab_dev = nullptr;
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ready = nullptr;
	size_t img_bytes = 0, lab_bytes = 0;
	int batch_size = 0, image_size = 0, device = 0;
};
static std::map<Batch *, Prefetcher *> g_prefetch;
static std::mutex g_pf_mu;

static const char *shard_dir() {
	const char *dir = getenv("RESNET_B200_SHARD_DIR");
	return dir ? dir : "/mnt/storage/data/vision/imagenet/2012/train_data_shards";
}

// reads batch `b` of shard `s` into host buffers; false when the files are missing / short
static bool read_batch(int s, int b, int batch_size, int image_size, float *img, int *lab) {
	char path[1024];
	snprintf(path, sizeof(path), "%s/%03d.images", shard_dir(), s);
	int fd = open(path, O_RDONLY);
	if (fd < 0) return false;
	const size_t nb = (size_t)batch_size * image_size * sizeof(float);
	size_t got = 0;
	while (got < nb) {
		ssize_t r = pread(fd, (char *)img + got, nb - got, (off_t)((size_t)b * nb + got));
		if (r <= 0) break;
		got += (size_t)r;
	}
	close(fd);
	if (got != nb) return false;
	snprintf(path, sizeof(path), "%s/%03d.labels", shard_dir(), s);
	fd = open(path, O_RDONLY);
	if (fd < 0) return false;
	const size_t lb = (size_t)batch_size * sizeof(int);
	ssize_t r = pread(fd, lab, lb, (off_t)((size_t)b * lb));
	close(fd);
	return r == (ssize_t)lb;
}

static void worker(Prefetcher *p) {
	cudaSetDevice(p->device);
	std::unique_lock<std::mutex> lk(p->mu);
	for (;;) {
		p->cv.wait(lk, [&] { return p->stop || p->busy; });
		if (p->stop) return;
		const int s = p->req_shard, b = p->req_batch;
		lk.unlock();
		// The pinned buffer about to be overwritten may still be the source of an H2D copy queued on copy_stream (a host that runs
		// several steps ahead, e.g. with set_pred_copy(0), never waits for those copies; neither does a prefetch miss, which leaves
		// the buffers unswapped): drain the copy stream first.  It also holds the wait on the compute stream's D2D out of the
		// staging device buffers, so this thread stays at most one batch ahead of the GPU.
		cudaStreamSynchronize(p->copy_stream);
		bool ok = read_batch(s, b, p->batch_size, p->image_size, p->img_pinned, p->lab_pinned);
		if (ok) {
			cudaMemcpyAsync(p->img_dev, p->img_pinned, p->img_bytes, cudaMemcpyHostToDevice, p->copy_stream);
			cudaMemcpyAsync(p->lab_dev, p->lab_pinned, p->lab_bytes, cudaMemcpyHostToDevice, p->copy_stream);
			cudaEventRecord(p->ready, p->copy_stream);
		}
		lk.lock();
		p->ok = ok; p->have_shard = s; p->have_batch = b; p->busy = false;
		p->cv.notify_all();
	}
}

static Prefetcher *prefetcher_of(Batch *bb) {
	std::lock_guard<std::mutex> g(g_pf_mu);
	auto it = g_prefetch.find(bb);
	if (it != g_prefetch.end()) return it->second;
	Prefetcher *p = new Prefetcher();
	p->batch_size = bb->n_images; p->image_size = bb->image_size;
	p->img_bytes = (size_t)bb->n_images * bb->image_size * sizeof(float);
	p->lab_bytes = (size_t)bb->n_images * sizeof(int);
	cudaGetDevice(&p->device);
	RB_CUDA(cudaMallocHost(&p->img_pinned, p->img_bytes));
	RB_CUDA(cudaMallocHost(&p->lab_pinned, p->lab_bytes));
	RB_CUDA(cudaMalloc(&p->img_dev, p->img_bytes));
	RB_CUDA(cudaMalloc(&p->lab_dev, p->lab_bytes));
	RB_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
	RB_CUDA(cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming));
	p->th = std::thread(worker, p);
	g_prefetch[bb] = p;
	return p;
}

// stops and joins the prefetch thread of `bb` and frees its buffers (resnet_b200_destroy_trainer)
void loader_release(Batch *bb) {
	Prefetcher *p = nullptr;
	{
		std::lock_guard<std::mutex> g(g_pf_mu);
		auto it = g_prefetch.find(bb);
		if (it == g_prefetch.end()) return;
		p = it->second;
		g_prefetch.erase(it);
	}
	{
		std::lock_guard<std::mutex> lk(p->mu);
		p->stop = true;
	}
	p->cv.notify_all();
	if (p->th.joinable()) p->th.join();
	cudaStreamSynchronize(p->copy_stream);
	cudaFreeHost(p->img_pinned); cudaFreeHost(p->lab_pinned);
	cudaFree(p->img_dev); cudaFree(p->lab_dev);
	cudaEventDestroy(p->ready);
	cudaStreamDestroy(p->copy_stream);
	delete p;
}

// the reference's cursor step (resnet.cu:1260-1295): next batch of the shard, next shard when the shard is exhausted
static void advance_cursor(int &shard, int &batch, int batch_size, int shard_n_images) {
	batch += 1;
	if (batch * batch_size >= shard_n_images) { shard += 1; batch = 0; }
}

// Which (shard, batch) a load_new_batch call delivers, given the cursor the previous call left (the reference's convention:
// cur_batch_in_shard = batch after the one delivered last; -1 / -1 before the first call) -- pure host logic, shared by
// load_new_batch and the CPU test hook resnet_b200_loader_plan.
static void next_delivery(int &shard, int &batch, bool init_loaded, int rank, int world, int batch_size, int shard_n_images) {
	if (init_loaded) return;  // a restored cursor is rank-local and is taken as is (reference: resnet.cu:1266-1295)
	if (shard == -1) {
		shard = 0; batch = 0;
		for (int i = 0; i < rank; i++) advance_cursor(shard, batch, batch_size, shard_n_images);
		return;
	}
	if (batch * batch_size >= shard_n_images) { shard += 1; batch = 0; }
	for (int i = 1; i < world; i++) advance_cursor(shard, batch, batch_size, shard_n_images);
}

}  // namespace rb

using namespace rb;

// host-only view of the traversal for tests: the (shard, batch) pairs `n_calls` consecutive load_new_batch calls deliver on
// rank `rank` of `world`, starting from a fresh cursor
extern "C" int resnet_b200_loader_plan(int rank, int world, int batch_size, int shard_n_images, int n_calls, int *out_shard, int *out_batch) {
	int shard = -1, batch = -1;
	for (int i = 0; i < n_calls; i++) {
		next_delivery(shard, batch, false, rank, world, batch_size, shard_n_images);
		out_shard[i] = shard; out_batch[i] = batch;
		batch += 1;  // what load_new_batch stores in cur_batch_in_shard
	}
	return 0;
}

extern "C" void load_new_batch(Train_ResNet *trainer, Class_Metadata *class_metadata, Batch *bb) {
	(void)class_metadata;
	Engine *e = engine_of(trainer);
	cudaStream_t st = e ? e->stream : 0;
	const int batch_size = bb->n_images;
	// which (shard, batch) this call delivers -- the reference's traversal (resnet.cu:1260-1295)
	// Data parallel (dp.cu): the replicas share ONE global batch sequence -- the single-GPU traversal -- and rank r of `world` takes
	// positions r, r + world, r + 2 world, ...: the first call skips `rank` batches, every later call advances by `world`
	// (SURVEY.md 8e "rank r reads batches r, r+G, ...").  A restored cursor (init_loaded) is rank-local and is taken as is.
	int rank = 0, world = 1;
	if (e) dp_rank_world(e, &rank, &world);
	int shard = bb->cur_shard_id, batch = bb->cur_batch_in_shard;
	next_delivery(shard, batch, trainer->init_loaded != 0, rank, world, batch_size, bb->shard_n_images);
	trainer->init_loaded = 0;
	Prefetcher *p = prefetcher_of(bb);
	bool delivered = false;
	{
		std::unique_lock<std::mutex> lk(p->mu);
		p->cv.wait(lk, [&] { return !p->busy; });
		if (p->ok && p->have_shard == shard && p->have_batch == batch) {
			// prefetched: order the compute stream after the H2D copy, copy device-to-device, swap the pinned host buffers
			RB_CUDA(cudaStreamWaitEvent(st, p->ready, 0));
			RB_CUDA(cudaMemcpyAsync(bb->images, p->img_dev, p->img_bytes, cudaMemcpyDeviceToDevice, st));
			RB_CUDA(cudaMemcpyAsync(bb->correct_classes, p->lab_dev, p->lab_bytes, cudaMemcpyDeviceToDevice, st));
			std::swap(bb->images_float_cpu, p->img_pinned);
			std::swap(bb->correct_classes_cpu, p->lab_pinned);
			// the staging device buffers are read by the D2D copies above: the next H2D into them must come after
			RB_CUDA(cudaEventRecord(p->ready, st));
			RB_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ready, 0));
			p->ok = false;
			delivered = true;
		}
	}
	if (!delivered) {  // first call, resume, or prefetch miss: synchronous path (what the reference always does)
		if (!read_batch(shard, batch, batch_size, bb->image_size, bb->images_float_cpu, bb->correct_classes_cpu)) {
			set_error("load_new_batch: cannot read batch %d of shard %03d under %s", batch, shard, shard_dir());
			return;
		}
		RB_CUDA(cudaMemcpyAsync(bb->images, bb->images_float_cpu, p->img_bytes, cudaMemcpyHostToDevice, st));
		RB_CUDA(cudaMemcpyAsync(bb->correct_classes, bb->correct_classes_cpu, p->lab_bytes, cudaMemcpyHostToDevice, st));
		RB_CUDA(cudaStreamSynchronize(st));
	}
	bb->cur_shard_id = shard;
	bb->cur_batch_in_shard = batch + 1;
	trainer->cur_dump_id += 1;
	// kick off the prefetch of the batch the NEXT call will ask for
	int nshard = shard, nbatch = batch;
	for (int i = 0; i < world; i++) advance_cursor(nshard, nbatch, batch_size, bb->shard_n_images);
	{
		std::lock_guard<std::mutex> lk(p->mu);
		p->req_shard = nshard; p->req_batch = nbatch; p->busy = true; p->ok = false;
	}
	p->cv.notify_all();
}
