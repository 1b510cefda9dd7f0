// dump.cu -- the reference's training-dump / checkpoint directory format: dump_trainer,
// overwrite_trainer_hyperparams, overwrite_model_params.
//
// reference: resnet.cu:2250-2316 (dump_parameters), 2319-2349 (dump_batch_norm_cache), 2351-2513 (dump_conv_block_activation),
// 2515-2680 (dump_activations), 2682-2753 (trainer_metadata.txt / trainer_checkpoint.txt), 2755-2772 (dump_trainer),
// 2778-2875 (restore).  Same relative paths, file names, element order and text-line order, so the reference's
// analyze_trainer_dump.ipynb and its own overwrite_* read our dumps and we read its dumps.
//
// Differences that are deliberate:
//  * the root directory is $RESNET_B200_DUMP_ROOT (default: the reference's hard-coded
//    /mnt/storage/data/vision/imagenet/training_dumps) and missing directories are created (the reference fopen()s into
//    directories it assumes exist and dereferences the NULL FILE* otherwise);
//  * files are always fp32, whatever the trainer stores on the device: bf16 tensors are widened on the way out;
//  * buffers the trainer does not materialise outside keep-all mode (x-hat caches, pre-ReLU sums: NULL pointers in the public
//    structs, as in the reference's own resnet_clean.h) are skipped;
//  * failures are recorded in resnet_b200_last_error() instead of crashing.
#include "engine.h"
#include <errno.h>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace rb {

static std::string dump_root() {
	const char *r = getenv("RESNET_B200_DUMP_ROOT");
	return std::string(r && *r ? r : "/mnt/storage/data/vision/imagenet/training_dumps");
}
static bool mkdirs(const std::string &path) {
	for (size_t i = 1; i <= path.size(); i++) {
		if (i == path.size() || path[i] == '/') {
			std::string sub = path.substr(0, i);
			if (mkdir(sub.c_str(), 0777) != 0 && errno != EEXIST) { set_error("dump: cannot create directory %s (%s)", sub.c_str(), strerror(errno)); return false; }
		}
	}
	return true;
}
static std::string dump_dir_of(int dump_id, const char *special_dir) {
	char id[16];
	snprintf(id, sizeof(id), "%08d", dump_id);
	return dump_root() + "/" + (special_dir ? special_dir : "") + "/" + id + "/";
}

// n elements from device memory to `path` as fp32 (or int32); act = the tensor is stored in the engine's activation type
static void write_buffer(Engine *e, const std::string &path, const void *dev, size_t n, bool act, bool is_int = false) {
	if (!dev) return;  // not materialised in this mode
	const bool bf = act && e->bf16;
	std::vector<float> host(n);
	if (bf) {
		std::vector<uint16_t> raw(n);
		RB_CUDA(cudaMemcpy(raw.data(), dev, n * sizeof(uint16_t), cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < n; i++) { uint32_t u = (uint32_t)raw[i] << 16; memcpy(&host[i], &u, 4); }
	} else {
		RB_CUDA(cudaMemcpy(host.data(), dev, n * 4, cudaMemcpyDeviceToHost));  // float and int are both 4 bytes
	}
	(void)is_int;
	FILE *fp = fopen(path.c_str(), "wb");
	if (!fp) { set_error("dump: cannot open %s for writing (%s)", path.c_str(), strerror(errno)); return; }
	if (fwrite(host.data(), 4, n, fp) != n) set_error("dump: short write to %s", path.c_str());
	fclose(fp);
}
static void read_buffer(const std::string &path, float *dev, size_t n) {
	std::vector<float> host(n);
	FILE *fp = fopen(path.c_str(), "rb");
	if (!fp) { set_error("restore: cannot open %s (%s)", path.c_str(), strerror(errno)); return; }
	const size_t got = fread(host.data(), 4, n, fp);
	fclose(fp);
	if (got != n) { set_error("restore: %s holds %zu of %zu floats", path.c_str(), got, n); return; }
	RB_CUDA(cudaMemcpy(dev, host.data(), n * 4, cudaMemcpyHostToDevice));
}

// reference: resnet.cu:2250-2316
static void dump_parameters(Engine *e, const std::string &root) {
	Train_ResNet *t = e->trainer;
	Params *trees[4] = {t->model->params, t->backprop_buffer->param_derivs, t->backprop_buffer->prev_means, t->backprop_buffer->prev_vars};
	const char *names[4] = {"model_params", "gradients", "means", "vars"};
	for (int k = 0; k < 4; k++) {
		if (!mkdirs(root + names[k])) return;
		for (int i = trees[k]->n_locations - 1; i >= 0; i--) {
			char f[32];
			snprintf(f, sizeof(f), "/%03d.buffer", i);
			write_buffer(e, root + names[k] + f, trees[k]->locations[i], (size_t)t->model->params->sizes[i], false);
		}
	}
}
// reference: resnet.cu:2319-2349 (means.buffer, vars.buffer)
static void dump_bn_cache(Engine *e, const std::string &dir, Cache_BatchNorm *c) {
	if (!c || !c->means || !mkdirs(dir)) return;
	write_buffer(e, dir + "means.buffer", c->means, (size_t)c->feature_size, false);
	write_buffer(e, dir + "vars.buffer", c->vars, (size_t)c->feature_size, false);
}
// reference: resnet.cu:2351-2513
static void dump_block(Engine *e, const std::string &root, Activation_ConvBlock *b, int ind, bool deriv) {
	char nn[16];
	snprintf(nn, sizeof(nn), "%02d/", ind);
	const std::string sub = deriv ? "activation_derivs/" : "activations/";
	const std::string dir = root + sub + "conv_blocks/" + nn, bn = root + sub + "batch_norms/" + nn;
	if (!mkdirs(dir)) return;
	const size_t N = (size_t)e->N, S = (size_t)b->incoming_spatial_dim, st = (size_t)b->stride;
	const size_t red = S * S * b->reduced_depth * N, spa = red / (st * st), exp_ = S * S * b->expanded_depth * N / (st * st);
	write_buffer(e, dir + "reduction_applied.buffer", b->post_reduced, red, true);
	dump_bn_cache(e, bn + "reduced/", b->norm_post_reduced);
	write_buffer(e, dir + "reduction_activated.buffer", b->post_reduced_activated, red, true);
	write_buffer(e, dir + "spatial_applied.buffer", b->post_spatial, spa, true);
	dump_bn_cache(e, bn + "spatial/", b->norm_post_spatial);
	write_buffer(e, dir + "spatial_activated.buffer", b->post_spatial_activated, spa, true);
	write_buffer(e, dir + "expanded_applied.buffer", b->post_expanded, exp_, true);
	dump_bn_cache(e, bn + "expanded/", b->norm_post_expanded);
	write_buffer(e, dir + "expanded_post_norm.buffer", b->post_expanded_norm_vals, exp_, true);
	if (b->transformed_residual) {
		write_buffer(e, dir + "transformed_residual.buffer", b->transformed_residual, exp_, true);
		dump_bn_cache(e, bn + "projected/", b->norm_post_projection);
	}
	write_buffer(e, dir + "combined_output.buffer", b->output, exp_, true);
	write_buffer(e, dir + "output_activated.buffer", b->output_activated, exp_, true);
}
// reference: resnet.cu:2515-2680
static void dump_activations(Engine *e, const std::string &root, Activations *A, bool deriv) {
	Train_ResNet *t = e->trainer;
	Dims *d = t->model->dims;
	const std::string dir = root + (deriv ? "activation_derivs/" : "activations/");
	if (!mkdirs(dir)) return;
	const size_t N = (size_t)e->N;
	if (!deriv) write_buffer(e, dir + "input.buffer", t->cur_batch->images, (size_t)t->cur_batch->image_size * N, false);
	const size_t S1 = (size_t)(d->input / d->init_conv_stride);
	const size_t n0 = N * d->init_conv_filters * S1 * S1, np = n0 / ((size_t)d->init_maxpool_stride * d->init_maxpool_stride);
	write_buffer(e, dir + "init_conv_applied.buffer", A->init_conv_applied, n0, true);
	dump_bn_cache(e, dir + "batch_norms/init/", A->norm_init_conv);
	write_buffer(e, dir + "init_conv_activated.buffer", A->init_conv_activated, n0, true);
	if (!deriv) write_buffer(e, dir + "max_inds.buffer", A->max_inds, np, false, true);
	write_buffer(e, dir + "init_convblock_input.buffer", A->init_convblock_input, np, true);
	for (int i = 0; i < A->n_conv_blocks; i++) dump_block(e, root, A->activation_conv_blocks[i], i, deriv);
	write_buffer(e, dir + "final_avg_pool.buffer", A->final_conv_output_pooled, N * d->final_depth, false);
	write_buffer(e, dir + "fc_output.buffer", A->linear_output, N * d->output, false);
	write_buffer(e, dir + "softmax.buffer", deriv ? t->backprop_buffer->output_layer_deriv : t->forward_buffer->pred, N * d->output, false);
	if (!deriv) write_buffer(e, dir + "correct_classes.buffer", t->cur_batch->correct_classes, N, false, true);
}
// reference: resnet.cu:2682-2731
static void dump_meta(Engine *e, const std::string &root) {
	Train_ResNet *t = e->trainer;
	FILE *fp = fopen((root + "trainer_metadata.txt").c_str(), "w");
	if (!fp) { set_error("dump: cannot open %strainer_metadata.txt", root.c_str()); return; }
	fprintf(fp, "%d\n%d\n%d\n%d\n", t->batch_size, t->cur_batch->image_size, t->cur_batch->image_dim, t->cur_batch->shard_n_images);
	fprintf(fp, "%f\n%f\n%f\n%f\n%f\n%f\n%f\n", t->learning_rate, t->weight_decay, t->base_mean_decay, t->base_var_decay, t->cur_mean_decay,
	        t->cur_var_decay, t->eps);
	fprintf(fp, "%d\n%d\n%d\n", t->n_epochs, t->cur_dump_id, t->cur_epoch);
	for (int i = 0; i < t->cur_epoch; i++) fprintf(fp, i ? ",%f" : "%f", t->loss_per_epoch[i]);
	fprintf(fp, "\n");
	for (int i = 0; i < t->cur_epoch; i++) fprintf(fp, i ? ",%f" : "%f", t->accuracy_per_epoch[i]);
	fprintf(fp, "\n");
	fclose(fp);
}
// reference: resnet.cu:2733-2753.  The reference prints the two decay products with %f (6 decimals), which rounds
// beta2^t to 1.000000 for the first ~500 steps; we print them with %.9g so a restore reproduces the bias correction -- atof
// in the reference's own overwrite_trainer_hyperparams reads either form.
static void dump_checkpoint(Engine *e, const std::string &root) {
	Train_ResNet *t = e->trainer;
	FILE *fp = fopen((root + "trainer_checkpoint.txt").c_str(), "w");
	if (!fp) { set_error("dump: cannot open %strainer_checkpoint.txt", root.c_str()); return; }
	fprintf(fp, "%d\n%d\n", t->cur_batch->cur_shard_id, t->cur_batch->cur_batch_in_shard);
	fprintf(fp, "%.9g\n%.9g\n", t->cur_mean_decay, t->cur_var_decay);
	fprintf(fp, "%d\n%d\n", t->cur_dump_id, t->cur_epoch);
	fclose(fp);
}

}  // namespace rb

using namespace rb;

extern "C" {

// reference: resnet.cu:2755-2772
void dump_trainer(int dump_id, Train_ResNet *t, const char *special_dir) {
	Engine *e = engine_of(t);
	if (!e) { set_error("dump_trainer: unknown trainer"); return; }
	RB_CUDA(cudaStreamSynchronize(e->stream));
	const std::string root = dump_dir_of(dump_id, special_dir);
	if (!mkdirs(root)) return;
	dump_parameters(e, root);
	dump_activations(e, root, t->forward_buffer->activations, false);
	dump_activations(e, root, t->backprop_buffer->activation_derivs, true);
	dump_meta(e, root);
	dump_checkpoint(e, root);
}

// reference: resnet.cu:2778-2817 (line order fixed by dump_trainer_checkpoint)
void overwrite_trainer_hyperparams(Train_ResNet *t, int dump_id, const char *special_dir) {
	const std::string path = dump_dir_of(dump_id, special_dir) + "trainer_checkpoint.txt";
	FILE *fp = fopen(path.c_str(), "r");
	if (!fp) { set_error("restore: cannot open %s (%s)", path.c_str(), strerror(errno)); return; }
	char line[256];
	double vals[6] = {0, 0, 0, 0, 0, 0};
	int n = 0;
	while (n < 6 && fgets(line, sizeof(line), fp)) vals[n++] = atof(line);
	fclose(fp);
	if (n != 6) { set_error("restore: %s has %d of 6 lines", path.c_str(), n); return; }
	t->cur_batch->cur_shard_id = (int)vals[0];
	t->cur_batch->cur_batch_in_shard = (int)vals[1];
	t->cur_mean_decay = (float)vals[2];
	t->cur_var_decay = (float)vals[3];
	t->cur_dump_id = (int)vals[4];
	t->cur_epoch = (int)vals[5];
	t->init_loaded = 1;
}

// reference: resnet.cu:2821-2875 (model_params, means, vars; gradients are not restored)
void overwrite_model_params(Train_ResNet *t, int dump_id, const char *special_dir) {
	Engine *e = engine_of(t);
	if (!e) { set_error("overwrite_model_params: unknown trainer"); return; }
	RB_CUDA(cudaStreamSynchronize(e->stream));
	const std::string root = dump_dir_of(dump_id, special_dir);
	Params *trees[3] = {t->model->params, t->backprop_buffer->prev_means, t->backprop_buffer->prev_vars};
	const char *names[3] = {"model_params", "means", "vars"};
	for (int i = t->model->params->n_locations - 1; i >= 0; i--)
		for (int k = 0; k < 3; k++) {
			char f[32];
			snprintf(f, sizeof(f), "/%03d.buffer", i);
			read_buffer(root + names[k] + f, trees[k]->locations[i], (size_t)t->model->params->sizes[i]);
		}
	// forward_pass re-packs the weights from locations[] at the start of every step, so nothing else needs refreshing
}

}  // extern "C"
