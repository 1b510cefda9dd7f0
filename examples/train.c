/*
 * examples/train.c -- a plain-C host driver for libresnet_b200.so written against include/resnet.h only (plus three runtime services of
 * include/resnet_b200.h: error string, RNG handle, device sync).  It drives the same entry points, in the same order, as the
 * reference's main() (reference: resnet.cu:3222-3429): populate_class_info -> init_dimensions -> init_resnet -> init_general_batch ->
 * init_trainer -> [overwrite_*] -> per iteration { load_new_batch, forward_pass, host loss / accuracy from pred_cpu,
 * backwards_pass, update_parameters } -> dump_trainer(77777777).
 *
 * The reference hard-codes its dataset paths and hyper-parameters; here they come from the command line and the environment
 * (RESNET_B200_SHARD_DIR, RESNET_B200_DUMP_ROOT), and the network size can be scaled down so the driver runs in seconds in a test:
 *
 *   train <steps> <batch> <input_dim> <n_blocks> <n_classes> <shard_n_images> [resume_dump_id]
 *   (ImageNet run of the reference: train <iters> 32 224 16 1000 32768)
 *
 * Shards are the reference's `%03d.images` / `%03d.labels` files (build_training_shards.c).  Built by __graft_entry__.build()
 * with gcc -- no nvcc, no CUDA headers: everything behind the C ABI.  Exit code 0 and a final "train ok" line when every step ran
 * without a recorded error.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "resnet_b200.h" /* includes resnet.h */

static void die_on_error(const char * where) {
	const char * err = resnet_b200_last_error();
	if (err && err[0]) {
		fprintf(stderr, "train: error after %s: %s\n", where, err);
		exit(2);
	}
}

int main(int argc, char * argv[]) {
	if (argc < 7) {
		fprintf(stderr, "usage: %s <steps> <batch> <input_dim> <n_blocks> <n_classes> <shard_n_images> [resume_dump_id]\n", argv[0]);
		return 1;
	}
	const int steps = atoi(argv[1]), batch_size = atoi(argv[2]), input_dim = atoi(argv[3]), n_blocks = atoi(argv[4]);
	const int n_classes = atoi(argv[5]), shard_n_images = atoi(argv[6]);
	const int resume_id = argc > 7 ? atoi(argv[7]) : -1;

	/* class metadata (label / synset / count text files) is optional here: the hot path never reads it (resnet.cu:1235 ignores it too) */
	Class_Metadata * class_metadata = NULL;
	const char * meta = getenv("RESNET_B200_CLASS_METADATA_DIR");
	if (meta) {
		char a[1024], b[1024], c[1024];
		snprintf(a, sizeof(a), "%s/id_to_label_mapping.txt", meta);
		snprintf(b, sizeof(b), "%s/id_to_synset_mapping.txt", meta);
		snprintf(c, sizeof(c), "%s/id_to_img_count_mapping.txt", meta);
		class_metadata = populate_class_info(a, b, c, n_classes);
	}

	/* model dimensions: the reference's ResNet-50 layout rule (stride-2 blocks after 3, 4, 6 blocks of a stage) scaled to n_blocks */
	int * reductions = (int *) calloc(n_blocks, sizeof(int));
	int final_depth = 256;
	if (n_blocks == 16) { reductions[3] = reductions[7] = reductions[13] = 1; }
	else if (n_blocks == 50) { reductions[3] = reductions[11] = reductions[47] = 1; }
	else if (n_blocks >= 3) { reductions[1] = 1; }
	for (int i = 0; i < n_blocks; i++) if (reductions[i]) final_depth *= 2;
	Dims * dims = init_dimensions(input_dim, 7, 64, 2, 3, 2, n_blocks, reductions, final_depth, n_classes);

	void * gen = resnet_b200_rng_create(1234ULL); /* curandGenerator_t*, seeded like resnet.cu:3264-3267 */
	ResNet * model = init_resnet(dims, gen);
	Batch * batch = init_general_batch(batch_size, input_dim * input_dim * 3, input_dim, shard_n_images);
	const char * dump_dir = "train_c";
	Train_ResNet * trainer = init_trainer(model, batch, batch_size, 0.0001f, 0.0f, 0.9f, 0.999f, 0.0000001f, 1, dump_dir);
	die_on_error("init_trainer");
	if (resume_id != -1) {
		overwrite_trainer_hyperparams(trainer, resume_id, dump_dir);
		overwrite_model_params(trainer, resume_id, dump_dir);
		die_on_error("overwrite_*");
	}

	float epoch_loss = 0.f, epoch_n_wrong = 0.f, first_loss = 0.f, last_loss = 0.f;
	for (int iter = 0; iter < steps; iter++) {
		load_new_batch(trainer, class_metadata, trainer->cur_batch);
		die_on_error("load_new_batch");
		forward_pass(trainer); /* returns with pred_cpu valid */
		die_on_error("forward_pass");

		/* loss and accuracy on the host, exactly the reference's loops (resnet.cu:3363-3383) */
		const float * pred = trainer->forward_buffer->pred_cpu;
		const int * correct = trainer->cur_batch->correct_classes_cpu;
		float batch_loss = 0.f;
		int batch_n_wrong = 0;
		for (int s = 0; s < batch_size; s++) batch_loss += -1.f * logf(pred[s * n_classes + correct[s]]);
		for (int s = 0; s < batch_size; s++) {
			const float v = pred[s * n_classes + correct[s]];
			for (int c = 0; c < n_classes; c++)
				if (c != correct[s] && pred[s * n_classes + c] >= v) { batch_n_wrong++; break; }
		}
		epoch_loss += batch_loss;
		epoch_n_wrong += (float) batch_n_wrong;
		if (iter == 0) first_loss = batch_loss / batch_size;
		last_loss = batch_loss / batch_size;
		printf("Epoch: 0, Batch: %d ----- Avg. Loss: %.4f, Accuracy: %.2f%% (shard %d, next batch %d, dump id %d)\n", iter, batch_loss / batch_size,
		       100.f * (batch_size - batch_n_wrong) / batch_size, trainer->cur_batch->cur_shard_id, trainer->cur_batch->cur_batch_in_shard, trainer->cur_dump_id);

		backwards_pass(trainer);
		die_on_error("backwards_pass");
		update_parameters(trainer);
		die_on_error("update_parameters");
	}
	resnet_b200_sync();
	trainer->loss_per_epoch[0] = epoch_loss;
	trainer->accuracy_per_epoch[0] = steps > 0 ? ((float) steps * batch_size - epoch_n_wrong) / ((float) steps * batch_size) : 0.f;
	trainer->cur_epoch += 1;
	dump_trainer(77777777, trainer, trainer->dump_dir); /* the reference's final dump id (resnet.cu:3423-3424) */
	die_on_error("dump_trainer");
	if (!(first_loss == first_loss) || !(last_loss == last_loss) || isinf(last_loss)) { fprintf(stderr, "train: non-finite loss\n"); return 3; }
	printf("train ok: %d steps, first loss %.4f, last loss %.4f, tensor cores %d\n", steps, first_loss, last_loss, resnet_b200_uses_tensor_cores(trainer));
	resnet_b200_destroy_trainer(trainer);
	resnet_b200_rng_destroy(gen);
	return 0;
}
