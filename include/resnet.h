/*
 * resnet.h -- drop-in boundary of the B200-native ResNet training hot path.
 *
 * The struct schema below is byte-compatible (same type names, same field names, same
 * field order and types) with the reference's resnet.h (reference: resnet.h:4-215), so a
 * host driver written against the reference (reference: resnet.cu:3222-3429) compiles and
 * links against libresnet_b200.so unchanged.  The reference keeps its entry points in the
 * same translation unit as main() and declares no prototypes; here they are declared
 * extern "C" so the library can be bound from C, C++ or ctypes.
 *
 * All device tensors are fp32, activations NHWC, conv weights [Cout][Cin][kh][kw]
 * exactly as in the reference's resnet.cu (reference: resnet.cu:140,145,155).
 *
 * Fields the B200 path does not materialise (they are recomputable and the reference's
 * own resnet_clean.h drops them) are allocated only when the library runs in
 * "keep-all" mode (RESNET_B200_KEEP_ALL=1); otherwise they are NULL.  See DESIGN.md.
 */
#ifndef RESNET_B200_RESNET_H
#define RESNET_B200_RESNET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- class metadata (reference: resnet.h:4-9) ---- */
typedef struct {
	char ** labels;
	char ** synsets;
	int * counts;
	int n_classes;
} Class_Metadata;

/* ---- network hyper-shape (reference: resnet.h:11-33) ---- */
typedef struct {
	int input;                        /* 224 */
	int init_kernel_dim;              /* 7 */
	int init_conv_filters;            /* 64 */
	int init_conv_stride;             /* 2 */
	int init_maxpool_dim;             /* 3 */
	int init_maxpool_stride;          /* 2 */
	int n_conv_blocks;                /* 16 for ResNet-50, 50 for ResNet-152 */
	int * is_block_spatial_reduction; /* 1 where the block strides by 2 and doubles depth */
	int final_depth;                  /* 2048 */
	int output;                       /* 1000 */
} Dims;

/* ---- parameters (reference: resnet.h:35-88) ---- */
typedef struct {
	int spatial_dim;
	int depth;
	float * gamma;
	float * beta;
} BatchNorm;

typedef struct {
	int incoming_filters;
	int incoming_spatial_dim;
	int reduced_depth;
	int expanded_depth;
	int stride;
	float * depth_reduction;            /* 1x1: [reduced][incoming] */
	BatchNorm * norm_depth_reduction;
	float * spatial;                    /* 3x3: [reduced][reduced][3][3], stride = block stride */
	BatchNorm * norm_spatial;
	float * depth_expansion;            /* 1x1: [expanded][reduced] */
	BatchNorm * norm_expansion;
	float * projection;                 /* NULL for identity shortcut; 1x1 (stride 1) or 3x3 (stride 2) */
	BatchNorm * norm_projection;
} ConvBlock;

typedef struct {
	float * init_conv_layer;            /* 7x7: [64][3][7][7] */
	BatchNorm * norm_init_conv;
	ConvBlock ** conv_blocks;
	float * fully_connected;            /* [final_depth][output], no bias */
	float ** locations;                 /* flat view, order of reference resnet.cu:839-943 */
	int * sizes;
	int n_locations;                    /* 16 + 9 * n_conv_blocks (reference: resnet.cu:819) */
} Params;

/* ---- activations (reference: resnet.h:90-157) ---- */
typedef struct {
	int input_size;
	int feature_size;
	float * means;
	float * vars;                       /* biased batch variance */
	float * normalized_temp;            /* x-hat; keep-all mode only */
	float * normalized;                 /* gamma * x-hat + beta; keep-all mode only */
} Cache_BatchNorm;

typedef struct {
	int incoming_filters;
	int incoming_spatial_dim;
	int reduced_depth;
	int expanded_depth;
	int stride;
	float * post_reduced;
	Cache_BatchNorm * norm_post_reduced;
	float * post_reduced_activated;
	float * post_spatial;
	Cache_BatchNorm * norm_post_spatial;
	float * post_spatial_activated;
	float * post_expanded;
	Cache_BatchNorm * norm_post_expanded;
	float * post_expanded_norm_vals;    /* keep-all mode only */
	float * transformed_residual;       /* NULL for identity shortcut */
	Cache_BatchNorm * norm_post_projection;
	float * post_projection_norm_vals;  /* keep-all mode only */
	float * output;                     /* keep-all mode only */
	float * output_activated;
} Activation_ConvBlock;

typedef struct {
	float * init_conv_applied;
	Cache_BatchNorm * norm_init_conv;
	float * init_conv_activated;
	int * max_inds;                     /* flat index into init_conv_activated (reference: resnet.cu:459-468) */
	float * init_convblock_input;
	Activation_ConvBlock ** activation_conv_blocks;
	int n_conv_blocks;
	float * final_conv_output_pooled;
	float * linear_output;
} Activations;

typedef struct {
	Dims * dims;
	Params * params;
} ResNet;

/* ---- step buffers (reference: resnet.h:160-174) ---- */
typedef struct {
	Activations * activations;
	float * pred;                       /* device, softmax, batch x output */
	float * pred_cpu;                   /* host copy, valid when forward_pass returns */
} Forward_Buffer;

typedef struct {
	float * output_layer_deriv;
	Params * param_derivs;
	Params * prev_means;                /* Adam first moment */
	Params * prev_vars;                 /* Adam second moment */
	Activations * activation_derivs;
} Backprop_Buffer;

/* ---- batch + trainer (reference: resnet.h:176-215) ---- */
typedef struct {
	int image_dim;
	int image_size;
	int n_images;
	int cur_shard_id;
	int cur_batch_in_shard;
	int shard_n_images;
	float * full_shard_images;
	int * full_shard_correct_classes;
	float * images_float_cpu;           /* pinned */
	float * images;                     /* device, NHWC fp32 */
	int * correct_classes_cpu;          /* pinned */
	int * correct_classes;              /* device */
} Batch;

typedef struct {
	ResNet * model;
	Batch * cur_batch;
	Forward_Buffer * forward_buffer;
	Backprop_Buffer * backprop_buffer;
	float learning_rate;
	float weight_decay;
	float base_mean_decay;
	float base_var_decay;
	float cur_mean_decay;
	float cur_var_decay;
	float eps;
	int batch_size;
	int n_epochs;
	int cur_dump_id;
	int cur_epoch;
	float * loss_per_epoch;
	float * accuracy_per_epoch;
	int init_loaded;
	const char * dump_dir;
} Train_ResNet;

/* ------------------------------------------------------------------------------------------
 * Entry points the reference's host driver calls (reference: resnet.cu:3222-3429).
 * Same names, argument order and meaning.  `gen` is the caller's curandGenerator_t*
 * (passed as void* so that C callers need not include curand.h); weights are drawn with
 * curandGenerateNormal in the reference's order, so seed 1234 reproduces its init bytes.
 * ------------------------------------------------------------------------------------------ */

/* reference: resnet.cu:1363 */
Class_Metadata * populate_class_info(char * label_filename, char * synset_filename, char * class_size_filename, int n_classes);
/* reference: resnet.cu:666 */
Dims * init_dimensions(int input, int init_kernel_dim, int init_conv_filters, int init_conv_stride, int init_maxpool_dim,
                       int init_maxpool_stride, int n_conv_blocks, int * is_block_spatial_reduction, int final_depth, int output);
/* reference: resnet.cu:951 */
ResNet * init_resnet(Dims * dims, void * gen /* curandGenerator_t* */);
/* reference: resnet.cu:1196 */
Batch * init_general_batch(int n_images, int image_size, int image_dim, int shard_n_images);
/* reference: resnet.cu:1157 */
Train_ResNet * init_trainer(ResNet * model, Batch * cur_batch, int batch_size, float learning_rate, float weight_decay,
                            float mean_decay, float var_decay, float eps, int n_epochs, const char * dump_dir);
/* reference: resnet.cu:1235 */
void load_new_batch(Train_ResNet * trainer, Class_Metadata * class_metadata, Batch * batch_buffer);
/* reference: resnet.cu:1526 */
void forward_pass(Train_ResNet * trainer);
/* reference: resnet.cu:1777 (block wiring as in resnet_clean.cu:2459-2958, which has the
 * spatial BatchNorm backward that resnet.cu:2060-2083 forgets to launch) */
void backwards_pass(Train_ResNet * trainer);
/* reference: resnet.cu:2910 */
void update_parameters(Train_ResNet * trainer);
/* reference: resnet.cu:2755 -- writes <root>/<special_dir>/<dump_id %08d>/{model_params,gradients,means,vars}/%03d.buffer,
 * activations/..., activation_derivs/..., trainer_metadata.txt, trainer_checkpoint.txt in the reference's layout (fp32 files
 * whatever the device storage type).  <root> = $RESNET_B200_DUMP_ROOT, default the reference's hard-coded
 * /mnt/storage/data/vision/imagenet/training_dumps; directories are created. */
void dump_trainer(int dump_id, Train_ResNet * trainer, const char * special_dir);
/* reference: resnet.cu:2778 -- cur_shard_id, cur_batch_in_shard, cur_mean_decay, cur_var_decay, cur_dump_id, cur_epoch */
void overwrite_trainer_hyperparams(Train_ResNet * trainer, int dump_id, const char * special_dir);
/* reference: resnet.cu:2821 -- model_params, means, vars of every location */
void overwrite_model_params(Train_ResNet * trainer, int dump_id, const char * special_dir);

#ifdef __cplusplus
}
#endif

#endif /* RESNET_B200_RESNET_H */
