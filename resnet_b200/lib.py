"""ctypes binding of libresnet_b200.so: struct mirrors of include/resnet.h (reference: resnet.h:4-215) and
prototypes of include/resnet.h + include/resnet_b200.h.  Loading fails loudly when the CUDA library is
missing -- there is no CPU or PyTorch fallback for the product path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libresnet_b200.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)


class Dims(C.Structure):
    _fields_ = [("input", C.c_int), ("init_kernel_dim", C.c_int), ("init_conv_filters", C.c_int), ("init_conv_stride", C.c_int),
                ("init_maxpool_dim", C.c_int), ("init_maxpool_stride", C.c_int), ("n_conv_blocks", C.c_int),
                ("is_block_spatial_reduction", i32p), ("final_depth", C.c_int), ("output", C.c_int)]


class BatchNorm(C.Structure):
    _fields_ = [("spatial_dim", C.c_int), ("depth", C.c_int), ("gamma", f32p), ("beta", f32p)]


class ConvBlock(C.Structure):
    _fields_ = [("incoming_filters", C.c_int), ("incoming_spatial_dim", C.c_int), ("reduced_depth", C.c_int),
                ("expanded_depth", C.c_int), ("stride", C.c_int), ("depth_reduction", f32p),
                ("norm_depth_reduction", C.POINTER(BatchNorm)), ("spatial", f32p), ("norm_spatial", C.POINTER(BatchNorm)),
                ("depth_expansion", f32p), ("norm_expansion", C.POINTER(BatchNorm)), ("projection", f32p),
                ("norm_projection", C.POINTER(BatchNorm))]


class Params(C.Structure):
    _fields_ = [("init_conv_layer", f32p), ("norm_init_conv", C.POINTER(BatchNorm)), ("conv_blocks", C.POINTER(C.POINTER(ConvBlock))),
                ("fully_connected", f32p), ("locations", C.POINTER(f32p)), ("sizes", i32p), ("n_locations", C.c_int)]


class Cache_BatchNorm(C.Structure):
    _fields_ = [("input_size", C.c_int), ("feature_size", C.c_int), ("means", f32p), ("vars", f32p), ("normalized_temp", f32p),
                ("normalized", f32p)]


class Activation_ConvBlock(C.Structure):
    _fields_ = [("incoming_filters", C.c_int), ("incoming_spatial_dim", C.c_int), ("reduced_depth", C.c_int),
                ("expanded_depth", C.c_int), ("stride", C.c_int), ("post_reduced", f32p),
                ("norm_post_reduced", C.POINTER(Cache_BatchNorm)), ("post_reduced_activated", f32p), ("post_spatial", f32p),
                ("norm_post_spatial", C.POINTER(Cache_BatchNorm)), ("post_spatial_activated", f32p), ("post_expanded", f32p),
                ("norm_post_expanded", C.POINTER(Cache_BatchNorm)), ("post_expanded_norm_vals", f32p),
                ("transformed_residual", f32p), ("norm_post_projection", C.POINTER(Cache_BatchNorm)),
                ("post_projection_norm_vals", f32p), ("output", f32p), ("output_activated", f32p)]


class Activations(C.Structure):
    _fields_ = [("init_conv_applied", f32p), ("norm_init_conv", C.POINTER(Cache_BatchNorm)), ("init_conv_activated", f32p),
                ("max_inds", i32p), ("init_convblock_input", f32p),
                ("activation_conv_blocks", C.POINTER(C.POINTER(Activation_ConvBlock))), ("n_conv_blocks", C.c_int),
                ("final_conv_output_pooled", f32p), ("linear_output", f32p)]


class ResNet(C.Structure):
    _fields_ = [("dims", C.POINTER(Dims)), ("params", C.POINTER(Params))]


class Forward_Buffer(C.Structure):
    _fields_ = [("activations", C.POINTER(Activations)), ("pred", f32p), ("pred_cpu", f32p)]


class Backprop_Buffer(C.Structure):
    _fields_ = [("output_layer_deriv", f32p), ("param_derivs", C.POINTER(Params)), ("prev_means", C.POINTER(Params)),
                ("prev_vars", C.POINTER(Params)), ("activation_derivs", C.POINTER(Activations))]


class Batch(C.Structure):
    _fields_ = [("image_dim", C.c_int), ("image_size", C.c_int), ("n_images", C.c_int), ("cur_shard_id", C.c_int),
                ("cur_batch_in_shard", C.c_int), ("shard_n_images", C.c_int), ("full_shard_images", f32p),
                ("full_shard_correct_classes", i32p), ("images_float_cpu", f32p), ("images", f32p), ("correct_classes_cpu", i32p),
                ("correct_classes", i32p)]


class Train_ResNet(C.Structure):
    _fields_ = [("model", C.POINTER(ResNet)), ("cur_batch", C.POINTER(Batch)), ("forward_buffer", C.POINTER(Forward_Buffer)),
                ("backprop_buffer", C.POINTER(Backprop_Buffer)), ("learning_rate", C.c_float), ("weight_decay", C.c_float),
                ("base_mean_decay", C.c_float), ("base_var_decay", C.c_float), ("cur_mean_decay", C.c_float),
                ("cur_var_decay", C.c_float), ("eps", C.c_float), ("batch_size", C.c_int), ("n_epochs", C.c_int),
                ("cur_dump_id", C.c_int), ("cur_epoch", C.c_int), ("loss_per_epoch", f32p), ("accuracy_per_epoch", f32p),
                ("init_loaded", C.c_int), ("dump_dir", C.c_char_p)]


# every symbol include/resnet.h and include/resnet_b200.h declare (tests check the library exports all of them)
RESNET_H_SYMBOLS = ["populate_class_info", "init_dimensions", "init_resnet", "init_general_batch", "init_trainer",
                    "load_new_batch", "forward_pass", "backwards_pass", "update_parameters", "dump_trainer",
                    "overwrite_trainer_hyperparams", "overwrite_model_params"]
RESNET_B200_H_SYMBOLS = [
    "resnet_b200_last_error", "resnet_b200_clear_error", "resnet_b200_set_device", "resnet_b200_malloc", "resnet_b200_free",
    "resnet_b200_malloc_host", "resnet_b200_free_host", "resnet_b200_memcpy_h2d", "resnet_b200_memcpy_d2h",
    "resnet_b200_memcpy_d2d", "resnet_b200_memset", "resnet_b200_sync", "resnet_b200_rng_create", "resnet_b200_rng_destroy",
    "resnet_b200_stage_batch", "resnet_b200_stage_batch_device", "resnet_b200_prefetch_batch", "resnet_b200_commit_batch", "resnet_b200_trainer_sync", "resnet_b200_timer_begin",
    "resnet_b200_timer_end_ms", "resnet_b200_loss_accuracy", "resnet_b200_set_pred_copy", "resnet_b200_fetch_pred", "resnet_b200_epoch_stats", "resnet_b200_launch_count", "resnet_b200_profile", "resnet_b200_profile_read", "resnet_b200_profile_read2", "resnet_b200_uses_tensor_cores",
    "resnet_b200_destroy_trainer", "resnet_b200_conv_forward", "resnet_b200_conv_bench", "resnet_b200_conv_backward", "resnet_b200_batchnorm_forward",
    "resnet_b200_batchnorm_backward", "resnet_b200_maxpool_forward", "resnet_b200_maxpool_backward",
    "resnet_b200_avgpool_forward", "resnet_b200_avgpool_backward", "resnet_b200_matmul", "resnet_b200_softmax_ce",
    "resnet_b200_adam", "resnet_b200_set_dtype", "resnet_b200_trainer_dtype", "resnet_b200_set_op_dtype", "resnet_b200_convert",
    "resnet_b200_selfcheck", "resnet_b200_selfcheck_read", "resnet_b200_dp_unique_id", "resnet_b200_dp_init", "resnet_b200_dp_world_size", "resnet_b200_loader_plan"]

_lib = None


def load():
    """Returns the ctypes handle; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError("libresnet_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the product path has no CPU / PyTorch fallback)")
    L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
    vp, ci, cf, cll, csz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t
    T = C.POINTER(Train_ResNet)

    def proto(name, res, args):
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args

    proto("init_dimensions", C.POINTER(Dims), [ci] * 7 + [i32p, ci, ci])
    proto("init_resnet", C.POINTER(ResNet), [C.POINTER(Dims), vp])
    proto("init_general_batch", C.POINTER(Batch), [ci, ci, ci, ci])
    proto("init_trainer", T, [C.POINTER(ResNet), C.POINTER(Batch), ci, cf, cf, cf, cf, cf, ci, C.c_char_p])
    proto("load_new_batch", None, [T, vp, C.POINTER(Batch)])
    proto("forward_pass", None, [T])
    proto("backwards_pass", None, [T])
    proto("update_parameters", None, [T])
    proto("populate_class_info", vp, [C.c_char_p, C.c_char_p, C.c_char_p, ci])
    proto("dump_trainer", None, [ci, T, C.c_char_p])
    proto("overwrite_trainer_hyperparams", None, [T, ci, C.c_char_p])
    proto("overwrite_model_params", None, [T, ci, C.c_char_p])
    proto("resnet_b200_last_error", C.c_char_p, [])
    proto("resnet_b200_clear_error", None, [])
    proto("resnet_b200_set_device", ci, [ci])
    proto("resnet_b200_malloc", vp, [csz])
    proto("resnet_b200_free", None, [vp])
    proto("resnet_b200_malloc_host", vp, [csz])
    proto("resnet_b200_free_host", None, [vp])
    proto("resnet_b200_memcpy_h2d", ci, [vp, vp, csz])
    proto("resnet_b200_memcpy_d2h", ci, [vp, vp, csz])
    proto("resnet_b200_memcpy_d2d", ci, [vp, vp, csz])
    proto("resnet_b200_memset", ci, [vp, ci, csz])
    proto("resnet_b200_sync", ci, [])
    proto("resnet_b200_rng_create", vp, [C.c_ulonglong])
    proto("resnet_b200_rng_destroy", None, [vp])
    proto("resnet_b200_stage_batch", ci, [T, vp, vp])
    proto("resnet_b200_stage_batch_device", ci, [T, vp, vp])
    proto("resnet_b200_prefetch_batch", ci, [T, vp, vp])
    proto("resnet_b200_commit_batch", ci, [T])
    proto("resnet_b200_trainer_sync", ci, [T])
    proto("resnet_b200_timer_begin", ci, [T])
    proto("resnet_b200_timer_end_ms", cf, [T])
    proto("resnet_b200_loss_accuracy", ci, [T, f32p, i32p])
    proto("resnet_b200_set_pred_copy", ci, [T, ci])
    proto("resnet_b200_fetch_pred", ci, [T])
    proto("resnet_b200_epoch_stats", ci, [T, C.POINTER(C.c_double), C.POINTER(cll), C.POINTER(cll), ci])
    proto("resnet_b200_launch_count", cll, [])
    proto("resnet_b200_profile", None, [ci])
    proto("resnet_b200_profile_read", ci, [ci, C.POINTER(C.c_double), C.POINTER(cll), C.POINTER(C.c_double)])
    proto("resnet_b200_profile_read2", ci, [ci, C.POINTER(C.c_double), C.POINTER(cll), C.POINTER(C.c_double), C.POINTER(C.c_double)])
    proto("resnet_b200_uses_tensor_cores", ci, [T])
    proto("resnet_b200_destroy_trainer", None, [T])
    proto("resnet_b200_conv_forward", ci, [ci] * 6 + [vp, vp, vp, ci])
    proto("resnet_b200_conv_backward", ci, [ci] * 7 + [vp, vp, vp, vp, vp, ci])
    proto("resnet_b200_conv_bench", cf, [ci] * 10 + [C.c_char_p, ci])
    proto("resnet_b200_batchnorm_forward", ci, [ci, ci, ci, cf, vp, vp, vp, vp, vp, vp, ci, vp, ci])
    proto("resnet_b200_batchnorm_backward", ci, [ci, ci, ci, cf] + [vp] * 9 + [ci])
    proto("resnet_b200_maxpool_forward", ci, [vp, ci, ci, ci, ci, ci, vp, vp])
    proto("resnet_b200_maxpool_backward", ci, [vp, vp, ci, ci, ci, ci, ci, vp])
    proto("resnet_b200_avgpool_forward", ci, [vp, ci, ci, ci, vp])
    proto("resnet_b200_avgpool_backward", ci, [vp, ci, ci, ci, vp])
    proto("resnet_b200_matmul", ci, [vp, vp, ci, ci, ci, ci, ci, vp])
    proto("resnet_b200_softmax_ce", ci, [vp, vp, ci, ci, vp, vp])
    proto("resnet_b200_adam", ci, [vp, vp, vp, vp, cll, cf, cf, cf, cf, cf, cf, cf])
    proto("resnet_b200_set_dtype", ci, [ci])
    proto("resnet_b200_trainer_dtype", ci, [T])
    proto("resnet_b200_set_op_dtype", ci, [ci])
    proto("resnet_b200_convert", ci, [vp, vp, cll, ci])
    proto("resnet_b200_selfcheck", ci, [ci])
    proto("resnet_b200_selfcheck_read", ci, [T, ci, f32p, C.POINTER(cll), C.c_char_p, ci])
    proto("resnet_b200_dp_unique_id", ci, [vp])
    proto("resnet_b200_dp_init", ci, [T, vp, ci, ci, cll])
    proto("resnet_b200_dp_world_size", ci, [T])
    proto("resnet_b200_loader_plan", ci, [ci, ci, ci, ci, ci, i32p, i32p])
    proto("resnet_b200_dp_plan", ci, [C.POINTER(cll), ci, cll, i32p, ci, cll, C.POINTER(cll), C.POINTER(cll), i32p, ci])
    _lib = L
    return L
