"""ctypes loader for liblnb200.so, the C-ABI CUDA library declared in include/lidarnerf_b200.h.

There is NO CPU fallback: if the library is missing the import fails loudly, and every entry point raises
RuntimeError on a non-zero status (CUDA error or rejected argument), like TORCH_CHECK in the reference
bindings (e.g. gridencoder.cu:608-624).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LNB200_LIB") or os.path.join(_HERE, "lib", "liblnb200.so")   # override: diagnostic builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found. Build it with `python lidar-nerf_b200/build.py` (needs nvcc; sm_100a only). "
        "lidar-nerf_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)
lib.lnb_strerror.restype = C.c_char_p
lib.lnb_strerror.argtypes = [C.c_int]
lib.lnb_arch.restype = C.c_char_p
lib.lnb_launch_count.restype = C.c_uint64
for _n in ("lnb_field_fused_weight_bytes", "lnb_field_fused_weight_bytes_bf16"):
    getattr(lib, _n).restype = C.c_size_t
    getattr(lib, _n).argtypes = [C.c_uint32] * 6
lib.lnb_ffmlp_backward_workspace_bytes.restype = C.c_size_t
lib.lnb_ffmlp_backward_workspace_bytes.argtypes = [C.c_uint32] * 4
for _n in ("lnb_lidar_to_pano_workspace_bytes", "lnb_pano_to_lidar_workspace_bytes"):
    getattr(lib, _n).restype = C.c_size_t
    getattr(lib, _n).argtypes = [C.c_uint32] * 2

u32, f32, i32, vp, sz = C.c_uint32, C.c_float, C.c_int, C.c_void_p, C.c_size_t

# every exported symbol of include/lidarnerf_b200.h (tests check the .so exports all of them)
SYMBOLS = [
    "lnb_strerror", "lnb_version", "lnb_arch", "lnb_launch_count",
    "lnb_near_far_from_aabb", "lnb_sph_from_ray", "lnb_morton3D", "lnb_morton3D_invert", "lnb_packbits",
    "lnb_march_rays_train", "lnb_composite_rays_train_forward", "lnb_composite_rays_train_backward",
    "lnb_composite_rays_train_forward_ex", "lnb_composite_rays_train_backward_ex",
    "lnb_march_rays", "lnb_composite_rays",
    "lnb_grid_encode_forward", "lnb_grid_encode_backward", "lnb_grad_total_variation",
    "lnb_freq_encode_forward", "lnb_freq_encode_backward",
    "lnb_sh_encode_forward", "lnb_sh_encode_backward",
    "lnb_ffmlp_forward", "lnb_ffmlp_inference", "lnb_ffmlp_backward_workspace_bytes", "lnb_ffmlp_backward",
    "lnb_allocate_splitk", "lnb_free_splitk", "lnb_adam_step", "lnb_adam_set_hyper", "lnb_adam_step_dev", "lnb_dp_adam_exchange",
    "lnb_grid_encode_forward_ex", "lnb_grid_encode_backward_ex", "lnb_ffmlp_backward_accumulate", "lnb_ffmlp_forward_ex",
    "lnb_march_rays_train_ex", "lnb_zero_sample_tail_ex", "lnb_field_supported", "lnb_field_ray_terms",
    "lnb_field_forward", "lnb_field_head_backward",
    "lnb_zero_sample_tail", "lnb_field_head_input", "lnb_field_head_rgb", "lnb_lidar_loss",
    "lnb_field_head_out_grad", "lnb_field_sigma_out_grad", "lnb_lidar_rays", "lnb_lidar_composite_step",
    "lnb_chamfer_forward", "lnb_chamfer_backward", "lnb_lidar_to_pano_workspace_bytes", "lnb_lidar_to_pano",
    "lnb_pano_to_lidar_workspace_bytes", "lnb_pano_to_lidar",
    "lnb_field_head_backward_rows", "lnb_ffmlp_backward_accumulate_rows", "lnb_grid_encode_backward_rows",
    "lnb_field_fused_weight_bytes", "lnb_field_pack_weights", "lnb_field_fused_forward", "lnb_field_set_l2_window",
    "lnb_lidar_loss_ex", "lnb_lidar_composite_forward", "lnb_lidar_composite_backward", "lnb_dp_adam_exchange_mc", "lnb_lidar_batch", "lnb_packbits_dev",
    "lnb_field_supported_bf16", "lnb_field_ray_terms_bf16", "lnb_field_forward_bf16", "lnb_field_head_backward_bf16", "lnb_field_head_backward_rows_bf16", "lnb_field_fused_weight_bytes_bf16", "lnb_field_pack_weights_bf16", "lnb_field_fused_forward_bf16", "lnb_ffmlp_forward_ex_bf16", "lnb_ffmlp_inference_bf16", "lnb_ffmlp_backward_accumulate_bf16", "lnb_ffmlp_backward_accumulate_rows_bf16",
]


def check(status, what):
    if status != 0:
        raise RuntimeError(f"{what}: {lib.lnb_strerror(int(status)).decode()} (status {int(status)})")


def launch_count():
    return int(lib.lnb_launch_count())
