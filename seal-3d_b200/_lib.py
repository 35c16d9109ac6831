"""ctypes binding of the C-ABI (include/seal3d_b200.h -> libseal3d_b200.so).

There is NO fallback: if the CUDA library is missing the import of any op raises.  Tensors are
passed as raw device pointers (``tensor.data_ptr()``), sizes as integers, the stream as the current
torch CUDA stream handle -- exactly what a cgo / JNI / pybind shim would pass.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libseal3d_b200.so")

P, U32, U64, F32, I32, SZ = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int, C.c_size_t

# name -> argument ctypes (without the trailing stream, which every launch function takes last)
_SIGS = {
    "s3d_near_far_from_aabb": [P, P, P, U32, F32, P, P],
    "s3d_sph_from_ray": [P, P, F32, U32, P],
    "s3d_morton3D": [P, U32, P],
    "s3d_morton3D_invert": [P, U32, P],
    "s3d_packbits": [P, U32, F32, P],
    "s3d_march_rays_train": [P, P, P, F32, F32, U32, U32, U32, U32, U32, P, P, P, P, P, P, P, P],
    "s3d_march_rays_train_count": [P, P, P, F32, F32, U32, U32, U32, U32, P, P, P, P, P],
    "s3d_march_rays_train_write": [P, P, P, F32, F32, U32, U32, U32, U32, U32, P, P, P, P, P, P, P],
    "s3d_composite_rays_train_forward": [P, P, P, P, U32, U32, F32, P, P, P],
    "s3d_composite_rays_train_backward": [P, P, P, P, P, P, P, P, U32, U32, F32, P, P],
    "s3d_march_rays": [U32, U32, P, P, P, P, F32, F32, U32, U32, U32, P, P, P, P, P, P, P],
    "s3d_composite_rays": [U32, U32, F32, P, P, P, P, P, P, P, P],
    "s3d_grid_encode_forward": [P, P, P, P, U32, U32, U32, U32, F32, U32, P, U32, I32, U32, I32],
    "s3d_grid_encode_backward": [P, P, P, P, P, U32, U32, U32, U32, F32, U32, P, P, U32, I32, U32, I32],
    "s3d_grad_total_variation": [P, P, P, P, F32, U32, U32, U32, U32, F32, U32, U32, I32, I32],
    "s3d_grid_level_scales": [U32, F32, U32, P],
    "s3d_sh_encode_forward": [P, P, U32, U32, U32, P],
    "s3d_sh_encode_backward": [P, P, U32, U32, U32, P, P],
    "s3d_freq_encode_forward": [P, U32, U32, U32, U32, P],
    "s3d_freq_encode_backward": [P, P, U32, U32, U32, U32, P],
    "s3d_ffmlp_forward": [P, P, U32, U32, U32, U32, U32, U32, U32, P, P],
    "s3d_ffmlp_inference": [P, P, U32, U32, U32, U32, U32, U32, U32, P, P],
    "s3d_linear_forward": [P, P, U32, U32, U32, P],
    "s3d_linear_backward": [P, P, P, U32, U32, U32, P, P],
    "s3d_ffmlp_backward": [P, P, P, P, U32, U32, U32, U32, U32, U32, U32, I32, P, P, P],
    "s3d_seal_bbox_map_to_origin": [P, P, U32, P, P, P, P, P, U32, P, U32, P, P, P, P, P, P],
    "s3d_seal_map_color": [P, P, U32, P, P, F32, P],
    "s3d_seal_force_fill_bitfield": [P, P, P, U32, U32],
    "s3d_seal_brush_map_to_origin": [P, U32, P, U32, P, U32, P, P, P, P, U32, F32, I32, P, P],
    "s3d_seal_anchor_map_to_origin": [P, U32, P, U32, P, U32, P, P, P, P, F32, F32, P, P, P, P],
    "s3d_seal_map_color_image": [P, P, P, U32, P, P, U32, U32, P, P, P, P, F32, P],
    "s3d_pretrain_loss": [P, P, P, P, U32, P, P, P],
    "s3d_finetune_loss": [P, P, P, P, P, U32, F32, P, P, P],
    "s3d_adam_step": [P, P, P, P, P, U64, F32, F32, F32, F32, U32, F32, I32, I32, P],
    "s3d_grad_scaler_check": [P, U64, P, F32, F32, I32],
    "s3d_grad_scaler_update": [P, F32, F32, U32],
    "s3d_ema_update": [P, P, U64, F32],
    "s3d_cast_f32_to_f16": [P, P, U64],
    "s3d_density_grid_ema": [P, P, U32, F32, P],
    "s3d_get_rays": [P, U32, F32, F32, F32, F32, U32, U32, P, U32, U32, P, P],
    "s3d_mark_untrained_grid": [P, P, U32, F32, F32, U32, U32, F32, P],
    "s3d_density_cells_to_xyz": [P, U32, U32, F32, U32, P],
    "s3d_density_scatter": [P, P, U32, F32, P],
    "s3d_distill_rays": [P, P, P, P, P, P, P, P, U32, U32, F32, F32, F32, P, P, P, P],
    "s3d_density_pick_cells": [P, U32, U32, U32, U32, P, P],
    "s3d_density_grid_update": [P, P, U32, F32, F32, P],
    "s3d_packbits_dev_thresh": [P, U32, P, P],
    "s3d_mean_count": [P, U32, P],
    "s3d_ngp_interleave_tables": [P, P, P, U64],
    "s3d_ngp_encode": [P, U32, F32, P, U32, P, U32, F32, U32, P, I32],
    "s3d_ngp_pair_tables": [P, P, P, U64],
    "s3d_ngp_encode_pair": [P, P, P, U32, F32, P, P, U32, F32, U32, P, P],
    "s3d_ngp_pair_forward": [P, P, P, P, P, U32, F32, P, P, U32, F32, U32, P, P, P, P, P, P, P, P, P, P, F32, F32, P, P, P, P, P],
    "s3d_ngp_mlp_forward": [P, P, U32, P, P, P, P, P, F32, P, P, P, I32],
    "s3d_ngp_mlp_backward": [P, P, U32, P, P, P, P, P, F32, P, P, P, F32, P, P, P, P, P, I32],
    "s3d_ngp_scatter": [P, P, U32, F32, P, P, U32, F32, U32, F32],
    "s3d_ngp_scatter_levels": [P, P, U32, F32, P, P, U32, F32, U32, F32, U32, U32],
    "s3d_ngp_scatter_fixed": [P, P, U32, F32, P, P, U32, F32, U32, F32, P],
    "s3d_ngp_mlp_backward_fixed": [P, P, U32, P, P, P, P, P, F32, P, P, P, F32, P, P, P, P, P, I32, P],
    "s3d_fixed_to_float": [P, P, U64, P],
    "s3d_ngp_scatter_count": [P, U32, F32, P, U32, F32, U32, P],
    "s3d_ngp_peer_adam_tables": [P, P, P, P, U32, U32, P, P, U32, U64, U64, F32, F32, F32, F32, U32, F32],
    "s3d_peer_sum": [P, U32, P, U64],
    "s3d_ngp_adam_tables": [P, P, P, P, P, P, U32, U64, F32, F32, F32, F32, U32, F32, P],
    "s3d_vm_forward": [P, U32, P, P, P, P, P, P, P, P, U32, I32, P],
    "s3d_vm_backward": [P, U32, P, P, P, P, P, P, P, P, U32, I32, P, P, P, P, P, P, P],
    "s3d_vm_forward_f16": [P, U32, P, P, P, P, P, P, P, P, U32, P],
    "s3d_vm_backward_f16": [P, U32, P, P, P, P, P, P, P, P, U32, P, P, P, P, P, P, P],
    "s3d_tensorf_head_encode": [P, U32, U32, P, U32, U32, U32, P],
    "s3d_tensorf_head_encode_backward": [P, P, U32, U32, U32, U32, U32, P],
    "s3d_vm_resize": [P, U32, U32, P, U32, U32, U32],
}
_NO_STREAM = {"s3d_allocate_splitk": [SZ], "s3d_free_splitk": [], "s3d_march_set_clip": [I32]}

_lib = None
LAUNCHES = 0  # kernels launched through this binding (bench.py reports the per-run delta)
PROFILE = None  # set to a list to record (name, start_event, end_event) per call (bench.py's per-kernel breakdown)
_KERNELS_PER_CALL = {"s3d_linear_backward": 3, "s3d_density_pick_cells": 4, "s3d_density_grid_update": 2, "s3d_march_rays_train": 9, "s3d_march_rays_train_count": 7, "s3d_ffmlp_backward": 2, "s3d_seal_map_color": 2, "s3d_seal_anchor_map_to_origin": 2, "s3d_seal_map_color_image": 2}


class S3DError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S3DError("seal3d_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU / PyTorch fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for n, a in _SIGS.items():
            f = getattr(_lib, n)
            f.argtypes = a + [P]
            f.restype = C.c_int
        for n, a in _NO_STREAM.items():
            f = getattr(_lib, n)
            f.argtypes = a
            f.restype = C.c_int
    return _lib


def exported_symbols():
    return sorted(list(_SIGS) + list(_NO_STREAM))


def _conv(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a


def host_f32(values):
    """small host-side constant block (float32 array) for the few by-value arguments of the C-ABI"""
    import numpy as np
    arr = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(-1))
    return arr, arr.ctypes.data


def host_ptrs(values):
    """host array of device pointers (ints or tensors) for the *_peers arguments; keep the returned array alive across the call"""
    import numpy as np
    arr = np.ascontiguousarray(np.asarray([v.data_ptr() if isinstance(v, torch.Tensor) else int(v) for v in values], dtype=np.uint64))
    return arr, arr.ctypes.data


def host_i32(values):
    import numpy as np
    arr = np.ascontiguousarray(np.asarray(values, dtype=np.int32).reshape(-1))
    return arr, arr.ctypes.data


def call(name, *args):
    """Launch `name` on the current torch CUDA stream; raises on a non-zero return code."""
    global LAUNCHES
    f = getattr(lib(), name)
    LAUNCHES += _KERNELS_PER_CALL.get(name, 1)
    stream = torch.cuda.current_stream().cuda_stream
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = f(*[_conv(a) for a in args], stream)
        e1.record()
        PROFILE.append((name, e0, e1))
    else:
        rc = f(*[_conv(a) for a in args], stream)
    if rc != 0:
        if rc > 0:
            raise S3DError("%s: CUDA error %d (%s)" % (name, rc, _cuda_err(rc)))
        raise S3DError("%s: %s" % (name, {-22: "invalid argument", -95: "configuration not supported by this build"}.get(rc, rc)))


def call_nostream(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise S3DError("%s failed with %d" % (name, rc))


def _cuda_err(rc):
    try:
        rt = C.CDLL("libcudart.so")
        rt.cudaGetErrorString.restype = C.c_char_p
        return rt.cudaGetErrorString(rc).decode()
    except Exception:
        return "?"


def check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise S3DError("expected CUDA tensors (the reference raises the same TORCH_CHECK)")
        if t is not None and not t.is_contiguous():
            raise S3DError("expected contiguous tensors")
