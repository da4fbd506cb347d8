"""ctypes binding of libyoho_b200.so (include/yoho_b200.h).  No torch types cross the boundary: tensors
are passed as raw device pointers + sizes.  There is NO fallback: if the library is missing or no CUDA
device is visible, construction fails loudly.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyoho_b200.so")

_c_f = ctypes.POINTER(ctypes.c_float)


class yoho_bn_host(ctypes.Structure):
    _fields_ = [("weight_host", _c_f), ("bias_host", _c_f), ("running_mean_host", _c_f), ("running_var_host", _c_f)]


class yoho_conv_host(ctypes.Structure):
    _fields_ = [("weight_host", _c_f), ("bias_host", _c_f)]


class yoho_part1_weights(ctypes.Structure):
    _fields_ = [("conv_in", yoho_conv_host), ("bn_a", yoho_bn_host), ("conv_a", yoho_conv_host),
                ("bn_b", yoho_bn_host), ("conv_b", yoho_conv_host), ("bn_out", yoho_bn_host),
                ("conv_out", yoho_conv_host)]


class yoho_part2_weights(ctypes.Structure):
    _fields_ = [("bn_init", yoho_bn_host), ("conv_init", yoho_conv_host), ("bn_a", yoho_bn_host),
                ("conv_a", yoho_conv_host), ("bn_b", yoho_bn_host), ("conv_b", yoho_conv_host),
                ("fc1", yoho_conv_host), ("bn1", yoho_bn_host), ("fc2", yoho_conv_host), ("bn2", yoho_bn_host),
                ("fc3", yoho_conv_host)]


class yoho_fourier_irrep(ctypes.Structure):
    _fields_ = [("d", ctypes.c_int), ("off", ctypes.c_int), ("w_a_host", _c_f), ("w_b_host", _c_f),
                ("idx_host", ctypes.POINTER(ctypes.c_int32)), ("omap_host", ctypes.POINTER(ctypes.c_int32)),
                ("w_in_host", _c_f), ("w_out_host", _c_f)]


class yoho_pair_io(ctypes.Structure):
    _fields_ = [("featA", ctypes.c_void_p), ("featB", ctypes.c_void_p), ("kpsA", ctypes.c_void_p), ("kpsB", ctypes.c_void_p),
                ("Ka", ctypes.c_int), ("Kb", ctypes.c_int), ("have_part1", ctypes.c_int),
                ("c_iters", ctypes.c_int), ("o_iters", ctypes.c_int),
                ("c_dist", ctypes.c_double), ("o_dist", ctypes.c_double), ("seed", ctypes.c_uint64),
                ("eqvA", ctypes.c_void_p), ("eqvB", ctypes.c_void_p), ("descA", ctypes.c_void_p), ("descB", ctypes.c_void_p),
                ("pairs", ctypes.c_void_p), ("n_pairs", ctypes.c_void_p), ("dr_index", ctypes.c_void_p),
                ("k0", ctypes.c_void_p), ("k1", ctypes.c_void_p), ("hyp", ctypes.c_void_p), ("c_status", ctypes.c_void_p),
                ("T_c", ctypes.c_void_p), ("c_best", ctypes.c_void_p), ("c_inl", ctypes.c_void_p), ("c_mask", ctypes.c_void_p),
                ("quat", ctypes.c_void_p), ("trans", ctypes.c_void_p), ("order", ctypes.c_void_p),
                ("T_o", ctypes.c_void_p), ("o_best", ctypes.c_void_p), ("o_inl", ctypes.c_void_p), ("o_mask", ctypes.c_void_p)]


# name -> (restype, argtypes); every symbol declared in include/yoho_b200.h
_vp = ctypes.c_void_p
_i = ctypes.c_int
SYMBOLS = {
    "yoho_abi_version": (_i, []),
    "yoho_last_error": (ctypes.c_char_p, []),
    "yoho_ctx_create": (_i, [_i, _vp, _vp, _vp, ctypes.POINTER(_vp)]),
    "yoho_ctx_destroy": (_i, [_vp]),
    "yoho_part1_load": (_i, [_vp, ctypes.POINTER(yoho_part1_weights)]),
    "yoho_part2_load": (_i, [_vp, ctypes.POINTER(yoho_part2_weights)]),
    "yoho_part1_load_fourier": (_i, [_vp, _vp, _i, _vp]),
    "yoho_set_gconv_impl": (_i, [_vp, _i]),
    "yoho_set_tuning": (_i, [_vp, _i, _i]),
    "yoho_part1_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "yoho_group_mean": (_i, [_vp, _vp, _i, _vp, _vp]),
    "yoho_nn1": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "yoho_mutual_nn": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "yoho_rot_argmax": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "yoho_part2_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "yoho_gather_kps": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "yoho_c_draw": (_i, [_vp, _vp, _i, _i, ctypes.c_uint64, _vp, _vp, _vp]),
    "yoho_c_ransac": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yoho_o_order": (_i, [_vp, _i, ctypes.c_uint64, _vp, _vp]),
    "yoho_o_score": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yoho_register_pair": (_i, [_vp, _vp, _vp, _vp]),
    "yoho_register_pair_begin": (_i, [_vp, _vp, _vp]),
    "yoho_register_pair_end": (_i, [_vp, _vp, _vp, _vp]),
    "yoho_lift_group_features": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yoho_fmr_batch": (_i, [_vp, _vp, _vp, _vp, _vp, _i, ctypes.c_double, _vp, _vp]),
    "yoho_registration_errors": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "yoho_gconv_train_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "yoho_gconv_train_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "yoho_rot_correlation_backward": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "yoho_launch_count": (ctypes.c_int64, [_vp]),
    "yoho_debug_layer": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "yoho_profile_enable": (_i, [_vp, _i]),
    "yoho_profile_read": (_i, [_vp, _vp, _vp, _vp]),
}
PROF_CLASSES = 9
PROF_NAMES = ["p1_L1_32x256", "p1_L2_256x512", "p1_L3_512x256", "p1_L4_256x32",
              "p2_init_128x256", "p2_a_256x512", "p2_b_512x256", "p2_head_1x1", "p1_fourier_transforms"]

_lib = None


class YohoError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree library and bind every symbol of the header.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise YohoError(f"{LIB_PATH} is missing: run `python -m yoho_b200.build` (or __graft_entry__.build()). "
                        "yoho_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load_library().yoho_last_error().decode("utf-8", "replace")
        raise YohoError(f"yoho_b200 call failed (status {rc}): {msg}")


def _hp(a):
    return a.ctypes.data_as(_c_f)


def _conv(sd, key):
    w = np.ascontiguousarray(np.asarray(sd[key + ".weight"], dtype=np.float32))
    b = np.ascontiguousarray(np.asarray(sd[key + ".bias"], dtype=np.float32))
    return yoho_conv_host(_hp(w), _hp(b)), (w, b)


def _bn(sd, key):
    arrs = [np.ascontiguousarray(np.asarray(sd[key + "." + n], dtype=np.float32))
            for n in ("weight", "bias", "running_mean", "running_var")]
    return yoho_bn_host(*[_hp(a) for a in arrs]), arrs


def _to_numpy_sd(sd):
    out = {}
    for k, v in sd.items():
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def part1_struct(sd):
    """state-dict (reference key names) -> yoho_part1_weights.  Returns (struct, keepalive)."""
    sd = _to_numpy_sd(sd)
    p, blk = "PartI_net.", "PartI_net.SO3_Conv_layers.0."
    keep = []
    w = yoho_part1_weights()
    for field, key, kind in (("conv_in", p + "Conv_in.0", _conv), ("bn_a", blk + "comb_layer_in.0", _bn),
                             ("conv_a", blk + "comb_layer_in.2", _conv), ("bn_b", blk + "comb_layer_out.0", _bn),
                             ("conv_b", blk + "comb_layer_out.2", _conv), ("bn_out", p + "Conv_out.comb_layer.0", _bn),
                             ("conv_out", p + "Conv_out.comb_layer.2", _conv)):
        s, k = kind(sd, key)
        setattr(w, field, s)
        keep.append(k)
    return w, keep


def part2_struct(sd):
    sd = _to_numpy_sd(sd)
    blk, fc = "PartII_SO3_Conv_layers.0.", "PartII_To_R_FC."
    keep = []
    w = yoho_part2_weights()
    for field, key, kind in (("bn_init", "Conv_init.comb_layer.0", _bn), ("conv_init", "Conv_init.comb_layer.2", _conv),
                             ("bn_a", blk + "comb_layer_in.0", _bn), ("conv_a", blk + "comb_layer_in.2", _conv),
                             ("bn_b", blk + "comb_layer_out.0", _bn), ("conv_b", blk + "comb_layer_out.2", _conv),
                             ("fc1", fc + "0", _conv), ("bn1", fc + "1", _bn), ("fc2", fc + "3", _conv),
                             ("bn2", fc + "4", _bn), ("fc3", fc + "6", _conv)):
        s, k = kind(sd, key)
        setattr(w, field, s)
        keep.append(k)
    return w, keep
