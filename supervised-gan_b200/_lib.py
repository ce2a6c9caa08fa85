"""ctypes binding of libsgk.so (C ABI: include/sgk.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Nothing here imports oracle/.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsgk.so")

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
ACT = {None: 0, "none": 0, "relu": 1, "lrelu": 2, "tanh": 3, "sigmoid": 4}
FP32, TF32, BF16 = 0, 1, 2
PRECISION = {"fp32": FP32, "tf32": TF32, "bf16": BF16}
OP_FWD, OP_DGRAD, OP_WGRAD = 0, 1, 2


class SgkConvDesc(Structure):
    _fields_ = [(n, c_int32) for n in ("N", "Cin", "Hin", "Win", "Cout", "Hout", "Wout", "k", "stride", "pad",
                                       "transposed", "precision")]


class SgkAdamTensor(Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", c_int64)]


# name -> (restype, argtypes); every symbol include/sgk.h declares
P = c_void_p
class SgkPackJob(ctypes.Structure):
    """include/sgk.h: struct SgkPackJob."""
    _fields_ = [("desc", SgkConvDesc), ("op", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("w_raw", ctypes.c_void_p), ("w_packed", ctypes.c_void_p)]


SIGNATURES = {
    "sgk_version": (c_int, []),
    "sgk_last_error": (c_char_p, []),
    "sgk_launch_count": (ctypes.c_longlong, []),
    "sgk_trace_kernels": (None, [ctypes.c_int]),
    "sgk_traced_kernels": (ctypes.c_char_p, []),
    "sgk_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "sgk_conv_packed_weight_elems": (c_size_t, [POINTER(SgkConvDesc), c_int]),
    "sgk_conv_pack_weight": (c_int, [POINTER(SgkConvDesc), c_int, P, P, P]),
    "sgk_conv_pack_weight_multi": (c_int, [POINTER(SgkPackJob), c_int, P]),
    "sgk_conv_fwd": (c_int, [POINTER(SgkConvDesc), P, P, P, P, c_int, c_float, P]),
    "sgk_conv_dgrad": (c_int, [POINTER(SgkConvDesc), P, P, P, P]),
    "sgk_conv_wgrad_workspace_bytes": (c_size_t, [POINTER(SgkConvDesc)]),
    "sgk_conv_wgrad": (c_int, [POINTER(SgkConvDesc), P, P, P, P, P, c_size_t, P]),
    "sgk_conv_wgrad_act": (c_int, [POINTER(SgkConvDesc), P, P, P, c_int, c_float, P, P, P, c_size_t, P]),
    "sgk_layout_nchw_to_nhwc": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "sgk_layout_nhwc_to_nchw": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "sgk_norm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sgk_norm_act_fwd": (c_int, [P, P, P, P, P, P, P, c_float, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                 P, c_size_t, P]),
    "sgk_norm_act_bwd": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, c_size_t, P]),
    "sgk_act_bwd": (c_int, [P, P, P, c_size_t, c_int, c_float, P]),
    "sgk_act_fwd": (c_int, [P, P, c_size_t, c_int, c_float, P]),
    "sgk_bias_grad": (c_int, [P, P, c_size_t, c_int, P, c_size_t, P]),
    "sgk_bias_grad_workspace_bytes": (c_size_t, [c_size_t, c_int]),
    "sgk_tap_rows": (c_int, []),
    "sgk_tap_weight_pack": (c_int, [P, P, c_int, c_int, c_int, P]),
    "sgk_tap_weight_unpack": (c_int, [P, P, c_int, c_int, c_int, P]),
    "sgk_tap_fold_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "sgk_tap_unfold": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_pad_nhwc": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_concat2_nhwc": (c_int, [P, c_int, P, c_int, P, c_size_t, P]),
    "sgk_split2_nhwc": (c_int, [P, P, c_int, P, c_int, c_size_t, P]),
    "sgk_axpy": (c_int, [P, P, c_float, P, c_size_t, P]),
    "sgk_mul": (c_int, [P, P, P, c_size_t, P]),
    "sgk_scale_by_dev_scalar": (c_int, [P, P, P, c_size_t, P]),
    "sgk_gauss_decimate_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_gauss_decimate_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_gauss_decimate_sep_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_bilinear_up2_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "sgk_bilinear_up2_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "sgk_avgpool_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_avgpool_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_loss_workspace_bytes": (c_size_t, [c_size_t]),
    "sgk_gan_loss": (c_int, [P, c_size_t, c_int, c_float, P, P, P, c_size_t, P]),
    "sgk_l1_loss": (c_int, [P, P, P, c_size_t, P, P, P, c_size_t, P]),
    "sgk_bce_pair_loss": (c_int, [P, P, c_size_t, P, P, P, c_size_t, P]),
    "sgk_image_pool_query": (c_int, [P, P, P, P, c_int, ctypes.c_longlong, c_int, P]),
    "sgk_image_transform_u8": (c_int, [P, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P, c_int, P]),
    "sgk_h2d_rows_async": (c_int, [P, c_size_t, P, c_size_t, c_size_t, c_size_t, P]),
    "sgk_l1_weight_map": (c_int, [P, P, c_int, c_int, ctypes.c_longlong, P, c_int, P]),
    "sgk_reflection_pad_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_reflection_pad_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "sgk_tensor2im_u8": (c_int, [P, P, c_int, c_int, c_int, P]),
    "sgk_ce_const_loss": (c_int, [P, c_int, c_int, ctypes.c_longlong, c_int, P, P, P, c_size_t, P]),
    "sgk_adam_multi_tensor": (c_int, [POINTER(SgkAdamTensor), c_int, P, P, P]),
    "sgk_multi_tensor_pack": (c_int, [POINTER(c_void_p), POINTER(c_int64), c_int, P, P]),
    "sgk_multi_tensor_unpack": (c_int, [P, POINTER(c_void_p), POINTER(c_int64), c_int, P]),
    "sgk_adam_block_elems": (c_int, []),
}

_lib = None


def load():
    """Loads libsgk.so (built in-tree by supervised-gan_b200/build.py).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libsgk.so not found at %s -- build it with `python supervised-gan_b200/build.py` "
                           "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.sgk_version() != 1:
        raise RuntimeError("libsgk.so version mismatch")
    _lib = lib
    return lib


EINVAL, EUNSUPPORTED, ECUDA, EWORKSPACE = -1, -2, -3, -4   # include/sgk.h SgkStatus


def check(rc, what=""):
    if rc != 0:
        msg = load().sgk_last_error()
        raise RuntimeError("libsgk %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
