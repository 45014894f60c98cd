"""ctypes binding of lib/libkpd_b200.so -- the only compute backend of this package.

There is deliberately NO fallback: if the shared library is missing or a CUDA device of compute capability
10.x is not present, the hot-path entry points raise.  Declarations mirror include/kpd_b200.h.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KPD_B200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libkpd_b200.so"))

ABI_VERSION = 1
NUM_STATE_TENSORS = 318
PREC_TF32, PREC_BF16, PREC_FP32_SIMT = 0, 1, 2
PRECISIONS = {"fp32": PREC_TF32, "tf32": PREC_TF32, "bf16": PREC_BF16, "fp32_strict": PREC_FP32_SIMT}


class NativeLibraryError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("in_channels", C.c_int32), ("num_keypoints", C.c_int32),
                ("reserved", C.c_int32)]


class WorkspaceLayout(C.Structure):
    _fields_ = [("total_bytes", C.c_size_t),
                ("batch", C.c_int32), ("size", C.c_int32), ("max_rois", C.c_int32), ("precision", C.c_int32),
                ("feat", C.c_size_t * 4), ("lateral0", C.c_size_t), ("fpn0", C.c_size_t),
                ("pool_avg", C.c_size_t), ("pool_max", C.c_size_t), ("scores", C.c_size_t), ("topk", C.c_size_t),
                ("roi_count", C.c_size_t), ("roi_box", C.c_size_t), ("roi_batch", C.c_size_t),
                ("roi_slot", C.c_size_t), ("img_count", C.c_size_t), ("roi_feat", C.c_size_t),
                ("head_cw", C.c_size_t), ("head_sw", C.c_size_t), ("head_in", C.c_size_t),
                ("head_a", C.c_size_t), ("head_b", C.c_size_t), ("head_c", C.c_size_t),
                ("scratch", C.c_size_t), ("elt_bytes", C.c_size_t)]


_P = C.c_void_p
_SIGNATURES = {
    "kpd_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(ModelDesc), C.POINTER(_P), C.c_int, C.c_int]),
    "kpd_plan_destroy": (None, [_P]),
    "kpd_plan_workspace_layout": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(WorkspaceLayout)]),
    "kpd_forward": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "kpd_backbone_fpn0": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P, C.POINTER(WorkspaceLayout), _P]),
    "kpd_fpn_levels123": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(WorkspaceLayout), C.POINTER(_P), _P]),
    "kpd_channel_topk": (C.c_int, [_P, C.c_int, _P, C.POINTER(WorkspaceLayout), _P]),
    "kpd_channel_attention": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "kpd_roi_head": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P,
                               C.POINTER(WorkspaceLayout), _P]),
    "kpd_roi_align": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "kpd_heatmap_head": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _P, _P, C.POINTER(WorkspaceLayout), _P]),
    "kpd_dense_conv": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "kpd_decode": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float,
                             _P, _P, _P, _P, _P, _P, _P, _P]),
    "kpd_person_heads": (C.c_int, [_P, C.c_int, _P, C.c_int64, _P, _P, _P]),
    "kpd_box_iou": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "kpd_nms": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, _P]),
    "kpd_detect_persons": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_float, C.c_float,
                                     C.c_int, C.c_int, _P, _P, _P, _P]),
    "kpd_detect_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "kpd_itransform_scratch_bytes": (C.c_size_t, [C.c_int] * 7),
    "kpd_itransform_stage_offsets": (C.c_int, [C.c_int] * 7 + [C.POINTER(C.c_size_t)]),
    "kpd_itransform": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "kpd_conv2d": (C.c_int, [_P] * 6 + [C.c_int] * 11 + [_P]),
    "kpd_adaptive_avgpool": (C.c_int, [_P, _P] + [C.c_int] * 6 + [_P]),
    "kpd_linear": (C.c_int, [_P] * 4 + [C.c_int] * 4 + [_P]),
    "kpd_layernorm": (C.c_int, [_P, _P, _P, C.c_float, _P, C.c_int, C.c_int, C.c_int, _P]),
    "kpd_mul_map": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, _P]),
    "kpd_upsample_add": (C.c_int, [_P, _P] + [C.c_int] * 6 + [_P]),
    "kpd_last_error": (C.c_char_p, []),
    "kpd_abi_version": (C.c_int, []),
    "kpd_launch_count": (C.c_int64, [C.c_int]),
    "kpd_profile_enable": (C.c_int, [C.c_int]),
    "kpd_profile_read": (C.c_int, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load(path=None):
    """dlopen the library once and type its entry points.  Raises NativeLibraryError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise NativeLibraryError(
            f"{path} not found: build it with `python keypoint-detection_b200/build.py` "
            "(this package has no CPU or PyTorch fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI mismatch, let it propagate
        fn.restype, fn.argtypes = res, args
    if lib.kpd_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI version mismatch: library {lib.kpd_abi_version()}, binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().kpd_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def ptr(t):
    """device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise NativeLibraryError("a CUDA device (B200, sm_100) is required: this package has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise NativeLibraryError(f"tensors must live on a CUDA device, got {dev}")
    return dev


class Plan:
    """Owns a kpd_plan (packed, BN-folded weights on one device)."""

    def __init__(self, state_tensors, in_channels, num_keypoints, device_index):
        lib = load()
        assert len(state_tensors) == NUM_STATE_TENSORS, len(state_tensors)
        host = [t.detach().to("cpu").contiguous() for t in state_tensors]
        host = [t if t.dtype in (torch.float32, torch.int64) else t.float() for t in host]
        arr = (_P * len(host))(*[C.c_void_p(t.data_ptr()) for t in host])
        desc = ModelDesc(ABI_VERSION, in_channels, num_keypoints, 0)
        handle = _P()
        check(lib.kpd_plan_create(C.byref(handle), C.byref(desc), arr, len(host), device_index), "kpd_plan_create")
        self._h = handle
        self.device_index = device_index
        self._layouts = {}

    def layout(self, batch, size, max_rois, precision):
        key = (batch, size, max_rois, precision)
        if key not in self._layouts:
            L = WorkspaceLayout()
            check(load().kpd_plan_workspace_layout(self._h, batch, size, max_rois, precision, C.byref(L)),
                  "kpd_plan_workspace_layout")
            self._layouts[key] = L
        return self._layouts[key]

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.kpd_plan_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
