"""ctypes binding of libb200aug.so (include/b200aug.h).  There is no fallback: if the CUDA library is missing or
fails to load, importing this module raises, and so does every transform that needs it."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200AUG_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libb200aug.so")  # (env: experiment builds)

ABI_VERSION = 7
MAX_FIELDS, NUM_OPS, NUM_NOISE = 8, 6, 4

F_HALF_PIXEL, F_ROI_FROM_LANDMARKS, F_FOCUS, F_FLIPROT, F_NORMALIZE, F_PHOTOMETRIC, F_WHITEN, F_INSERT_BACKTRANSFORM = 1, 2, 4, 8, 16, 32, 64, 128
CAT_GENERAL, CAT_QUAT, CAT_XYS, CAT_ROI, CAT_POINTS, CAT_BACKTRANSFORM = 0, 1, 2, 3, 4, 5
PHASE_ALL, PHASE_PLAN, PHASE_MAIN = 0, 1, 2
DOWN_AREA, DOWN_GAUSSIAN, DOWN_HAMMING = 0, 1, 2
UP_LINEAR, UP_CUBIC, UP_LANCZOS = 0, 1, 2
PREFILTER_MAX_TAPS = 63
S_OK, S_EMPTY_BOX, S_UNSUPPORTED, S_ROWBUF = 0, 1, 2, 3
OP_EQUALIZE, OP_POSTERIZE, OP_GAMMA, OP_CONTRAST, OP_BRIGHTNESS, OP_BLUR = range(6)


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32), ("reserved", C.c_int32)]


class Field(C.Structure):
    _fields_ = [("category", C.c_int32), ("count", C.c_int32), ("dim", C.c_int32), ("reserved", C.c_int32),
                ("inp", C.c_void_p), ("out", C.c_void_p)]


class PhotoParams(C.Structure):
    _fields_ = [("n_order", C.c_int32), ("order", C.c_int32 * NUM_OPS), ("clip", C.c_int32),
                ("apply", C.c_void_p), ("bits", C.c_void_p), ("gamma", C.c_void_p), ("contrast", C.c_void_p),
                ("brightness", C.c_void_p), ("noise_apply", C.c_void_p), ("noise_std", C.c_float * NUM_NOISE), ("noise_clip", C.c_int32 * NUM_NOISE),
                ("seed", C.c_uint64), ("sample_offset", C.c_uint64)]


class FusedArgs(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("batch", C.c_int32), ("out_w", C.c_int32), ("out_h", C.c_int32),
                ("flags", C.c_uint32), ("rowbuf_capacity", C.c_int32),
                ("src_table", C.c_void_p), ("src_uniform", Src), ("src_stride", C.c_int64),
                ("scales", C.c_void_p), ("angles", C.c_void_p), ("cos_sin", C.c_void_p), ("translations", C.c_void_p),
                ("beyond_border_shift", C.c_float), ("explicit_view_roi", C.c_void_p), ("explicit_tr", C.c_void_p), ("do_flip", C.c_void_p), ("rot_dir", C.c_void_p),
                ("n_fields", C.c_int32), ("roi_field", C.c_int32), ("landmark_field", C.c_int32), ("cluster_size", C.c_int32),
                ("fields", Field * MAX_FIELDS),
                ("view_roi_out", C.c_void_p), ("tr_out", C.c_void_p), ("backtransform_out", C.c_void_p),
                ("image_u8_out", C.c_void_p), ("image_f32_out", C.c_void_p), ("status_out", C.c_void_p),
                ("trace_out", C.c_void_p), ("order", C.c_void_p), ("workspace", C.c_void_p), ("workspace_stride", C.c_int64),
                ("plans", C.c_void_p), ("plan_stride", C.c_int64), ("warp_ctas", C.c_int32), ("phase", C.c_int32),
                ("downfilter", C.c_int32), ("upfilter", C.c_int32), ("hamming_taps", C.c_void_p), ("hamming_sym_mask", C.c_uint64),
                ("remap_tabs", C.c_void_p),
                ("photo", PhotoParams)]


EXPORTS = ("b200aug_abi_version", "b200aug_strerror", "b200aug_last_cuda_error", "b200aug_fused_smem_bytes", "b200aug_fused_occupancy",
           "b200aug_workspace_stride", "b200aug_plan_stride", "b200aug_plan_buffer_bytes", "b200aug_hamming_table", "b200aug_remap_table", "b200aug_upload_row_bands", "b200aug_upload_boxes", "b200aug_fused_forward", "b200aug_apply_affine2d", "b200aug_photometric_f32", "b200aug_corrected_rotation",
           "b200aug_quat_matrix", "b200aug_head_roi", "b200aug_jpeg_info", "b200aug_decode_jpeg_gray", "b200aug_jpeg_last_status", "b200aug_jpeg_backend")


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} is missing: build it with `python neuralnet-tracker-traincode_b200/build.py` "
                          "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    lib.b200aug_abi_version.restype = C.c_int
    lib.b200aug_strerror.restype = C.c_char_p
    lib.b200aug_strerror.argtypes = [C.c_int]
    lib.b200aug_last_cuda_error.restype = C.c_int
    lib.b200aug_fused_smem_bytes.restype = C.c_size_t
    lib.b200aug_fused_smem_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.b200aug_workspace_stride.restype = C.c_int64
    lib.b200aug_workspace_stride.argtypes = [C.c_int]
    lib.b200aug_plan_stride.restype = C.c_int64
    lib.b200aug_plan_stride.argtypes = [C.c_int, C.c_int]
    lib.b200aug_plan_buffer_bytes.restype = C.c_int64
    lib.b200aug_plan_buffer_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.b200aug_upload_row_bands.restype = C.c_int
    lib.b200aug_upload_boxes.restype = C.c_int
    lib.b200aug_upload_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.b200aug_upload_row_bands.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b200aug_fused_forward.restype = C.c_int
    lib.b200aug_fused_forward.argtypes = [C.POINTER(FusedArgs), C.c_void_p]
    lib.b200aug_apply_affine2d.restype = C.c_int
    lib.b200aug_apply_affine2d.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(Field), C.c_void_p]
    lib.b200aug_photometric_f32.restype = C.c_int
    lib.b200aug_photometric_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(PhotoParams),
                                            C.c_float, C.c_void_p]
    lib.b200aug_corrected_rotation.restype = C.c_int
    lib.b200aug_corrected_rotation.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int64, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200aug_quat_matrix.restype = C.c_int
    lib.b200aug_quat_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200aug_jpeg_info.restype = C.c_int
    lib.b200aug_jpeg_info.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.b200aug_decode_jpeg_gray.restype = C.c_int
    lib.b200aug_decode_jpeg_gray.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b200aug_jpeg_last_status.restype = C.c_int
    lib.b200aug_jpeg_backend.restype = C.c_int
    lib.b200aug_hamming_table.restype = C.c_int
    lib.b200aug_fused_occupancy.restype = C.c_int
    lib.b200aug_fused_occupancy.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.b200aug_remap_table.restype = C.c_int
    lib.b200aug_remap_table.argtypes = [C.c_int, C.c_void_p]
    lib.b200aug_head_roi.restype = C.c_int
    lib.b200aug_head_roi.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                     C.c_void_p, C.c_int32, C.c_void_p]
    lib.b200aug_hamming_table.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    if lib.b200aug_abi_version() != ABI_VERSION:
        raise NativeError(f"ABI mismatch: library {lib.b200aug_abi_version()} vs binding {ABI_VERSION}")
    return lib


lib = _load()


def check(rc: int, what: str):
    if rc != 0:
        msg = lib.b200aug_strerror(rc).decode()
        extra = f" (cudaError {lib.b200aug_last_cuda_error()})" if rc == 4 else ""
        raise NativeError(f"{what}: {msg}{extra}")
