"""ctypes binding of liblsq_b200.so (include/lsq_b200.h).  No torch types cross the boundary: only
device pointers, sizes and the current CUDA stream handle.  Missing library -> loud failure."""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# LSQ_B200_LIB: development override (kernel variants built next to the product library for A/B runs)
LIB_PATH = os.environ.get('LSQ_B200_LIB') or os.path.join(_HERE, 'liblsq_b200.so')
_lock = threading.Lock()
_lib = None

MAX_PLANES = 8

EXPORTS = [
    'lsq_abi_version', 'lsq_last_error', 'lsq_reduce_workspace_bytes', 'lsq_row_absmean', 'lsq_solve_v1',
    'lsq_fakequant', 'lsq_ste_backward', 'lsq_act_geometry', 'lsq_act_planes_bytes', 'lsq_encode_act',
    'lsq_wpack_bytes', 'lsq_pack_weights', 'lsq_bconv2d_fwd', 'lsq_bconv2d_tc_supported',
    'lsq_row_absmean_ex', 'lsq_solve_v1_ex', 'lsq_encode_act_ex', 'lsq_bconv2d_fwd_ex', 'lsq_stem_fwd',
    'lsq_stem_image_bytes', 'lsq_stem_workspace_bytes', 'lsq_stem_supported', 'lsq_stem_is_fused', 'lsq_stem_pack_weights', 'lsq_u8_expand', 'lsq_stem_fwd_u8', 'lsq_plane_mean',
    'lsq_pwconv_supported', 'lsq_pwconv_image_bytes', 'lsq_pwconv_pack_weights', 'lsq_pwconv_fwd',
    'lsq_solve_v1_multi', 'lsq_row_absmean_multi', 'lsq_wbits_bytes', 'lsq_unpack_weights',
    'lsq_quantize_act_workspace_bytes', 'lsq_quantize_act',
]


class ActGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ('n', 'c', 'h', 'w', 'kh', 'kw', 'stride', 'pad', 'ho', 'wo', 'cw', 'nphase', 'hv', 'wv', 'ph',
                 'pitch', 'rows_per_sample', 'lead')] + [('vtot', C.c_int64)]


class Prologue(C.Structure):
    _fields_ = [('d_ch_scale', C.c_void_p), ('d_ch_shift', C.c_void_p), ('channels', C.c_int32), ('inner', C.c_int64)]


class Epilogue(C.Structure):
    _fields_ = [('d_residual', C.c_void_p), ('d_prelu', C.c_void_p), ('n_prelu', C.c_int32), ('act', C.c_int32),
                ('residual_after_act', C.c_int32)]


class RowTensor(C.Structure):
    _fields_ = [('d_x', C.c_void_p), ('d_out', C.c_void_p), ('rows', C.c_int32), ('len', C.c_int32)]


class LsqError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise LsqError(
                    f'{LIB_PATH} is missing: the CUDA library is the only implementation of this path '
                    '(there is no CPU fallback). Build it with `python -m ml_quant_b200.build`.')
            L = C.CDLL(LIB_PATH)
            vp, f32, i64, i32, sz = C.c_void_p, C.c_float, C.c_int64, C.c_int, C.c_size_t
            gp = C.POINTER(ActGeom)
            L.lsq_abi_version.restype = i32
            L.lsq_last_error.restype = C.c_char_p
            L.lsq_reduce_workspace_bytes.restype = sz
            L.lsq_reduce_workspace_bytes.argtypes = [i64, i64]
            L.lsq_row_absmean.argtypes = [vp, i64, i64, f32, vp, i32, vp, vp, sz, vp]
            L.lsq_solve_v1.argtypes = [vp, i64, i64, i32, i32, f32, vp, vp, vp]
            L.lsq_fakequant.argtypes = [vp, i64, i64, f32, vp, i32, i32, vp, vp]
            L.lsq_ste_backward.argtypes = [vp, vp, vp, i64, vp]
            L.lsq_act_geometry.argtypes = [i32] * 8 + [gp]
            L.lsq_act_planes_bytes.restype = sz
            L.lsq_act_planes_bytes.argtypes = [gp, i32]
            L.lsq_encode_act.argtypes = [vp, gp, f32, vp, i32, i32, vp, vp, vp, sz, vp]
            L.lsq_wpack_bytes.restype = sz
            L.lsq_wpack_bytes.argtypes = [i32] * 4
            L.lsq_pack_weights.argtypes = [vp, i32, i32, i32, i32, vp, vp]
            L.lsq_bconv2d_fwd.argtypes = [vp, gp, i32, vp, vp, vp, vp, i32, vp, i32, vp]
            L.lsq_bconv2d_tc_supported.argtypes = [gp, i32, i32]
            pp, ep = C.POINTER(Prologue), C.POINTER(Epilogue)
            L.lsq_row_absmean_ex.argtypes = [vp, i64, i64, f32, vp, i32, vp, vp, sz, pp, vp]
            L.lsq_solve_v1_ex.argtypes = [vp, i64, i64, i32, i32, f32, vp, vp, pp, vp]
            L.lsq_encode_act_ex.argtypes = [vp, gp, f32, vp, i32, i32, vp, vp, vp, sz, pp, vp]
            L.lsq_bconv2d_fwd_ex.argtypes = [vp, gp, i32, vp, vp, vp, vp, i32, vp, i32, ep, vp]
            L.lsq_stem_fwd.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
            L.lsq_stem_fwd.restype = i32
            L.lsq_stem_image_bytes.restype = sz
            L.lsq_stem_image_bytes.argtypes = []
            L.lsq_stem_workspace_bytes.restype = sz
            L.lsq_stem_workspace_bytes.argtypes = [i32, i32, i32]
            L.lsq_stem_supported.restype = i32
            L.lsq_stem_is_fused.restype = i32
            L.lsq_plane_mean.restype = i32
            L.lsq_plane_mean.argtypes = [vp, i64, i32, vp, vp]
            L.lsq_u8_expand.restype = i32
            L.lsq_u8_expand.argtypes = [vp, i64, i32, i64, vp, vp, vp]
            L.lsq_stem_fwd_u8.restype = i32
            L.lsq_stem_fwd_u8.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
            L.lsq_stem_is_fused.argtypes = [i32, i32, i32]
            L.lsq_stem_supported.argtypes = [i32, i32, i32]
            L.lsq_stem_pack_weights.restype = i32
            L.lsq_stem_pack_weights.argtypes = [vp, vp, vp]
            L.lsq_pwconv_supported.restype = i32
            L.lsq_pwconv_supported.argtypes = [i32, i32]
            L.lsq_pwconv_image_bytes.restype = sz
            L.lsq_pwconv_image_bytes.argtypes = [i32, i32]
            L.lsq_pwconv_pack_weights.restype = i32
            L.lsq_pwconv_pack_weights.argtypes = [vp, i32, i32, vp, vp]
            L.lsq_pwconv_fwd.restype = i32
            L.lsq_pwconv_fwd.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp]
            L.lsq_wbits_bytes.restype = sz
            L.lsq_wbits_bytes.argtypes = [i32] * 4
            L.lsq_unpack_weights.restype = i32
            L.lsq_unpack_weights.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
            tp = C.POINTER(RowTensor)
            L.lsq_solve_v1_multi.restype = i32
            L.lsq_solve_v1_multi.argtypes = [tp, i32, i32, i32, f32, vp]
            L.lsq_row_absmean_multi.restype = i32
            L.lsq_row_absmean_multi.argtypes = [tp, i32, f32, vp]
            L.lsq_quantize_act_workspace_bytes.restype = sz
            L.lsq_quantize_act_workspace_bytes.argtypes = [gp]
            L.lsq_quantize_act.restype = i32
            L.lsq_quantize_act.argtypes = [vp, gp, f32, i32, i32, vp, vp, vp, sz, pp, vp, vp]
            for name in ('lsq_row_absmean', 'lsq_solve_v1', 'lsq_fakequant', 'lsq_ste_backward', 'lsq_act_geometry',
                         'lsq_encode_act', 'lsq_pack_weights', 'lsq_bconv2d_fwd', 'lsq_bconv2d_tc_supported',
                         'lsq_row_absmean_ex', 'lsq_solve_v1_ex', 'lsq_encode_act_ex', 'lsq_bconv2d_fwd_ex'):
                getattr(L, name).restype = i32
            if L.lsq_abi_version() != 1:
                raise LsqError('liblsq_b200.so ABI version mismatch')
            _lib = L
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().lsq_last_error().decode('utf-8', 'replace')
        raise LsqError(f'{what} failed ({status}): {msg}')
