"""ctypes binding of libarpde.so (the C ABI declared in include/arpde.h).

There is deliberately no fallback: if the shared library is missing, or a call is made with
tensors that are not on a CUDA device, this module raises.  The test oracle is never used here.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libarpde.so')

if not os.path.exists(LIB_PATH):
    raise ImportError(
        'libarpde.so not found at %s -- build it with `bash ar-pde-cnn_b200/csrc/build.sh` '
        '(or __graft_entry__.build()); there is no CPU/PyTorch fallback path.' % LIB_PATH)

lib = C.CDLL(LIB_PATH)

f32p, f64p, i64p, vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p   # raw device/host addresses
i32, i64, f64 = C.c_int, C.c_int64, C.c_double


class Op(C.Structure):
    """arpde_op_t"""
    _fields_ = [(n, C.c_int32) for n in ('in_buf', 'out_buf', 'cin', 'cout', 'kh', 'kw', 'stride', 'up',
                                         'out_coff', 'act', 'w_param', 'bn_param', 'bn_off')]


class Buf(C.Structure):
    """arpde_buf_t"""
    _fields_ = [('offset', C.c_int64), ('stats_off', C.c_int64), ('ctot', C.c_int32), ('h', C.c_int32),
                ('w', C.c_int32), ('pad_', C.c_int32)]


_CN2D = [f32p, i64, f32p, i64, f32p, f64p, i32, i32, i32, f64, f64, f64, i32, vp]
_PROTOS = {
    'arpde_version': [],
    'arpde_cn_burgers2d_fwd': _CN2D,
    'arpde_cn_burgers2d_fwd_generic': _CN2D,
    'arpde_cn_1d_fwd': [i32, f32p, i64, f32p, i64, f32p, f64p, i32, i32, f64, f64, f64, i32, i32, i32, i32, vp],
    'arpde_cn_burgers2d_bwd': [f32p, i64, f32p, i64, f32p, f32p, f32p, i32, i32, i32, f64, f64, f64, i32, vp],
    'arpde_cn_burgers2d_bwd_generic': [f32p, i64, f32p, i64, f32p, f32p, f32p, i32, i32, i32, f64, f64, f64, i32, vp],
    'arpde_cn_1d_bwd': [i32, f32p, i64, f32p, i64, f32p, f32p, f32p, i32, i32, f64, f64, f64, i32, i32, i32, i32, vp],
    'arpde_stencil_bulk_plan': [i32, i32, i32, i32, i32, vp, vp, vp, vp],
    'arpde_filter1d': [f32p, f32p, i64, i32, vp, vp],
    'arpde_filter2d': [f32p, f32p, i64, i32, i32, vp, vp],
    'arpde_neg_log_joint_fwd': [f32p, f32p, i64, i32, f32p, vp, vp, i32, f64, f64, f64, f64, f64, f32p, f64p, vp],
    'arpde_neg_log_joint_bwd': [f32p, f32p, i64, i32, f32p, vp, vp, i32, f64, f64, f64, f64, f64, f32p, f64p,
                                f32p, f32p, f32p, f32p, vp],
    'arpde_conv_fwd': [f32p, i64, f32p, i64, f32p, f32p, i64, f32p, f64p] + [i32] * 14 + [vp],
    'arpde_bn_stats': [f32p, i64, i32, i32, i64, f64p, vp],
    'arpde_bn_finalize_train': [f64p, f64, f32p, f32p, f32p, f32p, i64p, f64, f64, f32p, f32p, f32p, f32p, i32, vp],
    'arpde_bn_fold_eval': [f32p, f32p, f32p, f32p, f64, f32p, f32p, i32, vp],
    'arpde_affine_act': [f32p, i64, f32p, f32p, f32p, i32, i32, i32, i32, i32, i32, i32, vp],
    'arpde_densed_forward': [vp, i32, vp, i32, vp, vp, i32, f32p, i64, i32, i32, i32, i32, f32p, i32, i32, f32p, f32p,
                             i32, f64p, i64, f32p, i32, f64, f64, f32p, i64, vp],
    'arpde_conv_fwd_tc': [f32p, i64, f32p, i64, f32p, f32p, i64, f32p, f64p] + [i32] * 14 + [f32p, i64, vp],
    'arpde_densed_backward': [vp, i32, vp, i32, vp, i32, f32p, i64, i32, i32, i32, f32p, f32p, f32p, f32p, i32, f32p,
                              f32p, f32p, vp, f32p, f32p, f32p, f64p, i32, f32p, i64, vp],
    'arpde_densed_forward_dp': [vp, i32, vp, i32, vp, vp, i32, f32p, i64, i32, i32, i32, i32, f32p, i32, i32, f32p, f32p,
                                i32, f64p, i64, f32p, i32, f64, f64, f32p, i64, vp, vp, i32, vp],
    'arpde_densed_backward_dp': [vp, i32, vp, i32, vp, i32, f32p, i64, i32, i32, i32, f32p, f32p, f32p, f32p, i32, f32p,
                                 f32p, f32p, vp, f32p, f32p, f32p, f64p, i32, f32p, i64, vp, vp, i32, vp],
    'arpde_conv_tc_plan': [i32] * 8 + [vp],
    'arpde_error_metrics': [vp, i64, i64, vp, i64, i64, i32, i32, i64, i64, vp, vp, vp],
    'arpde_adam_step': [vp, vp, vp, vp, vp, vp, vp, vp, i32, f64, f64, f64, vp],
    'arpde_fourier_ic2d': [vp, vp, vp, i32, i32, i32, vp],
    'arpde_conv_wgrad': [vp, i64, vp, vp, vp, i64, vp, vp] + [i32] * 10 + [vp, i64, i32, vp],
    'arpde_conv_wgrad_tc': [vp, i64, vp, vp, vp, i64, vp, vp] + [i32] * 10 + [vp, i64, vp],
    'arpde_conv_wgrad_tc_plan': [i32] * 8 + [vp],
    'arpde_rollout': [vp, i32, vp, i32, vp, vp, i32, f32p, i32, i32, i32, i32, i32, i32, i32, i32, i32, f32p, f32p,
                      i32, f64, f32p, i64, vp],
    'arpde_rollout_layerwise': [vp, i32, vp, i32, vp, vp, i32, f32p, i32, i32, i32, i32, i32, i32, i32, i32, i32, f32p, f32p,
                                i32, f64, f32p, i64, vp],
    'arpde_fd_burgers1d': [vp, vp, i32, i32, f64, f64, f64, i32, i32, i32, i32, vp],
    'arpde_peer_allreduce_f64': [vp, vp, i64],
    'arpde_stem_pair_fwd': [vp, i64, vp, i64, vp, vp, i64, vp, i64, vp] + [i32] * 10 + [vp, i64, vp, i64, vp],
    'arpde_stem_pair_producer': [vp, i64, vp, i64, vp, vp, i64] + [i32] * 7 + [vp, i64, vp, i64, vp],
    'arpde_stem_pair_consumer': [vp, vp, i64, vp] + [i32] * 8 + [vp, i64, vp],
    'arpde_conv_fwd_h': [vp, i64, vp, i64, vp, vp, i64, vp] + [i32] * 11 + [vp, i64, vp],
    'arpde_swag_sample': [vp, vp, vp, vp, i32, f32p, f32p, f32p, i32, i32, i32, f64, f64, i32, vp],
    'arpde_predictive_moments': [f32p, i64, i64, i32, i32, f64, f32p, f32p, vp],
}
for _name, _args in _PROTOS.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = C.c_int
lib.arpde_last_error.restype = C.c_char_p
lib.arpde_last_error.argtypes = []
# size queries return a count, not a status
lib.arpde_densed_tc_scratch_floats.argtypes = [vp, i32, i32]
lib.arpde_densed_tc_scratch_floats.restype = C.c_int64
lib.arpde_conv_fwd_tc_scratch_floats.argtypes = [i32, i32, i32, i32, i32]
lib.arpde_conv_fwd_tc_scratch_floats.restype = C.c_int64
lib.arpde_launch_count.argtypes = [i32]
lib.arpde_launch_count.restype = C.c_int64
lib.arpde_conv_wgrad_tc_scratch.argtypes = [i32] * 9
lib.arpde_conv_wgrad_tc_scratch.restype = C.c_int64
lib.arpde_conv_z_floats.argtypes = [i32, i32, i32]
lib.arpde_conv_z_floats.restype = C.c_int64
lib.arpde_conv_fwd_h_scratch_floats.argtypes = [i32] * 5
lib.arpde_conv_fwd_h_scratch_floats.restype = C.c_int64

EXPORTS = sorted(list(_PROTOS) + ['arpde_last_error', 'arpde_launch_count', 'arpde_densed_tc_scratch_floats', 'arpde_conv_fwd_tc_scratch_floats',
                                   'arpde_conv_wgrad_tc_scratch', 'arpde_conv_z_floats', 'arpde_conv_fwd_h_scratch_floats'])


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError('libarpde %s failed (code %d): %s' % (what, rc, lib.arpde_last_error().decode()))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('arpde_b200 runs on CUDA devices only (got a %s tensor); there is no CPU path' % t.device)


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    check(getattr(lib, name)(*args), name)


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)     # arpde_allreduce_fn
PEER_MAX = 8


class PeerCtx(C.Structure):
    """arpde_peer_ctx_t"""
    _fields_ = [('slots', C.c_void_p * PEER_MAX), ('flags', C.c_void_p * PEER_MAX), ('world', C.c_int32), ('rank', C.c_int32),
                ('nslots', C.c_int32), ('slot_doubles', C.c_int32), ('seq', C.c_uint64), ('error_flag', C.c_void_p), ('stream', C.c_void_p)]


def allreduce_callback(sync, *tensors):
    """ctypes callback for arpde_densed_*_dp: `sync.all_reduce(view)` on the float64 slice of one of `tensors` that the
    library points at (the library only ever hands back addresses inside the scratch tensors it was given)."""
    spans = [(t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), t) for t in tensors if t is not None and t.dtype == torch.float64]

    def cb(user, buf, n):
        try:
            for lo, hi, t in spans:
                if lo <= buf and buf + 8 * n <= hi:
                    off = (buf - lo) // 8
                    sync.all_reduce(t.view(-1)[off:off + n])
                    return 0
            return 1
        except Exception:          # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return 2
    return ALLREDUCE_FN(cb)


def ptr_array(ptrs):
    """host array of addresses (void* const*)"""
    return (C.c_void_p * len(ptrs))(*ptrs)


def i64_array(vals):
    return (C.c_int64 * len(vals))(*vals)
