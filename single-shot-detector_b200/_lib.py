"""ctypes binding of libssdk.so (include/ssdk.h) -- the thin C-ABI layer under the Python mirror.

There is NO fallback: if the shared library has not been built, or no CUDA device is present, every
compute call raises.  Device memory, streams and (for multi-GPU) torch.distributed come from PyTorch;
all arithmetic happens in the hand-written sm_100a kernels of csrc/.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SSDK_LIB') or os.path.join(_HERE, 'lib', 'libssdk.so')   # SSDK_LIB: an alternative build (tuning experiments)

c_int, c_i64, c_double, c_void_p = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
P = c_void_p  # every tensor argument is passed as a raw address

# name -> argtypes (after the leading ctx pointer where has_ctx); mirrors include/ssdk.h declaration by declaration
_SIGNATURES = {
    'ssdk_version': (None, []),
    'ssdk_last_error': (ctypes.c_char_p, []),
    'ssdk_ctx_create': (c_int, [c_int, P, ctypes.POINTER(P)]),
    'ssdk_ctx_set_stream': (c_int, [P, P]),
    'ssdk_ctx_set_option': (c_int, [P, c_int, c_int]),
    'ssdk_ctx_destroy': (c_int, [P]),
    'ssdk_ctx_workspace_bytes': (c_i64, [P]),
    'ssdk_ctx_launch_count': (c_i64, [P]),
    'ssdk_ctx_synchronize': (c_int, [P]),
    'ssdk_ctx_async_error': (c_int, [P, ctypes.POINTER(c_int)]),
    'ssdk_ctx_round_times': (c_int, [P, P, c_int]),
    'ssdk_ctx_set_profiling': (c_int, [P, c_int]),
    'ssdk_ctx_profile_read': (c_int, [P, P, P, c_int, c_int]),
    'ssdk_num_anchors': (c_int, [c_int, c_int, P, c_int, c_int, ctypes.POINTER(c_i64), P]),
    'ssdk_anchors': (c_int, [P, c_int, c_int, P, P, P, c_int, c_int, P, P]),
    'ssdk_area': (c_int, [P, P, c_i64, P]),
    'ssdk_intersection': (c_int, [P, P, c_i64, P, c_i64, P]),
    'ssdk_iou': (c_int, [P, P, c_i64, P, c_i64, P]),
    'ssdk_encode': (c_int, [P, P, P, c_i64, P]),
    'ssdk_decode': (c_int, [P, P, P, c_i64, P]),
    'ssdk_batch_decode': (c_int, [P, P, P, c_i64, c_i64, P]),
    'ssdk_match_boxes': (c_int, [P, P, c_i64, P, P, c_int, c_int, c_double, c_double, c_int, P]),
    'ssdk_create_targets': (c_int, [P, P, c_i64, P, P, c_int, c_int, P, P, P]),
    'ssdk_training_targets': (c_int, [P, P, c_i64, P, P, P, c_int, c_int, c_double, c_double, P, P, P]),
    'ssdk_training_targets_count': (c_int, [P, P, c_i64, P, P, P, c_int, c_int, c_double, c_double, P, P, P, P]),
    'ssdk_localization_loss': (c_int, [P, P, P, P, c_i64, c_i64, P]),
    'ssdk_focal_loss': (c_int, [P, P, P, P, c_i64, c_i64, c_int, c_double, c_double, P]),
    'ssdk_ssd_loss': (c_int, [P, P, P, P, P, P, c_i64, c_i64, c_int, c_double, c_double, P, P, P]),
    'ssdk_loss_finalize': (c_int, [P, P, P]),
    'ssdk_ssd_loss_forward_backward': (c_int, [P, P, P, P, P, P, c_i64, c_i64, c_int, c_double, c_double, P, P, P, P, P]),
    'ssdk_count_matches': (c_int, [P, P, c_i64, P]),
    'ssdk_ssd_loss_backward': (c_int, [P, P, P, P, P, P, c_i64, c_i64, c_int, c_double, c_double, P, P, P, P]),
    'ssdk_ssd_targets_and_loss': (c_int, [P, P, P, P, P, P, P, c_int, c_i64, c_int, c_int, c_double, c_double,
                                          c_double, c_double, P, P, P, P, P, P]),
    'ssdk_ssd_loss_step': (c_int, [P, P, P, P, P, P, P, c_int, c_i64, c_int, c_int, c_double, c_double, c_double, c_double,
                                   c_int, P, P, P, P, P]),
    'ssdk_head_ssd_loss_step': (c_int, [P, P, P, P, P, P, c_int, c_i64, c_int, c_int, c_double, c_double, c_double, c_double,
                                        c_int, P, P, P, P, P]),
    'ssdk_ssd_targets_and_loss_host': (c_int, [P, P, P, P, P, P, P, c_int, c_i64, c_int, c_int, c_double, c_double,
                                               c_double, c_double, P, P]),
    'ssdk_postprocess': (c_int, [P, P, P, P, c_int, c_int, c_i64, c_int, c_double, c_double, c_int, P, P, P, P, P]),
    'ssdk_detect': (c_int, [P, P, P, P, c_int, c_int, c_i64, c_int, c_double, c_double, c_int, P, c_double, P, P, P, P]),
    'ssdk_detect_coco': (c_int, [P, P, P, P, c_int, c_int, c_i64, c_int, c_double, c_double, c_int, P, c_double, P, P, P, P, P, P, P, P, P, P]),
    'ssdk_detect_by_label': (c_int, [P, P, P, P, c_int, c_int, c_i64, c_int, c_double, c_double, c_int, P, c_double, P, P, P, P, P]),
    'ssdk_postprocess_host': (c_int, [P, P, P, P, c_int, c_int, c_i64, c_int, c_double, c_double, c_int, P, P, P, P]),
    'ssdk_head_concat': (c_int, [P, P, c_int, c_int, P, P]),
    'ssdk_head_ssd_loss': (c_int, [P, P, P, P, P, c_int, c_i64, c_int, c_double, c_double, P]),
    'ssdk_head_ssd_targets_and_loss': (c_int, [P, P, P, P, P, P, c_int, c_i64, c_int, c_int, c_double, c_double, c_double, c_double, P, P, P, P]),
    'ssdk_head_ssd_loss_forward_backward': (c_int, [P, P, P, P, P, c_int, c_i64, c_int, c_double, c_double, P, P, P, P]),
    'ssdk_head_detect': (c_int, [P, P, P, c_int, c_int, c_i64, c_int, c_double, c_double, c_int, P, c_double, P, P, P, P, P]),
    'ssdk_level_summaries': (c_int, [P, P, P, c_int, c_i64, P, c_int, c_double, P, P, P]),
    'ssdk_ioa': (c_int, [P, P, c_i64, P, c_i64, P]),
    'ssdk_change_coordinate_frame': (c_int, [P, P, c_i64, P, P]),
    'ssdk_prune_completely_outside_window': (c_int, [P, P, c_i64, P, P, P, P]),
    'ssdk_prune_non_overlapping_boxes': (c_int, [P, P, c_i64, P, c_i64, c_double, P, P, P]),
    'ssdk_crop_boxes': (c_int, [P, P, P, P, c_int, c_int, c_double, P, P, P]),
    'ssdk_comm_local_handle': (c_int, [P, P]),
    'ssdk_comm_connect': (c_int, [P, c_int, c_int, P]),
    'ssdk_comm_world': (c_int, [P]),
    'ssdk_comm_all_reduce_sum': (c_int, [P, P, c_int]),
    'ssdk_comm_loss_finalize': (c_int, [P, P, P]),
    'ssdk_comm_error': (c_int, [P, ctypes.POINTER(c_i64)]),
    'ssdk_comm_disconnect': (c_int, [P]),
}

SSDK_MAX_LEVELS = 8
SSDK_CHANNELS_LAST, SSDK_CHANNELS_FIRST = 0, 1


class SsdkHead(ctypes.Structure):
    """struct ssdk_head of include/ssdk.h: per-level tower outputs (detector/box_predictor.py:67-104)."""
    _fields_ = [('num_levels', ctypes.c_int32), ('anchors_per_location', ctypes.c_int32), ('data_format', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('height', ctypes.c_int32 * SSDK_MAX_LEVELS),
                ('width', ctypes.c_int32 * SSDK_MAX_LEVELS), ('class_predictions', c_void_p * SSDK_MAX_LEVELS),
                ('encoded_boxes', c_void_p * SSDK_MAX_LEVELS)]


class SsdkHeadGrads(ctypes.Structure):
    """struct ssdk_head_grads of include/ssdk.h."""
    _fields_ = [('class_predictions', c_void_p * SSDK_MAX_LEVELS), ('encoded_boxes', c_void_p * SSDK_MAX_LEVELS)]

SSDK_INPUT_SCORES, SSDK_INPUT_LOGITS = 0, 1
SSDK_BOXES_ENCODED, SSDK_BOXES_DECODED = 0, 2
SSDK_POST_SCAN_ONLY, SSDK_POST_FINISH_ONLY = 16, 32

_lib = None
_lock = threading.Lock()
_contexts = {}


class SsdkError(RuntimeError):
    pass


def load():
    """dlopen libssdk.so and declare every entry point; raises if the extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise SsdkError(
                    'CUDA extension not built: %s is missing. Run `python -c "import __graft_entry__ as g; g.build()"` '
                    '(or make -C single-shot-detector_b200/csrc). There is no CPU fallback.' % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in _SIGNATURES.items():
                fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
                fn.restype = c_int if restype is None else restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def check(status):
    if status != 0:
        msg = load().ssdk_last_error().decode('utf-8', 'replace')
        if status in (-1, -2):
            raise ValueError('ssdk: ' + msg)
        raise SsdkError('ssdk error %d: %s' % (status, msg))


def context(device_index):
    """One library context per (thread, device)."""
    key = (threading.get_ident(), int(device_index))
    ctx = _contexts.get(key)
    if ctx is None:
        lib = load()
        h = P()
        check(lib.ssdk_ctx_create(int(device_index), None, ctypes.byref(h)))
        ctx = _contexts[key] = h
    return ctx


KERNEL_IDS = ['anchors', 'match', 'force_match', 'ssd_loss', 'loss_reduce', 'filter', 'filter_dense', 'nms', 'pack', 'other',
              'ssd_loss_backward', 'head_flat', 'head_rows', 'head_concat', 'comm', 'train_step', 'nms_rounds']


def set_profiling(enable, device_index=0):
    check(load().ssdk_ctx_set_profiling(context(device_index), 1 if enable else 0))


def profile_read(device_index=0, reset=True):
    """{kernel name: (total ms, launches)} measured with CUDA events on the context's stream."""
    n = len(KERNEL_IDS)
    ms = (c_double * n)()
    calls = (c_i64 * n)()
    check(load().ssdk_ctx_profile_read(context(device_index), ctypes.cast(ms, P), ctypes.cast(calls, P), n, 1 if reset else 0))
    return {k: (ms[i], int(calls[i])) for i, k in enumerate(KERNEL_IDS)}


SSDK_OPT_FUSED_TRAIN_STEP, SSDK_OPT_MATCH_CTAS_PER_SM, SSDK_OPT_MATCH_FLAT_SHARE_PCT, SSDK_OPT_TRAIN_CTAS_PER_SM = 1, 2, 3, 4
SSDK_OPT_PROGRAMMATIC_LAUNCH = 5
SSDK_OPT_TRAIN_DYNAMIC_CHUNKS = 6
SSDK_STEP_ALL_REDUCE = 1
SSDK_ASYNC_ROUNDS_TIMEOUT = 1


def round_times(device_index=0, n=32):
    """[rounds run, timestamps in ns ...] of the dense-segment stage of the last post-processing call (ssdk_ctx_round_times)."""
    buf = (c_i64 * n)()
    check(load().ssdk_ctx_round_times(context(device_index), ctypes.cast(buf, P), n))
    return [int(v) for v in buf]


def async_error(device_index=0):
    """Sticky asynchronous error word of the context (0 = none); synchronises its stream."""
    e = c_int(0)
    check(load().ssdk_ctx_async_error(context(device_index), ctypes.byref(e)))
    return int(e.value)


def set_option(option, value, device_index=0):
    check(load().ssdk_ctx_set_option(context(device_index), int(option), int(value)))


def launch_count(device_index=0):
    return int(load().ssdk_ctx_launch_count(context(device_index)))


def workspace_bytes(device_index=0):
    return int(load().ssdk_ctx_workspace_bytes(context(device_index)))
