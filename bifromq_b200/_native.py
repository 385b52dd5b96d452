"""ctypes binding of libbfq_gpumatch.so (the C-ABI of include/bfq_gpumatch.h).

There is no fallback: if the library is missing this module raises, and every match call needs a CUDA device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFQ_LIB") or os.path.join(_HERE, "libbfq_gpumatch.so")   # BFQ_LIB: A/B experiments only
WORKLOAD_LIB_PATH = os.path.join(_HERE, "libbfq_workload.so")


class NativeError(RuntimeError):
    pass


class BfqRange(C.Structure):
    _fields_ = [("first", C.c_uint32), ("count", C.c_uint32)]


class BfqThrottled(C.Structure):
    _fields_ = [("topic", C.c_uint32), ("rank", C.c_uint32), ("kind", C.c_uint32)]


class BfqDeviceResult(C.Structure):
    _fields_ = [("d_span_begin", C.c_void_p), ("d_span_count", C.c_void_p), ("d_route_count", C.c_void_p),
                ("d_ranges", C.c_void_p), ("d_throttled", C.c_void_p), ("n_ranges", C.c_int64),
                ("n_throttled", C.c_int64), ("n_routes", C.c_int64), ("n_overflow_topics", C.c_int64),
                ("n_flagged_topics", C.c_int64), ("n_launches", C.c_int64), ("n_topics", C.c_int64), ("n_distinct_topics", C.c_int64),
                ("tier0_ms", C.c_double), ("generation", C.c_uint64), ("lease", C.c_void_p)]


class BfqGathered(C.Structure):
    _fields_ = [("d_route_count", C.c_void_p), ("d_span_count", C.c_void_p), ("d_ranges", C.c_void_p),
                ("topic_base", C.POINTER(C.c_int64)), ("range_base", C.POINTER(C.c_int64)), ("topic_count", C.POINTER(C.c_int64)),
                ("range_count", C.POINTER(C.c_int64)), ("n_topics_total", C.c_int64),
                ("n_ranges_total", C.c_int64), ("bytes_received", C.c_int64), ("world", C.c_int32)]


class BfqFanoutResult(C.Structure):
    _fields_ = [("d_pack_offsets", C.c_void_p), ("d_pack_topic", C.c_void_p), ("d_pack_rank", C.c_void_p), ("d_pack_member", C.c_void_p),
                ("n_pairs", C.c_int64), ("n_deliverers", C.c_int32), ("ordered_share_id", C.c_int32), ("generation", C.c_uint64)]


EXCHANGE_ID_BYTES, EXCHANGE_COUNTS, EXCHANGE_RANGES = 128, 1, 2
_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_SIGNATURES = {
    "bfq_last_error": (C.c_char_p, []),
    "bfq_index_create": (_i32, [_i32, C.POINTER(_vp)]),
    "bfq_index_destroy": (None, [_vp]),
    "bfq_index_reset": (_i32, [_vp]),
    "bfq_index_load": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64]),
    "bfq_index_apply": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64]),
    "bfq_index_commit": (_i32, [_vp]),
    "bfq_index_generation": (_i32, [_vp, C.POINTER(C.c_uint64)]),
    "bfq_index_set_option": (_i32, [_vp, C.c_char_p, _i64]),
    "bfq_index_stats": (_i32, [_vp, _vp, _i32]),
    "bfq_host_build_stats": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32]),
    "bfq_index_last_kernel_ms": (_i32, [_vp, C.POINTER(C.c_double)]),
    "bfq_route_lookup": (_i32, [_vp, _i64, _vp, _i64, C.POINTER(_i64), _vp, _i64, C.POINTER(_i64)]),
    "bfq_route_kind": (_i32, [_vp, _i64, C.POINTER(_i32)]),
    "bfq_route_kinds": (_i32, [_vp, _vp, _i64, _vp]),
    "bfq_match": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, C.POINTER(_vp)]),
    "bfq_result_num_topics": (_i64, [_vp]),
    "bfq_result_span_begin": (_vp, [_vp]),
    "bfq_result_span_count": (_vp, [_vp]),
    "bfq_result_route_count": (_vp, [_vp]),
    "bfq_result_ranges": (_vp, [_vp, C.POINTER(_i64)]),
    "bfq_result_throttled": (_vp, [_vp, C.POINTER(_i64)]),
    "bfq_result_expand": (_i64, [_vp, _vp, _vp, _i64]),
    "bfq_result_route_lookup": (_i32, [_vp, _i64, _vp, _i64, C.POINTER(_i64), _vp, _i64, C.POINTER(_i64)]),
    "bfq_result_route_kinds": (_i32, [_vp, _vp, _i64, _vp]),
    "bfq_result_generation": (C.c_uint64, [_vp]),
    "bfq_result_timings": (_i32, [_vp, _vp, _i32]),
    "bfq_result_free": (None, [_vp]),
    "bfq_match_device": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, C.POINTER(BfqDeviceResult)]),
    "bfq_match_device_async": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, C.POINTER(BfqDeviceResult)]),
    "bfq_device_result_wait": (_i32, [C.POINTER(BfqDeviceResult)]),
    "bfq_device_result_release": (None, [C.POINTER(BfqDeviceResult)]),
    "bfq_expand_device": (_i32, [C.POINTER(BfqDeviceResult), _vp, _vp, _i64, _vp, C.POINTER(_i64)]),
    "bfq_range_lookup": (_i32, [_i32, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bfq_fanout_device": (_i32, [C.POINTER(BfqDeviceResult), _vp, _vp, _i64, _vp, C.POINTER(BfqFanoutResult)]),
    "bfq_fanout_deliverer": (_i32, [_vp, _i32, C.POINTER(_i32), _vp, _i64, C.POINTER(_i64)]),
    "bfq_exchange_unique_id": (_i32, [_vp, _i32]),
    "bfq_exchange_create": (_i32, [_i32, _i32, _i32, _vp, C.POINTER(_vp)]),
    "bfq_exchange_destroy": (None, [_vp]),
    "bfq_exchange_gather": (_i32, [_vp, C.POINTER(BfqDeviceResult), _i32, _vp, C.POINTER(BfqGathered)]),
    "bfq_receiver_url": (_i64, [_i32, C.c_char_p, _i64, C.c_char_p, _i64, _vp, _i64]),
    "bfq_route_key": (_i64, [C.c_char_p, _i64, C.c_char_p, _i64, C.c_char_p, _i64, _vp, _i64]),
    "bfq_tenant_begin_key": (_i64, [C.c_char_p, _i64, _vp, _i64]),
    "bfq_retain_key": (_i64, [C.c_char_p, _i64, C.c_char_p, _i64, _vp, _i64]),
    "bfq_retain_key_prefix": (_i64, [C.c_char_p, _i64, C.c_char_p, _i64, _vp, _i64]),
    "bfq_is_valid_topic": (_i32, [C.c_char_p, _i64, _i32, _i32, _i32]),
    "bfq_is_valid_topic_filter": (_i32, [C.c_char_p, _i64, _i32, _i32, _i32]),
    "bfq_rindex_create": (_i32, [_i32, C.POINTER(_vp)]),
    "bfq_rindex_destroy": (None, [_vp]),
    "bfq_rindex_reset": (_i32, [_vp]),
    "bfq_rindex_add": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "bfq_rindex_load_keys": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "bfq_rresult_retain_keys": (_i64, [_vp, _vp, _vp, _i64, _vp]),
    "bfq_rindex_remove": (_i32, [_vp, C.c_char_p, _i64, C.c_char_p, _i64]),
    "bfq_rindex_commit": (_i32, [_vp]),
    "bfq_rindex_lookup": (_i32, [_vp, _i64, _vp, _i64, C.POINTER(_i64), _vp, _i64, C.POINTER(_i64)]),
    "bfq_rmatch": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, C.POINTER(_vp)]),
    "bfq_rresult_num_filters": (_i64, [_vp]),
    "bfq_rresult_offsets": (_vp, [_vp]),
    "bfq_rresult_ids": (_vp, [_vp, C.POINTER(_i64)]),
    "bfq_rresult_total_matches": (_vp, [_vp]),
    "bfq_rresult_timings": (_i32, [_vp, _vp, _i32]),
    "bfq_rresult_free": (None, [_vp]),
}

_lib = None


def load_library(path=LIB_PATH):
    """Load the CUDA library and bind every symbol include/bfq_gpumatch.h declares. Raises NativeError when the
    extension has not been built (run `python -c "import __graft_entry__ as g; g.build()"`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise NativeError("%s is missing: build it with __graft_entry__.build() (no CPU fallback exists)" % path)
    lib_ = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib_, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib_
    return _lib


class _LazyLib:
    def __getattr__(self, name):
        return getattr(load_library(), name)


lib = _LazyLib()


def check(rc):
    if rc != 0:
        raise NativeError("bfq error %d: %s" % (rc, load_library().bfq_last_error().decode("utf-8", "replace")))


def as_blob(strings):
    """list[str|bytes] -> (uint8 array, int64 offsets[n+1])"""
    bs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    joined = b"".join(bs)
    data = np.frombuffer(joined, dtype=np.uint8).copy() if joined else np.zeros(1, np.uint8)
    return data, off


def ptr(a):
    return a.ctypes.data if a is not None else None
