"""C-ABI contract on a GPU box: error codes instead of crashes, results owned by the handle, no fallback."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N():
    import bifromq_b200
    from bifromq_b200 import _native
    bifromq_b200.load_library()
    return _native


def test_error_codes(N):
    lib = N.lib
    h = C.c_void_p()
    assert lib.bfq_index_create(0, C.byref(h)) == 0
    # match before the first commit
    r = C.c_void_p()
    tb, toff = N.as_blob(["t"])
    pb, poff = N.as_blob(["a"])
    tt = np.zeros(1, np.int32)
    rc = lib.bfq_match(h, N.ptr(tb), N.ptr(toff), 1, N.ptr(pb), N.ptr(poff), N.ptr(tt), 1, None, None, C.byref(r))
    assert rc == -4 and b"commit" in lib.bfq_last_error()
    # bad arguments
    assert lib.bfq_match(None, None, None, 0, None, None, None, 0, None, None, C.byref(r)) == -1
    assert lib.bfq_index_create(9999, C.byref(C.c_void_p())) == -1
    assert lib.bfq_index_load(h, None, None, None, None, 3) == -1
    # an empty index is a valid snapshot: every topic matches nothing
    assert lib.bfq_index_commit(h) == 0
    assert lib.bfq_match(h, N.ptr(tb), N.ptr(toff), 1, N.ptr(pb), N.ptr(poff), N.ptr(tt), 1, None, None, C.byref(r)) == 0
    assert lib.bfq_result_num_topics(r) == 1
    off = np.zeros(2, np.int64)
    assert lib.bfq_result_expand(r, off.ctypes.data, None, 0) == 0
    lib.bfq_result_free(r)
    # out-of-range tenant index: empty result, not a crash
    tt[0] = 7
    assert lib.bfq_match(h, N.ptr(tb), N.ptr(toff), 1, N.ptr(pb), N.ptr(poff), N.ptr(tt), 1, None, None, C.byref(r)) == 0
    lib.bfq_result_free(r)
    # lookups out of range
    kl = C.c_int64(0)
    assert lib.bfq_route_lookup(h, 0, None, 0, C.byref(kl), None, 0, C.byref(kl)) == -5
    kind = C.c_int32(0)
    assert lib.bfq_route_kind(h, -1, C.byref(kind)) == -5
    lib.bfq_index_destroy(h)
    # inverse index
    rh = C.c_void_p()
    assert lib.bfq_rindex_create(0, C.byref(rh)) == 0
    rr = C.c_void_p()
    fb, foff = N.as_blob(["#"])
    assert lib.bfq_rmatch(rh, N.ptr(tb), N.ptr(toff), 1, N.ptr(fb), N.ptr(foff), N.ptr(np.zeros(1, np.int32)), 1, None, C.byref(rr)) == -4
    assert lib.bfq_rindex_commit(rh) == 0
    assert lib.bfq_rmatch(rh, N.ptr(tb), N.ptr(toff), 1, N.ptr(fb), N.ptr(foff), N.ptr(np.zeros(1, np.int32)), 1, None, C.byref(rr)) == 0
    assert lib.bfq_rresult_num_filters(rr) == 1
    lib.bfq_rresult_free(rr)
    assert lib.bfq_rindex_lookup(rh, 5, None, 0, C.byref(kl), None, 0, C.byref(kl)) == -5
    lib.bfq_rindex_destroy(rh)


def test_duplicate_topics_and_repeated_calls(N):
    """the same topic twice in a batch gets two independent answers; buffers are reused across calls of different sizes"""
    import bifromq_b200
    from bifromq_b200 import schema
    idx = bifromq_b200.GpuRouteIndex(0)
    idx.load_pairs([(schema.route_key("t", "a/+", schema.receiver_url(0, "r", "d")), schema.incarnation_bytes(1)),
                    (schema.route_key("t", "a/#", schema.receiver_url(1, "r2", "d")), schema.incarnation_bytes(1))])
    idx.commit()
    for n in (1, 5, 2000, 3, 70000, 2):
        topics = ["a/b", "a/b", "x"] * n
        res = idx.match_topics(["t"], topics)
        off, ranks = res.expand()
        assert np.diff(off).tolist() == [2, 2, 0] * n
        res.close()
