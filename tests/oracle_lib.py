"""ctypes driver for the CPU oracle (oracle/_build/liboracle.so).

Test infrastructure only: the product package (bifromq_b200) never imports this module.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])
    return _SO


def _load():
    if not os.path.exists(_SO):
        build_oracle()
    lib = C.CDLL(_SO)
    vp, i32, i64, u8p = C.c_void_p, C.c_int32, C.c_int64, C.c_char_p
    i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    sig = {
        "orc_java_hash": (i32, [u8p, i64]),
        "orc_java_compare": (i32, [u8p, i64, u8p, i64]),
        "orc_bucket": (i32, [u8p, i64]),
        "orc_parse": (i64, [u8p, i64, i32, u8p, i64]),
        "orc_is_valid_topic": (i32, [u8p, i64, i32, i32, i32]),
        "orc_is_valid_topic_filter": (i32, [u8p, i64, i32, i32, i32]),
        "orc_is_wildcard_topic_filter": (i32, [u8p, i64]),
        "orc_is_shared_subscription": (i32, [u8p, i64]),
        "orc_is_ordered_shared": (i32, [u8p, i64]),
        "orc_is_unordered_shared": (i32, [u8p, i64]),
        "orc_route_matcher_from": (i64, [u8p, i64, u8p, i64]),
        "orc_receiver_url": (i64, [i32, u8p, i64, u8p, i64, u8p, i64]),
        "orc_tenant_begin_key": (i64, [u8p, i64, u8p, i64]),
        "orc_range_lookup": (None, [u8p, i64, u8p, i64, i64, u8p, u8p, i64p, u8p, i64p, u8p]),
        "orc_retain_key": (i64, [u8p, i64, u8p, i64, u8p, i64]),
        "orc_retain_key_prefix": (i64, [u8p, i64, u8p, i64, u8p, i64]),
        "orc_level_hash_byte": (i32, [u8p, i64]),
        "orc_tenant_route_start_key": (i64, [u8p, i64, u8p, i64, u8p, i64]),
        "orc_route_key": (i64, [u8p, i64, u8p, i64, u8p, i64, u8p, i64]),
        "orc_upper_bound": (i64, [u8p, i64, u8p, i64]),
        "orc_route_group": (i64, [i64, u8p, i64p, u64p, u8p, i64]),
        "orc_build_match_route": (i64, [u8p, i64, u8p, i64, u8p, i64]),
        "orc_kv_new": (vp, []),
        "orc_kv_free": (None, [vp]),
        "orc_kv_put": (None, [vp, u8p, i64, u8p, i64]),
        "orc_kv_erase": (None, [vp, u8p, i64]),
        "orc_kv_load": (None, [vp, vp, i64p, vp, i64p, i64]),
        "orc_kv_size": (i64, [vp]),
        "orc_kv_freeze": (None, [vp]),
        "orc_kv_key": (i64, [vp, i64, u8p, i64]),
        "orc_kv_value": (i64, [vp, i64, u8p, i64]),
        "orc_kv_lower_bound": (i64, [vp, u8p, i64]),
        "orc_match_batch": (vp, [vp, i32, i32, vp, i64p, i64, vp, i64p, i32p, i64, i32, i32, i32]),
        "orc_result_free": (None, [vp]),
        "orc_result_total_routes": (i64, [vp]),
        "orc_result_routes": (None, [vp, i64p, i64p]),
        "orc_result_fanouts": (None, [vp, i32p, i32p]),
        "orc_result_num_events": (i64, [vp]),
        "orc_result_events": (None, [vp, i32p, i32p, i64p, i32p]),
        "orc_result_stats": (None, [vp, u64p]),
        "orc_expansion_list": (i64, [vp, i64p, i64, i32, u8p, i64]),
        "orc_expansion_seek": (i64, [vp, i64p, i64, i32, u8p, i64, u8p, i64]),
        "orc_topic_matches_filter": (i32, [u8p, i64, u8p, i64]),
        "orc_tli_new": (vp, []),
        "orc_tli_free": (None, [vp]),
        "orc_tli_add": (None, [vp, u8p, i64, u8p, i64, i64]),
        "orc_tli_remove": (None, [vp, u8p, i64, u8p, i64, i64]),
        "orc_tli_match": (i64, [vp, u8p, i64, u8p, i64, vp, i64, vp]),
        "orc_tli_get": (i64, [vp, u8p, i64, vp, i64]),
        "orc_tli_find_all": (i64, [vp, vp, i64]),
        "orc_tli_match_batch": (None, [vp, vp, i64p, vp, vp, i64p, i64, i32, i64p, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def _b(s):
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def _bytes_call(fn, *args):
    cap = 4096
    while True:
        buf = C.create_string_buffer(cap)
        n = fn(*args, buf, cap)
        if n < 0:
            return None
        if n <= cap:
            return buf.raw[:n]
        cap = n


class _Reader:
    def __init__(self, data):
        self.d, self.p = data, 0

    def u32(self):
        v = struct.unpack_from("<I", self.d, self.p)[0]
        self.p += 4
        return v

    def i32(self):
        v = struct.unpack_from("<i", self.d, self.p)[0]
        self.p += 4
        return v

    def u64(self):
        v = struct.unpack_from("<Q", self.d, self.p)[0]
        self.p += 8
        return v

    def bstr(self):
        n = self.u32()
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def levels(self):
        return [self.bstr().decode("utf-8") for _ in range(self.u32())]


def blob(strings):
    """list[str|bytes] -> (uint8 blob, int64 offsets[n+1])"""
    bs = [_b(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(x) for x in bs])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    if data.size == 0:
        data = np.zeros(1, np.uint8)
    return data, off


# ------------------------------------------------------------------ TopicUtil / JDK
def java_hash(s):
    b = _b(s)
    return lib.orc_java_hash(b, len(b))


def java_compare(a, b):
    a, b = _b(a), _b(b)
    return lib.orc_java_compare(a, len(a), b, len(b))


def bucket(s):
    b = _b(s)
    return lib.orc_bucket(b, len(b))


def parse(s, escaped=False):
    b = _b(s)
    return _Reader(_bytes_call(lib.orc_parse, b, len(b), int(escaped))).levels()


def is_valid_topic(s, max_level_length, max_level, max_length):
    b = _b(s)
    return bool(lib.orc_is_valid_topic(b, len(b), max_level_length, max_level, max_length))


def is_valid_topic_filter(s, max_level_length, max_level, max_length):
    b = _b(s)
    return bool(lib.orc_is_valid_topic_filter(b, len(b), max_level_length, max_level, max_length))


def is_wildcard_topic_filter(s):
    b = _b(s)
    return bool(lib.orc_is_wildcard_topic_filter(b, len(b)))


def route_matcher_from(tf):
    b = _b(tf)
    r = _Reader(_bytes_call(lib.orc_route_matcher_from, b, len(b)))
    return {"type": ["Normal", "UnorderedShare", "OrderedShare"][r.u32()], "filterLevels": r.levels(),
            "group": r.bstr().decode(), "mqttTopicFilter": r.bstr().decode()}


# ------------------------------------------------------------------ KVSchemaUtil
def receiver_url(sub_broker_id, receiver_id, deliverer_key):
    a, b = _b(receiver_id), _b(deliverer_key)
    return _bytes_call(lib.orc_receiver_url, sub_broker_id, a, len(a), b, len(b))


def tenant_begin_key(tenant):
    t = _b(tenant)
    return _bytes_call(lib.orc_tenant_begin_key, t, len(t))


def retain_key(tenant, topic):
    t, p = _b(tenant), _b(topic)
    return _bytes_call(lib.orc_retain_key, t, len(t), p, len(p))


def retain_key_prefix(tenant, topic_filter):
    t, f = _b(tenant), _b(topic_filter)
    return _bytes_call(lib.orc_retain_key_prefix, t, len(t), f, len(f))


def level_hash_byte(level):
    l = _b(level)
    return lib.orc_level_hash_byte(l, len(l))


def tenant_route_start_key(tenant, topic_filter):
    t, f = _b(tenant), _b(topic_filter)
    return _bytes_call(lib.orc_tenant_route_start_key, t, len(t), f, len(f))


def route_key(tenant, mqtt_topic_filter, receiver_url_=b""):
    """toNormalRouteKey / toGroupRouteKey depending on the $share/$oshare prefix."""
    t, f, u = _b(tenant), _b(mqtt_topic_filter), _b(receiver_url_)
    return _bytes_call(lib.orc_route_key, t, len(t), f, len(f), u, len(u))


def upper_bound(key):
    return _bytes_call(lib.orc_upper_bound, key, len(key))


def route_group(members):
    """members: dict receiverUrl(bytes) -> incarnation"""
    urls = list(members.keys())
    data, off = blob(urls)
    inc = np.array([members[u] for u in urls], dtype=np.uint64)
    if inc.size == 0:
        inc = np.zeros(1, np.uint64)
    return _bytes_call(lib.orc_route_group, len(urls), data.tobytes(), off, inc)


def incarnation_bytes(v):
    return struct.pack(">Q", v)


def build_match_route(key, value):
    raw = _bytes_call(lib.orc_build_match_route, key, len(key), value, len(value))
    if raw is None:
        raise ValueError("undecodable route")
    r = _Reader(raw)
    m = {"type": ["Normal", "Group"][r.u32()], "tenantId": r.bstr().decode(), "mqttTopicFilter": r.bstr().decode(),
         "filterLevels": r.levels(), "receiverUrl": r.bstr(), "incarnation": r.u64(), "subBrokerId": r.i32()}
    m["members"] = {}
    for _ in range(r.u32()):
        k = r.bstr()
        m["members"][k] = r.u64()
    return m


def matching_identity(m):
    """Equality key of a Matching (NormalMatching.java:30-41 / GroupMatching.java:32-39)."""
    if m["type"] == "Normal":
        return ("N", m["tenantId"], m["mqttTopicFilter"], m["receiverUrl"], m["incarnation"])
    return ("G", m["tenantId"], m["mqttTopicFilter"], tuple(sorted(m["members"].items())))


# ------------------------------------------------------------------ sorted KV + matchers
MODE_REFERENCE, MODE_BRUTE, MODE_TRIE = 0, 1, 2


class MatchOutcome:
    def __init__(self, offsets, ranks, pf, gf, events, stats):
        self.offsets, self.ranks, self.persistent_fanout, self.group_fanout = offsets, ranks, pf, gf
        self.events = events  # sorted list of (kind, topicIdx, rank, maxCount)
        self.stats = stats    # dict seeks, nexts, V, P, R, ranges

    def routes(self, i):
        return self.ranks[self.offsets[i]:self.offsets[i + 1]]

    def route_sets(self):
        return [tuple(self.routes(i).tolist()) for i in range(len(self.offsets) - 1)]


class KV:
    def __init__(self):
        self.h = lib.orc_kv_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_kv_free(self.h)
            self.h = None

    def put(self, k, v):
        lib.orc_kv_put(self.h, k, len(k), v, len(v))

    def erase(self, k):
        lib.orc_kv_erase(self.h, k, len(k))

    def load(self, keys, koff, vals, voff):
        lib.orc_kv_load(self.h, keys.ctypes.data, koff, vals.ctypes.data, voff, len(koff) - 1)

    def __len__(self):
        return lib.orc_kv_size(self.h)

    def freeze(self):
        lib.orc_kv_freeze(self.h)

    def key(self, rank):
        return _bytes_call(lib.orc_kv_key, self.h, rank)

    def value(self, rank):
        return _bytes_call(lib.orc_kv_value, self.h, rank)

    def items(self):
        self.freeze()
        return [(self.key(i), self.value(i)) for i in range(len(self))]

    def export(self):
        """-> (keys blob, key offsets, values blob, value offsets) in KV order"""
        it = self.items()
        k, ko = blob([x[0] for x in it])
        v, vo = blob([x[1] for x in it])
        return k, ko, v, vo

    def match_batch(self, tenants, topics, topic_tenant=None, max_persistent=2 ** 31 - 1, max_group=100,
                    mode=MODE_REFERENCE, singleton=False, nthreads=1):
        tb, toff = blob(tenants)
        pb, poff = blob(topics)
        n = len(topics)
        tt = np.zeros(max(n, 1), np.int32) if topic_tenant is None else np.ascontiguousarray(topic_tenant, dtype=np.int32)
        return self.match_blobs(tb, toff, pb, poff, tt, n, max_persistent, max_group, mode, singleton, nthreads)

    def match_blobs(self, tb, toff, pb, poff, tt, n, max_persistent, max_group, mode, singleton, nthreads):
        import time as _time
        _t0 = _time.perf_counter()
        r = lib.orc_match_batch(self.h, mode, int(singleton), tb.ctypes.data, toff, len(toff) - 1, pb.ctypes.data, poff,
                                tt, n, max_persistent, max_group, nthreads)
        self.last_match_seconds = _time.perf_counter() - _t0   # the matcher alone, without copying the result out
        try:
            total = lib.orc_result_total_routes(r)
            offsets = np.zeros(n + 1, np.int64)
            ranks = np.zeros(max(total, 1), np.int64)
            lib.orc_result_routes(r, offsets, ranks)
            pf, gf = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
            lib.orc_result_fanouts(r, pf, gf)
            ne = lib.orc_result_num_events(r)
            ek, et, ec = (np.zeros(max(ne, 1), np.int32) for _ in range(3))
            er = np.zeros(max(ne, 1), np.int64)
            lib.orc_result_events(r, ek, et, er, ec)
            st = np.zeros(7, np.uint64)
            lib.orc_result_stats(r, st)
        finally:
            lib.orc_result_free(r)
        events = sorted((int(ek[i]), int(et[i]), int(er[i]), int(ec[i])) for i in range(ne))
        stats = dict(zip(["seeks", "nexts", "V", "P", "R", "ranges", "backward_seeks"], (int(x) for x in st)))
        return MatchOutcome(offsets, ranks[:total], pf[:n], gf[:n], events, stats)

    def match_all(self, tenant, topics, max_persistent, max_group, mode=MODE_REFERENCE):
        """ITenantRouteMatcher.matchAll for one tenant -> dict topic -> list of decoded Matching dicts,
        plus the outcome object."""
        topics = list(topics)
        out = self.match_batch([tenant], topics, None, max_persistent, max_group, mode)
        res = {}
        for i, t in enumerate(topics):
            res[t] = [build_match_route(self.key(int(r)), self.value(int(r))) for r in out.routes(i)]
        return res, out


# ------------------------------------------------------------------ expansion set
def expansion_list(topics, is_global=False):
    data, off = blob(topics)
    raw = _bytes_call(lib.orc_expansion_list, data.ctypes.data, off, len(topics), int(is_global))
    r = _Reader(raw)
    out = []
    for _ in range(r.u32()):
        lv = r.levels()
        vals = [r.u32() for _ in range(r.u32())]
        out.append((lv, vals))
    return out


def expansion_seek(topics, filter_str, is_global=False):
    data, off = blob(topics)
    f = _b(filter_str) if filter_str is not None else b""
    fn = len(f) if filter_str is not None else -1
    raw = _bytes_call(lib.orc_expansion_seek, data.ctypes.data, off, len(topics), int(is_global), f, fn)
    return None if raw is None else _Reader(raw).levels()


def topic_matches_filter(topic, topic_filter):
    t, f = _b(topic), _b(topic_filter)
    return bool(lib.orc_topic_matches_filter(t, len(t), f, len(f)))


# ------------------------------------------------------------------ inverse index
class TopicLevelIndex:
    """TopicIndex (tenant=None) / RetainTopicIndex (tenant given) restatement."""

    def __init__(self):
        self.h = lib.orc_tli_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_tli_free(self.h)
            self.h = None

    @staticmethod
    def _t(tenant):
        return (b"", -1) if tenant is None else (_b(tenant), len(_b(tenant)))

    def add(self, topic, value, tenant=None):
        t, tn = self._t(tenant)
        b = _b(topic)
        lib.orc_tli_add(self.h, t, tn, b, len(b), value)

    def remove(self, topic, value, tenant=None):
        t, tn = self._t(tenant)
        b = _b(topic)
        lib.orc_tli_remove(self.h, t, tn, b, len(b), value)

    def match(self, topic_filter, tenant=None, with_visited=False):
        t, tn = self._t(tenant)
        b = _b(topic_filter)
        vis = C.c_uint64(0)
        cap = 1024
        while True:
            out = np.zeros(cap, np.int64)
            vis.value = 0
            n = lib.orc_tli_match(self.h, t, tn, b, len(b), out.ctypes.data, cap, C.addressof(vis))
            if n <= cap:
                res = out[:n].tolist()
                return (res, vis.value) if with_visited else res
            cap = n

    def get(self, topic):
        b = _b(topic)
        out = np.zeros(1024, np.int64)
        n = lib.orc_tli_get(self.h, b, len(b), out.ctypes.data, 1024)
        return out[:n].tolist()

    def find_all(self):
        cap = 1024
        while True:
            out = np.zeros(cap, np.int64)
            n = lib.orc_tli_find_all(self.h, out.ctypes.data, cap)
            if n <= cap:
                return out[:n].tolist()
            cap = n


# ------------------------------------------------------------------ fan-out grouping
def deliverer_of_receiver_url(receiver_url):
    """(subBrokerId, delivererKey) a NormalMatching is delivered through: DeliverExecutor.send
    (bifromq-dist/bifromq-dist-worker/src/main/java/org/apache/bifromq/dist/worker/DeliverExecutor.java:89-93) keys its DeliveryCall
    by matched.subBrokerId() and matched.delivererKey(), both cut out of the receiver url
    "<subBrokerId>\\0<receiverId>\\0<delivererKey>" (KVSchemaUtil.java:56-58, cache/ReceiverCache.java:32-36)."""
    broker, _receiver, deliverer_key = bytes(receiver_url).split(b"\0", 2)
    return int(broker), deliverer_key


def route_group_members_in_wire_order(value):
    """receiver urls of a RouteGroup value in the order the proto carries them — GroupMatching.receiverList keeps the map's
    iteration order (cache/GroupMatching.java:44-46), which for a parsed protobuf map is the wire order."""
    b, i, out = bytes(value), 0, []

    def varint():
        nonlocal i
        v, shift = 0, 0
        while True:
            c = b[i]
            i += 1
            v |= (c & 0x7F) << shift
            if not c & 0x80:
                return v
            shift += 7
    while i < len(b):
        assert varint() == (1 << 3) | 2
        end = varint() + i
        key = None
        while i < end:
            tag = varint()
            if tag == (1 << 3) | 2:
                n = varint()
                key = b[i:i + n]
                i += n
            else:
                varint()
        out.append(key)
    return out


# ------------------------------------------------------------------ dist-server range pruning
def range_lookup(tenant, topic, candidates):
    """TenantRangeLookupCache.lookup restated literally (oracle/capi.cc: orc_range_lookup). candidates: ordered list of None
    (no Fact) or (first, last) global filter level lists (either may be None). Returns the kept candidate indices."""
    flags, firsts, lasts = [], [], []
    enc = lambda lv: b"\0".join(_b(x) for x in lv)
    for c in candidates:
        if c is None:
            flags.append(0); firsts.append(b""); lasts.append(b"")
        else:
            first, last = c
            flags.append(1 | (2 if first is not None else 0) | (4 if last is not None else 0))
            firsts.append(enc(first) if first is not None else b"")
            lasts.append(enc(last) if last is not None else b"")
    fb, foff = blob(firsts)
    lb, loff = blob(lasts)
    keep = (C.c_uint8 * max(len(candidates), 1))()
    fl = bytes(flags) + b"\0"
    t, p = _b(tenant), _b(topic)
    lib.orc_range_lookup(t, len(t), p, len(p), len(candidates), fl, fb.tobytes() + b"\0", foff, lb.tobytes() + b"\0", loff,
                         C.cast(keep, C.c_char_p))
    return [k for k in range(len(candidates)) if keep[k]]
