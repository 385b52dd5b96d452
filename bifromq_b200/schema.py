"""Host-side mirror of the reference's route schema helpers, enough to re-hydrate a Matching from the stored
(key, value) pair a route rank maps back to — what the Java side does with its own
KVSchemaUtil.buildMatchRoute (bifromq-dist/bifromq-dist-worker-schema/.../schema/KVSchemaUtil.java:73-79,
cache/RouteDetailCache.java:53-109, cache/NormalMatching.java:30-41, cache/GroupMatching.java:32-39).
"""
import struct
from collections import namedtuple

from . import _native as N

# equality follows the reference: Normal = (tenantId, mqttTopicFilter, receiverUrl, incarnation),
# Group = (tenantId, mqttTopicFilter, members)
NormalMatching = namedtuple("NormalMatching", "tenant_id mqtt_topic_filter receiver_url incarnation")
GroupMatching = namedtuple("GroupMatching", "tenant_id mqtt_topic_filter members ordered")

FLAG_NORMAL, FLAG_UNORDERED, FLAG_ORDERED = 1, 2, 3


def _varint(b, p):
    v = shift = 0
    while True:
        c = b[p]
        p += 1
        v |= (c & 0x7F) << shift
        if not c & 0x80:
            return v, p
        shift += 7


def parse_route_group(b):
    """RouteGroup { map<string, uint64> members = 1; } -> tuple of sorted (receiverUrl bytes, incarnation)"""
    members, p = {}, 0
    while p < len(b):
        tag, p = _varint(b, p)
        if tag != 0x0A:
            raise ValueError("Unable to parse matching record")
        ln, p = _varint(b, p)
        end, key, val = p + ln, b"", 0
        while p < end:
            t, p = _varint(b, p)
            if t == 0x0A:
                kl, p = _varint(b, p)
                key = bytes(b[p:p + kl])
                p += kl
            elif t == 0x10:
                val, p = _varint(b, p)
            else:
                raise ValueError("Unable to parse matching record")
        members[key] = val
    return tuple(sorted(members.items()))


def build_match_route(key, value):
    tenant_len = struct.unpack_from(">H", key, 1)[0]
    receiver_len = struct.unpack_from(">H", key, len(key) - 2)[0]
    receiver_start = len(key) - 2 - receiver_len
    flag = key[receiver_start - 1]
    sep = receiver_start - 1 - 1 - 2
    tenant = key[3:3 + tenant_len].decode("utf-8")
    escaped = key[3 + tenant_len:sep]
    receiver = bytes(key[receiver_start:receiver_start + receiver_len])
    topic_filter = escaped.replace(b"\x00", b"/").decode("utf-8")
    if flag == FLAG_NORMAL:
        return NormalMatching(tenant, topic_filter, receiver, struct.unpack(">Q", value)[0])
    prefix = "$share/" if flag == FLAG_UNORDERED else "$oshare/"
    return GroupMatching(tenant, prefix + receiver.decode("utf-8") + "/" + topic_filter, parse_route_group(value),
                         flag == FLAG_ORDERED)


def sub_broker_id(m):
    return int(m.receiver_url.split(b"\x00")[0])


def _bytes_call(fn, *args):
    import ctypes as C
    cap = 512
    while True:
        buf = C.create_string_buffer(cap)
        n = fn(*args, C.addressof(buf), cap)
        if n < 0:
            raise N.NativeError("codec call failed: %d" % n)
        if n <= cap:
            return buf.raw[:n]
        cap = n


def _b(s):
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def receiver_url(sub_broker, receiver_id, deliverer_key):
    a, b = _b(receiver_id), _b(deliverer_key)
    return _bytes_call(N.lib.bfq_receiver_url, sub_broker, a, len(a), b, len(b))


def route_key(tenant, mqtt_topic_filter, receiver_url_=b""):
    t, f, u = _b(tenant), _b(mqtt_topic_filter), _b(receiver_url_)
    return _bytes_call(N.lib.bfq_route_key, t, len(t), f, len(f), u, len(u))


def retain_key(tenant, topic):
    """retainMessageKey (retain-store-schema KVSchemaUtil.java:44-50)"""
    t, p = _b(tenant), _b(topic)
    return _bytes_call(N.lib.bfq_retain_key, t, len(t), p, len(p))


def retain_key_prefix(tenant, topic_filter):
    """retainKeyPrefix of a topic filter (KVSchemaUtil.java:52-72)"""
    t, f = _b(tenant), _b(topic_filter)
    return _bytes_call(N.lib.bfq_retain_key_prefix, t, len(t), f, len(f))


def tenant_begin_key(tenant):
    t = _b(tenant)
    return _bytes_call(N.lib.bfq_tenant_begin_key, t, len(t))


def is_valid_topic(topic, max_level_length=40, max_level=16, max_length=255):
    t = _b(topic)
    return bool(N.lib.bfq_is_valid_topic(t, len(t), max_level_length, max_level, max_length))


def is_valid_topic_filter(tf, max_level_length=40, max_level=16, max_length=255):
    t = _b(tf)
    return bool(N.lib.bfq_is_valid_topic_filter(t, len(t), max_level_length, max_level, max_length))


def incarnation_bytes(v):
    return struct.pack(">Q", v)


def route_group_bytes(members):
    """dict receiverUrl(bytes) -> incarnation  ->  RouteGroup proto bytes"""
    def vi(v):
        o = bytearray()
        while v >= 0x80:
            o.append((v & 0x7F) | 0x80)
            v >>= 7
        o.append(v)
        return bytes(o)
    out = bytearray()
    for k, v in members.items():
        k = _b(k)
        e = b"\x0a" + vi(len(k)) + k + b"\x10" + vi(v)
        out += b"\x0a" + vi(len(e)) + e
    return bytes(out)
