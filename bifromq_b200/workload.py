"""Seeded synthetic workloads for BASELINE.json's configs (C1..C5), generated natively by libbfq_workload.so.

C1  1 tenant, 10k exact filters, 1k topics                       (plumbing / parity)
C2  1 tenant, 1M filters (50% with '+'), 100k topics
C3  1000 tenants x 10k filters mixed '+'/'#', 2% $share groups, 1M topics
C4  10M filters over 1000 tenants (Zipf sizes), Zipf fan-out and topic popularity, 1M topics
C5  inverse: 1M retained topics vs 100k wildcard filters
`scale` shrinks every size for tests. Arrays are numpy views into native memory owned by the Workload.
"""
import ctypes as C
import os

import numpy as np

from ._native import WORKLOAD_LIB_PATH, NativeError

_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(WORKLOAD_LIB_PATH):
            raise NativeError("%s is missing: build it with __graft_entry__.build()" % WORKLOAD_LIB_PATH)
        lib = C.CDLL(WORKLOAD_LIB_PATH)
        lib.bfqw_generate.restype = C.c_void_p
        lib.bfqw_generate.argtypes = [C.c_char_p, C.c_uint64, C.c_double, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        lib.bfqw_generate2.restype = C.c_void_p
        lib.bfqw_generate2.argtypes = [C.c_char_p, C.c_uint64, C.c_double, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
        lib.bfqw_free.argtypes = [C.c_void_p]
        lib.bfqw_info.argtypes = [C.c_void_p, C.c_void_p]
        for name in ["keys", "key_off", "vals", "val_off", "tenants", "tenant_off", "topics", "topic_off", "topic_tenant",
                     "filters", "filter_off", "filter_tenant"]:
            fn = getattr(lib, "bfqw_" + name)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _view(p, count, dtype):
    dtype = np.dtype(dtype)
    if count == 0 or not p:
        return np.zeros(0, dtype)
    return np.frombuffer((C.c_uint8 * (count * dtype.itemsize)).from_address(p), dtype=dtype)


class Workload:
    SEED = 0xB1F20

    def __init__(self, config, seed=SEED, scale=1.0, shard_index=0, shard_count=1, tenant_prefix="", nthreads=None, replicate_hot=False,
                 topic_mult=1):
        """shard_index / shard_count: keep the tenants this shard owns (fnv1a64(tenant) % shard_count) and their topics;
        replicate_hot: tenants above 1 / (4 shard_count) of the batch live on every shard, topics dealt round-robin;
        topic_mult: the publish batch is topic_mult times the config's size (N GPUs serving N times the traffic of one filter set)"""
        lib = _load()
        if nthreads is None:
            nthreads = max(1, min(32, os.cpu_count() or 1))
        self.config, self.seed, self.scale = config, seed, scale
        self._h = lib.bfqw_generate2(config.encode(), seed, float(scale), shard_index, shard_count, tenant_prefix.encode(), nthreads,
                                     1 if replicate_hot else 0, int(topic_mult))
        if not self._h:
            raise ValueError("unknown workload config %r" % config)
        info = np.zeros(8, np.int64)
        lib.bfqw_info(self._h, info.ctypes.data)
        (self.n_routes, self.n_tenants, self.n_topics, self.n_query_filters, self.n_filters, kb, vb, tb) = info.tolist()
        h = self._h
        self.key_off = _view(lib.bfqw_key_off(h), self.n_routes + 1, np.int64)
        self.val_off = _view(lib.bfqw_val_off(h), self.n_routes + 1, np.int64)
        self.keys = _view(lib.bfqw_keys(h), max(kb, 1), np.uint8)
        self.vals = _view(lib.bfqw_vals(h), max(vb, 1), np.uint8)
        self.tenant_off = _view(lib.bfqw_tenant_off(h), self.n_tenants + 1, np.int64)
        self.tenants_blob = _view(lib.bfqw_tenants(h), max(int(self.tenant_off[-1]), 1), np.uint8)
        self.topic_off = _view(lib.bfqw_topic_off(h), self.n_topics + 1, np.int64)
        self.topics = _view(lib.bfqw_topics(h), max(tb, 1), np.uint8)
        self.topic_tenant = _view(lib.bfqw_topic_tenant(h), max(self.n_topics, 1), np.int32)
        self.filter_off = _view(lib.bfqw_filter_off(h), self.n_query_filters + 1, np.int64)
        self.filters = _view(lib.bfqw_filters(h), max(int(self.filter_off[-1]), 1), np.uint8)
        self.filter_tenant = _view(lib.bfqw_filter_tenant(h), max(self.n_query_filters, 1), np.int32)

    @property
    def tenants(self):
        b = self.tenants_blob.tobytes()
        return [b[self.tenant_off[i]:self.tenant_off[i + 1]].decode() for i in range(self.n_tenants)]

    def topic(self, i):
        return self.topics[self.topic_off[i]:self.topic_off[i + 1]].tobytes()

    def topic_list(self):
        b = self.topics.tobytes()
        return [b[self.topic_off[i]:self.topic_off[i + 1]] for i in range(self.n_topics)]

    def query_filter_list(self):
        b = self.filters.tobytes()
        return [b[self.filter_off[i]:self.filter_off[i + 1]] for i in range(self.n_query_filters)]

    def close(self):
        if getattr(self, "_h", None):
            _load().bfqw_free(self._h)
            self._h = None

    def __del__(self):
        self.close()
