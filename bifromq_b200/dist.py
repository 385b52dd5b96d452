"""Multi-GPU plumbing: tenant sharding and whole-job aggregation (one process per GPU, torch.distributed).

The path shards by tenant with no data-path collective (routes and matches are independent per tenant: the tenant id is
the key prefix, bifromq-dist/bifromq-dist-worker-schema/.../schema/KVSchemaUtil.java:91-94; the reference itself never joins
across tenants, TenantRouteMatcher.java:81-86). The only collectives are the barrier and the max / sum reductions that turn
per-rank timings into a whole-job number.
"""
import numpy as np


def fnv1a64(data):
    if isinstance(data, str):
        data = data.encode("utf-8")
    h = 0xCBF29CE484222325
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def tenant_shard(tenant_id, world_size):
    """rank that owns a tenant: fnv1a64(tenantId) mod world — the same function
    libbfq_workload.so uses when it generates one shard."""
    return fnv1a64(tenant_id) % world_size


def hot_tenants(tenant_share, world_size):
    """tenants hosted by EVERY rank ("replicas" mode of SURVEY.md 8e, applied per tenant): those that carry more than
    1 / (4 world) of the traffic. tenant_share: per-tenant fraction of the batch (any positive weights). With Zipf tenant
    sizes the largest tenant alone is 13 % of BASELINE C4's batch: pure hash placement leaves one of 8 GPUs with 1.87x the
    mean load, replicating the 4 hot tenants brings it to 1.08x."""
    w = np.asarray(tenant_share, dtype=np.float64)
    if world_size <= 1 or w.sum() <= 0:
        return np.zeros(len(w), bool)
    return w / w.sum() > 1.0 / (4.0 * world_size)


def split_batch_by_owner(tenants, topic_tenant, world_size, hot=None):
    """indices of the topics each rank must match -> list of int64 arrays (the dist-server side of the sharding). A topic of
    a hot (replicated) tenant goes to rank (batch position mod world), any other to fnv1a64(tenant) mod world — the rule
    libbfq_workload.so applies when it generates one shard."""
    owner_of_tenant = np.array([tenant_shard(t, world_size) for t in tenants], dtype=np.int64)
    tt = np.asarray(topic_tenant, dtype=np.int64)
    owner = owner_of_tenant[tt] if len(tenants) else np.zeros(0, np.int64)
    if hot is not None and len(tt):
        is_hot = np.asarray(hot, bool)[tt]
        owner = np.where(is_hot, np.arange(len(tt), dtype=np.int64) % world_size, owner)
    return [np.nonzero(owner == r)[0] for r in range(world_size)]


def aggregate(step_ms_total, n_units, device=None):
    """whole-job view of a timed region: MAX over ranks of the elapsed time, SUM over ranks of the units processed.
    Works with any initialised backend (nccl on GPUs, gloo in the CPU tests); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(step_ms_total), float(n_units)
    t = torch.tensor([float(step_ms_total)], dtype=torch.float64, device=device)
    c = torch.tensor([float(n_units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.item(), c.item()


class _DeviceArray:
    """zero-copy view of device memory owned by the native library (valid until the next match on the index)"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, typestr="<i4", device=None):
    """torch tensor over n elements at a device pointer returned through the C-ABI (e.g. bfq_device_result.d_route_count)"""
    import torch
    return torch.as_tensor(_DeviceArray(ptr, n, typestr), device=device)


def all_gather_varlen(x):
    """all-gather of one 1-D tensor per rank with different lengths: the sizes first, then the payload padded to the
    longest (NCCL wants equal shapes). Returns the list of per-rank tensors, on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [x]
    world = dist.get_world_size()
    n = torch.tensor([x.numel()], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    buf = torch.zeros(m, dtype=x.dtype, device=x.device)
    buf[:x.numel()] = x
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return [parts[r][:sizes[r]] for r in range(world)]


def gather_fanout(local_fanout, owner_indices=None, n_total=None):
    """The one exchange step of the sharded path: every rank matched the topics of the tenants it owns
    and holds their fan-out counts — what BatchDistReply carries back per topic (TopicFanout{map<topic,uint32>},
    bifromq-dist/bifromq-dist-rpc-definition/.../DistWorkerCoProc.proto:34-131; the routes themselves are delivered
    by the worker that matched them and never travel). All ranks end with the fan-out of the whole batch.

    local_fanout: 1-D int32 tensor, this rank's topics in the order of owner_indices[rank];
    owner_indices: split_batch_by_owner(...) of the original batch (None: concatenate in rank order);
    returns a 1-D tensor of n_total fan-out counts in batch order."""
    import torch
    parts = all_gather_varlen(local_fanout)
    if owner_indices is None:
        return torch.cat(parts)
    if n_total is None:
        n_total = int(sum(len(ix) for ix in owner_indices))
    out = torch.zeros(n_total, dtype=local_fanout.dtype, device=local_fanout.device)
    for r, part in enumerate(parts):
        ix = torch.as_tensor(np.asarray(owner_indices[r], dtype=np.int64), device=local_fanout.device)
        out[ix] = part
    return out


def split_batch_replicas(n_topics, world_size):
    """"replicas" mode for a tenant too large to shard by tenant (BASELINE config C2: one tenant): every rank holds the
    whole index, the topic batch is cut into contiguous slices -> list of (begin, end) per rank."""
    base, extra = divmod(int(n_topics), int(world_size))
    out, at = [], 0
    for r in range(world_size):
        ln = base + (1 if r < extra else 0)
        out.append((at, at + ln))
        at += ln
    return out


class Gathered:
    """result of Exchange.gather: torch views (device) over the reassembled arrays + the per-rank slice bounds (host)"""

    def __init__(self, raw, device):
        self.raw = raw
        self.world = raw.world
        self.topic_base = [raw.topic_base[i] for i in range(raw.world + 1)]
        self.range_base = [raw.range_base[i] for i in range(raw.world + 1)]
        self.topic_count = [raw.topic_count[i] for i in range(raw.world)]
        self.range_count = [raw.range_count[i] for i in range(raw.world)]
        self.n_topics_total, self.n_ranges_total, self.bytes_received = raw.n_topics_total, raw.n_ranges_total, raw.bytes_received
        self._device = device

    def _slices(self, ptr, base, count, width):
        import torch
        whole = device_view(ptr, max(width * base[-1], width), "<u4", self._device)
        parts = [whole[width * base[r]:width * (base[r] + count[r])] for r in range(self.world)]
        return torch.cat(parts) if self.world > 1 else parts[0]

    def route_count(self):
        """matched routes per topic, the ranks' slices concatenated in rank order (padding removed)"""
        return self._slices(self.raw.d_route_count, self.topic_base, self.topic_count, 1)

    def span_count(self):
        return self._slices(self.raw.d_span_count, self.topic_base, self.topic_count, 1)

    def ranges(self):
        """[n_ranges_total, 2] uint32: first rank, count (| 0x80000000 for a multi-segment filter), slices concatenated"""
        return self._slices(self.raw.d_ranges, self.range_base, self.range_count, 2).view(-1, 2)


class Exchange:
    """bfq_exchange_* of the C-ABI (NCCL inside the library). The NCCL unique id travels out of band: here over the
    already-initialised torch.distributed group (any backend); a Java host would use its own RPC."""

    def __init__(self, device_index, rank=None, world=None):
        import ctypes as C

        import torch.distributed as dist

        from . import _native as N
        self._N, self._C = N, C
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
            world = dist.get_world_size() if dist.is_initialized() else 1
        ident = [None]
        if rank == 0:
            buf = (C.c_uint8 * N.EXCHANGE_ID_BYTES)()
            N.check(N.lib.bfq_exchange_unique_id(buf, N.EXCHANGE_ID_BYTES))
            ident[0] = bytes(buf)
        if world > 1:
            dist.broadcast_object_list(ident, src=0)
        self._id = (C.c_uint8 * N.EXCHANGE_ID_BYTES).from_buffer_copy(ident[0])
        h = C.c_void_p()
        N.check(N.lib.bfq_exchange_create(device_index, rank, world, self._id, C.byref(h)))
        self._h, self.rank, self.world, self.device_index = h, rank, world, device_index

    def gather(self, device_result, ranges=True, stream=0):
        """collective: every rank passes the DeviceResult of its own (completed) match"""
        import torch
        N, C = self._N, self._C
        out = N.BfqGathered()
        raw = device_result.raw if hasattr(device_result, "raw") else device_result
        N.check(N.lib.bfq_exchange_gather(self._h, C.byref(raw), N.EXCHANGE_RANGES if ranges else N.EXCHANGE_COUNTS, stream, C.byref(out)))
        return Gathered(out, torch.device("cuda", self.device_index))

    def close(self):
        if getattr(self, "_h", None):
            self._N.lib.bfq_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def range_lookup(tenants, topics, topic_tenant, candidates, device=0):
    """TenantRangeLookupCache.lookup for a batch (bfq_range_lookup). candidates[t] = the ordered candidate ranges of tenant t,
    each None (no Fact) or a pair (first, last) of global filter level lists (either may be None: empty range); level 0 of
    first / last is the tenant id. Returns, per topic, the list of kept candidate indices."""
    import numpy as np

    from . import _native as N
    tb, toff = N.as_blob(tenants)
    pb, poff = N.as_blob(topics)
    tt = np.ascontiguousarray(topic_tenant, dtype=np.int32)
    cand_off = np.zeros(len(tenants) + 1, np.int64)
    flags, firsts, lasts = [], [], []
    for t, cl in enumerate(candidates):
        cand_off[t + 1] = cand_off[t] + len(cl)
        for c in cl:
            if c is None:
                flags.append(0)
                firsts.append(b"")
                lasts.append(b"")
                continue
            first, last = c
            enc = lambda lv: b"\0".join(x.encode("utf-8") if isinstance(x, str) else bytes(x) for x in lv)
            flags.append(1 | (2 if first is not None else 0) | (4 if last is not None else 0))
            firsts.append(enc(first) if first is not None else b"")
            lasts.append(enc(last) if last is not None else b"")
    fl = np.asarray(flags + [0], np.uint8)
    fb, foff = N.as_blob(firsts)
    lb, loff = N.as_blob(lasts)
    keep_off = np.zeros(len(topics) + 1, np.int64)
    total = int(sum(cand_off[t + 1] - cand_off[t] for t in tt.tolist())) if len(topics) else 0
    keep = np.zeros(max(total, 1), np.uint8)
    N.check(N.lib.bfq_range_lookup(device, N.ptr(tb), N.ptr(toff), len(tenants), N.ptr(pb), N.ptr(poff), N.ptr(tt), len(topics),
                                   N.ptr(cand_off), N.ptr(fl), N.ptr(fb), N.ptr(foff), N.ptr(lb), N.ptr(loff), N.ptr(keep_off), N.ptr(keep)))
    return [np.nonzero(keep[keep_off[i]:keep_off[i + 1]])[0].tolist() for i in range(len(topics))]
