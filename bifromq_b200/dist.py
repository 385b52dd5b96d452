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


def split_batch_by_owner(tenants, topic_tenant, world_size):
    """indices of the topics each rank must match -> list of int64 arrays (the dist-server side of the sharding)"""
    owner_of_tenant = np.array([tenant_shard(t, world_size) for t in tenants], dtype=np.int64)
    owner = owner_of_tenant[np.asarray(topic_tenant, dtype=np.int64)] if len(tenants) else np.zeros(0, np.int64)
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
