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
