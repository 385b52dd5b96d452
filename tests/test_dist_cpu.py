"""N > 1 logic on CPU: two gloo ranks each own the tenants fnv1a64(tenant) % 2 == rank, match their shard with the
oracle (no GPU here), and the aggregated result equals the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib as O
    from bifromq_b200 import dist as D
    from bifromq_b200.workload import Workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = Workload("C3", scale=0.004, shard_index=rank, shard_count=world)
    assert all(D.tenant_shard(t, world) == rank for t in w.tenants)
    kv = O.KV()
    kv.load(w.keys, w.key_off, w.vals, w.val_off)
    tb, toff = O.blob(w.tenants)
    out = kv.match_blobs(tb, toff, w.topics, w.topic_off, np.ascontiguousarray(w.topic_tenant), w.n_topics, 2 ** 31 - 1, 100,
                         O.MODE_TRIE, False, 1)
    routes = int(np.diff(out.offsets).sum())
    # a fake per-rank time: rank r "took" (r + 1) * 10 ms -> whole-job time is the max
    ms, units = D.aggregate((rank + 1) * 10.0, w.n_topics)
    _, total_routes = D.aggregate(0.0, routes)
    # the exchange step: per-topic fan-out counts of every rank's topics, reassembled in rank order on every rank
    import torch
    fan = torch.from_numpy(np.diff(out.offsets).astype(np.int32))
    gathered = D.gather_fanout(fan)
    assert gathered.numel() == int(units) and int(gathered.sum()) == int(total_routes)
    parts = D.all_gather_varlen(fan)
    assert torch.equal(parts[rank], fan) and len(parts) == world
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, w.n_topics, routes, ms, units, total_routes, sorted(w.tenants)))


def test_two_rank_tenant_sharding_matches_single_process():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from bifromq_b200 import dist as D
    from bifromq_b200.workload import Workload
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = Workload("C3", scale=0.004)
    kv = O.KV()
    kv.load(full.keys, full.key_off, full.vals, full.val_off)
    tb, toff = O.blob(full.tenants)
    out = kv.match_blobs(tb, toff, full.topics, full.topic_off, np.ascontiguousarray(full.topic_tenant), full.n_topics,
                         2 ** 31 - 1, 100, O.MODE_TRIE, False, 2)
    total_routes = int(np.diff(out.offsets).sum())
    assert sum(r[1] for r in results) == full.n_topics
    assert sum(r[2] for r in results) == total_routes
    for r in results:
        assert r[3] == 20.0                      # MAX over ranks of the per-rank time
        assert r[4] == float(full.n_topics)      # SUM over ranks of the units
        assert r[5] == float(total_routes)
    assert sorted(results[0][6] + results[1][6]) == sorted(full.tenants)
    # the batch splitter used by a front-end agrees with the generator's shard function
    parts = D.split_batch_by_owner(full.tenants, full.topic_tenant[:full.n_topics], 2)
    assert [len(p) for p in parts] == [results[0][1], results[1][1]]


def test_fnv1a64_known_answers():
    from bifromq_b200 import dist as D
    assert D.fnv1a64(b"") == 0xCBF29CE484222325
    assert D.fnv1a64(b"a") == 0xAF63DC4C8601EC8C
    assert D.fnv1a64("foobar") == 0x85944171F73967E8


def _reassemble_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from bifromq_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tenants = ["tenant-%d" % i for i in range(40)]
    rng = np.random.default_rng(7)
    topic_tenant = rng.integers(0, len(tenants), 1000)
    truth = (np.arange(1000) * 7 + 3).astype(np.int32)          # the "fan-out" of topic i, known to its owner only
    parts = D.split_batch_by_owner(tenants, topic_tenant, world)
    mine = torch.from_numpy(truth[parts[rank]])
    whole = D.gather_fanout(mine, parts, 1000)
    ok = bool(np.array_equal(whole.numpy(), truth))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_fanout_exchange_reassembles_batch_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reassemble_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True), (1, True)]


def test_replica_split_covers_the_batch():
    sys.path.insert(0, ROOT)
    from bifromq_b200 import dist as D
    for n, wsz in ((10, 3), (0, 2), (7, 8), (1000000, 8)):
        cuts = D.split_batch_replicas(n, wsz)
        assert len(cuts) == wsz and cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert max(e - b for b, e in cuts) - min(e - b for b, e in cuts) <= 1


def test_hot_tenant_replication_balances_the_split():
    """strong scaling: tenants above 1 / (4 world) of the batch are hosted by every rank and their topics dealt round-robin by
    batch position; the generator's shards and the host-side split agree and the union is the whole batch"""
    sys.path.insert(0, ROOT)
    from bifromq_b200 import dist as D
    from bifromq_b200.workload import Workload
    world = 4
    full = Workload("C4", scale=0.02)
    tt = np.asarray(full.topic_tenant[:full.n_topics])
    share = np.bincount(tt, minlength=full.n_tenants).astype(float)
    hot = D.hot_tenants(share, world)
    assert 1 <= hot.sum() <= 6
    parts = D.split_batch_by_owner(full.tenants, tt, world, hot)
    assert sorted(np.concatenate(parts).tolist()) == list(range(full.n_topics))
    sizes = [len(p) for p in parts]
    plain = [len(p) for p in D.split_batch_by_owner(full.tenants, tt, world)]
    assert max(sizes) / np.mean(sizes) < max(plain) / np.mean(plain)
    names = full.tenants
    shards = [Workload("C4", scale=0.02, shard_index=r, shard_count=world, replicate_hot=True) for r in range(world)]
    assert sum(s.n_topics for s in shards) == full.n_topics
    for r, s in enumerate(shards):
        # the shard holds exactly the topics the split assigns to rank r, in batch order
        want = [(names[tt[i]], full.topic(int(i))) for i in parts[r]] if abs(len(parts[r]) - s.n_topics) == 0 else None
        # the generator decides "hot" from the tenants' planned sizes, the split above from the observed batch: they agree on
        # the big tenants, so the per-rank topic lists are identical
        assert want is not None
        sn = s.tenants
        got = [(sn[s.topic_tenant[i]], s.topic(i)) for i in range(s.n_topics)]
        assert got == want
