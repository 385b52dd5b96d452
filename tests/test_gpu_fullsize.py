"""Full-size run (BASELINE.json config C4: 10M filters, 1M-topic batch) checked through size-independent properties, plus an
oracle spot check on a bounded sample of tenants. The exhaustive bit-exact comparisons live in test_gpu_forward.py at sizes the
oracle finishes in seconds."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def test_c4_full_size_properties():
    import bifromq_b200
    from bifromq_b200 import workload
    w = workload.Workload("C4")
    assert w.n_filters > 9_900_000 and w.n_topics == 1_000_000
    idx = bifromq_b200.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    st = idx.stats()
    assert st["routes"] == w.n_routes and st["tenants"] == w.n_tenants
    tenants = idx.tenant_blob(w.tenants)
    tt = np.ascontiguousarray(w.topic_tenant[:w.n_topics])

    def run(topics, off, tenant_idx):
        r = idx.match(tenants, topics, off, tenant_idx)
        out = (r.span_begin.copy(), r.span_count.copy(), r.route_count.copy(), r.ranges.copy(), len(r.throttled))
        r.close()
        return out
    sb, sc, rc, rg, nthr = run(w.topics, w.topic_off, tt)
    # (1) dense topic-ordered ranges; a topic's range counts add up to its route count (no multi-segment filters in C4)
    assert nthr == 0
    assert sb[0] == 0 and (sb[1:].astype(np.int64) == sb[:-1].astype(np.int64) + sc[:-1]).all() and int(sb[-1]) + int(sc[-1]) == len(rg)
    assert st["multi_segment_filters"] == 0
    per_topic = np.add.reduceat(rg["count"].astype(np.int64), sb[sc > 0].astype(np.int64)) if (sc > 0).any() else np.zeros(0)
    assert (per_topic == rc[sc > 0]).all()
    assert (rg["first"].astype(np.int64) + rg["count"] <= w.n_routes).all()
    # (2) idempotence: the same batch again gives the same answer
    sb2, sc2, rc2, rg2, _ = run(w.topics, w.topic_off, tt)
    assert (sc2 == sc).all() and (rc2 == rc).all()
    key = lambda a: np.sort(a.view(np.uint64))   # ranges of one topic may come out in a different order
    assert (key(rg2) == key(rg)).all()
    # (3) permutation equivariance on a shuffled 200k-topic sub-batch
    rng = np.random.RandomState(1)
    pick = rng.permutation(w.n_topics)[:200_000]
    tl = [w.topic(int(i)) for i in pick[:50_000]]
    blob = np.frombuffer(b"".join(tl), dtype=np.uint8).copy()
    off = np.zeros(len(tl) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in tl])
    _, sc3, rc3, _, _ = run(blob, off, np.ascontiguousarray(tt[pick[:50_000]]))
    assert (sc3 == sc[pick[:50_000]]).all() and (rc3 == rc[pick[:50_000]]).all()
    # (4) oracle spot check: every topic of four tenants (incl. the largest) bit-exact
    names = w.tenants
    chosen = [0, 7, 123, 999]
    sel = np.nonzero(np.isin(tt, chosen))[0]
    kv = O.KV()
    kb = memoryview(w.keys)

    def lower_bound(key_bytes):
        lo, hi = 0, w.n_routes
        while lo < hi:
            mid = (lo + hi) // 2
            if bytes(kb[w.key_off[mid]:w.key_off[mid + 1]]) < key_bytes:
                lo = mid + 1
            else:
                hi = mid
        return lo
    base = {}
    for t in chosen:
        b = O.tenant_begin_key(names[t])
        lo, hi = lower_bound(b), lower_bound(O.upper_bound(b))
        base[t] = lo
        O.lib.orc_kv_load(kv.h, w.keys.ctypes.data, np.ascontiguousarray(w.key_off[lo:hi + 1]), w.vals.ctypes.data,
                          np.ascontiguousarray(w.val_off[lo:hi + 1]), hi - lo)
    kv.freeze()
    sub_names = [names[t] for t in chosen]
    # oracle ranks are positions inside the 4-tenant KV: translate to global ranks through each tenant's first rank
    order = sorted(chosen, key=lambda t: O.tenant_begin_key(names[t]))
    sizes, acc = {}, 0
    for t in order:
        b = O.tenant_begin_key(names[t])
        n_t = lower_bound(O.upper_bound(b)) - base[t]
        sizes[t] = (acc, n_t)
        acc += n_t
    topics = [w.topic(int(i)) for i in sel]
    sub_tt = np.array([chosen.index(int(x)) for x in tt[sel]], np.int32)
    want = kv.match_batch(sub_names, topics, sub_tt, mode=O.MODE_TRIE, nthreads=8)
    res = idx.match_topics(tenants, topics, np.ascontiguousarray(tt[sel]))
    offsets, ranks = res.expand()
    res.close()
    assert offsets.tolist() == want.offsets.tolist()
    local = want.ranks.copy()
    # map local oracle ranks to global ranks
    glob = np.empty_like(local)
    starts = np.array([sizes[t][0] for t in order]); ends = starts + np.array([sizes[t][1] for t in order])
    for t, s0, e0 in zip(order, starts, ends):
        m = (local >= s0) & (local < e0)
        glob[m] = local[m] - s0 + base[t]
    assert ranks.tolist() == glob.tolist()
    assert len(ranks) > 100_000
