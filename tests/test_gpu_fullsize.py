"""Full-size parity for BASELINE.json's forward configs on the GPU box: the CUDA path against the oracle, EXHAUSTIVELY (every
topic's surviving route ranks and every throttle event), at scale 1.0 — C4 (10M filters, 1M-topic batch), C2 (1 tenant, 1M
filters with 50 % '+', 100k topics) and C3 (1000 tenants x 10k filters, 1M topics) — under the reference's default caps
(MaxPersistentFanout = INT_MAX, MaxGroupFanout = 100) and under the stress caps (4, 4) BASELINE.md asks for, plus the
size-independent properties (dense spans, idempotence, permutation equivariance). The ordering / de-duplication kernels and
the 2^16-slot perfect-hash child arrays are only reached at these sizes. Comparisons are numpy array equalities (the C4
result is 3.6e8 ranks). The oracle is built once per config (its trie build is single-threaded: ~1 min for C4)."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
INT_MAX = 2 ** 31 - 1
THREADS = os.cpu_count() or 8


class Full:
    """one config at scale 1.0: workload, CUDA index and the oracle over the same KV"""

    def __init__(self, config):
        import bifromq_b200
        from bifromq_b200 import workload
        self.w = w = workload.Workload(config)
        self.idx = bifromq_b200.GpuRouteIndex(0)
        self.idx.load(w.keys, w.key_off, w.vals, w.val_off)
        self.idx.commit()
        self.kv = O.KV()
        self.kv.load(w.keys, w.key_off, w.vals, w.val_off)
        self.kv.freeze()
        self.names = w.tenants
        self.tenants = self.idx.tenant_blob(self.names)
        self.tb, self.toff = O.blob(self.names)
        self.tt = np.ascontiguousarray(w.topic_tenant[:w.n_topics]).astype(np.int32)

    def gpu(self, lo, hi, max_p, max_g):
        nt = len(self.names)
        off = np.ascontiguousarray(self.w.topic_off[lo:hi + 1])
        r = self.idx.match(self.tenants, self.w.topics, off, self.tt[lo:hi], [max_p] * nt, [max_g] * nt)
        offsets, ranks = r.expand()
        ev = r.throttled.copy()
        rc = r.route_count.copy()
        n_ranges = int(r.span_count.astype(np.int64).sum())
        r.close()
        return offsets, ranks, ev, rc, n_ranges

    def cpu(self, lo, hi, max_p, max_g):
        off = np.ascontiguousarray(self.w.topic_off[lo:hi + 1])
        return self.kv.match_blobs(self.tb, self.toff, self.w.topics, off, self.tt[lo:hi], hi - lo, max_p, max_g, O.MODE_TRIE, False, THREADS)

    def check(self, lo, hi, max_p, max_g):
        offsets, ranks, ev, rc, n_ranges = self.gpu(lo, hi, max_p, max_g)
        want = self.cpu(lo, hi, max_p, max_g)
        assert np.array_equal(offsets, want.offsets), "per-topic surviving route counts differ"
        assert np.array_equal(ranks, want.ranks), "surviving route ranks differ"
        got_ev = sorted(zip(ev["kind"].tolist(), ev["topic"].tolist(), ev["rank"].tolist()))
        assert got_ev == [(k, t, r) for k, t, r, _ in want.events], "throttle events differ"
        return want, rc, n_ranges, len(ranks), len(got_ev)


@pytest.fixture(scope="module")
def c4():
    return Full("C4")


def test_c4_full_size_exhaustive_default_caps(c4):
    """the whole 1M-topic batch, reference default caps: bit-exact; and the §8(d) counters of the SAME population — the
    roofline numerator bench.py prints — land on the judge's whole-batch figure (1162 B/topic) and on the GPU's own counts"""
    w = c4.w
    assert w.n_filters > 9_900_000 and w.n_topics == 1_000_000
    st = c4.idx.stats()
    assert st["routes"] == w.n_routes and st["tenants"] == w.n_tenants
    before = c4.idx.stats()
    want, rc, n_ranges, n_ranks, n_ev = c4.check(0, w.n_topics, INT_MAX, 100)
    after = c4.idx.stats()
    assert n_ranks > 300_000_000
    s = want.stats
    n = w.n_topics
    per_topic = (float(w.topic_off[n] - w.topic_off[0]) + 8 * n + 32 * s["V"] + 8 * s["P"] + 8 * s["ranges"]) / n
    assert 1150 < per_topic < 1175, per_topic
    assert s["ranges"] == n_ranges                      # matched filters with >= 1 route
    assert s["R"] == int(rc.astype(np.int64).sum())     # matched routes before caps
    # a third of the batch are repeats of an earlier (tenant, topic) pair: answered from the first occurrence (the host path
    # looks for them inside each of its four sub-batches: 2.3e5 found of the 3.3e5 the whole batch holds)
    assert after["duplicate_topics"] - before["duplicate_topics"] > 150_000
    assert after["overflow_topics"] == before["overflow_topics"]


def test_c4_full_size_stress_caps(c4):
    """BASELINE.md's stress run, both caps = 4: first 4 persistent / group routes in KV order survive, every later one is an
    event (200k topics: the event list is a large fraction of the 7e7 matched routes)"""
    want, rc, _, n_ranks, n_ev = c4.check(300_000, 500_000, 4, 4)
    assert n_ev > 1000


def test_c4_full_size_properties(c4):
    w, idx, tenants, tt = c4.w, c4.idx, c4.tenants, c4.tt

    def run(topics, off, tenant_idx):
        r = idx.match(tenants, topics, off, tenant_idx)
        out = (r.span_begin.copy(), r.span_count.copy(), r.route_count.copy(), r.ranges.copy(), len(r.throttled))
        r.close()
        return out
    sb, sc, rc, rg, nthr = run(w.topics, w.topic_off, tt)
    # (1) dense ranges: every distinct span is a slice of the range array, the distinct spans tile it exactly (a repeated
    # (tenant, topic) pair shares the span of its first occurrence), and a topic's range counts add up to its route count
    assert nthr == 0
    sb64 = sb.astype(np.int64)
    assert (sb64 + sc <= len(rg)).all()
    spans = np.unique(np.stack([sb64[sc > 0], sc[sc > 0].astype(np.int64)], axis=1), axis=0)
    assert spans[0, 0] == 0 and (spans[1:, 0] == spans[:-1, 0] + spans[:-1, 1]).all() and spans[-1, 0] + spans[-1, 1] == len(rg)
    assert idx.stats()["multi_segment_filters"] == 0
    csum = np.concatenate([[0], np.cumsum(rg["count"].astype(np.int64))])
    assert (csum[sb64 + sc] - csum[sb64] == rc).all()
    assert (rg["first"].astype(np.int64) + rg["count"] <= w.n_routes).all()
    # (2) idempotence: the same batch again gives the same answer
    sb2, sc2, rc2, rg2, _ = run(w.topics, w.topic_off, tt)
    assert (sc2 == sc).all() and (rc2 == rc).all()
    key = lambda a: np.sort(a.view(np.uint64))   # ranges of one topic may come out in a different order
    assert (key(rg2) == key(rg)).all()
    # (3) permutation equivariance on a shuffled 50k-topic sub-batch
    rng = np.random.RandomState(1)
    pick = rng.permutation(w.n_topics)[:50_000]
    tl = [w.topic(int(i)) for i in pick]
    blob = np.frombuffer(b"".join(tl), dtype=np.uint8).copy()
    off = np.zeros(len(tl) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in tl])
    _, sc3, rc3, _, _ = run(blob, off, np.ascontiguousarray(tt[pick]))
    assert (sc3 == sc[pick]).all() and (rc3 == rc[pick]).all()


def test_c4_delta_commits_at_full_size(c4):
    """bfq_index_commit's delta path at 10M filters: one SUB into a mid-sized tenant, one UNSUB, a brand-new tenant and a route
    into the LARGEST tenant — each commit rebuilds only the touched tenant, the result equals the oracle fed the same
    mutations (a 100k-topic slice that covers every tenant), and the small commits take milliseconds, not the 3 s of a
    full build"""
    import time
    from bifromq_b200 import schema
    w, idx, kv = c4.w, c4.idx, c4.kv
    names = c4.names
    before = idx.stats()
    mid = names[len(names) // 2]
    muts = [("add", schema.route_key(mid, "delta/+/x", schema.receiver_url(0, "newcomer", "d")), schema.incarnation_bytes(3)),
            ("add", schema.route_key("zz-new-tenant", "#", schema.receiver_url(1, "p", "d")), schema.incarnation_bytes(1)),
            ("del", bytes(w.keys[w.key_off[w.n_routes // 3]:w.key_off[w.n_routes // 3 + 1]]), None),
            ("add", schema.route_key(names[0], "#", schema.receiver_url(0, "catch", "d")), schema.incarnation_bytes(9))]
    times = []
    for kind, k, v in muts:
        if kind == "add":
            idx.apply(adds=[(k, v)])
            kv.put(k, v)
        else:
            idx.apply(dels=[k])
            kv.erase(k)
        t0 = time.perf_counter()
        idx.commit()
        times.append(time.perf_counter() - t0)
    after = idx.stats()
    assert after["delta_commits"] - before["delta_commits"] == len(muts) and after["full_commits"] == before["full_commits"]
    assert after["routes"] == w.n_routes + 2
    kv.freeze()
    # tenants list grew: the new tenant sorts last, topics of every tenant still resolve
    all_names = names + ["zz-new-tenant"]
    tenants = idx.tenant_blob(all_names)
    tb, toff = O.blob(all_names)
    lo, hi = 100_000, 200_000
    off = np.ascontiguousarray(w.topic_off[lo:hi + 1])
    nt = len(all_names)
    r = idx.match(tenants, w.topics, off, c4.tt[lo:hi], [INT_MAX] * nt, [100] * nt)
    offsets, ranks = r.expand()
    r.close()
    want = kv.match_blobs(tb, toff, w.topics, off, c4.tt[lo:hi], hi - lo, INT_MAX, 100, O.MODE_TRIE, False, THREADS)
    assert np.array_equal(offsets, want.offsets) and np.array_equal(ranks, want.ranks)
    # a commit costs the rebuild of the touched tenant plus a device-side copy: milliseconds for an ordinary tenant (tools/
    # commit_bench.py records 5 ms at this size), well under a second for the 1.4M-route tenant — never a full build. The two
    # small commits are looked at together: a single one can hit a slow cudaMalloc of the 2.7 GB snapshot copy (the commit
    # traces under profiles/ show 0.2 - 0.4 s outliers in "device allocations")
    assert min(times[0], times[1]) < 0.25 and max(times) < 2.5, times
    print("delta commit seconds:", [round(t, 4) for t in times])


@pytest.mark.parametrize("config", ["C2", "C3"])
def test_c2_c3_full_size_exhaustive(config):
    f = Full(config)
    n = f.w.n_topics
    assert n == (100_000 if config == "C2" else 1_000_000)
    _, _, _, n_ranks, _ = f.check(0, n, INT_MAX, 100)
    assert n_ranks > n // 4
    f.check(0, min(n, 200_000), 4, 4)


def test_c5_full_size_exhaustive():
    """BASELINE config C5 at scale 1.0 (1M retained topics, 100k wildcard SUBSCRIBE filters): with limit = unlimited every
    filter's id set equals the oracle's TopicLevelTrie restatement; with the default limit of 10 every filter returns
    min(total, 10) ids and each of them is a member of its full match set (which `limit` ids come back is unpinned in the
    reference too: it iterates a HashSet, RetainStoreCoProc.java:177-188). Also concurrently from four threads on one handle."""
    import threading
    from bifromq_b200 import retain, workload
    w = workload.Workload("C5")
    assert w.n_topics >= 990_000 and w.n_query_filters == 100_000
    idx = retain.GpuTopicMatchIndex(0)
    tenants = w.tenants
    ids = idx.add_blobs(tenants, w.topics, w.topic_off, w.topic_tenant[:w.n_topics])
    idx.commit()
    orc = O.TopicLevelIndex()
    tl = w.topic_list()
    for i in range(w.n_topics):
        orc.add(tl[i], int(ids[i]), tenants[w.topic_tenant[i]])
    n = w.n_query_filters
    ft = np.ascontiguousarray(w.filter_tenant[:n])
    res = idx.match_blobs(tenants, w.filters, w.filter_off, ft)
    fl = w.query_filter_list()
    hits = 0
    for i in range(n):
        want = orc.match(fl[i], tenants[ft[i]])
        got = np.sort(res.matches(i))
        assert len(got) == len(want) and got.tolist() == want, fl[i]
        hits += bool(want)
    assert hits > 0.5 * n and int(res.totals.sum()) > 1_000_000
    lim = np.full(n, 10, np.int64)
    res10 = idx.match_blobs(tenants, w.filters, w.filter_off, ft, lim)
    assert (np.diff(res10.offsets) == np.minimum(res.totals, 10)).all() and (res10.totals == res.totals).all()
    for i in range(0, n, 7):
        assert np.isin(res10.matches(i), res.matches(i)).all()
    # four threads, different slices of the filter batch, one handle: every result is its own (results own their arrays)
    outs, errs = {}, []

    def worker(k):
        try:
            b, e = k * (n // 4), (k + 1) * (n // 4)
            off = np.ascontiguousarray(w.filter_off[b:e + 1])
            for _ in range(3):
                outs[k] = idx.match_blobs(tenants, w.filters, off, np.ascontiguousarray(ft[b:e]))
        except Exception as ex:   # pragma: no cover
            errs.append(ex)
    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for k in range(4):
        b = k * (n // 4)
        assert (outs[k].totals == res.totals[b:b + n // 4]).all()
        for i in range(0, n // 4, 101):
            assert np.sort(outs[k].matches(i)).tolist() == np.sort(res.matches(b + i)).tolist()
