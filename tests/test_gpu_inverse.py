"""GPU parity tests of the inverse match (topic filter -> indexed topics; retain-store and TopicIndex): CUDA path
through the C-ABI vs the CPU oracle's restatement of TopicLevelTrie.lookup + the reference selectors. Vectors from
DWT/TopicIndexTest.java:41-139, RST/index/RetainTopicIndexTest.java:42-117, RST/RetainMatchTest.java:38-124
(DWT = bifromq-dist/bifromq-dist-worker/src/test/java/org/apache/bifromq/dist/worker,
 RST = bifromq-retain/bifromq-retain-store/src/test/java/org/apache/bifromq/retain/store)."""
import random

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_golden import INDEXED, INVERSE_CASES, TOPIC_INDEX_ONLY

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import bifromq_b200
    from bifromq_b200 import retain, workload
    bifromq_b200.load_library()

    class NS:
        pass
    ns = NS()
    ns.retain, ns.workload = retain, workload
    return ns


def test_topic_index_match_vectors(R):  # TopicIndexTest.testMatch :41-73
    idx = R.retain.GpuTopicIndex(0)
    for t in INDEXED:
        idx.add(t, t)
    cases = dict(INVERSE_CASES)
    cases.update(TOPIC_INDEX_ONLY)
    for f, want in cases.items():
        assert idx.match(f) == set(want), f
    for t in INDEXED:  # testGet :75-88
        assert idx.get(t) == {t}


def test_topic_index_remove_multivalue_edge(R):  # TopicIndexTest.testRemove/testMultiValue/testEdgeCases :90-139
    idx = R.retain.GpuTopicIndex(0)
    for t in INDEXED:
        idx.add(t, t)
    for t in INDEXED:
        idx.remove(t, t)
        assert idx.match(t) == set()
    assert idx.match("#") == set()
    idx.add("a", "a1")
    idx.add("a", "a1")
    idx.add("a", "a2")
    assert idx.get("a") == {"a1", "a2"} and idx.match("+") == {"a1", "a2"}
    idx.remove("a", "a3")
    assert idx.get("a") == {"a1", "a2"}
    idx.remove("a", "a2")
    assert idx.match("a") == {"a1"}
    idx.remove("a", "a1")
    assert idx.match("a") == set()
    idx2 = R.retain.GpuTopicIndex(0)
    idx2.add("/", "/")
    idx2.add("/", "/")
    assert idx2.match("#") == {"/"}


def test_retain_topic_index_vectors(R):  # RetainTopicIndexTest.testMatch/testRemove/testEdgeCases :42-117
    tenant = "tenantA"
    idx = R.retain.GpuRetainTopicIndex(0)
    for t in INDEXED:
        idx.add(tenant, t, 1, 2)
    for f, want in INVERSE_CASES.items():
        assert {m[1] for m in idx.match(tenant, f)} == set(want), f
    assert idx.match("tenantB", "#") == set()
    assert {m[1] for m in idx.find_all()} == set(INDEXED)  # RetainTopicIndexTest.testFindAll :77-82
    for t in INDEXED:
        idx.remove(tenant, t)
        assert idx.match(tenant, t) == set()
    assert idx.match(tenant, "#") == set()
    idx.add(tenant, "/")
    idx.add(tenant, "/")
    assert {m[1] for m in idx.match(tenant, "#")} == {"/"}


def test_retain_match_vectors_and_limit(R):  # RetainMatchTest.wildcardTopicFilter/matchLimit :38-124
    tenant = "tenantA"
    msgs = ["/a/b/c", "/a/b/", "/c/", "a"]
    idx = R.retain.GpuRetainTopicIndex(0)
    for t in msgs:
        idx.add(tenant, t)
    cases = {"#": [0, 1, 2, 3], "+": [3], "+/#": [0, 1, 2, 3], "+/+/#": [0, 1, 2], "+/+/+": [2], "/#": [0, 1, 2],
             "/c/#": [2], "/a/+": [], "/a/#": [0, 1], "/a/+/+": [0, 1], "/a/+/#": [0, 1], "/+/b/": [1],
             "/+/b/#": [0, 1], "/a/b/c/#": [0], "/a/b/#": [0, 1]}
    for f, want in cases.items():
        assert {m[1] for m in idx.match(tenant, f, 10)} == {msgs[i] for i in want}, f
    assert len(idx.match(tenant, "#", 0)) == 0
    assert len(idx.match(tenant, "#", 1)) == 1
    full = idx.match(tenant, "#")
    assert idx.match(tenant, "#", 2) <= full and len(idx.match(tenant, "#", 2)) == 2


def _random_case(rng, n_topics, n_filters, vocab, depth):
    def lvl(i):
        r = rng.random()
        if r < 0.08:
            return ""
        if r < 0.2 and i == 0:
            return "$" + rng.choice(vocab)
        return rng.choice(vocab)
    tenants = ["tA", "tB"]
    topics = sorted({(rng.choice(tenants), "/".join(lvl(i) for i in range(rng.randint(1, depth)))) for _ in range(n_topics)})
    filters = []
    for _ in range(n_filters):
        n = rng.randint(1, depth)
        lv = []
        for i in range(n):
            r = rng.random()
            lv.append("+" if r < 0.3 else ("#" if r < 0.45 and i == n - 1 else lvl(i)))
        filters.append((rng.choice(tenants + ["tC"]), "/".join(lv)))
    return tenants + ["tC"], topics, filters


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_inverse_vs_oracle(R, seed):
    rng = random.Random(seed)
    tenants, topics, filters = _random_case(rng, 1500, 400, ["a", "b", "c", "dd", "x" * 30], 5)
    idx = R.retain.GpuTopicMatchIndex(0)
    orc = O.TopicLevelIndex()
    from bifromq_b200 import _native as N
    blob, off = N.as_blob([t[1] for t in topics])
    tt = np.array([tenants.index(t[0]) for t in topics], np.int32)
    ids = idx.add_blobs(tenants, blob, off, tt)
    for (tenant, topic), i in zip(topics, ids.tolist()):
        orc.add(topic, i, tenant)
    idx.commit()
    fblob, foff = N.as_blob([f[1] for f in filters])
    ft = np.array([tenants.index(f[0]) for f in filters], np.int32)
    res = idx.match_blobs(tenants, fblob, foff, ft)
    n_nonempty = 0
    for i, (tenant, f) in enumerate(filters):
        want = orc.match(f, tenant)
        got = sorted(res.matches(i).tolist())
        assert got == want, (tenant, f)
        assert res.totals[i] == len(want)
        n_nonempty += bool(want)
    assert n_nonempty > 50
    # with a limit: any `limit`-subset of the full match set (the reference pins only the count, RetainMatchTest.java:113-124)
    lim = np.array([rng.choice([0, 1, 3, 10]) for _ in filters], np.int64)
    res2 = idx.match_blobs(tenants, fblob, foff, ft, lim)
    for i, (tenant, f) in enumerate(filters):
        want = set(orc.match(f, tenant))
        got = res2.matches(i).tolist()
        assert len(got) == min(len(want), int(lim[i])) and set(got) <= want and len(set(got)) == len(got)
        assert res2.totals[i] == len(want)
    # removal
    for tenant, topic in topics[::3]:
        idx.remove(tenant, topic)
        orc.remove(topic, ids[topics.index((tenant, topic))], tenant)
    idx.commit()
    res3 = idx.match_blobs(tenants, fblob, foff, ft)
    for i, (tenant, f) in enumerate(filters):
        assert sorted(res3.matches(i).tolist()) == orc.match(f, tenant)


def test_inverse_wide_fanout_goes_through_tier2(R):
    # 200 first-level names x 3 second-level names: "+/+/x" walks an interval frontier, "+/b/+" produces 200 single-node intervals
    idx = R.retain.GpuTopicMatchIndex(0)
    orc = O.TopicLevelIndex()
    topics = ["n%03d/%s/%s" % (i, b, c) for i in range(200) for b in "abc" for c in "xyz"] + ["$s/a/x", "n000", "n001/a"]
    ids = idx.add("t", topics)
    for t, i in zip(topics, ids.tolist()):
        orc.add(t, i, "t")
    idx.commit()
    filters = ["+/+/x", "+/b/+", "+/b/#", "+/+/+", "#", "+/#", "+/+/#", "n005/+/+", "+/a", "+", "$s/#", "$s/+/+", "+/b/x/#", "+/+/x/#"]
    res = idx.match("t", filters)
    for i, f in enumerate(filters):
        assert sorted(res.matches(i).tolist()) == orc.match(f, "t"), f


def test_c5_config_scaled_vs_oracle(R):
    w = R.workload.Workload("C5", scale=0.01)
    idx = R.retain.GpuTopicMatchIndex(0)
    tenants = w.tenants
    ids = idx.add_blobs(tenants, w.topics, w.topic_off, w.topic_tenant)
    idx.commit()
    orc = O.TopicLevelIndex()
    tl = w.topic_list()
    for i in range(w.n_topics):
        orc.add(tl[i], int(ids[i]), tenants[w.topic_tenant[i]])
    res = idx.match_blobs(tenants, w.filters, w.filter_off, w.filter_tenant)
    fl = w.query_filter_list()
    hits = 0
    for i in range(w.n_query_filters):
        want = orc.match(fl[i], tenants[w.filter_tenant[i]])
        assert sorted(res.matches(i).tolist()) == want
        hits += bool(want)
    assert hits > 0.5 * w.n_query_filters
    lim = np.full(w.n_query_filters, 10, np.int64)
    res10 = idx.match_blobs(tenants, w.filters, w.filter_off, w.filter_tenant, lim)
    assert (np.diff(res10.offsets) == np.minimum(res.totals, 10)).all()


def test_retain_store_feed_from_raw_keys_and_batched_get_keys(R):
    """RetainStoreCoProc.load() (RS/RetainStoreCoProc.java:279-296) rebuilds the index from a range scan; bfq_rindex_load_keys takes
    the scan's raw KEYS (the topic is in the key, no value parsing), and bfq_rresult_retain_keys returns the retainMessageKey of
    every matched topic as one batch — the keys of RetainStoreCoProc.match's follow-up reader.get calls (:177-188). Same
    answers as feeding (tenant, topic) strings; junk keys are skipped; the keys equal the oracle's retainMessageKey."""
    w = R.workload.Workload("C5", scale=0.01)
    tenants = w.tenants
    tl = w.topic_list()
    keys = [O.retain_key(tenants[w.topic_tenant[i]], tl[i]) for i in range(w.n_topics)]
    keys_sorted = sorted(keys)                       # a range scan delivers them in key order
    junk = [b"\x00\x00\x01t", b"garbage", keys_sorted[0][:-1] + b"\x00extra-level"]
    from bifromq_b200 import _native as N
    kb, ko = N.as_blob(keys_sorted + junk)
    a = R.retain.GpuTopicMatchIndex(0)
    ids = a.load_keys(kb, ko)
    assert (ids[:len(keys_sorted)] >= 0).all() and (ids[len(keys_sorted):] == -1).all()
    a.commit()
    b = R.retain.GpuTopicMatchIndex(0)
    b.add_blobs(tenants, w.topics, w.topic_off, w.topic_tenant[:w.n_topics])
    b.commit()
    ra = a.match_blobs(tenants, w.filters, w.filter_off, w.filter_tenant[:w.n_query_filters], with_retain_keys=True)
    rb = b.match_blobs(tenants, w.filters, w.filter_off, w.filter_tenant[:w.n_query_filters])
    assert ra.offsets.tolist() == rb.offsets.tolist() and int(ra.totals.sum()) > w.n_query_filters
    blob, koff = ra.retain_keys
    assert len(koff) == len(ra.ids) + 1

    def topic_of(idx, i):
        t, p = idx.lookup(int(i))
        return t, p
    for f in range(0, w.n_query_filters, 23):
        sa = sorted(topic_of(a, i) for i in ra.matches(f))
        sb = sorted(topic_of(b, i) for i in rb.matches(f))
        assert sa == sb
        for j in range(int(ra.offsets[f]), int(ra.offsets[f + 1])):
            t, p = topic_of(a, ra.ids[j])
            assert bytes(blob[koff[j]:koff[j + 1]]) == O.retain_key(t, p)
