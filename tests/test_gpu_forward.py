"""GPU parity tests of the forward match (publish topic -> routes): the CUDA path, called through the C-ABI,
against the CPU oracle on the same bytes. Bit-exact: identical sets of route ranks per topic, identical
throttle events. Cases named after the reference tests they transcribe
(bifromq-dist/bifromq-dist-worker/src/test/java/org/apache/bifromq/dist/worker/cache/TenantRouteMatcherTest.java).
"""
import random

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

TENANT_ID, OTHER_TENANT = "tenantA", "tenantB"


@pytest.fixture(scope="module")
def B():
    import bifromq_b200
    from bifromq_b200 import schema, workload
    bifromq_b200.load_library()

    class NS:
        pass
    ns = NS()
    ns.pkg, ns.schema, ns.workload = bifromq_b200, schema, workload
    return ns


class Events:
    def __init__(self):
        self.events = []

    def report(self, e):
        self.events.append(e)


def make_index(B, pairs):
    idx = B.pkg.GpuRouteIndex(0)
    idx.load_pairs(pairs)
    idx.commit()
    return idx


def normal(B, pairs, tenant, tf, broker, receiver, deliverer, inc):
    url = B.schema.receiver_url(broker, receiver, deliverer)
    pairs.append((B.schema.route_key(tenant, tf, url), B.schema.incarnation_bytes(inc)))
    return B.schema.NormalMatching(tenant, tf, url, inc)


def group(B, pairs, tenant, tf, grp, members, ordered=False):
    full = ("$oshare/" if ordered else "$share/") + grp + "/" + tf
    pairs.append((B.schema.route_key(tenant, full), B.schema.route_group_bytes(members)))
    return B.schema.GroupMatching(tenant, full, tuple(sorted(members.items())), ordered)


# ------------------------------------------------------------------ TenantRouteMatcherTest.java:89-342
def test_match_all_returns_empty_when_no_tenant_data(B):  # :89-111
    pairs = []
    normal(B, pairs, OTHER_TENANT, "sensors/+/temp", 1, "receiverX", "delivererX", 1)
    ev = Events()
    matcher = B.pkg.GpuTenantRouteMatcher(TENANT_ID, make_index(B, pairs), ev)
    topics = {"sensors/device1/temp", "sensors/device1/humidity"}
    matched = matcher.match_all(topics, 10, 10)
    assert set(matched) == topics
    for routes in matched.values():
        assert routes.routes() == set()
        assert routes.persistent_fanout() == 0 and routes.group_fanout() == 0
    assert ev.events == []


def test_match_all_across_multiple_topics(B):  # :113-146
    pairs = []
    temp = normal(B, pairs, TENANT_ID, "sensors/+/temp", 1, "receiverA", "delivererA", 1)
    hum = normal(B, pairs, TENANT_ID, "sensors/+/humidity", 1, "receiverB", "delivererB", 2)
    ev = Events()
    matcher = B.pkg.GpuTenantRouteMatcher(TENANT_ID, make_index(B, pairs), ev)
    topics = {"sensors/device1/temp", "sensors/device1/humidity", "sensors/device2/temp"}
    matched = matcher.match_all(topics, 10, 10)
    assert set(matched) == topics
    assert temp in matched["sensors/device1/temp"].routes()
    assert temp in matched["sensors/device2/temp"].routes()
    assert hum in matched["sensors/device1/humidity"].routes()
    for t in topics:
        assert matched[t].persistent_fanout() == 1 and matched[t].group_fanout() == 0
    assert ev.events == []


def test_reuse_cached_filter_matches_for_repeated_subscriptions(B):  # :148-175
    pairs = []
    first = normal(B, pairs, TENANT_ID, "devices/+/status", 1, "receiverA", "delivererA", 1)
    second = normal(B, pairs, TENANT_ID, "devices/+/status", 2, "receiverB", "delivererB", 1)
    ev = Events()
    matcher = B.pkg.GpuTenantRouteMatcher(TENANT_ID, make_index(B, pairs), ev)
    matched = matcher.match_all({"devices/a/status", "devices/b/status"}, 5, 5)
    for routes in matched.values():
        assert routes.routes() == {first, second}
        assert routes.persistent_fanout() == 1  # only subBrokerId == 1 counts as persistent
        assert routes.group_fanout() == 0
    assert ev.events == []


def test_match_all_with_shared_subscription(B):  # :177-204
    pairs = []
    members = {B.schema.receiver_url(1, "receiverA", "delivererA"): 10, B.schema.receiver_url(2, "receiverB", "delivererB"): 11}
    g = group(B, pairs, TENANT_ID, "alerts/+/+/temperature", "groupAlpha", members)
    ev = Events()
    matcher = B.pkg.GpuTenantRouteMatcher(TENANT_ID, make_index(B, pairs), ev)
    matched = matcher.match_all({"alerts/site1/device1/temperature", "alerts/site1/device2/temperature"}, 10, 10)
    for routes in matched.values():
        assert g in routes.routes()
        assert routes.persistent_fanout() == 0 and routes.group_fanout() == 1
    assert ev.events == []


def test_skip_non_matching_routes(B):  # :206-236 (the probe/seek counters have no GPU analogue)
    pairs = []
    for i in range(21):
        normal(B, pairs, TENANT_ID, "invalid/%d" % i, 1, "noise%d" % i, "deliverer%d" % i, i)
    valid = normal(B, pairs, TENANT_ID, "metrics/+/cpu", 1, "receiverA", "delivererA", 1)
    ev = Events()
    matcher = B.pkg.GpuTenantRouteMatcher(TENANT_ID, make_index(B, pairs), ev)
    result = matcher.match_all({"metrics/server1/cpu"}, 10, 10)["metrics/server1/cpu"]
    assert result.routes() == {valid}
    assert result.persistent_fanout() == 1 and result.group_fanout() == 0
    assert ev.events == []


def test_isolate_routes_by_tenant(B):  # :238-268
    pairs = []
    mine = normal(B, pairs, TENANT_ID, "devices/+/signal", 1, "receiverA", "delivererA", 1)
    other = normal(B, pairs, OTHER_TENANT, "devices/+/signal", 1, "receiverB", "delivererB", 1)
    idx = make_index(B, pairs)
    ev = Events()
    r = B.pkg.GpuTenantRouteMatcher(TENANT_ID, idx, ev).match_all({"devices/a/signal"}, 10, 10)
    assert r["devices/a/signal"].routes() == {mine}
    r = B.pkg.GpuTenantRouteMatcher(OTHER_TENANT, idx, ev).match_all({"devices/a/signal"}, 10, 10)
    assert r["devices/a/signal"].routes() == {other}
    assert ev.events == []


def test_trigger_persistent_fanout_throttling(B):  # :270-301
    pairs = []
    normal(B, pairs, TENANT_ID, "alarms/+/critical", 1, "receiverA", "delivererA", 1)
    second = normal(B, pairs, TENANT_ID, "alarms/+/critical", 1, "receiverB", "delivererB", 2)
    ev = Events()
    matcher = B.pkg.GpuTenantRouteMatcher(TENANT_ID, make_index(B, pairs), ev)
    routes = matcher.match_all({"alarms/device1/critical"}, 1, 10)["alarms/device1/critical"]
    assert routes.persistent_fanout() == 1 and routes.group_fanout() == 0
    assert len(routes.routes()) == 1
    assert len(ev.events) == 1
    e = ev.events[0]
    assert isinstance(e, B.pkg.PersistentFanoutThrottled)
    assert (e.tenant_id, e.topic, e.mqtt_topic_filter, e.max_count) == (TENANT_ID, "alarms/device1/critical", second.mqtt_topic_filter, 1)


def test_trigger_group_fanout_throttling(B):  # :303-342
    pairs = []
    first = group(B, pairs, TENANT_ID, "jobs/+/progress", "groupA", {B.schema.receiver_url(1, "receiverA", "delivererA"): 1})
    group(B, pairs, TENANT_ID, "jobs/+/progress", "groupB", {B.schema.receiver_url(1, "receiverB", "delivererB"): 1})
    ev = Events()
    matcher = B.pkg.GpuTenantRouteMatcher(TENANT_ID, make_index(B, pairs), ev)
    routes = matcher.match_all({"jobs/job1/progress"}, 10, 1)["jobs/job1/progress"]
    assert routes.group_fanout() == 1
    assert sum(1 for m in routes.routes() if isinstance(m, B.schema.GroupMatching)) == 1
    assert len(ev.events) == 1
    e = ev.events[0]
    assert isinstance(e, B.pkg.GroupFanoutThrottled)
    # "second comes before first in lexicographical order by bucketing key"
    assert (e.tenant_id, e.topic, e.mqtt_topic_filter, e.max_count) == (TENANT_ID, "jobs/job1/progress", first.mqtt_topic_filter, 1)


# ------------------------------------------------------------------ expansion-set fixtures (Fixtures.java:31-105)
LOCAL_FIXTURES = {
    "a": ["#", "+", "+/#", "a", "a/#"],
    "$sys/a": ["$sys/#", "$sys/+", "$sys/+/#", "$sys/a", "$sys/a/#"],
    "/": ["/", "//#", "/#", "/+", "/+/#", "#", "+/", "+//#", "+/#", "+/+", "+/+/#"],
    "a/b": ["#", "+/#", "+/+", "+/+/#", "+/b", "+/b/#", "a/#", "a/+", "a/+/#", "a/b", "a/b/#"],
}


def test_fixture_expansion_sets(B):
    # one route per filter over a vocabulary that covers every fixture filter plus non-matching neighbours
    import itertools
    vocab = ["", "a", "b", "$sys", "+", "#"]
    filters = set()
    for n in range(1, 4):
        for combo in itertools.product(vocab, repeat=n):
            if "#" in combo[:-1]:
                continue
            filters.add("/".join(combo))
    filters.discard("")
    filters = sorted(filters)
    pairs = []
    for i, f in enumerate(filters):
        normal(B, pairs, "t", f, 0, "r%d" % i, "d", 1)
    idx = make_index(B, pairs)
    matcher = B.pkg.GpuTenantRouteMatcher("t", idx)
    res = matcher.match_all(list(LOCAL_FIXTURES), 100, 100)
    for topic, want in LOCAL_FIXTURES.items():
        got = sorted(m.mqtt_topic_filter for m in res[topic].routes())
        assert got == sorted(want), topic


# ------------------------------------------------------------------ randomized cross-checks vs the oracle
def oracle_kv_from_pairs(pairs):
    kv = O.KV()
    for k, v in pairs:
        kv.put(k, v)
    kv.freeze()
    return kv


def compare_with_oracle(B, idx, kv, tenants, topics, tt, max_p, max_g, mode=O.MODE_TRIE):
    nt = len(tenants)
    res = idx.match_topics(tenants, topics, tt, [max_p] * nt, [max_g] * nt)
    offsets, ranks = res.expand()
    want = kv.match_batch(tenants, topics, tt, max_p, max_g, mode)
    assert offsets.tolist() == want.offsets.tolist()
    assert ranks.tolist() == want.ranks.tolist()
    got_events = sorted((int(k), int(t), int(r)) for t, r, k in res.throttled.tolist())
    assert got_events == sorted((k, t, r) for k, t, r, _ in want.events)
    # route_count is the pre-cap match count
    uncapped = kv.match_batch(tenants, topics, tt, 2 ** 31 - 1, 2 ** 31 - 1, mode)
    assert res.route_count.tolist() == np.diff(uncapped.offsets).tolist()
    res.close()
    return want


def random_pairs(B, rng, n_filters, vocab, depth, empties=True):
    def level(i, allow_empty):
        r = rng.random()
        if r < 0.08 and allow_empty:
            return ""
        if r < 0.15 and i == 0:
            return "$" + rng.choice(vocab)
        return rng.choice(vocab)

    def filt():
        n = rng.randint(1, depth)
        lv = []
        for i in range(n):
            r = rng.random()
            if r < 0.25:
                lv.append("+")
            elif r < 0.35 and i == n - 1:
                lv.append("#")
            else:
                lv.append(level(i, empties))
        return "/".join(lv)
    tenants = ["tA", "tB", "t"]
    pairs = {}
    for _ in range(n_filters):
        tenant = rng.choice(tenants)
        f = filt()
        if rng.random() < 0.15:
            members = {B.schema.receiver_url(rng.choice([0, 1]), "m%d" % rng.randint(0, 5), "d"): rng.randint(1, 9)
                       for _ in range(rng.randint(1, 3))}
            full = rng.choice(["$share/", "$oshare/"]) + "g%d" % rng.randint(0, 3) + "/" + f
            pairs[B.schema.route_key(tenant, full)] = B.schema.route_group_bytes(members)
        else:
            for _ in range(rng.choice([1, 1, 1, 2, 5])):
                url = B.schema.receiver_url(rng.choice([0, 1, 1, 2]), "r%d" % rng.randint(0, 400), "d%d" % rng.randint(0, 3))
                pairs[B.schema.route_key(tenant, f, url)] = B.schema.incarnation_bytes(rng.randint(0, 99))
    topics = ["/".join(level(i, empties) for i in range(rng.randint(1, depth))) for _ in range(300)]
    tt = np.array([rng.randrange(len(tenants)) for _ in topics], dtype=np.int32)
    return sorted(pairs.items()), tenants, topics, tt


@pytest.mark.parametrize("seed,caps", [(1, (2 ** 31 - 1, 100)), (2, (2, 1)), (3, (0, 0)), (4, (4, 4)), (5, (1, 2 ** 31 - 1))])
def test_random_small_vocab_vs_oracle(B, seed, caps):
    rng = random.Random(seed)
    pairs, tenants, topics, tt = random_pairs(B, rng, 600, ["a", "b", "c", "dd", "e1"], 5)
    idx = make_index(B, pairs)
    kv = oracle_kv_from_pairs(pairs)
    want = compare_with_oracle(B, idx, kv, tenants, topics, tt, caps[0], caps[1], O.MODE_TRIE)
    # and against the brute-force predicate (independent of any trie)
    brute = kv.match_batch(tenants, topics, tt, caps[0], caps[1], O.MODE_BRUTE)
    assert brute.route_sets() == want.route_sets()
    assert sum(len(r) for r in want.route_sets()) > 0
    assert idx.stats()["multi_segment_filters"] >= 0


@pytest.mark.parametrize("seed,caps", [(11, (2 ** 31 - 1, 100)), (12, (2, 1)), (13, (0, 0))])
def test_locality_order_path_vs_oracle(B, seed, caps, monkeypatch):
    # BFQ_ORDER=1: every batch (not only those of >= 32768 topics) is matched in locality order — the order keys, the radix
    # sort and the order-indirected chunk claims of tier 0 against the oracle, incl. empty levels, '$' topics, unknown tenants
    monkeypatch.setenv("BFQ_ORDER", "1")
    rng = random.Random(seed)
    pairs, tenants, topics, tt = random_pairs(B, rng, 600, ["a", "b", "c", "dd", "e1"], 5)
    topics += ["", "/", "//", "$sys", "x" * 70 + "/a", "/".join("l%d" % i for i in range(14))]
    tt = np.concatenate([tt, np.array([0, 1, 2, 0, 1, 2], np.int32)])
    idx = make_index(B, pairs)
    kv = oracle_kv_from_pairs(pairs)
    want = compare_with_oracle(B, idx, kv, tenants, topics, tt, caps[0], caps[1], O.MODE_TRIE)
    assert sum(len(r) for r in want.route_sets()) > 0
    # a tenant the index has never seen, mixed into the same batch
    compare_with_oracle(B, idx, kv, tenants + ["nobody"], topics, np.where(np.arange(len(topics)) % 7 == 0, 3, tt).astype(np.int32),
                        caps[0], caps[1], O.MODE_TRIE)
    idx.close()


def test_reference_literal_algorithm_agrees_without_empty_levels(B):
    rng = random.Random(9)
    pairs, tenants, topics, tt = random_pairs(B, rng, 500, ["a", "b", "c", "dd"], 4, empties=False)
    idx = make_index(B, pairs)
    kv = oracle_kv_from_pairs(pairs)
    want = compare_with_oracle(B, idx, kv, tenants, topics, tt, 3, 2, O.MODE_REFERENCE)
    assert want.stats["backward_seeks"] == 0


def test_multi_segment_filter_interleaving(B):
    # keys of "a//b" sort inside the bucket range of "a": the routes of "a" are not one contiguous rank run
    pairs = []
    for i in range(300):
        normal(B, pairs, "t", "a", i % 2, "r%d" % i, "d", 1)
    for i in range(60):
        normal(B, pairs, "t", "a//b", 1, "x%d" % i, "d", 1)
    for i in range(20):
        normal(B, pairs, "t", "a/", 0, "y%d" % i, "d", 1)
    pairs.sort()
    idx = make_index(B, pairs)
    assert idx.stats()["multi_segment_filters"] >= 1
    kv = oracle_kv_from_pairs(pairs)
    topics = ["a", "a//b", "a/", "b"]
    for caps in [(2 ** 31 - 1, 2 ** 31 - 1), (7, 3), (0, 0)]:
        compare_with_oracle(B, idx, kv, ["t"], topics, np.zeros(4, np.int32), caps[0], caps[1], O.MODE_BRUTE)


def test_wide_fanouts_perfect_hash_and_tag_table(B):
    # child arrays of every kind under one tenant: 1 child (fingerprint), 2..16, 17..1000 (perfect hash, arrays up to 2^16
    # slots) and > ~1000 children (global tag table, two accesses) — at the first level and below a '+' / exact parent
    pairs = []
    widths = {"w1": 1, "w3": 3, "w16": 16, "w17": 17, "w40": 40, "w300": 300, "w1500": 1500}
    for name, n in widths.items():
        for i in range(n):
            normal(B, pairs, "t", "%s/c%04d" % (name, i), i % 2, "r%s%d" % (name, i), "d", 1)
        normal(B, pairs, "t", "%s/+" % name, 0, "plus" + name, "d", 1)
        normal(B, pairs, "t", "%s/#" % name, 1, "hash" + name, "d", 1)
    for i in range(2500):   # wide at the tenant root too
        normal(B, pairs, "t2", "dev%05d/state" % i, 0, "s%d" % i, "d", 1)
    normal(B, pairs, "t2", "+/state", 1, "all", "d", 1)
    pairs.sort()
    idx = make_index(B, pairs)
    kv = oracle_kv_from_pairs(pairs)
    topics, tt = [], []
    for name, n in widths.items():
        for i in sorted({0, 1, n // 2, n - 1, n, n + 7}):
            topics.append("%s/c%04d" % (name, i))
            tt.append(0)
        topics += [name, name + "/c0000/x"]
        tt += [0, 0]
    for i in (0, 1, 1234, 2499, 2500, 99999):
        topics += ["dev%05d/state" % i, "dev%05d" % i]
        tt += [1, 1]
    for caps in [(2 ** 31 - 1, 2 ** 31 - 1), (1, 1)]:
        want = compare_with_oracle(B, idx, kv, ["t", "t2"], topics, np.array(tt, np.int32), caps[0], caps[1], O.MODE_BRUTE)
    assert sum(len(r) for r in want.route_sets()) > 0


def test_long_levels_and_long_topics(B):
    # levels longer than the 24 inline token bytes (continuation chunks) and topics longer than the 256 B stage
    l25, l24, l48, l49, l100 = "x" * 25, "y" * 24, "z" * 48, "w" * 49, "v" * 100
    filters = [l25, l24, l48, l49, l100, l25 + "/" + l48, "+/" + l48, l100 + "/#", l49 + "/+", "x" * 24 + "/b", "x" * 26,
               "/".join(["seg%02d" % i for i in range(40)]), "/".join(["seg%02d" % i for i in range(39)]) + "/#",
               "你好" * 9, "你好" * 8]
    pairs = []
    for i, f in enumerate(filters):
        normal(B, pairs, "t", f, 0, "r%d" % i, "d", 1)
    pairs.sort()
    idx = make_index(B, pairs)
    assert idx.stats()["long_token_chunks"] > 0
    kv = oracle_kv_from_pairs(pairs)
    topics = [l25, l24, l48, l49, l100, l25 + "/" + l48, l24 + "/" + l48, l100 + "/a/b", l49 + "/q", "x" * 26, "x" * 27,
              "/".join(["seg%02d" % i for i in range(40)]), "/".join(["seg%02d" % i for i in range(41)]), "你好" * 9, "你好" * 8,
              "x" * 24, "x" * 23]
    want = compare_with_oracle(B, idx, kv, ["t"], topics, np.zeros(len(topics), np.int32), 10, 10, O.MODE_BRUTE)
    n_matched = [len(r) for r in want.route_sets()]
    assert n_matched == [1, 1, 1, 1, 2, 2, 1, 1, 1, 1, 0, 2, 1, 1, 1, 0, 0]


def test_frontier_overflow_goes_through_tier2(B):
    # every {a,+}^8 filter matches a/a/a/a/a/a/a/a: 256 ranges and a frontier of up to 128 nodes -> tier 2
    import itertools
    pairs = []
    i = 0
    for combo in itertools.product(["a", "+"], repeat=8):
        normal(B, pairs, "t", "/".join(combo), i % 2, "r%d" % i, "d", 1)
        i += 1
    for n in range(1, 8):
        normal(B, pairs, "t", "/".join(["a"] * n) + "/#", 1, "h%d" % n, "d", 1)
    pairs.sort()
    idx = make_index(B, pairs)
    kv = oracle_kv_from_pairs(pairs)
    topics = ["/".join(["a"] * 8), "/".join(["a"] * 7 + ["b"]), "/".join(["b"] * 8), "a/a"]
    before = idx.stats()["overflow_topics"]
    want = compare_with_oracle(B, idx, kv, ["t"], topics, np.zeros(len(topics), np.int32), 2 ** 31 - 1, 2 ** 31 - 1, O.MODE_BRUTE)
    assert len(want.route_sets()[0]) == 256 + 7
    assert idx.stats()["overflow_topics"] > before
    compare_with_oracle(B, idx, kv, ["t"], topics, np.zeros(len(topics), np.int32), 5, 0, O.MODE_BRUTE)


def test_apply_and_commit_incremental(B):
    pairs = []
    a = normal(B, pairs, "t", "a/+", 1, "r1", "d", 1)
    b = normal(B, pairs, "t", "a/b", 0, "r2", "d", 2)
    idx = make_index(B, pairs)
    m = B.pkg.GpuTenantRouteMatcher("t", idx)
    assert m.match_all(["a/b"], 10, 10)["a/b"].routes() == {a, b}
    extra = []
    c = normal(B, extra, "t", "a/#", 1, "r3", "d", 3)
    idx.apply(adds=extra, dels=[pairs[0][0]])
    # not visible until commit (matches see whole snapshots)
    assert m.match_all(["a/b"], 10, 10)["a/b"].routes() == {a, b}
    idx.commit()
    assert m.match_all(["a/b"], 10, 10)["a/b"].routes() == {b, c}
    idx.reset()
    idx.commit()
    assert m.match_all(["a/b"], 10, 10)["a/b"].routes() == set()


@pytest.mark.parametrize("seed,caps", [(21, (2 ** 31 - 1, 100)), (22, (2, 1)), (23, (4, 4))])
def test_delta_commits_random_sub_unsub_stream(B, seed, caps):
    """DistWorkerCoProc.batchAddRoute / batchRemoveRoute (DW/DistWorkerCoProc.java:304-513) apply SUBs and UNSUBs one raft entry
    at a time: a random stream of small apply + commit rounds — adds, overwrites, deletes, tenants appearing and vanishing,
    filters with empty levels (multi-segment rank runs), group routes — goes through the DELTA path of bfq_index_commit (only
    the touched tenants are rebuilt, the rest of the snapshot is copied and rank-shifted on the device). After every commit
    the whole answer (ranks, caps events, route lookups) equals the oracle fed the same mutations, and results taken before a
    commit keep resolving against their own snapshot."""
    rng = random.Random(seed)
    pairs, tenants, topics, tt = random_pairs(B, rng, 500, ["a", "b", "c", "dd", "e1"], 5)
    pool, _, _, _ = random_pairs(B, random.Random(seed + 100), 700, ["a", "b", "c", "dd", "e1", "zz"], 5)
    tenants = tenants + ["tNew1", "tNew2"]
    extra_t = []
    for i, (k, v) in enumerate(pool[:60]):   # routes of two tenants that do not exist at first
        extra_t.append((k.replace(b"\x00\x02tA", b"\x00\x05tNew1", 1) if i % 2 else k.replace(b"\x00\x02tB", b"\x00\x05tNew2", 1), v))
    extra_t = [kv for kv in extra_t if b"tNew" in kv[0]]
    idx = make_index(B, pairs)
    live = dict(pairs)
    kv = oracle_kv_from_pairs(pairs)
    before = idx.stats()
    held = []
    n_rounds = 25
    for rnd in range(n_rounds):
        adds, dels = [], []
        for _ in range(rng.randint(1, 4)):
            r = rng.random()
            if r < 0.45:
                k, v = rng.choice(pool)
                adds.append((k, v))
            elif r < 0.55 and extra_t:
                adds.append(rng.choice(extra_t))
            elif r < 0.65:   # overwrite the value of a live route
                k = rng.choice(sorted(live))
                adds.append((k, B.schema.incarnation_bytes(rng.randint(100, 200)) if len(live[k]) == 8 else live[k]))
            elif live:
                dels.append(rng.choice(sorted(live)))
        if rnd == 12:   # a tenant disappears completely, later routes may bring it back
            dels += [k for k in live if b"\x00\x02tB" in k[:6]]
        dels = [k for k in dels if k not in dict(adds)]
        idx.apply(adds=adds, dels=dels)
        for k, v in adds:
            live[k] = v
            kv.put(k, v)
        for k in dels:
            if k in live:
                del live[k]
                kv.erase(k)
        # a result taken before the commit ...
        res_old = idx.match_topics(tenants, topics[:40], tt[:40])
        o_old, r_old = res_old.expand()
        keys_old = [res_old.route(int(x))[0] for x in r_old[:50]]
        idx.commit()
        kv.freeze()
        want = compare_with_oracle(B, idx, kv, tenants, topics, tt, caps[0], caps[1], O.MODE_TRIE)
        # ... still resolves its ranks against its own snapshot
        assert [res_old.route(int(x))[0] for x in r_old[:50]] == keys_old
        held.append(res_old)
        if len(held) > 3:
            held.pop(0).close()
        # rank -> key through the handle follows the new snapshot: spot-check against the sorted live keys
        order = sorted(live)
        for x in want.ranks[:20].tolist():
            assert idx.route(int(x))[0] == order[int(x)]
        assert idx.stats()["routes"] == len(live)
    after = idx.stats()
    assert after["delta_commits"] - before["delta_commits"] >= n_rounds - 2
    for r in held:
        r.close()
    # a full rebuild of the same state gives the same answers
    idx2 = make_index(B, sorted(live.items()))
    compare_with_oracle(B, idx2, kv, tenants, topics, tt, caps[0], caps[1], O.MODE_TRIE)


def test_load_rejects_unsorted_and_undecodable(B):
    idx = B.pkg.GpuRouteIndex(0)
    k1 = B.schema.route_key("t", "b", B.schema.receiver_url(0, "r", "d"))
    k2 = B.schema.route_key("t", "a", B.schema.receiver_url(0, "r", "d"))
    from bifromq_b200 import _native as N
    kb, ko = N.as_blob([k1, k2])
    vb, vo = N.as_blob([b"\0" * 8, b"\0" * 8])
    with pytest.raises(B.pkg.NativeError):
        idx.load(kb, ko, vb, vo)
    idx.reset()
    kb, ko = N.as_blob([b"\x00\x00\x01tgarbage"])
    vb, vo = N.as_blob([b"\0" * 8])
    idx.load(kb, ko, vb, vo)
    with pytest.raises(B.pkg.NativeError):
        idx.commit()


def test_empty_batch_and_unknown_tenant(B):
    pairs = []
    normal(B, pairs, "t", "#", 0, "r", "d", 1)
    idx = make_index(B, pairs)
    res = idx.match_topics(["t"], [])
    assert res.n_topics == 0
    res.close()
    res = idx.match_topics(["nobody", "t"], ["a", "a", "$sys", ""], np.array([0, 1, 1, 1], np.int32))
    off, ranks = res.expand()
    assert off.tolist() == [0, 0, 1, 1, 2]
    res.close()


# ------------------------------------------------------------------ BASELINE.json configs (scaled) vs the oracle
@pytest.mark.parametrize("config,scale", [("C1", 1.0), ("C2", 0.02), ("C3", 0.005), ("C4", 0.005)])
@pytest.mark.parametrize("caps", [(2 ** 31 - 1, 100), (4, 4)])
def test_baseline_configs_vs_oracle(B, config, scale, caps):
    w = B.workload.Workload(config, scale=scale)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    kv = O.KV()
    kv.load(w.keys, w.key_off, w.vals, w.val_off)
    kv.freeze()
    assert len(kv) == w.n_routes
    tenants = w.tenants
    nt = len(tenants)
    res = idx.match(tenants, w.topics, w.topic_off, w.topic_tenant, [caps[0]] * nt, [caps[1]] * nt)
    offsets, ranks = res.expand()
    tb, toff = O.blob(tenants)
    want = kv.match_blobs(tb, toff, w.topics, w.topic_off, np.ascontiguousarray(w.topic_tenant), w.n_topics, caps[0], caps[1],
                          O.MODE_TRIE, False, 8)
    assert offsets.tolist() == want.offsets.tolist()
    assert ranks.tolist() == want.ranks.tolist()
    got_events = sorted((int(k), int(t), int(r)) for t, r, k in res.throttled.tolist())
    assert got_events == sorted((k, t, r) for k, t, r, _ in want.events)
    hit = float((np.diff(offsets) > 0).mean())
    assert hit > 0.5, "workload should mostly hit (got %.2f)" % hit
    res.close()


@pytest.mark.parametrize("caps", [(2 ** 31 - 1, 2 ** 31 - 1), (3, 1)])
def test_pipelined_host_path_large_batch(B, caps):
    """batches >= 128k topics go through the 4-sub-batch, 3-stream pipeline of bfq_match: same results, same throttle
    events (absolute topic indices), dense ranges in topic order"""
    w = B.workload.Workload("C3", scale=0.02)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    reps = (140000 + w.n_topics - 1) // w.n_topics
    topics = w.topic_list() * reps
    tt = np.tile(np.asarray(w.topic_tenant[:w.n_topics]), reps).astype(np.int32)
    # shuffle so that sub-batches are not periodic copies of each other
    perm = np.random.RandomState(5).permutation(len(topics))
    topics = [topics[i] for i in perm]
    tt = np.ascontiguousarray(tt[perm])
    assert len(topics) >= 131072
    tenants = w.tenants
    nt = len(tenants)
    res = idx.match_topics(tenants, topics, tt, [caps[0]] * nt, [caps[1]] * nt)
    assert int(res.timings_ms["sub_batches"]) == 4   # number of pipelined sub-batches
    offsets, ranks = res.expand()
    # dense ranges: the distinct spans tile the range array (repeats of a (tenant, topic) pair share their first occurrence's)
    sb, sc = res.span_begin.astype(np.int64), res.span_count.astype(np.int64)
    spans = np.unique(np.stack([sb[sc > 0], sc[sc > 0]], axis=1), axis=0)
    assert spans[0, 0] == 0 and (spans[1:, 0] == spans[:-1, 0] + spans[:-1, 1]).all() and spans[-1, 0] + spans[-1, 1] == len(res.ranges)
    assert len(spans) < int((sc > 0).sum())   # the batch is a workload repeated: most topics are repeats
    kv = O.KV()
    kv.load(w.keys, w.key_off, w.vals, w.val_off)
    want = kv.match_batch(tenants, topics, tt, caps[0], caps[1], O.MODE_TRIE, False, 8)
    assert offsets.tolist() == want.offsets.tolist()
    assert ranks.tolist() == want.ranks.tolist()
    got_events = sorted((int(k), int(t), int(r)) for t, r, k in res.throttled.tolist())
    assert got_events == sorted((k, t, r) for k, t, r, _ in want.events)
    if caps[0] == 3:
        assert len(got_events) > 100 and max(e[1] for e in got_events) > 100000
    res.close()


@pytest.mark.parametrize("caps", [(2 ** 31 - 1, 2 ** 31 - 1), (3, 1)])
def test_device_resident_match_and_expand(B, caps):
    """bfq_match_device (batch already in HBM, result left there) + bfq_expand_device == the host path"""
    import torch
    w = B.workload.Workload("C3", scale=0.01)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    tenants = w.tenants
    nt, n = len(tenants), w.n_topics
    dev = torch.device("cuda", 0)
    d_topics = torch.from_numpy(np.ascontiguousarray(w.topics)).to(dev)
    d_off = torch.from_numpy(np.ascontiguousarray(w.topic_off)).to(dev)
    d_tt = torch.from_numpy(np.ascontiguousarray(w.topic_tenant[:n])).to(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = idx.match_device(tenants, d_topics.data_ptr(), d_off.data_ptr(), d_tt.data_ptr(), n, [caps[0]] * nt, [caps[1]] * nt, stream)
    d_offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    total = out.expand(d_offsets.data_ptr(), None, 0, stream)
    d_ranks = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)
    assert out.expand(d_offsets.data_ptr(), d_ranks.data_ptr(), total, stream) == total
    torch.cuda.synchronize()
    res = idx.match(tenants, w.topics, w.topic_off, w.topic_tenant[:n], [caps[0]] * nt, [caps[1]] * nt)
    offsets, ranks = res.expand()
    assert d_offsets.cpu().numpy().tolist() == offsets.tolist()
    got = d_ranks.cpu().numpy()[:total]
    for i in range(n):   # unordered within a topic on the device
        assert sorted(got[offsets[i]:offsets[i + 1]].tolist()) == ranks[offsets[i]:offsets[i + 1]].tolist()
    assert out.n_throttled == len(res.throttled)
    assert out.generation == res.generation == idx.generation()
    res.close()
    out.release()


def test_async_device_matches_in_flight_together(B):
    """bfq_match_device_async: three batches enqueued back to back on one stream without a host sync, each on its own leased
    workspace; waited afterwards, each equals the host path on its batch"""
    import torch
    w = B.workload.Workload("C3", scale=0.05)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    tenants = w.tenants
    nt = len(tenants)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    cuts = [(0, w.n_topics), (0, w.n_topics // 2), (w.n_topics // 3, w.n_topics)]
    keep, outs = [], []
    for b, e in cuts:
        off = np.ascontiguousarray(w.topic_off[b:e + 1])
        d_topics = torch.from_numpy(np.ascontiguousarray(w.topics)).to(dev)
        d_off = torch.from_numpy(off).to(dev)
        d_tt = torch.from_numpy(np.ascontiguousarray(w.topic_tenant[b:e])).to(dev)
        keep.append((d_topics, d_off, d_tt))
        outs.append(idx.match_device(tenants, d_topics.data_ptr(), d_off.data_ptr(), d_tt.data_ptr(), e - b, [3] * nt, [1] * nt,
                                     stream, wait=False))
    for (b, e), out in zip(cuts, outs):
        out.wait()
        n = e - b
        d_offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        total = out.expand(d_offsets.data_ptr(), None, 0, stream)
        d_ranks = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)
        out.expand(d_offsets.data_ptr(), d_ranks.data_ptr(), total, stream)
        torch.cuda.synchronize()
        res = idx.match(tenants, w.topics, np.ascontiguousarray(w.topic_off[b:e + 1]), w.topic_tenant[b:e], [3] * nt, [1] * nt)
        offsets, ranks = res.expand()
        assert d_offsets.cpu().numpy().tolist() == offsets.tolist()
        got = d_ranks.cpu().numpy()[:total]
        for i in range(n):
            assert sorted(got[offsets[i]:offsets[i + 1]].tolist()) == ranks[offsets[i]:offsets[i + 1]].tolist()
        assert out.n_throttled == len(res.throttled) and out.n_distinct_topics <= n
        res.close()
    for out in outs:
        out.release()


@pytest.mark.parametrize("caps", [(2 ** 31 - 1, 100), (3, 1)])
def test_fanout_groups_routes_by_deliverer(B, caps):
    """bfq_fanout_device: the surviving (topic, route) pairs of a batch grouped by (subBrokerId, delivererKey) — what
    DeliverExecutorGroup.submit + DeliverExecutor.send do one route at a time (DW/DeliverExecutorGroup.java:112-231,
    DW/DeliverExecutor.java:89-93). Every pair of the device CSR appears exactly once, under the deliverer its route is
    delivered through (restated in oracle_lib.deliverer_of_receiver_url); a $share route is resolved to one member of its
    stored group (any member is a valid outcome of the reference's random pick), an $oshare route is left to the host under
    the reserved id."""
    import torch
    w = B.workload.Workload("C3", scale=0.02)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    tenants, n = w.tenants, w.n_topics
    nt = len(tenants)
    dev = torch.device("cuda", 0)
    d_topics = torch.from_numpy(np.ascontiguousarray(w.topics)).to(dev)
    d_off = torch.from_numpy(np.ascontiguousarray(w.topic_off)).to(dev)
    d_tt = torch.from_numpy(np.ascontiguousarray(w.topic_tenant[:n])).to(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = idx.match_device(tenants, d_topics.data_ptr(), d_off.data_ptr(), d_tt.data_ptr(), n, [caps[0]] * nt, [caps[1]] * nt, stream)
    d_offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    total = out.expand(d_offsets.data_ptr(), None, 0, stream)
    d_ranks = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)
    out.expand(d_offsets.data_ptr(), d_ranks.data_ptr(), total, stream)
    fo = out.fanout(d_offsets.data_ptr(), d_ranks.data_ptr(), total, stream)
    torch.cuda.synchronize()
    from bifromq_b200 import dist as D
    D_ = fo.n_deliverers
    pack_off = D.device_view(fo.d_pack_offsets, D_ + 1, "<i8", dev).cpu().numpy()
    pt = D.device_view(fo.d_pack_topic, max(total, 1), "<u4", dev).cpu().numpy()[:total]
    pr = D.device_view(fo.d_pack_rank, max(total, 1), "<u4", dev).cpu().numpy()[:total]
    pm = D.device_view(fo.d_pack_member, max(total, 1), "<u4", dev).cpu().numpy()[:total]
    assert fo.n_pairs == total and pack_off[0] == 0 and pack_off[-1] == total and (np.diff(pack_off) >= 0).all()
    assert fo.ordered_share_id == D_ - 1
    # every pair of the CSR exactly once
    offsets = d_offsets.cpu().numpy()
    ranks = d_ranks.cpu().numpy()[:total]
    want_pairs = np.stack([np.repeat(np.arange(n), np.diff(offsets)), ranks], axis=1)
    got_pairs = np.stack([pt.astype(np.int64), pr.astype(np.int64)], axis=1)
    assert sorted(map(tuple, want_pairs.tolist())) == sorted(map(tuple, got_pairs.tolist()))
    # every pair under the right deliverer
    deliverers = [idx.deliverer(d) for d in range(D_ - 1)]
    assert len(set(deliverers)) == len(deliverers)
    route_cache = {}
    n_group = n_ordered = 0
    for d in range(D_):
        for j in range(int(pack_off[d]), int(pack_off[d + 1])):
            r = int(pr[j])
            if r not in route_cache:
                k, v = idx.route(r)
                route_cache[r] = (O.build_match_route(k, v), v)
            m, v = route_cache[r]
            if m["type"] == "Normal":
                assert pm[j] == 0xFFFFFFFF and deliverers[d] == O.deliverer_of_receiver_url(m["receiverUrl"])
            elif m["mqttTopicFilter"].startswith("$oshare/"):
                assert d == fo.ordered_share_id
                n_ordered += 1
            else:
                members = O.route_group_members_in_wire_order(v)
                assert pm[j] < len(members) and deliverers[d] == O.deliverer_of_receiver_url(members[pm[j]])
                n_group += 1
    assert n_group > 0 and n_ordered > 0 and total > n
    out.release()


def test_exchange_gather_single_rank(B):
    """bfq_exchange_gather with a world of one (NCCL communicator of size 1): the reassembled arrays are the dense,
    topic-ordered form of the device result — equal to what the host path returns for the same batch"""
    import torch
    from bifromq_b200 import dist as D
    w = B.workload.Workload("C3", scale=0.05)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    tenants, n = w.tenants, w.n_topics
    dev = torch.device("cuda", 0)
    d_topics = torch.from_numpy(np.ascontiguousarray(w.topics)).to(dev)
    d_off = torch.from_numpy(np.ascontiguousarray(w.topic_off)).to(dev)
    d_tt = torch.from_numpy(np.ascontiguousarray(w.topic_tenant[:n])).to(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = idx.match_device(tenants, d_topics.data_ptr(), d_off.data_ptr(), d_tt.data_ptr(), n, stream=stream)
    x = D.Exchange(0, rank=0, world=1)
    for ranges in (True, False, True):
        g = x.gather(out, ranges=ranges, stream=stream)
        torch.cuda.synchronize()
        res = idx.match(tenants, w.topics, w.topic_off, w.topic_tenant[:n])
        assert g.world == 1 and g.topic_base[0] == 0 and g.topic_count == [n] and g.n_topics_total == n
        assert g.route_count().cpu().numpy().tolist() == res.route_count.tolist()
        if ranges:
            sc = g.span_count().cpu().numpy().astype(np.int64)
            assert sc.tolist() == res.span_count.tolist()
            assert g.range_base[0] == 0 and g.range_count == [int(sc.sum())]   # every topic's ranges in full (no shared spans)
            got = g.ranges().cpu().numpy()
            gb = np.concatenate([[0], np.cumsum(sc)])
            want = np.stack([res.ranges["first"], res.ranges["count"]], axis=1)
            sb = res.span_begin
            for i in range(0, n, 97):   # a topic's ranges may come out in a different order
                a, b = int(sb[i]), int(sb[i]) + int(res.span_count[i])
                assert sorted(map(tuple, got[gb[i]:gb[i + 1]].tolist())) == sorted(map(tuple, want[a:b].tolist()))
        res.close()
    x.close()
    out.release()


def test_concurrent_matches_on_one_handle(B):
    """ITenantRouteMatcher.matchAll is called from the shared topic-matcher pool (DW/DistWorkerCoProcFactory.java:74-85): six
    threads match DIFFERENT batches on ONE handle at the same time, hold their results while the others run, and every
    result equals the oracle's answer for its own batch (results own their buffers)"""
    import threading
    w = B.workload.Workload("C3", scale=0.05)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    kv = O.KV()
    kv.load(w.keys, w.key_off, w.vals, w.val_off)
    tenants = w.tenants
    nt = len(tenants)
    all_topics = w.topic_list()
    T = 6
    per = len(all_topics) // T
    want, got, errs = {}, {}, []
    for k in range(T):
        b = k * per
        tt = np.ascontiguousarray(w.topic_tenant[b:b + per])
        want[k] = kv.match_batch(tenants, all_topics[b:b + per], tt, 3, 1, O.MODE_TRIE, False, 8)
    barrier = threading.Barrier(T)

    def worker(k):
        try:
            b = k * per
            tt = np.ascontiguousarray(w.topic_tenant[b:b + per])
            for rep in range(3):
                barrier.wait()
                res = idx.match_topics(tenants, all_topics[b:b + per], tt, [3] * nt, [1] * nt)
                barrier.wait()   # every thread now holds a result while the others' are alive too
                offsets, ranks = res.expand()
                ev = sorted((int(kk), int(t), int(r)) for t, r, kk in res.throttled.tolist())
                got[(k, rep)] = (offsets.tolist(), ranks.tolist(), ev)
                res.close()
        except Exception as e:   # pragma: no cover
            errs.append(e)
            barrier.abort()
    th = [threading.Thread(target=worker, args=(k,)) for k in range(T)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for (k, rep), (offsets, ranks, ev) in got.items():
        assert offsets == want[k].offsets.tolist() and ranks == want[k].ranks.tolist()
        assert ev == sorted((kk, t, r) for kk, t, r, _ in want[k].events)
    assert len(got) == 3 * T


def test_results_resolve_against_their_own_snapshot(B):
    """ranks are positions in KV order, so every add/remove shifts them: a result taken before a commit must keep resolving
    (expand, route lookup, kinds) against the snapshot it was produced from, while lookups through the handle follow the
    latest commit. The reference reads keys and values from one consistent KV reader (TenantRouteMatcher.java:81-98)."""
    import threading
    pairs = []
    for i in range(50):
        normal(B, pairs, "t", "a/%02d" % i, 0, "r%02d" % i, "d", 1)
    normal(B, pairs, "t", "a/+", 1, "wild", "d", 1)
    pairs.sort()
    idx = make_index(B, pairs)
    res_old = idx.match_topics(["t"], ["a/07", "a/33"])
    off_old, ranks_old = res_old.expand()
    keys_old = [res_old.route(int(r))[0] for r in ranks_old]
    gen_old = res_old.generation
    # a route that sorts FIRST in the tenant shifts every rank by one; concurrent lookups on the old result run meanwhile
    stop, errs = threading.Event(), []

    def reader():
        try:
            while not stop.is_set():
                o, r = res_old.expand()
                assert r.tolist() == ranks_old.tolist()
                assert [res_old.route(int(x))[0] for x in r] == keys_old
        except Exception as e:   # pragma: no cover
            errs.append(e)
    th = threading.Thread(target=reader)
    th.start()
    for rnd in range(5):
        adds = []
        normal(B, adds, "t", "!first%d" % rnd, 0, "x", "d", 1)
        idx.apply(adds=adds)
        idx.commit()
    stop.set()
    th.join()
    assert not errs, errs
    assert idx.generation() == gen_old + 5 and res_old.generation == gen_old
    res_new = idx.match_topics(["t"], ["a/07", "a/33"])
    off_new, ranks_new = res_new.expand()
    assert ranks_new.tolist() == [r + 5 for r in ranks_old.tolist()]
    assert [res_new.route(int(r))[0] for r in ranks_new] == keys_old          # same routes, new ranks
    assert [res_old.route(int(r))[0] for r in ranks_old] == keys_old          # the old result still resolves its own ranks
    assert [idx.route(int(r))[0] for r in ranks_new] == keys_old              # the handle follows the latest commit
    assert res_old.route_kinds(ranks_old).tolist() == res_new.route_kinds(ranks_new).tolist()
    res_old.close()
    res_new.close()


def test_duplicate_topics_share_one_walk(B):
    """a batch with many repeats of the same (tenant, topic) pairs (above the ordering threshold, so the dedup pass runs):
    every occurrence gets the answer of its own tenant, including cap events per occurrence"""
    w = B.workload.Workload("C3", scale=0.02)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    kv = O.KV()
    kv.load(w.keys, w.key_off, w.vals, w.val_off)
    base = w.topic_list()[:3000]
    btt = w.topic_tenant[:3000]
    rng = np.random.default_rng(5)
    pick = rng.integers(0, 3000, 40000)
    topics = [base[i] for i in pick]
    tt = np.ascontiguousarray(btt[pick]).astype(np.int32)
    # the same topic string under a DIFFERENT tenant must not be merged with it
    topics[:100] = [base[0]] * 100
    tt[:100] = np.arange(100) % len(w.tenants)
    before = idx.stats()["duplicate_topics"]
    compare_with_oracle(B, idx, kv, w.tenants, topics, tt, 3, 1, O.MODE_TRIE)
    assert idx.stats()["duplicate_topics"] - before >= 40000 - 3100


def test_caps_with_saturated_node_counters(B):
    """a node with >= 255 persistent / group routes saturates its one-byte caps counter: the topic must be handed to the
    exact caps kernel, which then decides (no drops when the real count is below the cap)"""
    pairs = []
    for i in range(400):
        normal(B, pairs, "t", "big/+", 1, "p%d" % i, "d", 1)
    for i in range(300):
        group(B, pairs, "t", "big/#", "g%03d" % i, {B.schema.receiver_url(0, "m", "d"): 1})
    for i in range(10):
        normal(B, pairs, "t", "big/x", 0, "n%d" % i, "d", 1)
    pairs.sort()
    idx = make_index(B, pairs)
    kv = oracle_kv_from_pairs(pairs)
    topics = ["big/x", "big/y", "small"]
    tt = np.zeros(3, np.int32)
    for caps in [(1000, 1000), (400, 300), (399, 299), (100, 7), (0, 0), (2 ** 31 - 1, 100)]:
        compare_with_oracle(B, idx, kv, ["t"], topics, tt, caps[0], caps[1], O.MODE_BRUTE)
    before = idx.stats()["flagged_topics"]
    res = idx.match_topics(["t"], topics, tt, [1000], [1000])
    assert len(res.throttled) == 0 and res.route_count.tolist() == [710, 700, 0]
    res.close()
    assert idx.stats()["flagged_topics"] == before + 2   # saturated counters force the exact path even though nothing drops


def test_matches_keep_running_during_commit(B):
    """bfq_index_commit rebuilds under the staging lock only: concurrent matches keep answering from the previous snapshot and
    every answer is a whole-snapshot answer (old or new, never a mix)"""
    import threading
    w = B.workload.Workload("C3", scale=0.02)
    idx = B.pkg.GpuRouteIndex(0)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    tenants = w.tenants
    topics = w.topic_list()[:2000]
    tt = np.ascontiguousarray(w.topic_tenant[:2000])

    def snapshot_answer():
        r = idx.match_topics(tenants, topics, tt)
        rc = r.route_count.copy()
        r.close()
        return rc.tolist()
    old = snapshot_answer()
    # the delta: a catch-all '#' route for every tenant -> every non-'$' topic gains exactly one route
    extra = [(B.schema.route_key(t, "#", B.schema.receiver_url(0, "catchall", "d")), B.schema.incarnation_bytes(1)) for t in tenants]
    idx.apply(adds=extra)
    new_expected = [c + 1 for c in old]
    seen, stop, errs = [], threading.Event(), []

    def matcher():
        try:
            while not stop.is_set():
                seen.append(snapshot_answer())
        except Exception as e:   # pragma: no cover
            errs.append(e)
    th = threading.Thread(target=matcher)
    th.start()
    idx.commit()
    stop.set()
    th.join()
    assert not errs
    assert snapshot_answer() == new_expected
    assert seen and all(s == old or s == new_expected for s in seen)
