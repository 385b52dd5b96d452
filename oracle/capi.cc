// oracle/capi.cc — CPU ORACLE (test infrastructure only; see oracle.h).
// Flat C entry points so tests/ and bench.py can drive the oracle through ctypes.
#include <atomic>
#include <cstring>
#include <thread>
#include <unordered_map>

#include "oracle.h"

using namespace orc;

namespace {
struct KVHandle {
    SortedKV kv;
    std::shared_ptr<FilterTrie> trie;
};
struct ResultHandle {
    MatchResult r;
};
std::string S(const uint8_t* p, int64_t n) { return std::string((const char*) p, (size_t) n); }
int64_t emit(const std::string& s, uint8_t* out, int64_t cap) {
    if ((int64_t) s.size() <= cap && out) memcpy(out, s.data(), s.size());
    return (int64_t) s.size();
}
void put_u32(std::string& o, uint32_t v) { o.append((const char*) &v, 4); }
void put_str(std::string& o, const std::string& s) {
    put_u32(o, (uint32_t) s.size());
    o += s;
}
void put_levels(std::string& o, const Levels& l) {
    put_u32(o, (uint32_t) l.size());
    for (const auto& s : l) put_str(o, s);
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- JDK / TopicUtil
int32_t orc_java_hash(const uint8_t* s, int64_t n) { return java_hash(S(s, n)); }
int32_t orc_java_compare(const uint8_t* a, int64_t an, const uint8_t* b, int64_t bn) { return java_compare(S(a, an), S(b, bn)); }
int32_t orc_bucket(const uint8_t* s, int64_t n) { return bucket(S(s, n)); }
int64_t orc_parse(const uint8_t* s, int64_t n, int32_t escaped, uint8_t* out, int64_t cap) {
    std::string o;
    put_levels(o, parse(S(s, n), escaped != 0));
    return emit(o, out, cap);
}
int32_t orc_is_valid_topic(const uint8_t* s, int64_t n, int32_t maxLevelLength, int32_t maxLevel, int32_t maxLength) {
    return is_valid_topic(S(s, n), maxLevelLength, maxLevel, maxLength);
}
int32_t orc_is_valid_topic_filter(const uint8_t* s, int64_t n, int32_t maxLevelLength, int32_t maxLevel, int32_t maxLength) {
    return is_valid_topic_filter(S(s, n), maxLevelLength, maxLevel, maxLength);
}
int32_t orc_is_wildcard_topic_filter(const uint8_t* s, int64_t n) { return is_wildcard_topic_filter(S(s, n)); }
int32_t orc_is_shared_subscription(const uint8_t* s, int64_t n) { return is_shared_subscription(S(s, n)); }
int32_t orc_is_ordered_shared(const uint8_t* s, int64_t n) { return is_ordered_shared(S(s, n)); }
int32_t orc_is_unordered_shared(const uint8_t* s, int64_t n) { return is_unordered_shared(S(s, n)); }
// RouteMatcher serialised as: u32 type, levels, str group, str mqttTopicFilter
int64_t orc_route_matcher_from(const uint8_t* s, int64_t n, uint8_t* out, int64_t cap) {
    RouteMatcher m = route_matcher_from(S(s, n));
    std::string o;
    put_u32(o, (uint32_t) m.type);
    put_levels(o, m.filterLevels);
    put_str(o, m.group);
    put_str(o, m.mqttTopicFilter);
    return emit(o, out, cap);
}

// ---------------------------------------------------------------- KVSchemaUtil
int64_t orc_receiver_url(int32_t subBrokerId, const uint8_t* rid, int64_t rn, const uint8_t* dk, int64_t dn, uint8_t* out, int64_t cap) {
    return emit(to_receiver_url(subBrokerId, S(rid, rn), S(dk, dn)), out, cap);
}
int64_t orc_tenant_begin_key(const uint8_t* t, int64_t tn, uint8_t* out, int64_t cap) {
    return emit(tenant_begin_key(S(t, tn)), out, cap);
}
// retain store schema
int64_t orc_retain_key(const uint8_t* t, int64_t tn, const uint8_t* topic, int64_t n, uint8_t* out, int64_t cap) {
    return emit(retain_message_key(S(t, tn), S(topic, n)), out, cap);
}
// KVSchemaUtilTest.toRetainMessageKeyPrefix (test helper, :88-94): levels = the filter's level count, one less under a final '#'
int64_t orc_retain_key_prefix(const uint8_t* t, int64_t tn, const uint8_t* tf, int64_t fn, uint8_t* out, int64_t cap) {
    const std::vector<std::string> lv = parse(S(tf, fn), false);
    const bool multi = !lv.empty() && lv.back() == "#";
    return emit(retain_key_prefix(S(t, tn), (int) (multi ? lv.size() - 1 : lv.size()), retain_filter_prefix(lv)), out, cap);
}
int32_t orc_level_hash_byte(const uint8_t* l, int64_t n) { return (int32_t) level_hash_byte(S(l, n)); }
int64_t orc_tenant_route_start_key(const uint8_t* t, int64_t tn, const uint8_t* filter, int64_t fn, uint8_t* out, int64_t cap) {
    return emit(tenant_route_start_key(S(t, tn), parse(S(filter, fn), false)), out, cap);
}
// route key of an MQTT topic filter ("a/+", "$share/g/a/#", ...): normal filters need receiverUrl
int64_t orc_route_key(const uint8_t* t, int64_t tn, const uint8_t* tf, int64_t fn, const uint8_t* url, int64_t un, uint8_t* out, int64_t cap) {
    RouteMatcher m = route_matcher_from(S(tf, fn));
    if (m.type == RouteMatcher::Normal) return emit(to_normal_route_key(S(t, tn), m, S(url, un)), out, cap);
    return emit(to_group_route_key(S(t, tn), m), out, cap);
}
int64_t orc_upper_bound(const uint8_t* k, int64_t kn, uint8_t* out, int64_t cap) {
    bool open = false;
    std::string u = upper_bound(S(k, kn), &open);
    if (open) return -1;
    return emit(u, out, cap);
}
int64_t orc_route_group(int64_t n, const uint8_t* urls, const int64_t* url_off, const uint64_t* inc, uint8_t* out, int64_t cap) {
    std::vector<std::pair<std::string, uint64_t>> members;
    for (int64_t i = 0; i < n; i++) members.emplace_back(S(urls + url_off[i], url_off[i + 1] - url_off[i]), inc[i]);
    return emit(encode_route_group(members), out, cap);
}
// Matching serialised as: u32 type(0 normal,1 group), str tenant, str mqttTopicFilter, levels, str receiverUrl,
//                         u64 incarnation, i32 subBrokerId, u32 nmembers, (str url, u64 inc)*
int64_t orc_build_match_route(const uint8_t* k, int64_t kn, const uint8_t* v, int64_t vn, uint8_t* out, int64_t cap) {
    Matching m;
    try {
        m = build_match_route(S(k, kn), S(v, vn));
    } catch (const std::exception&) {
        return -1;
    }
    std::string o;
    put_u32(o, (uint32_t) m.type);
    put_str(o, m.tenantId);
    put_str(o, m.mqttTopicFilter);
    put_levels(o, m.filterLevels);
    put_str(o, m.receiverUrl);
    o.append((const char*) &m.incarnation, 8);
    put_u32(o, (uint32_t) m.subBrokerId);
    put_u32(o, (uint32_t) m.members.size());
    for (const auto& e : m.members) {
        put_str(o, e.first);
        o.append((const char*) &e.second, 8);
    }
    return emit(o, out, cap);
}

// ---------------------------------------------------------------- sorted KV
void* orc_kv_new() { return new KVHandle(); }
void orc_kv_free(void* h) { delete (KVHandle*) h; }
void orc_kv_put(void* h, const uint8_t* k, int64_t kn, const uint8_t* v, int64_t vn) {
    auto* kv = (KVHandle*) h;
    kv->kv.put(S(k, kn), S(v, vn));
    kv->trie.reset();
}
void orc_kv_erase(void* h, const uint8_t* k, int64_t kn) {
    auto* kv = (KVHandle*) h;
    kv->kv.erase(S(k, kn));
    kv->trie.reset();
}
void orc_kv_load(void* h, const uint8_t* keys, const int64_t* koff, const uint8_t* vals, const int64_t* voff, int64_t n) {
    auto* kv = (KVHandle*) h;
    for (int64_t i = 0; i < n; i++) kv->kv.put(S(keys + koff[i], koff[i + 1] - koff[i]), S(vals + voff[i], voff[i + 1] - voff[i]));
    kv->trie.reset();
}
int64_t orc_kv_size(void* h) { return (int64_t) ((KVHandle*) h)->kv.size(); }
void orc_kv_freeze(void* h) { ((KVHandle*) h)->kv.freeze(); }
int64_t orc_kv_key(void* h, int64_t rank, uint8_t* out, int64_t cap) { return emit(((KVHandle*) h)->kv.key(rank), out, cap); }
int64_t orc_kv_value(void* h, int64_t rank, uint8_t* out, int64_t cap) { return emit(((KVHandle*) h)->kv.value(rank), out, cap); }
int64_t orc_kv_lower_bound(void* h, const uint8_t* k, int64_t kn) { return ((KVHandle*) h)->kv.lower_bound(S(k, kn)); }

// ---------------------------------------------------------------- forward match
// mode: 0 = literal reference algorithm, 1 = brute force predicate, 2 = filter-trie walk.
// singleton != 0: one matchAll call per topic (the production shape, DW/cache/TenantRouteCache.java:185-186).
void* orc_match_batch(void* h, int32_t mode, int32_t singleton, const uint8_t* tenants, const int64_t* tenant_off,
                      int64_t n_tenants, const uint8_t* topics, const int64_t* topic_off, const int32_t* topic_tenant,
                      int64_t n, int32_t maxP, int32_t maxG, int32_t nthreads) {
    auto* kvh = (KVHandle*) h;
    kvh->kv.freeze();
    if (mode == 2 && !kvh->trie) kvh->trie = build_filter_trie(kvh->kv);
    auto* res = new ResultHandle();
    MatchResult& R = res->r;
    R.routes.resize(n);
    R.persistentFanout.assign(n, 0);
    R.groupFanout.assign(n, 0);
    std::vector<std::string> tenantIds(n_tenants);
    for (int64_t t = 0; t < n_tenants; t++) tenantIds[t] = S(tenants + tenant_off[t], tenant_off[t + 1] - tenant_off[t]);
    // work items: (tenant, topic index list)
    std::vector<std::pair<int, std::vector<int64_t>>> items;
    if (singleton) {
        items.reserve(n);
        for (int64_t i = 0; i < n; i++) items.push_back({topic_tenant[i], {i}});
    } else {
        std::vector<std::vector<int64_t>> per(n_tenants);
        for (int64_t i = 0; i < n; i++) per[topic_tenant[i]].push_back(i);
        // topics are independent: for the per-topic matchers (brute force, trie walk) split a tenant's topics into
        // pieces so that one huge tenant does not serialise on a single thread; the literal algorithm keeps the
        // whole-tenant batch (its topic trie is per matchAll call)
        const size_t piece = mode == 0 ? (size_t) -1 : 512;
        for (int64_t t = 0; t < n_tenants; t++) {
            if (per[t].empty()) continue;
            if (per[t].size() <= piece) {
                items.push_back({(int) t, std::move(per[t])});
            } else {
                for (size_t b = 0; b < per[t].size(); b += piece)
                    items.push_back({(int) t, std::vector<int64_t>(per[t].begin() + b, per[t].begin() + std::min(per[t].size(), b + piece))});
            }
        }
    }
    std::atomic<size_t> cursor{0};
    const int T = std::max(1, nthreads);
    std::vector<MatchStats> tstats(T);
    std::vector<std::vector<ThrottleEvent>> tevents(T);
    auto worker = [&](int tid) {
        const size_t chunk = singleton ? 64 : 1;
        while (true) {
            size_t b = cursor.fetch_add(chunk);
            if (b >= items.size()) break;
            size_t e = std::min(items.size(), b + chunk);
            for (size_t w = b; w < e; w++) {
                const auto& it = items[w];
                std::vector<std::string> ts;
                ts.reserve(it.second.size());
                for (int64_t i : it.second) ts.push_back(S(topics + topic_off[i], topic_off[i + 1] - topic_off[i]));
                MatchResult r;
                const std::string& tenant = tenantIds[it.first];
                if (mode == 0) r = match_all_reference(kvh->kv, tenant, ts, maxP, maxG);
                else if (mode == 1) r = match_all_brute(kvh->kv, tenant, ts, maxP, maxG);
                else r = match_all_trie(*kvh->trie, kvh->kv, tenant, ts, maxP, maxG);
                for (size_t j = 0; j < it.second.size(); j++) {
                    int64_t gi = it.second[j];
                    R.routes[gi] = std::move(r.routes[j]);
                    R.persistentFanout[gi] = r.persistentFanout[j];
                    R.groupFanout[gi] = r.groupFanout[j];
                }
                for (auto ev : r.events) {
                    ev.topicIdx = (int) it.second[ev.topicIdx];
                    tevents[tid].push_back(ev);
                }
                auto& s = tstats[tid];
                s.seeks += r.stats.seeks; s.nexts += r.stats.nexts; s.backwardSeeks += r.stats.backwardSeeks;
                s.V += r.stats.V; s.P += r.stats.P; s.R += r.stats.R; s.ranges += r.stats.ranges;
            }
        }
    };
    if (T == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    for (int t = 0; t < T; t++) {
        R.events.insert(R.events.end(), tevents[t].begin(), tevents[t].end());
        R.stats.seeks += tstats[t].seeks; R.stats.nexts += tstats[t].nexts; R.stats.backwardSeeks += tstats[t].backwardSeeks;
        R.stats.V += tstats[t].V; R.stats.P += tstats[t].P; R.stats.R += tstats[t].R; R.stats.ranges += tstats[t].ranges;
    }
    return res;
}
void orc_result_free(void* r) { delete (ResultHandle*) r; }
int64_t orc_result_total_routes(void* r) {
    int64_t t = 0;
    for (const auto& v : ((ResultHandle*) r)->r.routes) t += (int64_t) v.size();
    return t;
}
void orc_result_routes(void* r, int64_t* offsets, int64_t* ranks) {
    const auto& R = ((ResultHandle*) r)->r;
    int64_t p = 0;
    for (size_t i = 0; i < R.routes.size(); i++) {
        offsets[i] = p;
        for (int64_t x : R.routes[i]) ranks[p++] = x;
    }
    offsets[R.routes.size()] = p;
}
void orc_result_fanouts(void* r, int32_t* pf, int32_t* gf) {
    const auto& R = ((ResultHandle*) r)->r;
    for (size_t i = 0; i < R.routes.size(); i++) {
        pf[i] = R.persistentFanout[i];
        gf[i] = R.groupFanout[i];
    }
}
int64_t orc_result_num_events(void* r) { return (int64_t) ((ResultHandle*) r)->r.events.size(); }
void orc_result_events(void* r, int32_t* kind, int32_t* topicIdx, int64_t* rank, int32_t* maxCount) {
    const auto& E = ((ResultHandle*) r)->r.events;
    for (size_t i = 0; i < E.size(); i++) {
        kind[i] = (int32_t) E[i].kind;
        topicIdx[i] = E[i].topicIdx;
        rank[i] = E[i].routeRank;
        maxCount[i] = E[i].maxCount;
    }
}
// out[7] = seeks, nexts, V, P, R, ranges, backwardSeeks
void orc_result_stats(void* r, uint64_t* out) {
    const auto& s = ((ResultHandle*) r)->r.stats;
    out[0] = s.seeks; out[1] = s.nexts; out[2] = s.V; out[3] = s.P; out[4] = s.R; out[5] = s.ranges; out[6] = s.backwardSeeks;
}

// ---------------------------------------------------------------- expansion set (TopicFilterIterator)
// Enumerate the whole expansion set of a topic batch in iterator order. Output:
// u32 count, then per filter: levels, u32 nvalues, u32 topicIdx*
int64_t orc_expansion_list(const uint8_t* topics, const int64_t* topic_off, int64_t n, int32_t isGlobal, uint8_t* out, int64_t cap) {
    TopicTrie trie(isGlobal != 0);
    for (int64_t i = 0; i < n; i++) trie.add_topic(parse(S(topics + topic_off[i], topic_off[i + 1] - topic_off[i]), false), (int) i);
    TopicFilterIterator it(trie);
    std::string body;
    uint32_t count = 0;
    for (; it.is_valid(); it.next()) {
        put_levels(body, it.key());
        auto v = it.value();
        put_u32(body, (uint32_t) v.size());
        for (int x : v) put_u32(body, (uint32_t) x);
        count++;
    }
    std::string o;
    put_u32(o, count);
    o += body;
    return emit(o, out, cap);
}
// TenantRangeLookupCache.lookup(CacheKey) literally (bifromq-dist/bifromq-dist-server/src/main/java/org/apache/bifromq/dist/server/
// scheduler/TenantRangeLookupCache.java:70-106) for ONE topic of a tenant over its ordered candidate ranges: flags[k] bit 0 = the
// candidate has a Fact, bit 1 = the Fact has firstGlobalFilterLevels, bit 2 = lastGlobalFilterLevels; first / last are the global
// filter levels joined by NUL (level 0 = tenant id). keep[k] = 1 iff the candidate is in the returned collection.
void orc_range_lookup(const uint8_t* tenant, int64_t tn, const uint8_t* topic, int64_t pn, int64_t n_cand, const uint8_t* flags,
                      const uint8_t* first_blob, const int64_t* first_off, const uint8_t* last_blob, const int64_t* last_off, uint8_t* keep) {
    TopicTrie trie(true);                                                          // TopicTrieNode.builder(true)   :71
    Levels global{S(tenant, tn)};
    for (auto& l : parse(S(topic, pn), false)) global.push_back(l);                // TopicUtil.parse(tenantId, topic, false)   :72
    trie.add_topic(global, 0);
    TopicFilterIterator it(trie);                                                   // :73-75
    auto joined = [](const Levels& l) {                                             // fastJoin(NUL, levels)
        std::string s;
        for (size_t i = 0; i < l.size(); i++) {
            if (i) s.push_back('\0');
            s += l[i];
        }
        return s;
    };
    for (int64_t k = 0; k < n_cand; k++) keep[k] = 0;
    for (int64_t k = 0; k < n_cand; k++) {                                          // :77
        if (!(flags[k] & 1)) {                                                      // no Fact: conservatively included   :79-82
            keep[k] = 1;
            continue;
        }
        if ((flags[k] & 6) != 6) continue;                                          // range is empty   :84-87
        const Levels first = parse(S(first_blob + first_off[k], first_off[k + 1] - first_off[k]), true);
        const Levels last = parse(S(last_blob + last_off[k], last_off[k + 1] - last_off[k]), true);
        it.seek(first);                                                             // :90
        if (it.is_valid()) {                                                        // :91
            const Levels key = it.key();
            if (key == first || java_compare(joined(key), joined(last)) <= 0) keep[k] = 1;   // :93-98
        } else {
            break;                                                                  // :99-102
        }
    }
}

// seek(filter) -> serialised key levels, or -1 when the cursor is invalid
int64_t orc_expansion_seek(const uint8_t* topics, const int64_t* topic_off, int64_t n, int32_t isGlobal,
                           const uint8_t* filter, int64_t fn, uint8_t* out, int64_t cap) {
    TopicTrie trie(isGlobal != 0);
    for (int64_t i = 0; i < n; i++) trie.add_topic(parse(S(topics + topic_off[i], topic_off[i + 1] - topic_off[i]), false), (int) i);
    TopicFilterIterator it(trie);
    it.seek(fn < 0 ? Levels{} : parse(S(filter, fn), false));
    if (!it.is_valid()) return -1;
    std::string o;
    put_levels(o, it.key());
    return emit(o, out, cap);
}
int32_t orc_topic_matches_filter(const uint8_t* t, int64_t tn, const uint8_t* f, int64_t fn) {
    return topic_matches_filter(parse(S(t, tn), false), parse(S(f, fn), false));
}

// ---------------------------------------------------------------- inverse match
void* orc_tli_new() { return new TopicLevelIndex(); }
void orc_tli_free(void* h) { delete (TopicLevelIndex*) h; }
static Levels tli_levels(const uint8_t* tenant, int64_t tn, const uint8_t* s, int64_t n) {
    Levels l;
    if (tn >= 0) l.push_back(S(tenant, tn));  // TopicUtil.parse(tenantId, topic, false) :198-202
    Levels rest = parse(S(s, n), false);
    l.insert(l.end(), rest.begin(), rest.end());
    return l;
}
// tn < 0 => TopicIndex (no tenant level); tn >= 0 => RetainTopicIndex (tenantId is level 0)
void orc_tli_add(void* h, const uint8_t* tenant, int64_t tn, const uint8_t* topic, int64_t n, int64_t value) {
    ((TopicLevelIndex*) h)->add(tli_levels(tenant, tn, topic, n), value);
}
void orc_tli_remove(void* h, const uint8_t* tenant, int64_t tn, const uint8_t* topic, int64_t n, int64_t value) {
    ((TopicLevelIndex*) h)->remove(tli_levels(tenant, tn, topic, n), value);
}
int64_t orc_tli_match(void* h, const uint8_t* tenant, int64_t tn, const uint8_t* filter, int64_t fn, int64_t* out, int64_t cap, uint64_t* visited) {
    auto v = ((TopicLevelIndex*) h)->match(tli_levels(tenant, tn, filter, fn), tn >= 0 ? 1 : 0, visited);
    if ((int64_t) v.size() <= cap && out) memcpy(out, v.data(), v.size() * 8);
    return (int64_t) v.size();
}
int64_t orc_tli_get(void* h, const uint8_t* topic, int64_t n, int64_t* out, int64_t cap) {
    auto v = ((TopicLevelIndex*) h)->get(parse(S(topic, n), false));
    if ((int64_t) v.size() <= cap && out) memcpy(out, v.data(), v.size() * 8);
    return (int64_t) v.size();
}
int64_t orc_tli_find_all(void* h, int64_t* out, int64_t cap) {
    auto v = ((TopicLevelIndex*) h)->find_all();
    if ((int64_t) v.size() <= cap && out) memcpy(out, v.data(), v.size() * 8);
    return (int64_t) v.size();
}
// batch inverse match over nthreads; per-filter limit (<0 = unlimited) truncates in ascending value order
void orc_tli_match_batch(void* h, const uint8_t* tenants, const int64_t* tenant_off, const int32_t* filter_tenant,
                         const uint8_t* filters, const int64_t* filter_off, int64_t n, int32_t nthreads,
                         int64_t* counts, uint64_t* visited_total) {
    auto* idx = (TopicLevelIndex*) h;
    std::atomic<int64_t> cursor{0};
    std::atomic<uint64_t> vis{0};
    auto worker = [&]() {
        uint64_t local = 0;
        while (true) {
            int64_t b = cursor.fetch_add(16);
            if (b >= n) break;
            for (int64_t i = b; i < std::min(n, b + 16); i++) {
                int t = filter_tenant ? filter_tenant[i] : -1;
                Levels l = t >= 0 ? tli_levels(tenants + tenant_off[t], tenant_off[t + 1] - tenant_off[t], filters + filter_off[i], filter_off[i + 1] - filter_off[i])
                                  : tli_levels(nullptr, -1, filters + filter_off[i], filter_off[i + 1] - filter_off[i]);
                counts[i] = (int64_t) idx->match(l, t >= 0 ? 1 : 0, &local).size();
            }
        }
        vis += local;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, nthreads); t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    if (visited_total) *visited_total = vis.load();
}

}  // extern "C"
