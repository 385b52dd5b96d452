// workload.cc — seeded synthetic workloads for BASELINE.json's configs C1..C5 (libbfq_workload.so).
//
// Emits exactly what the reference would hold / receive: route KV pairs in the dist-worker key encoding
// (codec.h; values = 8-byte BE incarnation or a RouteGroup proto), sorted in KV byte order, plus a batch of
// publish topics (or, for C5, retained topics and wildcard SUBSCRIBE filters). The oracle, the CPU baseline
// and the CUDA path all consume these same bytes. Level names use [a-z0-9_-], length U[3,12], depth U[3,8].
//
// Names come from an implicit per-tenant tree: node = hash of its path; child j of a node has a name derived
// from (node hash, j). Filters are paths with levels replaced by '+' / a trailing '#'; publish topics are
// 80% "hit" topics derived from a random filter's path (wildcard positions re-drawn) and 20% fresh paths.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "codec.h"

namespace {

using bfq::sv;

struct Rng {  // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t) (((next() >> 32) * (uint64_t) n) >> 32); }
    double uniform() { return (double) (next() >> 11) * (1.0 / 9007199254740992.0); }
};
inline uint64_t mix(uint64_t a, uint64_t b) {
    uint64_t z = a ^ (b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2));
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
const char ALPHABET[] = "abcdefghijklmnopqrstuvwxyz0123456789_-";
const int BRANCH[8] = {16, 16, 8, 8, 6, 4, 4, 4};

void append_name(std::string& out, uint64_t node_hash, uint32_t child) {
    Rng r(mix(node_hash, child));
    int len = 3 + (int) r.below(10);
    for (int i = 0; i < len; i++) out.push_back(ALPHABET[r.below(38)]);
}

struct Filter {
    uint8_t depth;       // number of levels in the base path
    uint8_t idx[8];      // child index per level
    uint8_t plus_mask;   // bit l set => level l is '+'
    uint8_t hash_tail;   // 1 => "/#" appended after `depth` levels
};

struct Blob {
    std::vector<uint8_t> data;
    std::vector<int64_t> off{0};
    void push(sv s) {
        data.insert(data.end(), s.begin(), s.end());
        off.push_back((int64_t) data.size());
    }
    int64_t n() const { return (int64_t) off.size() - 1; }
    void append(const Blob& o) {
        const int64_t base = (int64_t) data.size();
        data.insert(data.end(), o.data.begin(), o.data.end());
        for (size_t i = 1; i < o.off.size(); i++) off.push_back(base + o.off[i]);
    }
};

struct TenantGen {
    std::string id;
    uint64_t seed;
    std::vector<Filter> filters;

    // render the base path (no wildcards) with optional re-draw of the '+' levels and an extension below '#'
    std::string render_topic(const Filter& f, Rng* redraw) const {
        std::string s;
        uint64_t h = seed;
        for (int l = 0; l < f.depth; l++) {
            uint32_t c = f.idx[l];
            if (redraw && (f.plus_mask >> l & 1)) c = redraw->below((uint32_t) BRANCH[l]);
            if (l) s.push_back('/');
            append_name(s, h, c);
            h = mix(h, c);
        }
        if (redraw && f.hash_tail) {
            int extra = (int) redraw->below(3);
            for (int l = f.depth; l < f.depth + extra && l < 8; l++) {
                uint32_t c = redraw->below((uint32_t) BRANCH[l]);
                s.push_back('/');
                append_name(s, h, c);
                h = mix(h, c);
            }
        }
        return s;
    }
    std::string render_filter(const Filter& f) const {
        std::string s;
        uint64_t h = seed;
        for (int l = 0; l < f.depth; l++) {
            if (l) s.push_back('/');
            if (f.plus_mask >> l & 1) s.push_back('+');
            else append_name(s, h, f.idx[l]);
            h = mix(h, f.idx[l]);
        }
        if (f.hash_tail) s += "/#";
        return s;
    }
    Filter random_path(Rng& r, int depth) const {
        Filter f{};
        f.depth = (uint8_t) depth;
        for (int l = 0; l < depth; l++) f.idx[l] = (uint8_t) r.below((uint32_t) BRANCH[l]);
        return f;
    }
};

std::string u64be(uint64_t v) {
    std::string s(8, '\0');
    for (int i = 0; i < 8; i++) s[i] = (char) (v >> (56 - 8 * i));
    return s;
}
void put_varint(std::string& o, uint64_t v) {
    while (v >= 0x80) { o.push_back((char) ((v & 0x7F) | 0x80)); v >>= 7; }
    o.push_back((char) v);
}
std::string route_group(const std::vector<std::pair<std::string, uint64_t>>& members) {  // RouteGroup{map<string,uint64> members=1}
    std::string out;
    for (auto& m : members) {
        std::string e;
        e.push_back(0x0A); put_varint(e, m.first.size()); e += m.first;
        e.push_back(0x10); put_varint(e, m.second);
        out.push_back(0x0A); put_varint(out, e.size()); out += e;
    }
    return out;
}

enum Cfg { C1 = 1, C2, C3, C4, C5 };

struct Workload {
    Blob keys, vals, tenants, topics, filters;
    std::vector<int32_t> topic_tenant, filter_tenant;
    int64_t n_filters = 0;
};

uint64_t fnv1a64(sv s) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
    return h;
}

struct Plan {
    Cfg cfg;
    int64_t n_tenants, n_filters, n_topics, n_cfilters /*C5 query filters*/;
    double p_plus_filter, p_hash_filter, p_group;
    bool zipf;
};

Plan make_plan(Cfg c, double scale) {
    auto S = [&](double v, double lo) { return (int64_t) std::max(lo, std::floor(v * scale + 0.5)); };
    switch (c) {
        case C1: return {c, 1, S(10000, 50), S(1000, 20), 0, 0.0, 0.0, 0.0, false};
        case C2: return {c, 1, S(1000000, 200), S(100000, 50), 0, 0.5, 0.0, 0.0, false};
        case C3: return {c, S(1000, 4), S(10000000, 400), S(1000000, 100), 0, 0.25, 0.15, 0.02, false};
        case C4: return {c, S(1000, 4), S(10000000, 400), S(1000000, 100), 0, 0.25, 0.15, 0.02, true};
        default: return {C5, S(1000, 4), S(1000000, 200), S(1000000, 200), S(100000, 50), 0.5, 0.5, 0.0, false};
    }
}

void gen_filters(TenantGen& t, int64_t nf, const Plan& pl, Rng& r) {
    std::unordered_set<uint64_t> seen;
    seen.reserve((size_t) nf * 2);
    t.filters.reserve((size_t) nf);
    int64_t attempts = 0;
    while ((int64_t) t.filters.size() < nf && attempts < nf * 20 + 1000) {
        attempts++;
        int depth = 3 + (int) r.below(6);
        Filter f = t.random_path(r, depth);
        if (pl.cfg != C5) {
            double u = r.uniform();
            if (u < pl.p_plus_filter) {
                // each non-first level becomes '+' w.p. 0.5, at least one
                for (int l = 1; l < depth; l++)
                    if (r.below(2)) f.plus_mask |= (uint8_t) (1u << l);
                if (!f.plus_mask) f.plus_mask = (uint8_t) (1u << (1 + r.below((uint32_t) depth - 1)));
            } else if (u < pl.p_plus_filter + pl.p_hash_filter) {
                f.depth = (uint8_t) (1 + r.below((uint32_t) depth - 1));
                f.hash_tail = 1;
            }
        }
        uint64_t sig = f.depth | ((uint64_t) f.plus_mask << 8) | ((uint64_t) f.hash_tail << 16);
        uint64_t h = mix(t.seed, sig);
        for (int l = 0; l < f.depth; l++) h = mix(h, (f.plus_mask >> l & 1) ? 255u : f.idx[l]);
        if (!seen.insert(h).second) continue;
        t.filters.push_back(f);
    }
}

}  // namespace

struct bfqw {
    Workload w;
};

extern "C" {

// config: "C1".."C5"; scale shrinks every size (1.0 = BASELINE.json sizes); shard_index/shard_count keep only
// the tenants with fnv1a64(tenantId) % shard_count == shard_index (tenant sharding across GPUs);
// tenant_prefix namespaces tenant ids (weak-scaling runs give each rank its own tenants).
// replicate_hot != 0: a tenant that carries more than 1 / (4 shard_count) of the batch is hosted by EVERY shard and its
// topics are dealt round-robin by batch position (SURVEY.md 8e "replicas" mode, applied per hot tenant): with Zipf tenant
// sizes the largest tenant alone is 13 % of C4's batch, so pure hash placement caps 8 GPUs at ~4x.
// topic_mult > 1: the publish batch is topic_mult times the config's size (same distributions): N GPUs serving N times the
// publish traffic of ONE filter set.
bfqw* bfqw_generate2(const char* config, uint64_t seed, double scale, int32_t shard_index, int32_t shard_count,
                     const char* tenant_prefix, int32_t nthreads, int32_t replicate_hot, int32_t topic_mult);
bfqw* bfqw_generate(const char* config, uint64_t seed, double scale, int32_t shard_index, int32_t shard_count,
                    const char* tenant_prefix, int32_t nthreads) {
    return bfqw_generate2(config, seed, scale, shard_index, shard_count, tenant_prefix, nthreads, 0, 1);
}
bfqw* bfqw_generate2(const char* config, uint64_t seed, double scale, int32_t shard_index, int32_t shard_count,
                     const char* tenant_prefix, int32_t nthreads, int32_t replicate_hot, int32_t topic_mult) {
    Cfg cfg = (Cfg) (config && config[0] == 'C' ? config[1] - '0' : 0);
    if (cfg < C1 || cfg > C5) return nullptr;
    if (shard_count < 1) shard_count = 1;
    const Plan pl = make_plan(cfg, scale);
    const std::string prefix = tenant_prefix ? tenant_prefix : "";
    auto* out = new bfqw();
    Workload& W = out->w;

    // ---- tenants and their sizes
    std::vector<TenantGen> tg((size_t) pl.n_tenants);
    std::vector<int64_t> nf((size_t) pl.n_tenants);
    double hsum = 0;
    for (int64_t t = 0; t < pl.n_tenants; t++) hsum += 1.0 / (double) (t + 1);
    for (int64_t t = 0; t < pl.n_tenants; t++) {
        tg[(size_t) t].id = prefix + (pl.n_tenants == 1 ? "t0" : "tenant" + std::to_string(t));
        tg[(size_t) t].seed = mix(seed, (uint64_t) t + 1);
        nf[(size_t) t] = pl.zipf ? std::max<int64_t>(8, (int64_t) ((double) pl.n_filters / hsum / (double) (t + 1)))
                                 : std::max<int64_t>(1, pl.n_filters / pl.n_tenants);
    }
    std::vector<char> mine((size_t) pl.n_tenants), hot((size_t) pl.n_tenants, 0);
    {
        double wsum = 0;
        for (int64_t t = 0; t < pl.n_tenants; t++) wsum += pl.zipf ? (double) nf[(size_t) t] : 1.0;
        for (int64_t t = 0; t < pl.n_tenants; t++) {
            const double share = (pl.zipf ? (double) nf[(size_t) t] : 1.0) / wsum;   // topics are drawn in proportion to this
            hot[(size_t) t] = replicate_hot && shard_count > 1 && cfg != C5 && share > 1.0 / (4.0 * shard_count);
            mine[(size_t) t] = hot[(size_t) t] || (int32_t) (fnv1a64(tg[(size_t) t].id) % (uint64_t) shard_count) == shard_index;
        }
    }

    // ---- per-tenant generation (parallel): filters, then sorted route KV pairs
    struct TenantOut { Blob keys, vals; };
    std::vector<TenantOut> touts((size_t) pl.n_tenants);
    std::atomic<int64_t> cursor{0};
    auto worker = [&]() {
        while (true) {
            int64_t t = cursor.fetch_add(1);
            if (t >= pl.n_tenants) break;
            if (!mine[(size_t) t]) continue;
            TenantGen& T = tg[(size_t) t];
            Rng r(mix(T.seed, 0xF117E5));
            gen_filters(T, nf[(size_t) t], pl, r);
            if (cfg == C5) continue;  // C5 indexes topics, not routes
            std::vector<std::pair<std::string, std::string>> kvs;
            kvs.reserve(T.filters.size() + T.filters.size() / 8);
            const double top = std::min(10000.0, std::max(1.0, (double) T.filters.size() / 8.0));
            uint64_t rcount = 0;
            for (size_t j = 0; j < T.filters.size(); j++) {
                const std::string tf = T.render_filter(T.filters[j]);
                int64_t nroutes = 1;
                if (pl.zipf) nroutes = std::max<int64_t>(1, (int64_t) (top * std::pow((double) (j + 1), -1.1)));
                for (int64_t k = 0; k < nroutes; k++) {
                    const uint64_t rid = rcount++;
                    const int broker = cfg == C1 ? 0 : (rid % 4 == 0 ? 1 : 0);
                    std::string url = bfq::make_receiver_url(broker, "r" + std::to_string(rid), "d" + std::to_string(rid % 16));
                    kvs.emplace_back(bfq::make_route_key(T.id, tf, url), u64be(1 + rid % 7));
                }
                if (pl.p_group > 0 && r.uniform() < pl.p_group) {
                    const std::string g = "g" + std::to_string(r.below(64));
                    std::vector<std::pair<std::string, uint64_t>> members;
                    int nm = 1 + (int) r.below(8);
                    for (int m = 0; m < nm; m++)
                        members.emplace_back(bfq::make_receiver_url((int) r.below(2), "m" + std::to_string(r.below(100000)), "d0"), 1 + r.below(5));
                    kvs.emplace_back(bfq::make_route_key(T.id, (r.below(2) ? "$share/" : "$oshare/") + g + "/" + tf, ""), route_group(members));
                }
            }
            std::sort(kvs.begin(), kvs.end());
            kvs.erase(std::unique(kvs.begin(), kvs.end(), [](auto& a, auto& b) { return a.first == b.first; }), kvs.end());
            TenantOut& o = touts[(size_t) t];
            for (auto& kv : kvs) { o.keys.push(kv.first); o.vals.push(kv.second); }
        }
    };
    {
        std::vector<std::thread> th;
        for (int i = 0; i < std::max(1, nthreads); i++) th.emplace_back(worker);
        for (auto& x : th) x.join();
    }
    // ---- concatenate tenants in key order (tenant key prefix = u16 length, then id bytes)
    std::vector<int64_t> order;
    for (int64_t t = 0; t < pl.n_tenants; t++) if (mine[(size_t) t]) order.push_back(t);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        return bfq::make_tenant_begin_key(tg[(size_t) a].id) < bfq::make_tenant_begin_key(tg[(size_t) b].id);
    });
    std::vector<int32_t> tenant_pos((size_t) pl.n_tenants, -1);
    for (size_t i = 0; i < order.size(); i++) {
        const int64_t t = order[i];
        tenant_pos[(size_t) t] = (int32_t) i;
        W.tenants.push(tg[(size_t) t].id);
        W.keys.append(touts[(size_t) t].keys);
        W.vals.append(touts[(size_t) t].vals);
        W.n_filters += (int64_t) tg[(size_t) t].filters.size();
        touts[(size_t) t] = TenantOut();
    }

    // ---- publish topics (C1-C4) / retained topics + query filters (C5)
    Rng r(mix(seed, 0x70B1C5));
    std::vector<double> cum;
    if (pl.zipf) {
        double acc = 0;
        for (int64_t t = 0; t < pl.n_tenants; t++) { acc += (double) nf[(size_t) t]; cum.push_back(acc); }
    }
    auto pick_tenant = [&]() -> int64_t {
        if (!pl.zipf) return (int64_t) r.below((uint32_t) pl.n_tenants);
        double u = r.uniform() * cum.back();
        return (int64_t) (std::lower_bound(cum.begin(), cum.end(), u) - cum.begin());
    };
    if (cfg != C5) {
        const int64_t n_batch = pl.n_topics * (int64_t) std::max(1, topic_mult);
        for (int64_t i = 0; i < n_batch; i++) {
            const int64_t t = std::min<int64_t>(pick_tenant(), pl.n_tenants - 1);
            // draw the random numbers regardless of ownership so every shard sees the same batch
            const double hit = r.uniform(), pop = r.uniform();
            const uint64_t s1 = r.next();
            if (!mine[(size_t) t]) continue;
            if (hot[(size_t) t] && i % shard_count != shard_index) continue;   // a replicated tenant's topics: round-robin by batch position
            const TenantGen& T = tg[(size_t) t];
            Rng tr(s1);
            std::string topic;
            if (hit < 0.8 && !T.filters.empty()) {
                size_t j = pl.zipf ? (size_t) std::min<double>((double) T.filters.size() - 1, std::pow((double) T.filters.size(), pop) - 1.0)
                                   : (size_t) tr.below((uint32_t) T.filters.size());
                topic = T.render_topic(T.filters[j], &tr);
            } else {
                topic = T.render_topic(T.random_path(tr, 3 + (int) tr.below(6)), nullptr);
            }
            W.topics.push(topic);
            W.topic_tenant.push_back(tenant_pos[(size_t) t]);
        }
    } else {
        // retained topics = the exact base paths; query filters derived from them
        for (int64_t t : order) {
            const TenantGen& T = tg[(size_t) t];
            for (const Filter& f : T.filters) {
                W.topics.push(T.render_topic(f, nullptr));
                W.topic_tenant.push_back(tenant_pos[(size_t) t]);
            }
        }
        for (int64_t i = 0; i < pl.n_cfilters; i++) {
            const int64_t t = (int64_t) r.below((uint32_t) pl.n_tenants);
            const uint64_t s1 = r.next();
            if (!mine[(size_t) t]) continue;
            if (hot[(size_t) t] && i % shard_count != shard_index) continue;   // a replicated tenant's topics: round-robin by batch position
            const TenantGen& T = tg[(size_t) t];
            if (T.filters.empty()) continue;
            Rng tr(s1);
            Filter f = T.filters[tr.below((uint32_t) T.filters.size())];
            if (tr.below(2)) {
                for (int l = 1; l < f.depth; l++) if (tr.below(2)) f.plus_mask |= (uint8_t) (1u << l);
                if (!f.plus_mask) f.plus_mask = (uint8_t) (1u << (1 + tr.below((uint32_t) f.depth - 1)));
            } else {
                f.depth = (uint8_t) (1 + tr.below((uint32_t) f.depth - 1));
                f.hash_tail = 1;
            }
            W.filters.push(T.render_filter(f));
            W.filter_tenant.push_back(tenant_pos[(size_t) t]);
        }
    }
    if (W.topic_tenant.empty()) W.topic_tenant.push_back(0);
    if (W.filter_tenant.empty()) W.filter_tenant.push_back(0);
    for (Blob* b : {&W.keys, &W.vals, &W.tenants, &W.topics, &W.filters})
        if (b->data.empty()) b->data.push_back(0);
    return out;
}

void bfqw_free(bfqw* w) { delete w; }
// info[0..7] = n_routes, n_tenants, n_topics, n_query_filters, n_distinct_filters, key bytes, value bytes, topic bytes
void bfqw_info(bfqw* w, int64_t* info) {
    const Workload& W = w->w;
    info[0] = W.keys.n(); info[1] = W.tenants.n(); info[2] = W.topics.n(); info[3] = W.filters.n();
    info[4] = W.n_filters; info[5] = W.keys.off.back(); info[6] = W.vals.off.back(); info[7] = W.topics.off.back();
}
const uint8_t* bfqw_keys(bfqw* w) { return w->w.keys.data.data(); }
const int64_t* bfqw_key_off(bfqw* w) { return w->w.keys.off.data(); }
const uint8_t* bfqw_vals(bfqw* w) { return w->w.vals.data.data(); }
const int64_t* bfqw_val_off(bfqw* w) { return w->w.vals.off.data(); }
const uint8_t* bfqw_tenants(bfqw* w) { return w->w.tenants.data.data(); }
const int64_t* bfqw_tenant_off(bfqw* w) { return w->w.tenants.off.data(); }
const uint8_t* bfqw_topics(bfqw* w) { return w->w.topics.data.data(); }
const int64_t* bfqw_topic_off(bfqw* w) { return w->w.topics.off.data(); }
const int32_t* bfqw_topic_tenant(bfqw* w) { return w->w.topic_tenant.data(); }
const uint8_t* bfqw_filters(bfqw* w) { return w->w.filters.data.data(); }
const int64_t* bfqw_filter_off(bfqw* w) { return w->w.filters.off.data(); }
const int32_t* bfqw_filter_tenant(bfqw* w) { return w->w.filter_tenant.data(); }

}  // extern "C"
