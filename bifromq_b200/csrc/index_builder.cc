// index_builder.cc — see index_builder.h.
#include "index_builder.h"

#include <algorithm>
#include <cstring>

namespace bfq {

// ------------------------------------------------------------------------------------------------ staging
void Staging::reset() {
    base_.clear();
    delta_.clear();
    dirty_ = true;
}

bool Staging::load(const uint8_t* keys, const int64_t* koff, const uint8_t* vals, const int64_t* voff, int64_t n,
                   std::string* err) {
    if (n <= 0) return true;
    // fast path: appending a sorted run after the current base
    bool sorted_append = delta_.empty();
    if (sorted_append) {
        sv prev = base_.n() ? base_.key(base_.n() - 1) : sv();
        bool have_prev = base_.n() > 0;
        for (int64_t i = 0; i < n && sorted_append; i++) {
            sv k((const char*) keys + koff[i], (size_t) (koff[i + 1] - koff[i]));
            if (have_prev && !(prev < k)) sorted_append = false;
            prev = k;
            have_prev = true;
        }
    }
    if (sorted_append) {
        const int64_t kb = koff[0], vb = voff[0];
        const size_t k0 = base_.keys.size(), v0 = base_.vals.size();
        base_.keys.insert(base_.keys.end(), keys + kb, keys + koff[n]);
        base_.vals.insert(base_.vals.end(), vals + vb, vals + voff[n]);
        base_.koff.reserve(base_.koff.size() + n);
        base_.voff.reserve(base_.voff.size() + n);
        for (int64_t i = 1; i <= n; i++) {
            base_.koff.push_back((int64_t) k0 + koff[i] - kb);
            base_.voff.push_back((int64_t) v0 + voff[i] - vb);
        }
    } else {
        if (err) *err = "bfq_index_load: keys must be strictly ascending (and follow the keys already staged)";
        return false;
    }
    dirty_ = true;
    return true;
}

void Staging::upsert(sv k, sv v) {
    delta_[std::string(k)] = {true, std::string(v)};
    dirty_ = true;
}
void Staging::erase(sv k) {
    delta_[std::string(k)] = {false, std::string()};
    dirty_ = true;
}

const KVBlob& Staging::materialize() {
    if (!delta_.empty()) {
        KVBlob merged;
        merged.keys.reserve(base_.keys.size());
        merged.vals.reserve(base_.vals.size());
        int64_t i = 0;
        const int64_t n = base_.n();
        auto d = delta_.begin();
        while (i < n || d != delta_.end()) {
            int c;
            if (i >= n) c = 1;
            else if (d == delta_.end()) c = -1;
            else c = base_.key(i).compare(sv(d->first));
            if (c < 0) {
                merged.push(base_.key(i), base_.val(i));
                i++;
            } else {
                if (d->second.first) merged.push(d->first, d->second.second);
                if (c == 0) i++;
                ++d;
            }
        }
        base_ = std::move(merged);
        delta_.clear();
    }
    dirty_ = false;
    return base_;
}

// ------------------------------------------------------------------------------------------------ builder
namespace {

struct Target {
    uint32_t first = 0, count = 0, total = 0, pc = 0, gc = 0;
    int32_t multi = -1;
};
struct BNode {
    uint32_t parent = NONE;   // node index of the parent (NONE for tenant roots)
    uint32_t lenw = 0;
    uint32_t tok[TOKEN_WORDS] = {0, 0, 0, 0, 0, 0};
    uint32_t plus = NONE;     // node index of the '+' child
    uint32_t flags = 0;
    uint32_t root_ordinal = NONE;
    Target own, hash;
};

inline void make_tok(sv chunk, uint32_t* tok) {
    for (uint32_t k = 0; k < TOKEN_WORDS; k++) tok[k] = 0;
    for (size_t j = 0; j < chunk.size(); j++) tok[j >> 2] |= (uint32_t) (uint8_t) chunk[j] << (8 * (j & 3));
}

class Builder {
public:
    std::vector<BNode> nodes;
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> multi_lists;
    int64_t n_cont = 0;

    Builder() { table_.assign(1u << 16, NONE); }

    uint32_t new_root(uint32_t ordinal) {
        nodes.emplace_back();
        nodes.back().root_ordinal = ordinal;
        return (uint32_t) nodes.size() - 1;
    }
    // find-or-create the child of `parent` along one edge key
    uint32_t edge(uint32_t parent, uint32_t lenw, const uint32_t* tok, bool* created) {
        if ((nodes.size() + 1) * 2 > table_.size()) grow();
        uint64_t h = fmix64(token_hash(lenw, tok) + (uint64_t) parent * 0xC2B2AE3D27D4EB4Full);
        size_t mask = table_.size() - 1, s = (size_t) h & mask;
        while (true) {
            uint32_t idx = table_[s];
            if (idx == NONE) break;
            const BNode& c = nodes[idx];
            if (c.parent == parent && c.lenw == lenw && memcmp(c.tok, tok, sizeof(c.tok)) == 0) {
                *created = false;
                return idx;
            }
            s = (s + 1) & mask;
        }
        nodes.emplace_back();
        BNode& c = nodes.back();
        c.parent = parent;
        c.lenw = lenw;
        memcpy(c.tok, tok, sizeof(c.tok));
        uint32_t idx = (uint32_t) nodes.size() - 1;
        table_[s] = idx;
        *created = true;
        return idx;
    }
    // descend one filter level (not '#'); created_real reports whether the level's final node is new
    uint32_t descend(uint32_t node, sv level, bool* created_real) {
        uint32_t tok[TOKEN_WORDS];
        bool created;
        if (level.size() == 1 && level[0] == '+') {
            if (nodes[node].plus != NONE) {
                *created_real = false;
                return nodes[node].plus;
            }
            make_tok(sv(), tok);
            uint32_t c = edge(node, LEN_PLUS, tok, &created);
            nodes[node].plus = c;
            *created_real = created;
            return c;
        }
        nodes[node].flags |= FLAG_HAS_EXACT;
        size_t off = 0;
        uint32_t j = 0;
        while (level.size() - off > TOKEN_BYTES) {  // continuation chunks of a long token
            make_tok(level.substr(off, TOKEN_BYTES), tok);
            node = edge(node, LEN_CONT | j, tok, &created);
            if (created) n_cont++;
            nodes[node].flags |= FLAG_HAS_EXACT;
            off += TOKEN_BYTES;
            j++;
        }
        make_tok(level.substr(off), tok);
        uint32_t c = edge(node, (uint32_t) level.size(), tok, created_real);
        return c;
    }
    void add_route(Target& t, uint32_t rank, RouteKind kind) {
        if (t.total == 0) {
            t.first = rank;
            t.count = 1;
        } else if (t.first + t.count == rank) {
            t.count++;
        } else {
            if (t.multi < 0) {
                t.multi = (int32_t) multi_lists.size();
                multi_lists.emplace_back();
            }
            multi_lists[t.multi].push_back({t.first, t.count});
            t.first = rank;
            t.count = 1;
        }
        t.total++;
        if (kind == KIND_PERSISTENT) t.pc++;
        else if (kind == KIND_GROUP) t.gc++;
    }

private:
    std::vector<uint32_t> table_;
    void grow() {
        std::vector<uint32_t> nt(table_.size() * 2, NONE);
        size_t mask = nt.size() - 1;
        for (uint32_t idx = 0; idx < nodes.size(); idx++) {
            const BNode& c = nodes[idx];
            if (c.parent == NONE) continue;
            uint64_t h = fmix64(token_hash(c.lenw, c.tok) + (uint64_t) c.parent * 0xC2B2AE3D27D4EB4Full);
            size_t s = (size_t) h & mask;
            while (nt[s] != NONE) s = (s + 1) & mask;
            nt[s] = idx;
        }
        table_.swap(nt);
    }
};

inline uint32_t sat16(uint32_t v) { return v > 0xFFFFu ? 0xFFFFu : v; }

}  // namespace

bool build_flat_index(const KVBlob& kv, FlatIndex* out, std::string* err) {
    *out = FlatIndex();
    const int64_t n = kv.n();
    if (n >= (int64_t) 0x7FFFFFFF) {
        if (err) *err = "too many routes for 31-bit ranks";
        return false;
    }
    out->n_routes = n;
    out->rkind.resize((size_t) n);
    out->pfx_persistent.resize((size_t) n + 1);
    out->pfx_group.resize((size_t) n + 1);
    Builder b;
    b.nodes.reserve((size_t) std::min<int64_t>(n * 2 + 16, 1 << 28));

    sv cur_tenant;
    bool have_tenant = false;
    uint32_t cur_root = NONE;
    std::vector<sv> path_levels;          // levels of the previous key of this tenant
    std::vector<uint32_t> path_nodes;     // node reached after consuming path_levels[i]
    std::vector<uint32_t> depth_count;    // real nodes per depth of the current tenant
    int64_t tenant_nodes = 0;
    std::vector<sv> levels;
    auto close_tenant = [&]() {
        for (uint32_t c : depth_count) out->max_nodes_per_depth = std::max<int64_t>(out->max_nodes_per_depth, c);
        out->max_tenant_nodes = std::max(out->max_tenant_nodes, tenant_nodes);
        depth_count.clear();
        tenant_nodes = 0;
    };

    uint32_t pp = 0, pg = 0;
    for (int64_t r = 0; r < n; r++) {
        DecodedKey d;
        if (!decode_route_key(kv.key(r), &d)) {
            if (err) *err = "undecodable route key at rank " + std::to_string(r);
            return false;
        }
        out->rkind[(size_t) r] = (uint8_t) d.kind;
        out->pfx_persistent[(size_t) r] = pp;
        out->pfx_group[(size_t) r] = pg;
        if (d.kind == KIND_PERSISTENT) pp++;
        else if (d.kind == KIND_GROUP) pg++;

        if (!have_tenant || d.tenant != cur_tenant) {
            if (have_tenant) close_tenant();
            cur_tenant = d.tenant;
            have_tenant = true;
            auto it = out->tenant_ordinal.find(std::string(d.tenant));
            if (it == out->tenant_ordinal.end()) {
                uint32_t ord = (uint32_t) out->tenant_ordinal.size();
                out->tenant_ordinal.emplace(std::string(d.tenant), ord);
                cur_root = b.new_root(ord);
            } else {
                // cannot happen for sorted keys (the tenant id is the key prefix)
                if (err) *err = "tenant keys are not contiguous";
                return false;
            }
            path_levels.clear();
            path_nodes.clear();
        }
        levels.clear();
        for_each_level(d.escaped_filter, '\0', [&](sv l) { levels.push_back(l); });
        const bool multi_wild = levels.back().size() == 1 && levels.back()[0] == '#';
        const size_t walk = multi_wild ? levels.size() - 1 : levels.size();
        // reuse the longest common prefix with the previous key's path
        size_t k = 0;
        while (k < walk && k < path_levels.size() && path_levels[k] == levels[k]) k++;
        path_levels.resize(k);
        path_nodes.resize(k);
        uint32_t node = k ? path_nodes[k - 1] : cur_root;
        for (; k < walk; k++) {
            bool created = false;
            node = b.descend(node, levels[k], &created);
            if (created) {
                if (depth_count.size() <= k) depth_count.resize(k + 1, 0);
                depth_count[k]++;
                tenant_nodes++;
            }
            path_levels.push_back(levels[k]);
            path_nodes.push_back(node);
        }
        b.add_route(multi_wild ? b.nodes[node].hash : b.nodes[node].own, (uint32_t) r, d.kind);
    }
    if (have_tenant) close_tenant();
    out->pfx_persistent[(size_t) n] = pp;
    out->pfx_group[(size_t) n] = pg;
    out->n_cont_chunks = b.n_cont;

    // ---- flatten. Exact children of a node go either into a private perfect-hashed array (small fan-out) or into the
    // global tag table (big fan-out); '+' children get a slot of their own. Nodes are laid out in BFS order so a parent's
    // id is known before its children are keyed, and siblings are adjacent in memory.
    const size_t total_nodes = b.nodes.size();
    const size_t n_roots = out->tenant_ordinal.size();
    out->n_nodes = (int64_t) total_nodes;
    // group the exact children by parent (counting sort)
    std::vector<uint32_t> child_off(total_nodes + 1, 0);
    for (size_t i = 0; i < total_nodes; i++) {
        const BNode& nd = b.nodes[i];
        if (nd.parent != NONE && nd.lenw != LEN_PLUS) child_off[nd.parent + 1]++;
    }
    for (size_t i = 0; i < total_nodes; i++) child_off[i + 1] += child_off[i];
    std::vector<uint32_t> child_list(child_off[total_nodes]);
    {
        std::vector<uint32_t> fill(child_off.begin(), child_off.end() - 1);
        for (size_t i = 0; i < total_nodes; i++) {
            const BNode& nd = b.nodes[i];
            if (nd.parent != NONE && nd.lenw != LEN_PLUS) child_list[fill[nd.parent]++] = (uint32_t) i;
        }
    }
    // sizing pass: per parent choose {single, perfect hash of 2^lg slots with a seed, big}
    struct ChildPlan { uint8_t lg; uint8_t big; uint16_t seed; };
    std::vector<ChildPlan> plan(total_nodes, ChildPlan{0, 0, 0});
    uint64_t n_big_edges = 0, csr_slots = 0;
    {
        std::vector<uint32_t> t32, seen;
        for (size_t i = 0; i < total_nodes; i++) {
            const uint32_t c = child_off[i + 1] - child_off[i];
            if (b.nodes[i].plus != NONE) csr_slots++;
            if (c == 0) continue;
            ChildPlan& pl = plan[i];
            t32.clear();
            for (uint32_t j = child_off[i]; j < child_off[i + 1]; j++) {
                const BNode& ch = b.nodes[child_list[j]];
                t32.push_back(fold32(token_hash(ch.lenw, ch.tok)));
            }
            bool big = c > SMALL_FANOUT_MAX;
            if (!big && c == 1) {
                pl.lg = 0;
                pl.seed = (uint16_t) (t32[0] & 0xFFFFu);
            } else if (!big) {
                std::vector<uint32_t> sorted(t32);
                std::sort(sorted.begin(), sorted.end());
                if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) big = true;   // 32-bit fold collision
                uint32_t lg = 1;
                while ((1u << lg) < c) lg++;
                if (c > 4) lg++;
                bool found = false;
                for (; !big && !found && lg <= 8; lg++) {
                    for (uint32_t seed = 0; seed < 65536 && !found; seed++) {
                        uint64_t mask_lo = 0, mask_hi = 0, mask_2 = 0, mask_3 = 0;   // up to 256 positions
                        bool ok = true;
                        for (uint32_t v : t32) {
                            const uint32_t idx = child_index(v, seed, lg);
                            uint64_t& m = idx < 64 ? mask_lo : (idx < 128 ? mask_hi : (idx < 192 ? mask_2 : mask_3));
                            const uint64_t bit = 1ull << (idx & 63);
                            if (m & bit) { ok = false; break; }
                            m |= bit;
                        }
                        if (ok) {
                            found = true;
                            pl.lg = (uint8_t) lg;
                            pl.seed = (uint16_t) seed;
                        }
                    }
                    if (found) break;
                }
                if (!found) big = true;
            }
            if (big) {
                pl.big = 1;
                n_big_edges += c;
            } else {
                csr_slots += 1ull << pl.lg;
            }
        }
    }
    EdgeTable table;
    table.init(n_big_edges);
    const uint64_t csr_base = (uint64_t) table.n_blocks * BLOCK_SLOTS;
    if (csr_base + csr_slots >= 0x7FFFFFF0ull) {
        if (err) *err = "index too large for 31-bit slot ids";
        return false;
    }
    {
        Slot empty;
        memset(empty.w, 0, sizeof(empty.w));
        empty.w[W_PARENT] = EMPTY_PARENT;
        table.slots.resize((size_t) (csr_base + csr_slots), empty);
    }
    out->roots.assign(n_roots, Slot());
    std::vector<uint32_t> id_of(total_nodes, NONE), child_base(total_nodes, 0);
    // BFS placement
    {
        std::vector<uint32_t> order;
        order.reserve(total_nodes);
        for (size_t i = 0; i < total_nodes; i++)
            if (b.nodes[i].parent == NONE) {
                id_of[i] = ROOT_BASE + b.nodes[i].root_ordinal;
                order.push_back((uint32_t) i);
            }
        uint64_t cursor = csr_base;
        for (size_t qi = 0; qi < order.size(); qi++) {
            const uint32_t pi = order[qi];
            const BNode& P = b.nodes[pi];
            if (P.plus != NONE) {
                id_of[P.plus] = (uint32_t) cursor++;
                order.push_back(P.plus);
            }
            const ChildPlan& pl = plan[pi];
            const uint32_t c0 = child_off[pi], c1 = child_off[pi + 1];
            if (c1 == c0) continue;
            if (pl.big) {
                for (uint32_t j = c0; j < c1; j++) {
                    const BNode& ch = b.nodes[child_list[j]];
                    id_of[child_list[j]] = table.place(id_of[pi], ch.lenw, ch.tok);
                    order.push_back(child_list[j]);
                }
            } else {
                child_base[pi] = (uint32_t) cursor;
                for (uint32_t j = c0; j < c1; j++) {
                    const BNode& ch = b.nodes[child_list[j]];
                    const uint32_t idx = pl.lg ? child_index(fold32(token_hash(ch.lenw, ch.tok)), pl.seed, pl.lg) : 0u;
                    id_of[child_list[j]] = (uint32_t) (cursor + idx);
                    order.push_back(child_list[j]);
                }
                cursor += 1ull << pl.lg;
            }
        }
        if (order.size() != total_nodes || cursor != csr_base + csr_slots) {
            if (err) *err = "internal error: BFS placement did not cover the trie";
            return false;
        }
    }
    out->segs.clear();
    auto emit_target = [&](Target& t, uint32_t* first, uint32_t* count, uint32_t multi_flag, uint32_t* flags) {
        if (t.multi >= 0) {
            auto& lst = b.multi_lists[t.multi];
            lst.push_back({t.first, t.count});
            *first = (uint32_t) (out->segs.size() / 2);
            *count = t.total;
            *flags |= multi_flag;
            out->segs.push_back((uint32_t) lst.size());
            out->segs.push_back(t.total);
            for (auto& p : lst) {
                out->segs.push_back(p.first);
                out->segs.push_back(p.second);
            }
            out->n_multi++;
        } else {
            *first = t.first;
            *count = t.total;
        }
    };
    auto sat8 = [](uint32_t v) { return v > 255u ? 255u : v; };
    for (size_t i = 0; i < total_nodes; i++) {
        BNode& nd = b.nodes[i];
        Slot* rec;
        if (nd.parent == NONE) {
            rec = &out->roots[nd.root_ordinal];
            memset(rec->w, 0, sizeof(rec->w));
            rec->w[W_PARENT] = NONE;
        } else {
            rec = &table.slots[id_of[i]];
            rec->w[W_PARENT] = id_of[nd.parent];
            rec->w[W_LEN] = nd.lenw;
            for (uint32_t k = 0; k < TOKEN_WORDS; k++) rec->w[W_TOK + k] = nd.tok[k];
        }
        uint32_t flags = nd.flags & FLAG_HAS_EXACT;
        emit_target(nd.own, &rec->w[W_OWN_FIRST], &rec->w[W_OWN_COUNT], FLAG_OWN_MULTI, &flags);
        emit_target(nd.hash, &rec->w[W_HASH_FIRST], &rec->w[W_HASH_COUNT], FLAG_HASH_MULTI, &flags);
        rec->w[W_CAPS] = sat8(nd.own.pc) | (sat8(nd.own.gc) << 8) | (sat8(nd.hash.pc) << 16) | (sat8(nd.hash.gc) << 24);
        if (plan[i].big) flags |= FLAG_BIG;
        rec->w[W_META] = meta_pack(flags, plan[i].lg, plan[i].seed);
        rec->w[W_CHILD_BASE] = child_base[i];
        rec->w[W_PLUS] = nd.plus == NONE ? NONE : id_of[nd.plus];
    }
    if (out->segs.empty()) out->segs.assign(2, 0);
    {
        std::vector<uint32_t> nchild(total_nodes, 0);
        for (size_t i = 0; i < total_nodes; i++)
            if (b.nodes[i].parent != NONE && b.nodes[i].lenw != LEN_PLUS) nchild[b.nodes[i].parent]++;
        for (size_t i = 0; i < total_nodes; i++) out->child_hist[std::min<uint32_t>(nchild[i], 4)]++;
    }
    out->n_blocks = table.n_blocks;
    out->n_slots = (uint32_t) table.slots.size();
    out->overflowed_blocks = table.overflowed_blocks;
    out->slots = std::move(table.slots);
    out->tags = std::move(table.tags);
    return true;
}

}  // namespace bfq
