// index_builder.h — host side of the forward index: staging of raw route KV pairs and the flattening of
// all tenants' filter tries into the hash-table layout of trie_layout.h.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "codec.h"
#include "trie_layout.h"

namespace bfq {

// Sorted KV snapshot stored as two blobs (keys / values) with int64 offsets.
struct KVBlob {
    std::vector<uint8_t> keys, vals;
    std::vector<int64_t> koff{0}, voff{0};
    int64_t n() const { return (int64_t) koff.size() - 1; }
    sv key(int64_t i) const { return sv((const char*) keys.data() + koff[i], (size_t) (koff[i + 1] - koff[i])); }
    sv val(int64_t i) const { return sv((const char*) vals.data() + voff[i], (size_t) (voff[i + 1] - voff[i])); }
    void push(sv k, sv v) {
        keys.insert(keys.end(), k.begin(), k.end());
        vals.insert(vals.end(), v.begin(), v.end());
        koff.push_back((int64_t) keys.size());
        voff.push_back((int64_t) vals.size());
    }
    void clear() { keys.clear(); vals.clear(); koff.assign(1, 0); voff.assign(1, 0); }
};

// Everything the device needs, in host memory, plus build statistics.
struct FlatIndex {
    SlotVec slots;                            // blocked hash table (n_blocks * BLOCK_SLOTS)
    std::vector<uint8_t> tags;                // 16 tag bytes per block
    std::vector<Slot> roots;                  // one record per tenant (key words unused)
    std::vector<uint32_t> segs;               // segment table (pairs), see trie_layout.h
    std::vector<uint8_t> rkind;               // per rank RouteKind
    std::vector<uint32_t> pfx_persistent;     // [n_routes+1] exclusive prefix count of KIND_PERSISTENT
    std::vector<uint32_t> pfx_group;          // [n_routes+1] exclusive prefix count of KIND_GROUP
    std::unordered_map<std::string, uint32_t> tenant_ordinal;
    int64_t n_routes = 0, n_nodes = 0, max_nodes_per_depth = 0, max_tenant_nodes = 0, n_multi = 0, n_cont_chunks = 0;
    uint32_t n_slots = 0, n_blocks = 0;
    int64_t overflowed_blocks = 0;
    int64_t child_hist[5] = {0, 0, 0, 0, 0};   // nodes with 0, 1, 2, 3, >=4 exact children (diagnostic)
};

// Build the flat index from a sorted KV snapshot. Returns false and sets *err on undecodable input.
bool build_flat_index(const KVBlob& kv, FlatIndex* out, std::string* err);

// Host staging area behind bfq_index_load / bfq_index_apply / bfq_index_commit.
class Staging {
public:
    void reset();
    bool load(const uint8_t* keys, const int64_t* koff, const uint8_t* vals, const int64_t* voff, int64_t n, std::string* err);
    void upsert(sv k, sv v);
    void erase(sv k);
    // merge base + delta into a new sorted snapshot (becomes the new base); returns it
    const KVBlob& materialize();
    const KVBlob& base() const { return base_; }
    bool dirty() const { return dirty_; }
    bool has_delta() const { return !delta_.empty(); }
private:
    KVBlob base_;
    std::map<std::string, std::pair<bool, std::string>> delta_;  // key -> (present?, value)
    bool dirty_ = true;
};

}  // namespace bfq
