// fanout.h — fan-out expansion (fanout.cu): parameters of the device pass and the host-side deliverer interning.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "codec.h"
#include "index_builder.h"

namespace bfq {

constexpr uint32_t FO_GROUP_BIT = 0x80000000u;   // in rdeliv[rank]: the route is a shared subscription, low bits = group index

struct FanoutParams {
    // the surviving routes of a completed match as a device CSR (bfq_expand_device)
    int64_t n_topics;
    const int64_t* offsets;          // [n_topics + 1]
    int64_t n_pairs;                 // offsets[n_topics]
    const int64_t* ranks;            // [n_pairs]
    // per-snapshot tables
    const uint32_t* rdeliv;          // [n_routes] deliverer id, or FO_GROUP_BIT | group index
    const uint32_t* gmem_off;        // [n_groups + 1] members of group g: gmem_deliv[gmem_off[g] .. gmem_off[g + 1])
    const uint32_t* gmem_deliv;      // deliverer id of every member
    const uint8_t* gordered;         // [n_groups] 1 = $oshare (left to the host)
    uint32_t n_deliverers;           // ids are [0, n_deliverers); the last one is the reserved "ordered share" id
    // scratch
    uint32_t* tile_counts;           // [n_deliverers * n_tiles]
    uint32_t* tile_base;             // [n_deliverers * n_tiles]
    // outputs
    long long* pack_offsets;         // [n_deliverers + 1]
    uint32_t* pack_topic;            // [n_pairs]
    uint32_t* pack_rank;             // [n_pairs]
    uint32_t* pack_member;           // [n_pairs] member index of a shared subscription, 0xFFFFFFFF otherwise
};
// d_tmp == nullptr: query the scan scratch size
cudaError_t launch_fanout(const FanoutParams& p, void* d_tmp, size_t* tmp_bytes, cudaStream_t stream);
uint32_t fanout_max_deliverers();
int64_t fanout_tile();

// (subBrokerId, delivererKey) -> dense id, append-only and shared by every snapshot of an index (ids stay valid across commits)
struct DelivererTable {
    std::mutex mu;
    std::unordered_map<std::string, uint32_t> ids;
    std::vector<std::pair<int32_t, std::string>> list;
    uint32_t intern(int32_t broker, sv key);
};

// one tenant's routes resolved to deliverer ids (built once per tenant KV blob, reused by the snapshots that share the blob)
struct TenantFan {
    std::vector<uint32_t> rdeliv;        // per local rank; FO_GROUP_BIT | tenant-local group index for shared subscriptions
    std::vector<uint32_t> gmem_off;      // tenant-local
    std::vector<uint32_t> gmem_deliv;
    std::vector<uint8_t> gordered;
};
bool build_tenant_fan(const KVBlob& kv, DelivererTable* table, TenantFan* out, std::string* err);

}  // namespace bfq
