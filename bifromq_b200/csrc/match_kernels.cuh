// match_kernels.cuh — launch parameters of the forward-match kernels (see match_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "trie_layout.h"

namespace bfq {

// device counters, one block of uint64
enum : int {
    CTR_RANGES = 0,      // cursor into out_ranges (total ranges requested, may exceed capacity)
    CTR_OVERFLOW = 1,    // topics deferred to the tier-2 kernel
    CTR_FLAGGED = 2,     // topics whose matched persistent / group route counts may exceed their caps
    CTR_THROTTLED = 3,   // cursor into the throttled list
    CTR_ROUTES = 4,      // total matched routes (before caps)
    CTR_ERROR = 5,       // tier-2 scratch exhausted (cannot happen with correctly sized scratch)
    CTR_DEFER = 6,       // topics the lane-per-topic tier handed to the warp-per-topic tier
    CTR_CHUNK = 7,       // tier-0 work distribution: next unclaimed topic index
    CTR_COUNT = 8,
};

constexpr uint32_t SPAN_FLAGGED = 0x80000000u;   // in span_count: caps must be applied to this topic
constexpr uint32_t SPAN_OVERFLOW = 0x40000000u;  // in span_count: deferred to tier 2 (never visible to callers)
constexpr uint32_t SPAN_COUNT_MASK = 0x3FFFFFFFu;

struct MatchParams {
    // index snapshot
    const Slot* slots;              // blocked edge table (trie_layout.h)
    const uint4* tags;              // one 16-byte tag word per block
    const Slot* roots;
    uint32_t n_blocks;
    // topic batch
    const uint8_t* topics;          // blob
    const int64_t* topic_off;       // [n+1]
    const int32_t* topic_tenant;    // [n] index into the per-call tenant tables
    const int32_t* tenant_root;     // [n_tenants] root ordinal or -1 (tenant has no routes)
    const int32_t* max_pfanout;     // [n_tenants]
    const int32_t* max_gfanout;     // [n_tenants]
    int32_t n_tenants;
    int64_t n_topics;
    // tier 0: optional processing order (topic indices grouped by tenant and leading levels, see launch_order); nullptr =>
    // 0..n_topics. Only the order in which lanes pick topics changes; every output stays indexed by topic.
    const uint32_t* order;
    // tiers 1/2: list of topic indices to process (nullptr => all topics 0..n_topics)
    const uint32_t* work_list;
    int64_t n_work;
    // outputs
    uint32_t* span_begin;           // [n]
    uint32_t* span_count;           // [n]  count | SPAN_FLAGGED
    uint32_t* route_count;          // [n]
    uint2* ranges;                  // [ranges_cap] {first, count | RANGE_MULTI}
    uint64_t ranges_cap;
    uint64_t dyn_base;              // ranges[0, dyn_base) = tier-0 inline slots (INLINE_RANGES per topic); the cursor
                                    // CTR_RANGES allocates from ranges[dyn_base, ranges_cap) for tiers 1 and 2
    uint32_t* overflow_list;        // [n] topic indices deferred to tier 2
    uint32_t* defer_list;           // [n] topic indices deferred from tier 0 to tier 1
    uint32_t* flagged_list;         // [n] topic indices needing caps
    unsigned long long* counters;   // [CTR_COUNT]
    // tier-2 scratch (global memory frontier / range staging), per warp
    uint2* scratch;
    uint64_t scratch_frontier_cap;  // entries per frontier buffer
    uint64_t scratch_ranges_cap;    // entries of range staging
};

struct CapsParams {
    const uint32_t* flagged_list;
    int64_t n_flagged;
    const int32_t* topic_tenant;
    const int32_t* max_pfanout;
    const int32_t* max_gfanout;
    const uint32_t* span_begin;
    const uint32_t* span_count;
    const uint2* ranges;
    const uint32_t* segs;
    const uint8_t* rkind;
    const uint32_t* pfx_persistent;
    const uint32_t* pfx_group;
    uint3* throttled;               // {topic, rank, kind}
    uint32_t topic_base;            // added to the (sub-batch relative) topic index
    uint64_t throttled_cap;
    uint32_t* kept_count;           // [n] (optional) routes surviving per flagged topic
    unsigned long long* counters;
};

struct ExpandParams {
    int64_t n_topics;
    const uint32_t* span_begin;
    const uint32_t* span_count;     // with SPAN_FLAGGED bits
    const uint32_t* route_count;    // matched routes before caps
    const uint32_t* kept_count;     // surviving routes of cap-flagged topics (written by the caps kernel)
    const uint2* ranges;            // sparse
    const uint32_t* segs;
    unsigned long long* counts;     // scratch [n+1]
    int64_t* offsets;               // out [n+1]
    int64_t* ranks;                 // out
    int64_t rank_cap;
    // caps inputs for the flagged topics
    const uint32_t* flagged_list;
    int64_t n_flagged;
    const int32_t* topic_tenant;
    const int32_t* max_pfanout;
    const int32_t* max_gfanout;
    const uint8_t* rkind;
    const uint32_t* pfx_persistent;
    const uint32_t* pfx_group;
};
// device CSR of the surviving routes: phase 1 = per-topic counts + exclusive scan into offsets (offsets[n] = total),
// phase 2 = write the ranks (unordered within a topic). tmp as for launch_compact.
cudaError_t launch_expand(const ExpandParams& p, void* d_scan_tmp, size_t* tmp_bytes, cudaStream_t stream, int phase);

// Locality ordering for tier 0: 32-bit key = tenant index | hashes of the first three levels, radix-sorted (cub) with the
// topic index as payload -> order_out[n]. Topics that walk the same top of the trie are then matched by neighbouring
// lanes at the same time: their node reads hit L1/L2 instead of being ~18 random DRAM accesses per topic.
// keys/vals: scratch of 2 * n uint32 each; d_tmp / tmp_bytes as for launch_compact (query with d_tmp == nullptr).
struct OrderParams {
    int64_t n_topics;
    const uint8_t* topics;
    const int64_t* topic_off;
    const int32_t* topic_tenant;
    int32_t n_tenants;
    uint32_t* keys;                 // [2n]
    uint32_t* vals;                 // [2n]; the sorted order ends up in vals + n
};
cudaError_t launch_order(const OrderParams& q, void* d_tmp, size_t* tmp_bytes, cudaStream_t stream);

// tier 0: one LANE per topic (DFS, bounded smem); tier 1: one WARP per topic; tier 2: warp per topic, global scratch
void launch_match_lanes(const MatchParams& p, cudaStream_t stream);
void launch_match(const MatchParams& p, bool tier2, int n_warps_tier2, cudaStream_t stream);
void launch_caps(const CapsParams& p, cudaStream_t stream);

constexpr uint32_t INLINE_RANGES = 12;   // tier 0 writes topic t's ranges at ranges[t * INLINE_RANGES ...)

// Compaction for the host path: gathers the sparse (inline + dynamic) ranges into one dense array in topic order.
// d_scan_tmp / tmp_bytes: scratch for the exclusive scan (query the size with d_scan_tmp == nullptr).
struct CompactParams {
    int64_t n_topics;
    const uint32_t* span_begin;     // in
    const uint32_t* span_count;     // in (flag bits allowed)
    const uint2* ranges;            // in
    uint32_t* counts;               // scratch [n]
    uint32_t* new_begin;            // out [n]
    uint2* ranges_out;              // out [total]
    uint64_t ranges_out_cap;
    uint32_t out_base;              // index of ranges_out[0] in the concatenated result (added to new_begin in phase 2)
    unsigned long long* total_out;  // device scalar: total number of ranges (phase 1)
};
// phase 1: counts + exclusive scan + total; phase 2: gather (after the caller has read the total and placed ranges_out)
cudaError_t launch_compact(const CompactParams& p, void* d_scan_tmp, size_t* tmp_bytes, cudaStream_t stream, int phase);
int match_kernel_smem_bytes();

}  // namespace bfq
