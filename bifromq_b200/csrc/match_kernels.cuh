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
    CTR_NLEAD = 8,       // number of distinct (tenant, topic) pairs of the batch = length of the locality order
    CTR_FLAGGED2 = 9,    // flagged topics already handled by an earlier caps pass (the rare tier-2 path runs a second one)
    CTR_COUNT = 12,
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
    const unsigned long long* order_count;   // device scalar: entries of `order` (the batch's distinct topics); with order only
    int32_t max_ctas_per_sm;        // tier 0: cap of resident CTAs per SM (0 = as many as fit); one slot less leaves room for a
                                    // concurrently running exchange (NCCL) kernel
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
    int64_t n_flagged;              // < 0: flagged_list[counters[CTR_FLAGGED2] ... counters[CTR_FLAGGED]) (counts read on the device)
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

// Locality ordering + de-duplication for tier 0, all own kernels (no library sort):
//   prep     one thread per topic: 64-bit hash of (tenant, topic bytes) -> insert into an open-addressing table; the first
//            inserter of a (tenant, topic) pair is its LEADER, later identical ones (verified byte by byte) are followers that
//            only remember their leader. Leaders get an order key = tenant index | hashes of the level-0 / 0..1 / 0..2 prefixes
//            and count themselves into a bucket histogram (bucket = the key's leading hist_bits).
//   scan     exclusive prefix sum of the histogram (block-local scans; the last block to finish scans the block totals)
//   scatter  leaders -> order[bucket base + atomic cursor]: a counting sort, unstable inside a bucket (only grouping matters)
// Tier 0 then matches order[0 .. n_leaders) — topics that walk the same top of the trie are matched by neighbouring lanes at
// the same time — and finalize_kernel copies each follower's span from its leader (spans are indices into the sparse range
// array, so no range is copied). The reference never sees duplicates (matchAll takes a Set<String>,
// DW/cache/ITenantRouteMatcher.java:28-38; DW/cache/TenantRouteCache.java:100-139 serves repeats from its cache).
struct OrderParams {
    int64_t n_topics;
    const uint8_t* topics;
    const int64_t* topic_off;
    const int32_t* topic_tenant;
    int32_t n_tenants;
    uint32_t* keys;                 // [n] bucket of each leader
    uint32_t* leader;               // [n] out: i for a leader, else the index of the identical topic that leads
    uint32_t* order;                // [n] out: the leaders, grouped by bucket
    unsigned long long* hash_tab;   // [hash_mask + 1] pre-set to 0xFF bytes
    uint32_t hash_mask;
    uint32_t* hist;                 // [n_buckets] zeroed; n_buckets = 2^hist_bits, a multiple of 4096
    uint32_t* blk_tot;              // [n_buckets / 4096]
    uint32_t* blk_pfx;              // [n_buckets / 4096]
    uint32_t* ticket;               // zeroed
    int hist_bits;
    int dedup;                      // 0: every topic is its own leader
    unsigned long long* counters;   // CTR_NLEAD is written by the scan
};
size_t order_hist_buckets(int64_t n_topics, int32_t n_tenants);   // histogram entries launch_order will use
uint32_t order_hash_entries(int64_t n_topics);                    // dedup table entries (power of two)
cudaError_t launch_order(const OrderParams& q, cudaStream_t stream);

// followers copy their leader's span (and join the flagged list if it needs caps). second_pass: only the followers whose
// span still carries the tier-2 marker (their leader was finished by tier 2 after the first pass).
struct FinalizeParams {
    int64_t n_topics;
    const uint32_t* leader;
    uint32_t* span_begin;
    uint32_t* span_count;
    uint32_t* route_count;
    uint32_t* flagged_list;
    unsigned long long* counters;
    int second_pass;
};
void launch_finalize(const FinalizeParams& p, cudaStream_t stream);

// tier 0: one LANE per topic (DFS, bounded smem); tier 1: one WARP per topic; tier 2: warp per topic, global scratch
void launch_match_lanes(const MatchParams& p, cudaStream_t stream);
void launch_match(const MatchParams& p, bool tier2, int n_warps_tier2, cudaStream_t stream);
void launch_caps(const CapsParams& p, cudaStream_t stream);

constexpr uint32_t INLINE_RANGES = 12;   // tier 0 writes topic t's ranges at ranges[t * INLINE_RANGES ...)
constexpr uint32_t SPILL_RANGES = 64;    // ... and moves a topic with more of them, once, to a block of this many ranges

// Compaction for the host path: gathers the sparse (inline + dynamic) ranges into one dense array in topic order.
// d_scan_tmp / tmp_bytes: scratch for the exclusive scan (query the size with d_scan_tmp == nullptr).
struct CompactParams {
    int64_t n_topics;
    const uint32_t* span_begin;     // in
    const uint32_t* span_count;     // in (flag bits allowed)
    const uint2* ranges;            // in
    const uint32_t* leader;         // optional [n]: leader[i] != i marks a repeat of topic leader[i] (shares its dense span)
    uint32_t* counts;               // scratch [n]
    uint32_t* new_begin;            // scratch [n]: exclusive scan of counts
    uint32_t* final_begin;          // out [n] (phase 2): first range of topic i in the concatenated result
    uint32_t* final_count;          // out [n] (phase 2): its number of ranges
    uint2* ranges_out;              // out [total]
    uint64_t ranges_out_cap;
    uint32_t out_base;              // index of ranges_out[0] in the concatenated result (added to new_begin in phase 2)
    unsigned long long* total_out;  // device scalar: total number of ranges (phase 1)
};
// phase 1: counts + exclusive scan + total; phase 2: gather (after the caller has read the total and placed ranges_out)
cudaError_t launch_compact(const CompactParams& p, void* d_scan_tmp, size_t* tmp_bytes, cudaStream_t stream, int phase);
int match_kernel_smem_bytes();

// bfq_index_commit's delta path: the records of the tenants BEHIND a tenant that grew or shrank carry dense KV ranks, so
// their own / '#' first-rank words move by the difference. One streaming pass over the listed slot regions.
struct RankShiftRegion {
    uint64_t base, len;     // slots [base, base + len)
    int32_t delta;
};
void launch_rank_shift(Slot* slots, const RankShiftRegion* d_regions, int n_regions, cudaStream_t stream);
// dst[i] = src[i] + add (the prefix-count arrays of those tenants, copied to their shifted position)
void launch_copy_add(uint32_t* dst, const uint32_t* src, int64_t n, uint32_t add, cudaStream_t stream);

}  // namespace bfq
