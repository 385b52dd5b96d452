/* bfq_gpumatch.h — C-ABI of the B200 topic-filter matcher (libbfq_gpumatch.so).
 *
 * This is the drop-in boundary: everything a JNI shim needs to put the CUDA matcher behind
 * apache/bifromq's dist-worker / retain-store co-processors without touching their Java API.
 * Nothing native exists in the reference today (it is 100% Java), so each entry point cites the
 * Java interface or call site it replaces. Paths are relative to the reference root; DW/ =
 * bifromq-dist/bifromq-dist-worker/src/main/java/org/apache/bifromq/dist/worker/, DWS/ =
 * bifromq-dist/bifromq-dist-worker-schema/src/main/java/org/apache/bifromq/dist/worker/schema/,
 * RS/ = bifromq-retain/bifromq-retain-store/src/main/java/org/apache/bifromq/retain/store/,
 * U/ = bifromq-util/src/main/java/org/apache/bifromq/util/.
 *
 * Conventions: every function returns 0 (BFQ_OK) or a negative BFQ_E_* code; bfq_last_error()
 * gives the text. All buffers are caller-owned plain memory (host unless the name says device);
 * strings are (blob, int64 offsets[n+1]) pairs, never NUL-terminated. There is no CPU fallback: every
 * match runs on the GPU and the library fails (BFQ_E_CUDA) if no device is usable.
 *
 * Threading (what ITenantRouteMatcher's callers need: matchAll runs on the shared "topic-matcher" ForkJoinPool,
 * DW/DistWorkerCoProcFactory.java:74-85, while mutate() runs on the range's raft-apply thread): a handle may be used
 * from any number of threads at once. Every match leases its own workspace (streams, device scratch, pinned result
 * buffers) and pins the snapshot it ran on; the result keeps both until it is freed, so results of concurrent matches
 * never share memory and a commit never changes what an existing result resolves to. load/apply/commit serialise
 * among themselves and never block matches.
 *
 * Semantics: the matcher implements the match PREDICATE of the reference. It equals the reference's
 * literal merge-join (TenantRouteMatcher.java:96-156) whenever that neither skips routes after its 20 probes nor seeks
 * backwards, i.e. for filters/topics without empty levels next to a shared prefix; the two documented divergences are
 * pinned in tests/test_oracle_golden.py and cannot be cross-checked against a JVM in this repository.
 */
#ifndef BFQ_GPUMATCH_H
#define BFQ_GPUMATCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFQ_OK 0
#define BFQ_E_INVALID (-1)   /* bad argument (unsorted keys, undecodable route key, ...) */
#define BFQ_E_CUDA (-2)      /* CUDA runtime / device error */
#define BFQ_E_NOMEM (-3)
#define BFQ_E_STATE (-4)     /* e.g. match before the first commit */
#define BFQ_E_RANGE (-5)     /* index out of range */

typedef struct bfq_index bfq_index;     /* forward index: topic filters (routes) of many tenants  */
typedef struct bfq_result bfq_result;   /* result of one bfq_match call                            */
typedef struct bfq_rindex bfq_rindex;   /* inverse index: topics, matched BY filters (retain)      */
typedef struct bfq_rresult bfq_rresult;

/* text of the last error raised on this thread */
const char* bfq_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Forward index life cycle. One bfq_index per dist-worker KV range == one DistWorkerCoProc
 * (DW/DistWorkerCoProc.java:105-125). It replaces the per-tenant TenantRouteMatcher instances
 * (DW/cache/TenantRouteCacheFactory.java:67-71) and their RocksDB merge-join.
 * ---------------------------------------------------------------------------------------------- */
int32_t bfq_index_create(int32_t device_ordinal, bfq_index** out);
void bfq_index_destroy(bfq_index* h);

/* Drop all staged routes. Called from DistWorkerCoProc.reset(Boundary) (DW/DistWorkerCoProc.java:283-291)
 * before re-loading the range. */
int32_t bfq_index_reset(bfq_index* h);

/* Bulk-stage raw KV pairs exactly as stored by the reference: key = route key
 * (DWS/KVSchemaUtil.java:91-130), value = 8-byte BE incarnation (normal) or RouteGroup proto (shared).
 * Keys must be strictly ascending in unsigned byte order (a KV range scan is). Decoding is native
 * (replaces DWS/cache/RouteDetailCache.java:53-109 on the load path). */
int32_t bfq_index_load(bfq_index* h, const uint8_t* keys, const int64_t* key_off, const uint8_t* vals,
                       const int64_t* val_off, int64_t n);

/* Incremental feed from the post-persist Supplier of DistWorkerCoProc.mutate
 * (DW/DistWorkerCoProc.java:188-209; batchAddRoute :304-413, batchRemoveRoute :415-513):
 * upsert n_add pairs, delete n_del keys (any order). */
int32_t bfq_index_apply(bfq_index* h, const uint8_t* add_keys, const int64_t* add_key_off, const uint8_t* add_vals,
                        const int64_t* add_val_off, int64_t n_add, const uint8_t* del_keys, const int64_t* del_key_off,
                        int64_t n_del);

/* Publish the staged state as a new immutable device snapshot. The host-side rebuild and the upload run while matches
 * continue on the previous snapshot; the swap is atomic with respect to matches (they always see a whole snapshot).
 * Every snapshot has a generation (1, 2, ...); results report the generation they were produced from. The previous
 * snapshot is freed when the last match / result that pins it is gone. bfq_index_apply is all-or-nothing: an
 * undecodable key leaves the staging area untouched. */
int32_t bfq_index_commit(bfq_index* h);
int32_t bfq_index_generation(bfq_index* h, uint64_t* generation);   /* 0 before the first commit */
/* Tuning knobs (defaults are the measured best for a handle that has the GPU to itself):
 *   "tier0_ctas_per_sm"  cap of the lane-per-topic kernel's resident CTAs per SM (0 = as many as fit, 7 on a B200). One slot
 *                        less leaves room for kernels that must run BESIDE the matching: the exchange of the previous batch
 *                        (bfq_exchange_gather on another stream) in a multi-GPU pipeline;
 *   "order_min_topics"   batches of at least this many topics are de-duplicated and matched in locality order (default 32768);
 *   "dedup"              0: match repeated (tenant, topic) pairs separately. */
int32_t bfq_index_set_option(bfq_index* h, const char* name, int64_t value);

/* stats[k], k < n: 0 routes, 1 tenants, 2 trie nodes, 3 hash-table slots, 4 device bytes, 5 max nodes per
 * depth, 6 kernel launches so far, 7 overflow (tier-2) topics so far, 8 cap-flagged topics so far,
 * 9 multi-segment filters, 10 long-token chunks, 11 topics handed from the lane-per-topic tier to the
 * warp-per-topic tier so far, 12 duplicate (tenant, topic) pairs answered from their first occurrence so far */
int32_t bfq_index_stats(bfq_index* h, int64_t* stats, int32_t n);
/* device time of the tier-0 (lane-per-topic) match kernel of the latest completed match call on this handle, measured with
 * CUDA events recorded on the launching stream around the launch (for roofline accounting) */
int32_t bfq_index_last_kernel_ms(bfq_index* h, double* ms);

/* Host-only diagnostic: run the index builder on a sorted KV snapshot without touching a device and report
 * stats[0..7] = routes, tenants, trie nodes, hash slots, max nodes per depth, max nodes per tenant,
 * multi-segment filters, long-token chunks, 8 = tag-table blocks that overflowed, 9..13 = nodes with 0/1/2/3/>=4
 * exact children, 14 = staging microseconds, 15 = flatten microseconds, 16 = a checksum of the whole image the build would
 * upload, 17 = 1 if building from one concatenated KV blob gives that same image, 18 = tenants whose stand-alone image (what a delta
 * commit builds for a touched tenant) equals their part of the full image (used by CPU tests and to time the build). */
int32_t bfq_host_build_stats(const uint8_t* keys, const int64_t* key_off, const uint8_t* vals, const int64_t* val_off,
                             int64_t n, int64_t* stats, int32_t n_stats);

/* Map a route rank (position in the committed KV order) back to its stored key/value so the Java side
 * re-hydrates Matching objects with its own KVSchemaUtil.buildMatchRoute (DWS/KVSchemaUtil.java:73-79).
 * Lengths are returned even if the capacities are too small (nothing is copied then).
 * These three resolve against the CURRENT snapshot: ranks shift with every add/remove, so a rank taken from a match
 * result must be resolved with bfq_result_route_lookup / bfq_result_route_kinds (below), which use the snapshot the
 * result was produced from — the reference reads keys and values from one consistent KV reader too. */
int32_t bfq_route_lookup(bfq_index* h, int64_t rank, uint8_t* key_out, int64_t key_cap, int64_t* key_len,
                         uint8_t* val_out, int64_t val_cap, int64_t* val_len);
/* per-rank route kind: 0 normal, 1 normal persistent (subBrokerId == 1), 2 group (shared subscription) */
int32_t bfq_route_kind(bfq_index* h, int64_t rank, int32_t* kind);
int32_t bfq_route_kinds(bfq_index* h, const int64_t* ranks, int64_t n, uint8_t* kinds_out);

/* ------------------------------------------------------------------------------------------------
 * Forward match == ITenantRouteMatcher.matchAll(Set<String> topics, int maxPersistentFanout,
 * int maxGroupFanout) (DW/cache/ITenantRouteMatcher.java:28-38; implementation replaced:
 * DW/cache/TenantRouteMatcher.java:68-161 + caps of DW/cache/MatchedRoutes.java:87-141),
 * batched over tenants. Topic i belongs to tenant topic_tenant[i] (index into the tenants list);
 * max_pfanout/max_gfanout are per tenant (Setting.MaxPersistentFanout / MaxGroupFanout).
 * Host buffers in, host result out (H2D + kernels + D2H inside the call; batches of >= 128k topics are cut into four
 * sub-batches pipelined over three streams so the copies overlap the kernels). A topic whose topic_tenant[i] is outside
 * [0, n_tenants) simply matches nothing. Pinned (page-locked) host buffers give the best H2D rate.
 * ---------------------------------------------------------------------------------------------- */
int32_t bfq_match(bfq_index* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                  const uint8_t* topics, const int64_t* topic_off, const int32_t* topic_tenant, int64_t n_topics,
                  const int32_t* max_pfanout, const int32_t* max_gfanout, bfq_result** out);

/* Result layout. The arrays live in pinned memory leased to this result: they stay valid, and private to it, until
 * bfq_result_free() — other matches on the same handle (from any thread) and commits do not touch them:
 *   span_begin[i], span_count[i]   topic i's matched route RANGES are ranges[span_begin[i] ... +span_count[i])
 *   ranges[j] = {first rank, count} a run of consecutive route ranks (one matched filter's routes)
 *   route_count[i]                 routes matched by topic i before caps
 *   throttled[k] = {topic, rank, kind} routes dropped by the fan-out caps, kind 1 = PersistentFanoutThrottled,
 *                                  2 = GroupFanoutThrottled (MatchedRoutes.java:95-100,128-133); the caller
 *                                  emits the events. Surviving routes of topic i = its ranges minus these. */
typedef struct { uint32_t first; uint32_t count; } bfq_range;
typedef struct { uint32_t topic; uint32_t rank; uint32_t kind; } bfq_throttled;
int64_t bfq_result_num_topics(const bfq_result* r);
const uint32_t* bfq_result_span_begin(const bfq_result* r);
const uint32_t* bfq_result_span_count(const bfq_result* r);
const uint32_t* bfq_result_route_count(const bfq_result* r);
const bfq_range* bfq_result_ranges(const bfq_result* r, int64_t* n_ranges);
const bfq_throttled* bfq_result_throttled(const bfq_result* r, int64_t* n_throttled);
/* Convenience: flatten to CSR of surviving ranks, ascending per topic. offsets[n_topics+1]; returns the
 * total, copies only if it fits rank_cap (large results are filled by several host threads). */
int64_t bfq_result_expand(const bfq_result* r, int64_t* offsets, int64_t* ranks, int64_t rank_cap);
/* rank -> stored key/value and route kind, resolved against the snapshot THIS result was produced from */
int32_t bfq_result_route_lookup(const bfq_result* r, int64_t rank, uint8_t* key_out, int64_t key_cap, int64_t* key_len,
                                uint8_t* val_out, int64_t val_cap, int64_t* val_len);
int32_t bfq_result_route_kinds(const bfq_result* r, const int64_t* ranks, int64_t n, uint8_t* kinds_out);
uint64_t bfq_result_generation(const bfq_result* r);
/* timings of the call in milliseconds: 0 busy time of the H2D copy stream (overlapped with kernels), 1 device time
 * of the tier-0 kernel of the first sub-batch, 2 number of pipelined sub-batches, 3 wall time of the whole call */
int32_t bfq_result_timings(const bfq_result* r, double* ms, int32_t n);
void bfq_result_free(bfq_result* r);

/* Same match with the topic batch already resident in device memory and the result left there
 * (used by bench.py's kernel-only leg and by callers that pipeline batches). d_* are device pointers
 * (the kernels read d_topics in whole aligned words / 16-byte granules that hold at least one topic byte, so the
 * blob must lie in memory that is readable up to its enclosing 16-byte boundaries: any cudaMalloc'd buffer),
 * stream is a cudaStream_t (NULL = default stream).
 *   bfq_match_device_async  enqueues every kernel of the match on `stream` and returns without synchronising: all
 *                           counts the later kernels need are read on the device. The d_* result pointers are valid
 *                           for work enqueued on the same stream afterwards; the n_* fields are not filled yet.
 *   bfq_device_result_wait  waits for the match, fills the n_* fields and handles the rare cases the optimistic
 *                           enqueue cannot (topics that need the global-scratch tier, buffers that must grow: the
 *                           batch is then re-run and the d_* pointers may change — read them after the wait).
 *   bfq_match_device        = async + wait.
 * The result buffers belong to a workspace leased to this result: they stay valid until bfq_device_result_release,
 * whatever else runs on the handle. Several matches may be in flight on one handle (and one stream) at a time. */
typedef struct {
    const uint32_t* d_span_begin;   /* [n_topics] */
    const uint32_t* d_span_count;   /* [n_topics] */
    const uint32_t* d_route_count;  /* [n_topics] */
    const bfq_range* d_ranges;      /* sparse: topic i's ranges are d_ranges[d_span_begin[i] ... + d_span_count[i] & 0x3FFFFFFF);
                                       n_ranges is the extent of the array, not the number of ranges */
    const bfq_throttled* d_throttled; /* [n_throttled] */
    int64_t n_ranges, n_throttled, n_routes;
    int64_t n_overflow_topics, n_flagged_topics, n_launches;
    int64_t n_topics;               /* topics of the batch (length of the per-topic arrays) */
    int64_t n_distinct_topics;      /* (tenant, topic) pairs actually walked; duplicates share their first occurrence's span */
    double tier0_ms;                /* device time of the lane-per-topic kernel of this match (CUDA events on `stream`) */
    uint64_t generation;            /* snapshot the match ran on */
    void* lease;                    /* opaque; owned by the library until bfq_device_result_release */
} bfq_device_result;
int32_t bfq_match_device(bfq_index* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                         const uint8_t* d_topics, const int64_t* d_topic_off, const int32_t* d_topic_tenant,
                         int64_t n_topics, const int32_t* max_pfanout, const int32_t* max_gfanout, void* stream,
                         bfq_device_result* out);
int32_t bfq_match_device_async(bfq_index* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                               const uint8_t* d_topics, const int64_t* d_topic_off, const int32_t* d_topic_tenant,
                               int64_t n_topics, const int32_t* max_pfanout, const int32_t* max_gfanout, void* stream,
                               bfq_device_result* out);
int32_t bfq_device_result_wait(bfq_device_result* res);
void bfq_device_result_release(bfq_device_result* res);
/* Flatten a completed device result into a device CSR: d_offsets[n_topics+1] (int64) is always written, the surviving
 * ranks (caps applied, unordered within a topic) are written to d_ranks if the total fits rank_cap (pass
 * d_ranks = NULL to only size). Returns the total via *n_ranks. Resolves against the result's own snapshot and caps. */
int32_t bfq_expand_device(const bfq_device_result* res, int64_t* d_offsets, int64_t* d_ranks, int64_t rank_cap,
                          void* stream, int64_t* n_ranks);

/* ------------------------------------------------------------------------------------------------
 * Batched range pruning on the dist-server side: TenantRangeLookupCache.lookup
 * (bifromq-dist/bifromq-dist-server/src/main/java/org/apache/bifromq/dist/server/scheduler/TenantRangeLookupCache.java:70-106)
 * decides, per publish topic, which of the tenant's KV ranges can hold a matching route: a range with a Fact
 * {firstGlobalFilterLevels, lastGlobalFilterLevels} stays a candidate iff the topic's expansion set (every filter that matches
 * it) has a member in [first, last]; the reference runs its expansion iterator per topic and candidate behind a cache. Here one
 * kernel answers a whole batch (one thread per (topic, candidate): a lower-bound walk of the implicit expansion trie).
 *   tenants / topics / topic_tenant   as for bfq_match
 *   cand_off[n_tenants + 1]           tenant t's candidate ranges are [cand_off[t], cand_off[t + 1]), in boundary order
 *   cand_flags[c]                     bit 0: the range has a Fact, bit 1: it has first, bit 2: it has last
 *   first / last (blob, off[n_cand + 1])   the global filter levels joined by NUL bytes, level 0 = the tenant id
 *   keep_off_out[n_topics + 1], keep_out[keep_off_out[n_topics]]   topic i's row = one byte per candidate of its tenant:
 *                                     1 = the range is returned by the reference's lookup, 0 = it is not
 * Semantics are the reference's loop, literally: no Fact -> kept; a Fact without first or last -> empty range, skipped; the
 * first range whose seek runs past the end of the expansion set ends the scan. Stateless; needs a CUDA device.
 * ---------------------------------------------------------------------------------------------- */
int32_t bfq_range_lookup(int32_t device_ordinal, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                         const uint8_t* topics, const int64_t* topic_off, const int32_t* topic_tenant, int64_t n_topics,
                         const int64_t* cand_off, const uint8_t* cand_flags, const uint8_t* first_blob, const int64_t* first_off,
                         const uint8_t* last_blob, const int64_t* last_off, int64_t* keep_off_out, uint8_t* keep_out);

/* ------------------------------------------------------------------------------------------------
 * Fan-out expansion on the device: the step right behind the match. DeliverExecutorGroup.submit
 * (DW/DeliverExecutorGroup.java:112-231) walks every matched route of a message, resolves a shared subscription to one
 * member (:242-278) and hands each route to the deliverer of its (subBrokerId, delivererKey) (DW/DeliverExecutor.java:89-93),
 * which batches per deliverer. bfq_fanout_device does that grouping for a whole batch: it takes the device CSR of a completed
 * match (bfq_expand_device: surviving ranks per topic, caps applied) and returns every (topic, route) pair grouped by
 * deliverer id: pairs [d_pack_offsets[d], d_pack_offsets[d + 1]) belong to deliverer d; d_pack_topic / d_pack_rank give the
 * pair, d_pack_member the member index a $share subscription was resolved to (0xFFFFFFFF for ordinary routes; members in the
 * order of the stored RouteGroup). An unordered share picks member hash(topic position, rank) mod n (the reference picks
 * uniformly at random: any member is valid); ORDERED shares need each message's publisher (rendezvous hash of ClientInfo) and
 * are grouped, unresolved, under the last id (ordered_share_id) for the host. Ids are dense over the distinct
 * (subBrokerId, delivererKey) pairs of the index, stable across commits; bfq_fanout_deliverer gives the pair back.
 * The arrays live in the result's leased workspace: valid until bfq_device_result_release.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const int64_t* d_pack_offsets;   /* [n_deliverers + 1] */
    const uint32_t* d_pack_topic;    /* [n_pairs] topic position in the batch */
    const uint32_t* d_pack_rank;     /* [n_pairs] route rank (of the result's snapshot) */
    const uint32_t* d_pack_member;   /* [n_pairs] */
    int64_t n_pairs;
    int32_t n_deliverers;            /* ids [0, n_deliverers); the last one is ordered_share_id */
    int32_t ordered_share_id;
    uint64_t generation;
} bfq_fanout_result;
int32_t bfq_fanout_device(const bfq_device_result* res, const int64_t* d_offsets, const int64_t* d_ranks, int64_t n_pairs,
                          void* stream, bfq_fanout_result* out);
int32_t bfq_fanout_deliverer(bfq_index* h, int32_t id, int32_t* sub_broker_id, uint8_t* key_out, int64_t key_cap, int64_t* key_len);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: the one exchange step of the tenant-sharded path. Tenants are independent key ranges, so
 * every GPU (one process each) matches the topics of the tenants it hosts with NO data-path collective; what travels is
 * the reply, reassembled on every rank the way the dist-server reassembles the per-worker BatchDistReply messages
 * (bifromq-dist/bifromq-dist-server/src/main/java/org/apache/bifromq/dist/server/scheduler/BatchDistServerCall.java:186-205,
 * 245-271). bfq_exchange_gather all-gathers the device results of the ranks' matches over NCCL (NVLink / NVSwitch), in
 * rank order: per topic the matched-route count (what TopicFanout carries, DistWorkerCoProc.proto:34-131) and, with
 * BFQ_EXCHANGE_RANGES, the number of matched ranges plus the dense {first rank, count} ranges themselves (route ranks are
 * local to the rank that produced them: rank r's ranges refer to r's committed KV order). One host synchronisation per
 * call (NCCL needs the receive counts); no host-side copies. NCCL is taken from the process at run time (libnccl.so.2).
 *   rank 0: bfq_exchange_unique_id(id)  -> broadcast the 128 bytes to the other ranks out of band (the host's own RPC)
 *   every rank: bfq_exchange_create(device, rank, world, id, &x)       (collective: all ranks must call it)
 *   per batch, every rank: bfq_match_device(...) ; bfq_exchange_gather(x, &res, what, stream, &g)   (collective)
 * The gathered arrays live in device memory owned by the exchange, valid until the next gather on it; topic_base /
 * range_base (host, [world + 1]) give each rank's slice. A world of 1 is allowed (the gather is then a local compaction).
 * ---------------------------------------------------------------------------------------------- */
#define BFQ_EXCHANGE_ID_BYTES 128
#define BFQ_EXCHANGE_COUNTS 1
#define BFQ_EXCHANGE_RANGES 2
typedef struct bfq_exchange bfq_exchange;
typedef struct {
    const uint32_t* d_route_count;   /* matched routes per topic; rank r's slice starts at topic_base[r] */
    const uint32_t* d_span_count;    /* matched ranges per topic, same slices (NULL with BFQ_EXCHANGE_COUNTS) */
    const bfq_range* d_ranges;       /* rank r's slice starts at range_base[r]; dense inside a slice: a topic's ranges follow
                                        those of the topic before it (NULL with BFQ_EXCHANGE_COUNTS) */
    const int64_t* topic_base;       /* host [world + 1]: rank r's topics are [topic_base[r], topic_base[r] + topic_count[r]) — the  */
    const int64_t* range_base;       /* host [world + 1]   slices have one padded stride (the payload travels as ncclAllGather)    */
    const int64_t* topic_count;      /* host [world] */
    const int64_t* range_count;      /* host [world] */
    int64_t n_topics_total, n_ranges_total;
    int64_t bytes_received;          /* payload bytes this rank received from its peers */
    int32_t world;
} bfq_gathered;
int32_t bfq_exchange_unique_id(uint8_t* id_out, int32_t cap);
int32_t bfq_exchange_create(int32_t device_ordinal, int32_t rank, int32_t world, const uint8_t* id, bfq_exchange** out);
void bfq_exchange_destroy(bfq_exchange* x);
int32_t bfq_exchange_gather(bfq_exchange* x, const bfq_device_result* res, int32_t what, void* stream, bfq_gathered* out);

/* ------------------------------------------------------------------------------------------------
 * Route key codec + tokeniser, native restatement of DWS/KVSchemaUtil.java:56-130 and
 * U/TopicUtil.java:42-163,206-225 (exported so the Java side / tests can cross-check bytes).
 * Each returns the produced length (or BFQ_E_*), copying only if it fits cap.
 * ---------------------------------------------------------------------------------------------- */
int64_t bfq_receiver_url(int32_t sub_broker_id, const uint8_t* receiver_id, int64_t rn, const uint8_t* deliverer_key,
                         int64_t dn, uint8_t* out, int64_t cap);
/* mqtt_topic_filter may carry a $share/<g>/ or $oshare/<g>/ prefix (=> toGroupRouteKey, receiver_url ignored) */
int64_t bfq_route_key(const uint8_t* tenant, int64_t tn, const uint8_t* mqtt_topic_filter, int64_t fn,
                      const uint8_t* receiver_url, int64_t un, uint8_t* out, int64_t cap);
int64_t bfq_tenant_begin_key(const uint8_t* tenant, int64_t tn, uint8_t* out, int64_t cap);
/* retain store key layout (bifromq-retain/bifromq-retain-store-schema/src/main/java/org/apache/bifromq/retain/store/schema/
 * KVSchemaUtil.java:44-73, LevelHash.java:31-49): retainMessageKey(tenant, topic) and retainKeyPrefix of a topic filter */
int64_t bfq_retain_key(const uint8_t* tenant, int64_t tn, const uint8_t* topic, int64_t n, uint8_t* out, int64_t cap);
int64_t bfq_retain_key_prefix(const uint8_t* tenant, int64_t tn, const uint8_t* topic_filter, int64_t fn, uint8_t* out, int64_t cap);
int32_t bfq_is_valid_topic(const uint8_t* topic, int64_t n, int32_t max_level_length, int32_t max_level, int32_t max_length);
int32_t bfq_is_valid_topic_filter(const uint8_t* tf, int64_t n, int32_t max_level_length, int32_t max_level, int32_t max_length);

/* ------------------------------------------------------------------------------------------------
 * Inverse index == IRetainTopicIndex (RS/index/IRetainTopicIndex.java:27-35; implementation replaced:
 * RS/index/RetainTopicIndex.java:35-144 over U/index/TopicLevelTrie.java:190-249) and, with
 * tenant == NULL levels, DW/TopicIndex.java:39-156. Topics are staged with add/remove, published with
 * commit, and matched BY a batch of topic filters.
 * ---------------------------------------------------------------------------------------------- */
int32_t bfq_rindex_create(int32_t device_ordinal, bfq_rindex** out);
void bfq_rindex_destroy(bfq_rindex* h);
int32_t bfq_rindex_reset(bfq_rindex* h);
/* add n topics; topic i belongs to tenant topic_tenant[i] of the tenants list; returns via ids_out[i] the
 * stable topic id (>= 0) used in match results. Adding an existing (tenant, topic) returns its id. */
int32_t bfq_rindex_add(bfq_rindex* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                       const uint8_t* topics, const int64_t* topic_off, const int32_t* topic_tenant, int64_t n,
                       int64_t* ids_out);
/* The feed of RetainStoreCoProc.load() (RS/RetainStoreCoProc.java:279-296): raw retain-store KV KEYS of a range scan. The
 * reference parses every value (a TopicMessage proto) for the topic; the key carries it too, so the index is fed from the keys
 * alone. ids_out[i] = the topic's id, or -1 for bytes that are not a retain key (skipped). */
int32_t bfq_rindex_load_keys(bfq_rindex* h, const uint8_t* keys, const int64_t* key_off, int64_t n, int64_t* ids_out);
int32_t bfq_rindex_remove(bfq_rindex* h, const uint8_t* tenant, int64_t tn, const uint8_t* topic, int64_t n);
int32_t bfq_rindex_commit(bfq_rindex* h);
/* topic id -> (tenant, topic) strings */
int32_t bfq_rindex_lookup(bfq_rindex* h, int64_t id, uint8_t* tenant_out, int64_t tenant_cap, int64_t* tenant_len,
                          uint8_t* topic_out, int64_t topic_cap, int64_t* topic_len);
/* match n filters; filter i is scoped to tenant filter_tenant[i]; limit[i] < 0 = unlimited, else at most
 * limit[i] ids are returned for filter i (RS/RetainStoreCoProc.java:167-190 stops after `limit` messages;
 * which ones is unspecified there too — it iterates a HashSet). NOTE: the reference applies its expiry filter INSIDE that
 * loop (it keeps iterating until `limit` LIVE messages are found, RetainStoreCoProc.java:177-188), whereas this call
 * truncates to `limit` topic ids before the caller has looked at any message: a caller that drops expired messages must
 * ask for more than `limit` (total_matches tells how many exist) or pass limit < 0 and cut after its own expiry check.
 * Thread-safe: calls on one handle are serialised, every result owns its arrays. */
int32_t bfq_rmatch(bfq_rindex* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                   const uint8_t* filters, const int64_t* filter_off, const int32_t* filter_tenant, int64_t n_filters,
                   const int64_t* limit, bfq_rresult** out);
int64_t bfq_rresult_num_filters(const bfq_rresult* r);
const int64_t* bfq_rresult_offsets(const bfq_rresult* r);            /* [n_filters+1] */
const int64_t* bfq_rresult_ids(const bfq_rresult* r, int64_t* n);    /* topic ids, ascending per filter */
const int64_t* bfq_rresult_total_matches(const bfq_rresult* r);      /* [n_filters] matches before the limit */
/* ms[0..3]: wall time of the H2D section, the kernel section, the D2H section, the whole call; ms[4]: DEVICE time from
 * "inputs resident" to "ids expanded" (CUDA events on the call's stream); ms[5]: device time of rmatch_kernel alone;
 * ms[6]: rank ranges the kernel emitted (8 bytes each); ms[7]: filters that needed the global-scratch tier */
int32_t bfq_rresult_timings(const bfq_rresult* r, double* ms, int32_t n);
/* retainMessageKey of every id of the result, in result order, as one (blob, key_off[n_ids + 1]) batch: the keys of the
 * follow-up reader.get calls of RetainStoreCoProc.match (RS/RetainStoreCoProc.java:177-188). Returns the blob length (or a
 * negative BFQ_E_*); copies only if it fits blob_cap; key_off_out may be NULL. */
int64_t bfq_rresult_retain_keys(bfq_rindex* h, const bfq_rresult* r, uint8_t* blob_out, int64_t blob_cap, int64_t* key_off_out);
void bfq_rresult_free(bfq_rresult* r);

#ifdef __cplusplus
}
#endif
#endif /* BFQ_GPUMATCH_H */
