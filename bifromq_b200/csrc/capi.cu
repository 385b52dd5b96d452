// capi.cu — implementation of the C-ABI declared in include/bfq_gpumatch.h (forward index + match).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bfq_gpumatch.h"
#include "codec.h"
#include "index_builder.h"
#include "match_kernels.cuh"

using namespace bfq;

namespace bfq {
thread_local std::string g_last_error;
int32_t set_error(int32_t code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
}  // namespace bfq

namespace {

int32_t fail(int32_t code, const std::string& msg) { return bfq::set_error(code, msg); }
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(BFQ_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

// growable device / pinned-host buffers
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    size_t bytes() const { return cap * sizeof(T); }
};
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct bfq_result {
    bfq_index* owner = nullptr;
    int64_t n_topics = 0, n_ranges = 0, n_throttled = 0;
    const uint32_t *span_begin = nullptr, *span_count = nullptr, *route_count = nullptr;
    const bfq_range* ranges = nullptr;
    const bfq_throttled* throttled = nullptr;
    double ms[4] = {0, 0, 0, 0};
};

struct bfq_index {
    int device = 0;
    std::mutex mu;         // device snapshot + workspace: matches, lookups, the snapshot swap of commit
    std::mutex stage_mu;   // staging area: reset / load / apply and the (long) host-side rebuild of commit
    Staging staging;
    FlatIndex flat;          // host copy of the committed snapshot (segs / tenant map / stats are used on the host)
    KVBlob committed;        // committed KV (for bfq_route_lookup)
    bool have_snapshot = false;
    cudaStream_t stream = nullptr, copy_stream = nullptr, work_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_h2d[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evk[2] = {nullptr, nullptr};
    double last_kernel_ms = 0;
    size_t l2_window_bytes = 0;
    bool tenant_tab_valid = false;     // resolved tenant table cached on the device (invalidated by commit)
    uint64_t tenant_tab_fp = 0;
    int32_t tenant_tab_n = 0;
    // snapshot on device
    DevBuf<Slot> d_slots, d_roots;
    DevBuf<uint32_t> d_segs, d_pfxP, d_pfxG;
    DevBuf<uint8_t> d_rkind, d_tags;
    // per-call workspace
    DevBuf<uint8_t> d_topics;
    DevBuf<int64_t> d_topic_off;
    DevBuf<int32_t> d_topic_tenant, d_tenant_tab;   // tenant_tab = root | maxP | maxG, 3 x n_tenants
    DevBuf<uint32_t> d_span_begin, d_span_count, d_route_count, d_overflow, d_flagged, d_kept, d_defer;
    DevBuf<uint2> d_ranges, d_scratch, d_ranges_c;
    DevBuf<uint8_t> d_scan_tmp;
    DevBuf<uint32_t> d_cnt, d_new_begin;
    DevBuf<uint32_t> d_ord_keys, d_ord_vals;   // locality ordering of tier 0 (launch_order): 2n each
    DevBuf<uint8_t> d_ord_tmp;
    size_t ord_tmp_stride = 0;                 // bytes of sort scratch per sub-batch
    int64_t order_min = 32768;                 // batches smaller than this are matched in arrival order (BFQ_ORDER=0: never order)
    DevBuf<uint3> d_throttled;
    DevBuf<unsigned long long> d_counters;
    PinBuf<unsigned long long> h_counters;
    PinBuf<int32_t> h_tenant_tab;
    // pinned result buffers (leased to the bfq_result of the latest bfq_match)
    PinBuf<uint32_t> h_span_begin, h_span_count, h_route_count;
    PinBuf<uint2> h_ranges;
    PinBuf<uint3> h_throttled;
    // statistics
    int64_t launches = 0, overflow_topics = 0, flagged_topics = 0, deferred_topics = 0;
    // last device result (for bfq_expand_device)
    int64_t last_n_topics = 0, last_n_flagged = 0;
    int32_t last_n_tenants = 0;
    const int32_t* last_topic_tenant = nullptr;   // device pointer of the latest bfq_match_device call
    DevBuf<unsigned long long> d_exp_counts;

    ~bfq_index() {
        cudaSetDevice(device);
        d_slots.release(); d_roots.release(); d_segs.release(); d_pfxP.release(); d_pfxG.release(); d_rkind.release(); d_tags.release();
        d_topics.release(); d_topic_off.release(); d_topic_tenant.release(); d_tenant_tab.release();
        d_span_begin.release(); d_span_count.release(); d_route_count.release(); d_overflow.release();
        d_flagged.release(); d_kept.release(); d_defer.release(); d_exp_counts.release(); d_ranges_c.release(); d_scan_tmp.release(); d_cnt.release(); d_new_begin.release(); d_ranges.release(); d_scratch.release(); d_throttled.release();
        d_counters.release(); h_counters.release(); h_tenant_tab.release();
        h_span_begin.release(); h_span_count.release(); h_route_count.release(); h_ranges.release(); h_throttled.release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& e : evk) if (e) cudaEventDestroy(e);
        for (auto& e : ev_h2d) if (e) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (auto& w : work_stream) if (w) cudaStreamDestroy(w);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

// Shared core of bfq_match / bfq_match_device: topics are on the device; runs tier 1, tier 2 and caps.
struct CoreOut {
    int64_t n_ranges = 0, n_throttled = 0, n_overflow = 0, n_flagged = 0, n_launches = 0, n_deferred = 0;
    uint64_t want_dyn = 0, want_thr = 0;
    int64_t chunk_throttled[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

int32_t resolve_tenants(bfq_index* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                        const int32_t* max_p, const int32_t* max_g, cudaStream_t stream) {
    if (n_tenants < 0) return fail(BFQ_E_INVALID, "n_tenants < 0");
    if (n_tenants >= (1 << 30)) return fail(BFQ_E_RANGE, "more than 2^30 tenants in one batch");   // bit 30 of the lane's tenant word is a flag
    const size_t nt = (size_t) std::max(n_tenants, 1);
    // the same tenant list + caps usually accompany every batch: fingerprint it and keep the device table
    uint64_t fp = 0xcbf29ce484222325ull ^ (uint64_t) n_tenants;
    auto mixin = [&](const void* p, size_t n) {
        const uint8_t* b = (const uint8_t*) p;
        for (size_t i = 0; i < n; i++) fp = (fp ^ b[i]) * 0x100000001b3ull;
    };
    if (n_tenants > 0) {
        mixin(tenant_off, (size_t) (n_tenants + 1) * sizeof(int64_t));
        mixin(tenants + tenant_off[0], (size_t) (tenant_off[n_tenants] - tenant_off[0]));
        if (max_p) mixin(max_p, (size_t) n_tenants * 4);
        if (max_g) mixin(max_g, (size_t) n_tenants * 4);
        fp ^= (max_p ? 1u : 0u) | (max_g ? 2u : 0u);
    }
    if (h->tenant_tab_valid && h->tenant_tab_fp == fp && h->tenant_tab_n == n_tenants) return BFQ_OK;
    CUDA_TRY(cudaStreamSynchronize(stream));   // the pinned staging table may still be in flight
    CUDA_TRY(h->h_tenant_tab.reserve(3 * nt));
    CUDA_TRY(h->d_tenant_tab.reserve(3 * nt));
    for (int32_t i = 0; i < n_tenants; i++) {
        std::string t((const char*) tenants + tenant_off[i], (size_t) (tenant_off[i + 1] - tenant_off[i]));
        auto it = h->flat.tenant_ordinal.find(t);
        h->h_tenant_tab.p[i] = it == h->flat.tenant_ordinal.end() ? -1 : (int32_t) it->second;
        h->h_tenant_tab.p[nt + i] = max_p ? max_p[i] : 0x7FFFFFFF;
        h->h_tenant_tab.p[2 * nt + i] = max_g ? max_g[i] : 0x7FFFFFFF;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_tenant_tab.p, h->h_tenant_tab.p, 3 * nt * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));   // other streams of this handle read the table without an event
    h->tenant_tab_valid = true;
    h->tenant_tab_fp = fp;
    h->tenant_tab_n = n_tenants;
    return BFQ_OK;
}

// One sub-batch of a match: topics [begin, begin + n) of a batch of n_total. Device buffers are indexed by the position
// in the whole batch, so sub-batches of one call never overlap; each has its own counter block, its own slice of the
// dynamic range region and of the throttled list.
struct SubBatch {
    int64_t begin = 0, n = 0, n_total = 0;
    int chunk = 0;
    uint64_t dyn_off = 0, dyn_cap = 0;     // slice of ranges[n_total * INLINE_RANGES ...) for tiers 1/2
    uint64_t thr_off = 0, thr_cap = 0;     // slice of the throttled list
};
constexpr int MAX_CHUNKS = 8;
constexpr int32_t BFQ_RETRY_GROW = -100;   // internal: a slice was too small, redo the batch un-chunked with bigger buffers

int32_t prepare_workspace(bfq_index* h, int64_t n, int n_chunks) {
    const size_t nn = (size_t) std::max<int64_t>(n, 1);
    if (n >= (int64_t) 0x3FFFFFFF) return fail(BFQ_E_INVALID, "too many topics in one batch");
    CUDA_TRY(h->d_span_begin.reserve(nn));
    CUDA_TRY(h->d_span_count.reserve(nn));
    CUDA_TRY(h->d_route_count.reserve(nn));
    CUDA_TRY(h->d_overflow.reserve(nn));
    CUDA_TRY(h->d_flagged.reserve(nn));
    CUDA_TRY(h->d_kept.reserve(nn));
    CUDA_TRY(h->d_defer.reserve(nn));
    CUDA_TRY(h->d_counters.reserve(CTR_COUNT * MAX_CHUNKS));
    CUDA_TRY(h->h_counters.reserve(CTR_COUNT * MAX_CHUNKS));
    // ranges[0, n * INLINE_RANGES): tier-0 inline slots; the rest: cursor-allocated region of tiers 1 and 2
    const uint64_t dyn_base = (uint64_t) n * INLINE_RANGES;
    if (dyn_base >= 0xF0000000ull) return fail(BFQ_E_RANGE, "batch too large for 32-bit range indices; split the batch");
    const size_t min_dyn = std::max<size_t>((size_t) n_chunks << 18, nn);
    if (h->d_ranges.cap < dyn_base + min_dyn) CUDA_TRY(h->d_ranges.reserve((size_t) (dyn_base + std::max<size_t>(1 << 20, min_dyn))));
    if (n >= h->order_min) {
        CUDA_TRY(h->d_ord_keys.reserve(2 * nn));
        CUDA_TRY(h->d_ord_vals.reserve(2 * nn));
        OrderParams q{};
        q.n_topics = n;
        size_t tb = 0;
        CUDA_TRY(launch_order(q, nullptr, &tb, nullptr));
        tb = (tb + 255) / 256 * 256;
        h->ord_tmp_stride = std::max(h->ord_tmp_stride, tb);
        CUDA_TRY(h->d_ord_tmp.reserve(h->ord_tmp_stride * MAX_CHUNKS));
    }
    if (h->d_throttled.cap < ((size_t) n_chunks << 14)) CUDA_TRY(h->d_throttled.reserve(std::max<size_t>(1 << 16, (size_t) n_chunks << 14)));
    return BFQ_OK;
}

SubBatch whole_batch(bfq_index* h, int64_t n) {
    SubBatch sb;
    sb.begin = 0;
    sb.n = sb.n_total = n;
    sb.dyn_cap = h->d_ranges.cap - (uint64_t) n * INLINE_RANGES;
    sb.thr_cap = h->d_throttled.cap;
    return sb;
}

// Runs tier 0 + tier 1 (+ tier 2, + caps) for one sub-batch on `stream`; blocks until its kernels have finished.
int32_t match_core(bfq_index* h, const uint8_t* d_topics, const int64_t* d_topic_off, const int32_t* d_topic_tenant,
                   int32_t n_tenants, cudaStream_t stream, const SubBatch& sb, CoreOut* out) {
    const size_t nt = (size_t) std::max(n_tenants, 1);
    const int64_t n = sb.n, b = sb.begin;
    MatchParams p{};
    p.slots = h->d_slots.p;
    p.roots = h->d_roots.p;
    p.tags = reinterpret_cast<const uint4*>(h->d_tags.p);
    p.n_blocks = h->flat.n_blocks;
    p.topics = d_topics;
    p.topic_off = d_topic_off + b;
    p.topic_tenant = d_topic_tenant + b;
    p.tenant_root = h->d_tenant_tab.p;
    p.max_pfanout = h->d_tenant_tab.p + nt;
    p.max_gfanout = h->d_tenant_tab.p + 2 * nt;
    p.n_tenants = n_tenants;
    p.n_topics = n;
    p.span_begin = h->d_span_begin.p + b;
    p.span_count = h->d_span_count.p + b;
    p.route_count = h->d_route_count.p + b;
    p.overflow_list = h->d_overflow.p + b;
    p.defer_list = h->d_defer.p + b;
    p.flagged_list = h->d_flagged.p + b;
    unsigned long long* d_ctr = h->d_counters.p + (size_t) sb.chunk * CTR_COUNT;
    unsigned long long* hc = h->h_counters.p + (size_t) sb.chunk * CTR_COUNT;
    p.counters = d_ctr;
    // range indices are relative to the sub-batch's first inline slot
    p.ranges = h->d_ranges.p + (uint64_t) b * INLINE_RANGES;
    p.dyn_base = (uint64_t) (sb.n_total - b) * INLINE_RANGES + sb.dyn_off;
    p.ranges_cap = p.dyn_base + sb.dyn_cap;

    if (h->l2_window_bytes > 0) {
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = h->d_tags.p;
        attr.accessPolicyWindow.num_bytes = h->l2_window_bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
    }
    CUDA_TRY(cudaMemsetAsync(d_ctr, 0, CTR_COUNT * sizeof(unsigned long long), stream));
    p.work_list = nullptr;
    p.n_work = 0;
    p.order = nullptr;
    if (n >= h->order_min && h->d_ord_keys.cap >= 2 * (size_t) (b + n) && h->ord_tmp_stride > 0) {
        // group the topics by tenant and leading levels so that neighbouring lanes walk the same part of the trie
        OrderParams q{};
        q.n_topics = n;
        q.topics = d_topics;
        q.topic_off = p.topic_off;
        q.topic_tenant = p.topic_tenant;
        q.n_tenants = n_tenants;
        q.keys = h->d_ord_keys.p + 2 * b;
        q.vals = h->d_ord_vals.p + 2 * b;
        size_t tb = h->ord_tmp_stride;
        CUDA_TRY(launch_order(q, h->d_ord_tmp.p + (size_t) sb.chunk * h->ord_tmp_stride, &tb, stream));
        p.order = q.vals + n;
        out->n_launches += 1;   // order_keys_kernel (the radix-sort passes are cub's)
    }
    if (n > 0) {
        // tier 0 (one lane per topic) over the whole sub-batch, then tier 1 (one warp per topic) over whatever tier 0
        // deferred — its count is read on the device, so both launches go out back to back
        if (sb.chunk == 0) CUDA_TRY(cudaEventRecord(h->evk[0], stream));
        launch_match_lanes(p, stream);
        if (sb.chunk == 0) CUDA_TRY(cudaEventRecord(h->evk[1], stream));
        p.work_list = p.defer_list;
        p.n_work = -1;
        launch_match(p, false, 0, stream);
        out->n_launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(hc, d_ctr, CTR_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    out->n_overflow += (int64_t) hc[CTR_OVERFLOW];
    out->n_deferred += (int64_t) hc[CTR_DEFER];
    if (n > 0 && sb.chunk == 0) {
        float kms = 0;
        cudaEventElapsedTime(&kms, h->evk[0], h->evk[1]);
        h->last_kernel_ms = kms;
    }
    if (hc[CTR_OVERFLOW] > 0) {
        // ---- tier 2: per-warp global scratch sized from the index statistics (exact upper bounds)
        const uint64_t capF = (uint64_t) h->flat.max_nodes_per_depth + 2;
        const uint64_t capR = 2 * ((uint64_t) h->flat.max_tenant_nodes + 2) + 2;
        const uint64_t per_warp = 4 * capF + capR;   // uint2 units: two frontier buffers of uint4 entries + ranges
        uint64_t warps = std::min<uint64_t>(hc[CTR_OVERFLOW], std::max<uint64_t>(8, (1ull << 31) / (per_warp * sizeof(uint2))));
        warps = std::min<uint64_t>(warps, 148 * 8);
        warps = (warps + 7) / 8 * 8;
        CUDA_TRY(cudaDeviceSynchronize());   // the scratch is shared: no other sub-batch may be in tier 2 (or running at all)
        CUDA_TRY(h->d_scratch.reserve((size_t) (warps * per_warp)));
        p.scratch = h->d_scratch.p;
        p.scratch_frontier_cap = capF;
        p.scratch_ranges_cap = capR;
        p.work_list = p.overflow_list;
        p.n_work = (int64_t) hc[CTR_OVERFLOW];
        launch_match(p, true, (int) warps, stream);
        out->n_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(hc, d_ctr, CTR_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (hc[CTR_ERROR] != 0) return fail(BFQ_E_STATE, "tier-2 scratch exhausted (index statistics inconsistent)");
    }
    if (hc[CTR_RANGES] > sb.dyn_cap) {
        out->want_dyn = std::max<uint64_t>(out->want_dyn, hc[CTR_RANGES]);
        return BFQ_RETRY_GROW;
    }
    out->n_ranges += (int64_t) hc[CTR_RANGES];
    out->n_flagged += (int64_t) hc[CTR_FLAGGED];
    if (hc[CTR_FLAGGED] > 0) {
        CapsParams c{};
        c.flagged_list = p.flagged_list;
        c.n_flagged = (int64_t) hc[CTR_FLAGGED];
        c.topic_tenant = p.topic_tenant;
        c.max_pfanout = p.max_pfanout;
        c.max_gfanout = p.max_gfanout;
        c.span_begin = p.span_begin;
        c.span_count = p.span_count;
        c.ranges = p.ranges;
        c.segs = h->d_segs.p;
        c.rkind = h->d_rkind.p;
        c.pfx_persistent = h->d_pfxP.p;
        c.pfx_group = h->d_pfxG.p;
        c.kept_count = h->d_kept.p + b;
        c.counters = d_ctr;
        c.throttled = h->d_throttled.p + sb.thr_off;
        c.throttled_cap = sb.thr_cap;
        c.topic_base = (uint32_t) b;
        launch_caps(c, stream);
        out->n_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(hc, d_ctr, CTR_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (hc[CTR_THROTTLED] > sb.thr_cap) {
            out->want_thr = std::max<uint64_t>(out->want_thr, hc[CTR_THROTTLED]);
            return BFQ_RETRY_GROW;
        }
        out->chunk_throttled[sb.chunk] = (int64_t) hc[CTR_THROTTLED];
        out->n_throttled += (int64_t) hc[CTR_THROTTLED];
    }
    return BFQ_OK;
}

// un-chunked match with automatic buffer growth (device path, and the host path's fallback)
int32_t match_whole(bfq_index* h, const uint8_t* d_topics, const int64_t* d_topic_off, const int32_t* d_topic_tenant, int64_t n,
                    int32_t n_tenants, cudaStream_t stream, CoreOut* out) {
    for (int attempt = 0; attempt < 8; attempt++) {
        int32_t rc = prepare_workspace(h, n, 1);
        if (rc != BFQ_OK) return rc;
        *out = CoreOut();
        rc = match_core(h, d_topics, d_topic_off, d_topic_tenant, n_tenants, stream, whole_batch(h, n), out);
        if (rc != BFQ_RETRY_GROW) {
            if (rc == BFQ_OK) {
                h->launches += out->n_launches;
                h->overflow_topics += out->n_overflow;
                h->deferred_topics += out->n_deferred;
                h->flagged_topics += out->n_flagged;
                h->last_n_topics = n;
            }
            return rc;
        }
        CUDA_TRY(cudaDeviceSynchronize());
        if (out->want_dyn) {
            const size_t want = (size_t) ((uint64_t) n * INLINE_RANGES + out->want_dyn + out->want_dyn / 4 + 1024);
            if (want >= 0xFFFFFFF0ull) return fail(BFQ_E_RANGE, "more than 2^32 matched ranges in one batch; split the batch");
            CUDA_TRY(h->d_ranges.reserve(want));
        }
        if (out->want_thr) CUDA_TRY(h->d_throttled.reserve((size_t) (out->want_thr + out->want_thr / 4 + 1024)));
    }
    return fail(BFQ_E_STATE, "buffer sizing did not converge");
}

int64_t emit_bytes(const std::string& s, uint8_t* out, int64_t cap) {
    if (out && (int64_t) s.size() <= cap) memcpy(out, s.data(), s.size());
    return (int64_t) s.size();
}

}  // namespace

extern "C" {

const char* bfq_last_error(void) { return bfq::g_last_error.c_str(); }

int32_t bfq_index_create(int32_t device_ordinal, bfq_index** out) {
    if (!out) return fail(BFQ_E_INVALID, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(BFQ_E_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device_ordinal < 0 || device_ordinal >= count) return fail(BFQ_E_INVALID, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device_ordinal));
    auto* h = new bfq_index();
    h->device = device_ordinal;
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    for (auto& ev : h->ev)
        if (e == cudaSuccess) e = cudaEventCreate(&ev);
    for (auto& ev : h->evk)
        if (e == cudaSuccess) e = cudaEventCreate(&ev);
    for (auto& ev : h->ev_h2d)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    for (auto& w : h->work_stream)
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete h;
        return fail(BFQ_E_CUDA, cudaGetErrorString(e));
    }
    if (const char* eo = getenv("BFQ_ORDER")) {   // experiment switch: 0 = never order, N > 0 = order batches of >= N topics
        const long long v = atoll(eo);
        h->order_min = v <= 0 ? (int64_t) 1 << 62 : (int64_t) v;
    }
    *out = h;
    return BFQ_OK;
}

void bfq_index_destroy(bfq_index* h) { delete h; }

int32_t bfq_index_reset(bfq_index* h) {
    if (!h) return fail(BFQ_E_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> g(h->stage_mu);
    h->staging.reset();
    return BFQ_OK;
}

int32_t bfq_index_load(bfq_index* h, const uint8_t* keys, const int64_t* key_off, const uint8_t* vals,
                       const int64_t* val_off, int64_t n) {
    if (!h || n < 0 || (n > 0 && (!keys || !key_off || !vals || !val_off))) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->stage_mu);
    std::string err;
    if (!h->staging.load(keys, key_off, vals, val_off, n, &err)) return fail(BFQ_E_INVALID, err);
    return BFQ_OK;
}

int32_t bfq_index_apply(bfq_index* h, const uint8_t* add_keys, const int64_t* add_key_off, const uint8_t* add_vals,
                        const int64_t* add_val_off, int64_t n_add, const uint8_t* del_keys, const int64_t* del_key_off,
                        int64_t n_del) {
    if (!h || n_add < 0 || n_del < 0) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->stage_mu);
    for (int64_t i = 0; i < n_add; i++) {
        sv k((const char*) add_keys + add_key_off[i], (size_t) (add_key_off[i + 1] - add_key_off[i]));
        DecodedKey d;
        if (!decode_route_key(k, &d)) return fail(BFQ_E_INVALID, "undecodable route key in add set");
        h->staging.upsert(k, sv((const char*) add_vals + add_val_off[i], (size_t) (add_val_off[i + 1] - add_val_off[i])));
    }
    for (int64_t i = 0; i < n_del; i++)
        h->staging.erase(sv((const char*) del_keys + del_key_off[i], (size_t) (del_key_off[i + 1] - del_key_off[i])));
    return BFQ_OK;
}

int32_t bfq_index_commit(bfq_index* h) {
    if (!h) return fail(BFQ_E_INVALID, "handle is NULL");
    // The rebuild runs under the staging lock only: matches keep running on the previous snapshot meanwhile and the
    // new one is swapped in under the snapshot lock at the very end (two snapshots are resident for a moment).
    std::lock_guard<std::mutex> gs(h->stage_mu);
    CUDA_TRY(cudaSetDevice(h->device));
    const KVBlob& kv = h->staging.materialize();
    FlatIndex flat;
    std::string err;
    if (!build_flat_index(kv, &flat, &err)) return fail(BFQ_E_INVALID, err);
    DevBuf<Slot> n_slots, n_roots;
    DevBuf<uint32_t> n_segs, n_pfxP, n_pfxG;
    DevBuf<uint8_t> n_rkind, n_tags;
    auto drop_new = [&]() { n_slots.release(); n_roots.release(); n_segs.release(); n_pfxP.release(); n_pfxG.release(); n_rkind.release(); n_tags.release(); };
#define COMMIT_TRY(expr)                                                                            \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            drop_new();                                                                             \
            return fail(BFQ_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
        }                                                                                           \
    } while (0)
    COMMIT_TRY(n_slots.reserve(flat.slots.size()));
    COMMIT_TRY(n_tags.reserve(flat.tags.size()));
    COMMIT_TRY(n_roots.reserve(std::max<size_t>(flat.roots.size(), 1)));
    COMMIT_TRY(n_segs.reserve(flat.segs.size()));
    COMMIT_TRY(n_rkind.reserve(std::max<size_t>(flat.rkind.size(), 1)));
    COMMIT_TRY(n_pfxP.reserve(flat.pfx_persistent.size()));
    COMMIT_TRY(n_pfxG.reserve(flat.pfx_group.size()));
    COMMIT_TRY(cudaMemcpy(n_slots.p, flat.slots.data(), flat.slots.size() * sizeof(Slot), cudaMemcpyHostToDevice));
    COMMIT_TRY(cudaMemcpy(n_tags.p, flat.tags.data(), flat.tags.size(), cudaMemcpyHostToDevice));
    if (!flat.roots.empty())
        COMMIT_TRY(cudaMemcpy(n_roots.p, flat.roots.data(), flat.roots.size() * sizeof(Slot), cudaMemcpyHostToDevice));
    COMMIT_TRY(cudaMemcpy(n_segs.p, flat.segs.data(), flat.segs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (!flat.rkind.empty())
        COMMIT_TRY(cudaMemcpy(n_rkind.p, flat.rkind.data(), flat.rkind.size(), cudaMemcpyHostToDevice));
    COMMIT_TRY(cudaMemcpy(n_pfxP.p, flat.pfx_persistent.data(), flat.pfx_persistent.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    COMMIT_TRY(cudaMemcpy(n_pfxG.p, flat.pfx_group.data(), flat.pfx_group.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
#undef COMMIT_TRY
    // the host keeps only what it needs after the upload
    flat.slots.clear();
    flat.slots.shrink_to_fit();
    flat.tags.clear();
    flat.tags.shrink_to_fit();
    flat.roots.clear();
    flat.pfx_persistent.clear();
    flat.pfx_persistent.shrink_to_fit();
    flat.pfx_group.clear();
    flat.pfx_group.shrink_to_fit();
    KVBlob committed = kv;   // snapshot of the raw KV for bfq_route_lookup
    // ---- swap under the snapshot lock (no match is in flight while we hold it)
    std::unique_lock<std::mutex> g(h->mu);
    std::swap(h->d_slots, n_slots);
    std::swap(h->d_tags, n_tags);
    std::swap(h->d_roots, n_roots);
    std::swap(h->d_segs, n_segs);
    std::swap(h->d_rkind, n_rkind);
    std::swap(h->d_pfxP, n_pfxP);
    std::swap(h->d_pfxG, n_pfxG);
    h->flat = std::move(flat);
    h->committed = std::move(committed);
    h->have_snapshot = true;
    h->tenant_tab_valid = false;
    // Keep the tag array resident in L2 (persisting access window): every lookup starts with a tag read and the
    // array (~1/64 of the table) competes for L2 with the streaming slot traffic. BFQ_L2PERSIST=0 disables.
    {
        const char* e = getenv("BFQ_L2PERSIST");
        h->l2_window_bytes = 0;
        if (!e || atoi(e) != 0) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
            size_t want = std::min<size_t>(h->d_tags.bytes(), std::min<size_t>((size_t) max_persist, (size_t) max_window));
            if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) h->l2_window_bytes = want;
            cudaGetLastError();
        }
    }
    g.unlock();
    drop_new();   // the previous snapshot's buffers
    return BFQ_OK;
}

int32_t bfq_index_stats(bfq_index* h, int64_t* stats, int32_t n) {
    if (!h || !stats) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    const int64_t dev_bytes = (int64_t) (h->d_slots.bytes() + h->d_tags.bytes() + h->d_roots.bytes() + h->d_segs.bytes() + h->d_rkind.bytes() +
                                         h->d_pfxP.bytes() + h->d_pfxG.bytes());
    const int64_t v[12] = {h->flat.n_routes, (int64_t) h->flat.tenant_ordinal.size(), h->flat.n_nodes, (int64_t) h->flat.n_slots,
                           dev_bytes, h->flat.max_nodes_per_depth, h->launches, h->overflow_topics, h->flagged_topics,
                           h->flat.n_multi, h->flat.n_cont_chunks, h->deferred_topics};
    for (int32_t i = 0; i < n && i < 12; i++) stats[i] = v[i];
    return BFQ_OK;
}

int32_t bfq_host_build_stats(const uint8_t* keys, const int64_t* key_off, const uint8_t* vals, const int64_t* val_off,
                             int64_t n, int64_t* stats, int32_t n_stats) {
    if (n < 0 || !stats) return fail(BFQ_E_INVALID, "bad argument");
    Staging st;
    std::string err;
    const auto t0 = std::chrono::steady_clock::now();
    if (!st.load(keys, key_off, vals, val_off, n, &err)) return fail(BFQ_E_INVALID, err);
    const KVBlob& snapshot = st.materialize();
    const auto t1 = std::chrono::steady_clock::now();
    FlatIndex flat;
    if (!build_flat_index(snapshot, &flat, &err)) return fail(BFQ_E_INVALID, err);
    const auto t2 = std::chrono::steady_clock::now();
    // self-check: every placed node is found again from its parent's record the way the kernels look it up
    {
        EdgeTable t;
        t.slots = std::move(flat.slots);
        t.tags = std::move(flat.tags);
        t.n_blocks = flat.n_blocks;
        int64_t used = 0;
        for (uint32_t s = 0; s < flat.n_slots; s++) {
            const Slot& sl = t.slots[s];
            if (sl.w[W_PARENT] == EMPTY_PARENT) continue;
            used++;
            const uint32_t pid = sl.w[W_PARENT];
            const Slot& pr = pid >= ROOT_BASE ? flat.roots[pid - ROOT_BASE] : t.slots[pid];
            if (sl.w[W_LEN] == LEN_PLUS) {
                if (pr.w[W_PLUS] != s) return fail(BFQ_E_STATE, "'+' child is not linked from its parent");
                continue;
            }
            const uint32_t meta = pr.w[W_META];
            if (!(meta & FLAG_HAS_EXACT)) return fail(BFQ_E_STATE, "parent of an exact child lacks HAS_EXACT");
            uint32_t found;
            if (meta & FLAG_BIG) {
                found = t.find(pid, sl.w[W_LEN], &sl.w[W_TOK]);
            } else {
                const uint32_t lg = meta_log2size(meta), sd = meta >> 16, t32 = fold32(token_hash(sl.w[W_LEN], &sl.w[W_TOK]));
                if (lg == 0 && (t32 & 0xFFFFu) != sd) return fail(BFQ_E_STATE, "single-child fingerprint mismatch");
                found = pr.w[W_CHILD_BASE] + (lg ? child_index(t32, sd, lg) : 0u);
            }
            if (found != s) return fail(BFQ_E_STATE, "child lookup does not find a placed node");
        }
        if (used + (int64_t) flat.roots.size() != flat.n_nodes) return fail(BFQ_E_STATE, "node count mismatch");
    }
    const int64_t v[8] = {flat.n_routes, (int64_t) flat.tenant_ordinal.size(), flat.n_nodes, (int64_t) flat.n_slots,
                          flat.max_nodes_per_depth, flat.max_tenant_nodes, flat.n_multi, flat.n_cont_chunks};
    for (int32_t i = 0; i < n_stats && i < 8; i++) stats[i] = v[i];
    if (n_stats > 8) stats[8] = flat.overflowed_blocks;
    for (int32_t i = 9; i < n_stats && i < 9 + 5; i++) stats[i] = flat.child_hist[i - 9];
    if (n_stats > 14) stats[14] = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();   // staging
    if (n_stats > 15) stats[15] = std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();   // flatten
    return BFQ_OK;
}

int32_t bfq_index_last_kernel_ms(bfq_index* h, double* ms) {
    if (!h || !ms) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    *ms = h->last_kernel_ms;
    return BFQ_OK;
}

int32_t bfq_route_lookup(bfq_index* h, int64_t rank, uint8_t* key_out, int64_t key_cap, int64_t* key_len,
                         uint8_t* val_out, int64_t val_cap, int64_t* val_len) {
    if (!h) return fail(BFQ_E_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->have_snapshot) return fail(BFQ_E_STATE, "no committed snapshot");
    if (rank < 0 || rank >= h->committed.n()) return fail(BFQ_E_RANGE, "rank out of range");
    sv k = h->committed.key(rank), v = h->committed.val(rank);
    if (key_len) *key_len = (int64_t) k.size();
    if (val_len) *val_len = (int64_t) v.size();
    if (key_out && (int64_t) k.size() <= key_cap) memcpy(key_out, k.data(), k.size());
    if (val_out && (int64_t) v.size() <= val_cap) memcpy(val_out, v.data(), v.size());
    return BFQ_OK;
}

int32_t bfq_route_kind(bfq_index* h, int64_t rank, int32_t* kind) {
    if (!h || !kind) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->have_snapshot) return fail(BFQ_E_STATE, "no committed snapshot");
    if (rank < 0 || rank >= (int64_t) h->flat.rkind.size()) return fail(BFQ_E_RANGE, "rank out of range");
    *kind = h->flat.rkind[(size_t) rank];
    return BFQ_OK;
}

int32_t bfq_route_kinds(bfq_index* h, const int64_t* ranks, int64_t n, uint8_t* kinds_out) {
    if (!h || n < 0 || (n > 0 && (!ranks || !kinds_out))) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->have_snapshot) return fail(BFQ_E_STATE, "no committed snapshot");
    for (int64_t i = 0; i < n; i++) {
        if (ranks[i] < 0 || ranks[i] >= (int64_t) h->flat.rkind.size()) return fail(BFQ_E_RANGE, "rank out of range");
        kinds_out[i] = h->flat.rkind[(size_t) ranks[i]];
    }
    return BFQ_OK;
}

int32_t bfq_match(bfq_index* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                  const uint8_t* topics, const int64_t* topic_off, const int32_t* topic_tenant, int64_t n,
                  const int32_t* max_pfanout, const int32_t* max_gfanout, bfq_result** out) {
    if (!h || !out || n < 0 || n_tenants < 0) return fail(BFQ_E_INVALID, "bad argument");
    if (n > 0 && (!topics || !topic_off || !topic_tenant || !tenants || !tenant_off)) return fail(BFQ_E_INVALID, "NULL input");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->have_snapshot) return fail(BFQ_E_STATE, "bfq_match before the first bfq_index_commit");
    // topic_tenant[i] is range-checked on the device (an index outside [0, n_tenants) yields an empty result)
    CUDA_TRY(cudaSetDevice(h->device));
    auto t0 = std::chrono::steady_clock::now();
    const size_t nn = (size_t) std::max<int64_t>(n, 1);
    const int64_t blob_e = n ? topic_off[n] : 0;
    CUDA_TRY(h->d_topics.reserve((size_t) std::max<int64_t>(blob_e, 1) + 64));
    CUDA_TRY(h->d_topic_off.reserve(nn + 1));
    CUDA_TRY(h->d_topic_tenant.reserve(nn));
    CUDA_TRY(h->d_cnt.reserve(nn));
    CUDA_TRY(h->d_new_begin.reserve(nn));
    CUDA_TRY(h->h_span_begin.reserve(nn));
    CUDA_TRY(h->h_span_count.reserve(nn));
    CUDA_TRY(h->h_route_count.reserve(nn));

    // Large batches are cut into sub-batches that flow through three streams: all H2D copies on one, the kernels +
    // compaction + D2H of consecutive sub-batches alternating on two others, so the copy of sub-batch c+1 and the
    // result read-back of c-1 overlap the kernels of c (PCIe is full duplex; the copies dominate the host path).
    int C = n >= (1 << 17) ? 4 : 1;
    CoreOut co;
    int64_t rbase = 0, tbase = 0;
    for (int attempt = 0;; attempt++) {
        if (attempt == 8) return fail(BFQ_E_STATE, "buffer sizing did not converge");
        int32_t rc = prepare_workspace(h, n, C);
        if (rc != BFQ_OK) return rc;
        CUDA_TRY(h->d_ranges_c.reserve(h->d_ranges.cap));
        const uint64_t dyn_total = h->d_ranges.cap - (uint64_t) n * INLINE_RANGES;
        const uint64_t dyn_slice = dyn_total / (uint64_t) C, thr_slice = h->d_throttled.cap / (uint64_t) C;
        size_t tmp_bytes = 0;
        {
            CompactParams q{};
            q.n_topics = (n + C - 1) / C + 1;
            q.counts = h->d_cnt.p;
            q.new_begin = h->d_new_begin.p;
            CUDA_TRY(launch_compact(q, nullptr, &tmp_bytes, h->stream, 1));
            CUDA_TRY(h->d_scan_tmp.reserve(tmp_bytes * 2 + 512));   // one scratch per compute stream
        }
        co = CoreOut();
        rbase = tbase = 0;
        // ---- H2D of every sub-batch, back to back on the copy stream
        CUDA_TRY(cudaEventRecord(h->ev[0], h->copy_stream));
        rc = resolve_tenants(h, tenants, tenant_off, n_tenants, max_pfanout, max_gfanout, h->copy_stream);
        if (rc != BFQ_OK) return rc;
        int64_t bounds[MAX_CHUNKS + 1];
        for (int c = 0; c <= C; c++) bounds[c] = n * c / C;
        for (int c = 0; c < C && n > 0; c++) {
            const int64_t b = bounds[c], e = bounds[c + 1];
            const int64_t ob = topic_off[b], oe = topic_off[e];
            CUDA_TRY(cudaMemcpyAsync(h->d_topic_off.p + b, topic_off + b, (size_t) (e - b + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->copy_stream));
            CUDA_TRY(cudaMemcpyAsync(h->d_topic_tenant.p + b, topic_tenant + b, (size_t) (e - b) * sizeof(int32_t), cudaMemcpyHostToDevice, h->copy_stream));
            CUDA_TRY(cudaMemcpyAsync(h->d_topics.p + ob, topics + ob, (size_t) (oe - ob), cudaMemcpyHostToDevice, h->copy_stream));
            CUDA_TRY(cudaEventRecord(h->ev_h2d[c], h->copy_stream));
        }
        CUDA_TRY(cudaEventRecord(h->ev[1], h->copy_stream));
        bool retry = false;
        for (int c = 0; c < C && n > 0; c++) {
            cudaStream_t st = C == 1 ? h->stream : h->work_stream[c & 1];
            CUDA_TRY(cudaStreamWaitEvent(st, h->ev_h2d[c], 0));
            SubBatch sb;
            sb.begin = bounds[c];
            sb.n = bounds[c + 1] - bounds[c];
            sb.n_total = n;
            sb.chunk = c;
            sb.dyn_off = (uint64_t) c * dyn_slice;
            sb.dyn_cap = dyn_slice;
            sb.thr_off = (uint64_t) c * thr_slice;
            sb.thr_cap = thr_slice;
            rc = match_core(h, h->d_topics.p, h->d_topic_off.p, h->d_topic_tenant.p, n_tenants, st, sb, &co);
            if (rc == BFQ_RETRY_GROW) {
                retry = true;
                break;
            }
            if (rc != BFQ_OK) {
                cudaDeviceSynchronize();
                return rc;
            }
            // ---- compaction of this sub-batch: counts + scan + total, then gather into the dense result position
            unsigned long long* d_ctr = h->d_counters.p + (size_t) c * CTR_COUNT;
            unsigned long long* hc = h->h_counters.p + (size_t) c * CTR_COUNT;
            const uint64_t region = (uint64_t) sb.begin * INLINE_RANGES + sb.dyn_off;   // this sub-batch's private slice of d_ranges_c
            CompactParams cp{};
            cp.n_topics = sb.n;
            cp.span_begin = h->d_span_begin.p + sb.begin;
            cp.span_count = h->d_span_count.p + sb.begin;
            cp.ranges = h->d_ranges.p + (uint64_t) sb.begin * INLINE_RANGES;
            cp.counts = h->d_cnt.p + sb.begin;
            cp.new_begin = h->d_new_begin.p + sb.begin;
            cp.ranges_out = h->d_ranges_c.p + region;
            cp.ranges_out_cap = (uint64_t) sb.n * INLINE_RANGES + sb.dyn_cap;
            cp.total_out = d_ctr + CTR_ROUTES;
            uint8_t* scan_tmp = h->d_scan_tmp.p + (size_t) (c & 1) * (tmp_bytes + 256) / 256 * 256;
            CUDA_TRY(launch_compact(cp, scan_tmp, &tmp_bytes, st, 1));
            CUDA_TRY(cudaMemcpyAsync(hc, d_ctr, CTR_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            const int64_t total_c = (int64_t) hc[CTR_ROUTES], thr_c = co.chunk_throttled[c];
            // host result buffers grow by reallocation: wait for the copies in flight before moving them
            if ((size_t) (rbase + total_c) > h->h_ranges.cap || (size_t) (tbase + thr_c) > h->h_throttled.cap) {
                CUDA_TRY(cudaDeviceSynchronize());
                if ((size_t) (rbase + total_c) > h->h_ranges.cap) {
                    PinBuf<uint2> nb;
                    CUDA_TRY(nb.reserve((size_t) ((rbase + total_c) * (C - c > 1 ? 2 : 1) + (1 << 16))));
                    if (rbase) memcpy(nb.p, h->h_ranges.p, (size_t) rbase * sizeof(uint2));
                    h->h_ranges.release();
                    h->h_ranges = nb;
                }
                if ((size_t) (tbase + thr_c) > h->h_throttled.cap) {
                    PinBuf<uint3> nb;
                    CUDA_TRY(nb.reserve((size_t) ((tbase + thr_c) * 2 + 1024)));
                    if (tbase) memcpy(nb.p, h->h_throttled.p, (size_t) tbase * sizeof(uint3));
                    h->h_throttled.release();
                    h->h_throttled = nb;
                }
            }
            cp.out_base = (uint32_t) rbase;
            CUDA_TRY(launch_compact(cp, scan_tmp, &tmp_bytes, st, 2));
            co.n_launches += 4;
            CUDA_TRY(cudaMemcpyAsync(h->h_span_begin.p + sb.begin, h->d_new_begin.p + sb.begin, (size_t) sb.n * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h->h_span_count.p + sb.begin, h->d_cnt.p + sb.begin, (size_t) sb.n * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h->h_route_count.p + sb.begin, h->d_route_count.p + sb.begin, (size_t) sb.n * 4, cudaMemcpyDeviceToHost, st));
            if (total_c > 0)
                CUDA_TRY(cudaMemcpyAsync(h->h_ranges.p + rbase, h->d_ranges_c.p + region, (size_t) total_c * sizeof(uint2), cudaMemcpyDeviceToHost, st));
            if (thr_c > 0)
                CUDA_TRY(cudaMemcpyAsync(h->h_throttled.p + tbase, h->d_throttled.p + sb.thr_off, (size_t) thr_c * sizeof(uint3), cudaMemcpyDeviceToHost, st));
            rbase += total_c;
            tbase += thr_c;
        }
        if (!retry) break;
        // a slice of the range / throttled buffers was too small: grow them and redo the batch un-chunked
        CUDA_TRY(cudaDeviceSynchronize());
        if (co.want_dyn) {
            const size_t want = (size_t) ((uint64_t) n * INLINE_RANGES + (co.want_dyn + co.want_dyn / 4 + 1024) * (uint64_t) C);
            if (want >= 0xFFFFFFF0ull) return fail(BFQ_E_RANGE, "more than 2^32 matched ranges in one batch; split the batch");
            CUDA_TRY(h->d_ranges.reserve(want));
        }
        if (co.want_thr) CUDA_TRY(h->d_throttled.reserve((size_t) ((co.want_thr + co.want_thr / 4 + 1024) * (uint64_t) C)));
        C = 1;
    }
    CUDA_TRY(cudaStreamSynchronize(h->work_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(h->work_stream[1]));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
    h->launches += co.n_launches;
    h->overflow_topics += co.n_overflow;
    h->deferred_topics += co.n_deferred;
    h->flagged_topics += co.n_flagged;
    h->last_n_topics = n;
    co.n_ranges = rbase;
    co.n_throttled = tbase;
    if (co.n_throttled > 1) {
        uint3* th = h->h_throttled.p;
        std::sort(th, th + co.n_throttled, [](const uint3& a, const uint3& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    }
    auto* r = new bfq_result();
    r->owner = h;
    r->n_topics = n;
    r->n_ranges = co.n_ranges;
    r->n_throttled = co.n_throttled;
    r->span_begin = h->h_span_begin.p;
    r->span_count = h->h_span_count.p;
    r->route_count = h->h_route_count.p;
    r->ranges = reinterpret_cast<const bfq_range*>(h->h_ranges.p);
    r->throttled = reinterpret_cast<const bfq_throttled*>(h->h_throttled.p);
    float a = 0;
    cudaEventElapsedTime(&a, h->ev[0], h->ev[1]);
    r->ms[0] = a;                       // H2D stream busy time (overlapped with the kernels of earlier sub-batches)
    r->ms[1] = h->last_kernel_ms;       // tier-0 kernel of the first sub-batch
    r->ms[2] = (double) C;              // number of sub-batches
    r->ms[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = r;
    return BFQ_OK;
}

int64_t bfq_result_num_topics(const bfq_result* r) { return r ? r->n_topics : 0; }
const uint32_t* bfq_result_span_begin(const bfq_result* r) { return r->span_begin; }
const uint32_t* bfq_result_span_count(const bfq_result* r) { return r->span_count; }
const uint32_t* bfq_result_route_count(const bfq_result* r) { return r->route_count; }
const bfq_range* bfq_result_ranges(const bfq_result* r, int64_t* n_ranges) {
    if (n_ranges) *n_ranges = r->n_ranges;
    return r->ranges;
}
const bfq_throttled* bfq_result_throttled(const bfq_result* r, int64_t* n_throttled) {
    if (n_throttled) *n_throttled = r->n_throttled;
    return r->throttled;
}

int64_t bfq_result_expand(const bfq_result* r, int64_t* offsets, int64_t* ranks, int64_t rank_cap) {
    if (!r || !offsets) return BFQ_E_INVALID;
    const std::vector<uint32_t>& segs = r->owner->flat.segs;
    int64_t total = 0;
    int64_t ti = 0;  // cursor into the (topic, rank)-sorted throttled list
    std::vector<int64_t> tmp;
    for (int64_t t = 0; t < r->n_topics; t++) {
        offsets[t] = total;
        tmp.clear();
        const uint32_t b = r->span_begin[t], c = r->span_count[t];
        for (uint32_t j = 0; j < c; j++) {
            const bfq_range rg = r->ranges[b + j];
            if (rg.count & RANGE_MULTI) {
                const uint32_t nseg = segs[2 * (size_t) rg.first];
                for (uint32_t s = 0; s < nseg; s++) {
                    const uint32_t f = segs[2 * ((size_t) rg.first + 1 + s)], n = segs[2 * ((size_t) rg.first + 1 + s) + 1];
                    for (uint32_t x = 0; x < n; x++) tmp.push_back((int64_t) f + x);
                }
            } else {
                for (uint32_t x = 0; x < rg.count; x++) tmp.push_back((int64_t) rg.first + x);
            }
        }
        std::sort(tmp.begin(), tmp.end());
        while (ti < r->n_throttled && r->throttled[ti].topic < (uint32_t) t) ti++;
        for (int64_t x : tmp) {
            while (ti < r->n_throttled && r->throttled[ti].topic == (uint32_t) t && (int64_t) r->throttled[ti].rank < x) ti++;
            if (ti < r->n_throttled && r->throttled[ti].topic == (uint32_t) t && (int64_t) r->throttled[ti].rank == x) continue;
            if (ranks && total < rank_cap) ranks[total] = x;
            total++;
        }
    }
    offsets[r->n_topics] = total;
    return total;
}

int32_t bfq_result_timings(const bfq_result* r, double* ms, int32_t n) {
    if (!r || !ms) return fail(BFQ_E_INVALID, "bad argument");
    for (int32_t i = 0; i < n && i < 4; i++) ms[i] = r->ms[i];
    return BFQ_OK;
}
void bfq_result_free(bfq_result* r) { delete r; }

int32_t bfq_match_device(bfq_index* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                         const uint8_t* d_topics, const int64_t* d_topic_off, const int32_t* d_topic_tenant, int64_t n,
                         const int32_t* max_pfanout, const int32_t* max_gfanout, void* stream, bfq_device_result* out) {
    if (!h || !out || n < 0 || n_tenants < 0) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->have_snapshot) return fail(BFQ_E_STATE, "bfq_match_device before the first bfq_index_commit");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t) stream;
    int32_t rc = resolve_tenants(h, tenants, tenant_off, n_tenants, max_pfanout, max_gfanout, st);
    if (rc != BFQ_OK) return rc;
    CoreOut co;
    rc = match_whole(h, d_topics, d_topic_off, d_topic_tenant, n, n_tenants, st, &co);
    if (rc != BFQ_OK) return rc;
    h->last_topic_tenant = d_topic_tenant;
    h->last_n_tenants = n_tenants;
    h->last_n_flagged = co.n_flagged;
    out->d_span_begin = h->d_span_begin.p;
    out->d_span_count = h->d_span_count.p;
    out->d_route_count = h->d_route_count.p;
    out->d_ranges = reinterpret_cast<const bfq_range*>(h->d_ranges.p);
    out->d_throttled = reinterpret_cast<const bfq_throttled*>(h->d_throttled.p);
    out->n_ranges = (int64_t) ((uint64_t) n * INLINE_RANGES) + co.n_ranges;   // extent of the sparse range array
    out->n_throttled = co.n_throttled;
    out->n_routes = -1;
    out->n_overflow_topics = co.n_overflow;
    out->n_flagged_topics = co.n_flagged;
    out->n_launches = co.n_launches;
    return BFQ_OK;
}

int32_t bfq_expand_device(bfq_index* h, int64_t n_topics, int64_t* d_offsets, int64_t* d_ranks, int64_t rank_cap, void* stream,
                          int64_t* n_ranks) {
    if (!h || !d_offsets || n_topics < 0) return fail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->have_snapshot || n_topics != h->last_n_topics || !h->last_topic_tenant)
        return fail(BFQ_E_STATE, "bfq_expand_device must follow a bfq_match_device of the same batch");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t) stream;
    const size_t nt = (size_t) std::max(h->last_n_tenants, 1);
    CUDA_TRY(h->d_exp_counts.reserve((size_t) n_topics + 1));
    ExpandParams p{};
    p.n_topics = n_topics;
    p.span_begin = h->d_span_begin.p;
    p.span_count = h->d_span_count.p;
    p.route_count = h->d_route_count.p;
    p.kept_count = h->d_kept.p;
    p.ranges = h->d_ranges.p;
    p.segs = h->d_segs.p;
    p.counts = h->d_exp_counts.p;
    p.offsets = d_offsets;
    p.ranks = d_ranks;
    p.rank_cap = d_ranks ? rank_cap : 0;
    p.flagged_list = h->d_flagged.p;
    p.n_flagged = h->last_n_flagged;
    p.topic_tenant = h->last_topic_tenant;
    p.max_pfanout = h->d_tenant_tab.p + nt;
    p.max_gfanout = h->d_tenant_tab.p + 2 * nt;
    p.rkind = h->d_rkind.p;
    p.pfx_persistent = h->d_pfxP.p;
    p.pfx_group = h->d_pfxG.p;
    size_t tmp_bytes = 0;
    CUDA_TRY(launch_expand(p, nullptr, &tmp_bytes, st, 1));
    CUDA_TRY(h->d_scan_tmp.reserve(tmp_bytes + 256));
    CUDA_TRY(launch_expand(p, h->d_scan_tmp.p, &tmp_bytes, st, 1));
    long long total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, d_offsets + n_topics, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (n_ranks) *n_ranks = (int64_t) total;
    h->launches += 2;
    if (d_ranks && total <= rank_cap) {
        CUDA_TRY(launch_expand(p, h->d_scan_tmp.p, &tmp_bytes, st, 2));
        h->launches += 2;
    }
    return BFQ_OK;
}

// ---------------------------------------------------------------- codec exports
int64_t bfq_receiver_url(int32_t sub_broker_id, const uint8_t* receiver_id, int64_t rn, const uint8_t* deliverer_key,
                         int64_t dn, uint8_t* out, int64_t cap) {
    return emit_bytes(make_receiver_url(sub_broker_id, sv((const char*) receiver_id, (size_t) rn), sv((const char*) deliverer_key, (size_t) dn)), out, cap);
}
int64_t bfq_route_key(const uint8_t* tenant, int64_t tn, const uint8_t* tf, int64_t fn, const uint8_t* url, int64_t un,
                      uint8_t* out, int64_t cap) {
    return emit_bytes(make_route_key(sv((const char*) tenant, (size_t) tn), sv((const char*) tf, (size_t) fn), sv((const char*) url, (size_t) un)), out, cap);
}
int64_t bfq_tenant_begin_key(const uint8_t* tenant, int64_t tn, uint8_t* out, int64_t cap) {
    return emit_bytes(make_tenant_begin_key(sv((const char*) tenant, (size_t) tn)), out, cap);
}
int32_t bfq_is_valid_topic(const uint8_t* topic, int64_t n, int32_t a, int32_t b, int32_t c) {
    return is_valid_topic(sv((const char*) topic, (size_t) n), a, b, c) ? 1 : 0;
}
int32_t bfq_is_valid_topic_filter(const uint8_t* tf, int64_t n, int32_t a, int32_t b, int32_t c) {
    return is_valid_topic_filter(sv((const char*) tf, (size_t) n), a, b, c) ? 1 : 0;
}

}  // extern "C"
