// match_kernels.cu — forward match on sm_100a: a batch of publish topics against the flattened filter tries.
//
// Replaces the hot loop of TenantRouteMatcher.matchAll
// (bifromq-dist/bifromq-dist-worker/src/main/java/org/apache/bifromq/dist/worker/cache/TenantRouteMatcher.java:96-156)
// with a level-synchronous frontier walk: ONE WARP PER TOPIC.
//   * the topic bytes are staged in shared memory, '/' boundaries found with a warp ballot (device-side
//     TopicUtil.parse, bifromq-util/.../TopicUtil.java:206-225 — empty levels are kept);
//   * the frontier (trie nodes whose path matches the consumed prefix) lives in shared memory, one node
//     per lane; every lane probes the 64-byte hash slot of its node's exact child with four LDG.128 and,
//     independently, loads the '+' child record — all loads of a level are in flight together;
//   * '#' children are inlined in their parent record, so "prefix/#" (incl. the parent-level match) is
//     emitted at discovery without another access; the '$' rule masks the root's '+' and '#' only;
//   * results are emitted as RANGES of route ranks (one per matched filter) compacted with ballot/popc
//     into a shared staging area and flushed with one atomicAdd per topic.
// Tier 2 (kBig): the rare topic whose frontier or range count outgrows the shared buffers is re-run by the
// same code with per-warp buffers in global memory sized from the index statistics — never truncated.
#include "match_kernels.cuh"

namespace bfq {

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int STAGE_BYTES = 256;
constexpr uint32_t FR_CAP = 64;
constexpr uint32_t RG_CAP = 48;
constexpr uint32_t NONE31 = 0x7FFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct WarpSmem {
    uint8_t stage[STAGE_BYTES];
    uint32_t keyw[8];
    uint2 fr[2][FR_CAP];
    uint2 rg[RG_CAP];
};

__device__ __forceinline__ void load_slot(const Slot* s, uint32_t (&w)[16]) {
    const uint4* p = reinterpret_cast<const uint4*>(s);
    uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    w[12] = d.x; w[13] = d.y; w[14] = d.z; w[15] = d.w;
}

// open-addressing lookup of the edge (parent, lenw, k[0..5]); on success w holds the child record
__device__ __forceinline__ bool probe(const Slot* slots, uint32_t n_slots, uint32_t parent, uint32_t lenw,
                                      const uint32_t (&k)[6], uint64_t tokh, uint32_t (&w)[16], uint32_t& slot) {
    uint32_t s = home_slot(tokh, parent, n_slots);
    while (true) {
        load_slot(slots + s, w);
        if (w[W_PARENT] == EMPTY_PARENT) return false;
        if (w[W_PARENT] == parent && w[W_LEN] == lenw && w[2] == k[0] && w[3] == k[1] && w[4] == k[2] &&
            w[5] == k[3] && w[6] == k[4] && w[7] == k[5]) {
            slot = s;
            return true;
        }
        s = s + 1 == n_slots ? 0 : s + 1;
    }
}

__device__ __forceinline__ uint64_t caps_value(uint32_t c16) { return c16 == 0xFFFFu ? (1ull << 32) : (uint64_t) c16; }

template <bool kBig>
__device__ __forceinline__ void match_one(const MatchParams& p, WarpSmem& ws, uint32_t t, int lane, uint2* fr_a, uint2* fr_b,
                                          uint2* rg, uint32_t capF, uint32_t capR) {
    const int64_t tb = p.topic_off[t];
    const int len = (int) (p.topic_off[t + 1] - tb);
    const uint8_t* src = p.topics + tb;
    const bool staged = len <= STAGE_BYTES;
    __syncwarp();
    if (staged)
        for (int i = lane; i < len; i += 32) ws.stage[i] = src[i];
    __syncwarp();
    auto byte_at = [&](int i) -> uint32_t { return staged ? (uint32_t) ws.stage[i] : (uint32_t) src[i]; };

    const int tenant = p.topic_tenant[t];
    const int root_ord = p.tenant_root[tenant];
    uint32_t n_rg = 0;
    uint32_t acc_r = 0;
    uint64_t acc_p = 0, acc_g = 0;
    bool overflow = false;

    // warp-compacted append of one optional range per lane
    auto emit = [&](bool valid, uint32_t first, uint32_t count, bool multi, uint32_t caps) {
        const unsigned m = __ballot_sync(FULL, valid);
        if (m == 0) return;
        if (valid) {
            const uint32_t idx = n_rg + __popc(m & ((1u << lane) - 1));
            if (idx < capR) rg[idx] = make_uint2(first, multi ? (count | RANGE_MULTI) : count);
            acc_r += count;
            acc_p += caps_value(caps & 0xFFFFu);
            acc_g += caps_value(caps >> 16);
        }
        n_rg += __popc(m);
        if (n_rg > capR) overflow = true;
    };

    if (root_ord >= 0) {
        uint32_t rw[16];
        load_slot(p.roots + root_ord, rw);
        const bool sys = len > 0 && byte_at(0) == '$';
        // "#" at level 0 matches every non-'$' topic
        emit(lane == 0 && !sys && rw[W_HASH_COUNT] > 0, rw[W_HASH_FIRST], rw[W_HASH_COUNT], rw[W_FLAGS] & FLAG_HASH_MULTI,
             rw[W_HASH_CAPS]);
        uint2* fr_cur = fr_a;
        uint2* fr_next = fr_b;
        uint32_t n_fr = 0;
        {
            const uint32_t plus = (sys || rw[W_PLUS] == NONE) ? NONE31 : rw[W_PLUS];   // '+' at level 0 skips '$' topics
            const uint32_t has_exact = rw[W_FLAGS] & FLAG_HAS_EXACT;
            if (has_exact || plus != NONE31) {
                if (lane == 0) fr_cur[0] = make_uint2(ROOT_BASE + (uint32_t) root_ord, plus | (has_exact ? 0x80000000u : 0u));
                n_fr = 1;
            }
        }
        int pos = 0;
        while (n_fr > 0 && !overflow) {
            // ---- next token [pos, e)
            int e = len;
            for (int b = pos; b < len; b += 32) {
                const int i = b + lane;
                const unsigned m = __ballot_sync(FULL, i < len && byte_at(i) == '/');
                if (m) {
                    e = b + __ffs(m) - 1;
                    break;
                }
            }
            const bool last = e == len;
            const int tlen = e - pos;
            const int nchunks = tlen <= (int) TOKEN_BYTES ? 1 : (tlen + (int) TOKEN_BYTES - 1) / (int) TOKEN_BYTES;
            uint32_t n_next = 0;
            __syncwarp();
            for (uint32_t base = 0; base < n_fr && !overflow; base += 32) {
                const bool active = base + lane < n_fr;
                const uint2 fe = active ? fr_cur[base + lane] : make_uint2(0u, NONE31);
                const uint32_t plus = fe.y & NONE31;
                // '+' child record: independent of the token, issue its load first
                const bool has_plus = active && plus != NONE31;
                uint32_t pw[16];
                if (has_plus) load_slot(p.slots + plus, pw);
                // exact child: one probe per 24-byte chunk of the token (one chunk unless the level is > 24 B)
                bool alive = active && (fe.y >> 31);
                uint32_t node = fe.x, cid = 0;
                uint32_t cw[16];
                for (int c = 0; c < nchunks; c++) {
                    const int cpos = pos + c * (int) TOKEN_BYTES;
                    const int cend = min(e, cpos + (int) TOKEN_BYTES);
                    const uint32_t lenw = c == nchunks - 1 ? (uint32_t) tlen : (LEN_CONT | (uint32_t) c);
                    __syncwarp();
                    if (lane < (int) TOKEN_WORDS) {
                        uint32_t v = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int idx = cpos + lane * 4 + b;
                            if (idx < cend) v |= byte_at(idx) << (8 * b);
                        }
                        ws.keyw[lane] = v;
                    }
                    __syncwarp();
                    uint32_t k[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) k[j] = ws.keyw[j];
                    const uint64_t tokh = token_hash(lenw, k);
                    if (alive) {
                        alive = probe(p.slots, p.n_slots, node, lenw, k, tokh, cw, cid);
                        node = cid;
                    }
                }
                // ---- emit ranges of the discovered children and build the next frontier
                emit(alive && cw[W_HASH_COUNT] > 0, cw[W_HASH_FIRST], cw[W_HASH_COUNT], alive && (cw[W_FLAGS] & FLAG_HASH_MULTI),
                     cw[W_HASH_CAPS]);
                emit(has_plus && pw[W_HASH_COUNT] > 0, pw[W_HASH_FIRST], pw[W_HASH_COUNT],
                     has_plus && (pw[W_FLAGS] & FLAG_HASH_MULTI), pw[W_HASH_CAPS]);
                if (last) {
                    emit(alive && cw[W_OWN_COUNT] > 0, cw[W_OWN_FIRST], cw[W_OWN_COUNT], alive && (cw[W_FLAGS] & FLAG_OWN_MULTI),
                         cw[W_OWN_CAPS]);
                    emit(has_plus && pw[W_OWN_COUNT] > 0, pw[W_OWN_FIRST], pw[W_OWN_COUNT],
                         has_plus && (pw[W_FLAGS] & FLAG_OWN_MULTI), pw[W_OWN_CAPS]);
                } else {
                    const bool push_c = alive && ((cw[W_FLAGS] & FLAG_HAS_EXACT) || cw[W_PLUS] != NONE);
                    const bool push_p = has_plus && ((pw[W_FLAGS] & FLAG_HAS_EXACT) || pw[W_PLUS] != NONE);
                    const unsigned mc = __ballot_sync(FULL, push_c);
                    const unsigned mp = __ballot_sync(FULL, push_p);
                    const uint32_t lt = (1u << lane) - 1;
                    if (push_c) {
                        const uint32_t idx = n_next + __popc(mc & lt);
                        if (idx < capF)
                            fr_next[idx] = make_uint2(cid, (cw[W_PLUS] == NONE ? NONE31 : cw[W_PLUS]) |
                                                               ((cw[W_FLAGS] & FLAG_HAS_EXACT) ? 0x80000000u : 0u));
                    }
                    n_next += __popc(mc);
                    if (push_p) {
                        const uint32_t idx = n_next + __popc(mp & lt);
                        if (idx < capF)
                            fr_next[idx] = make_uint2(plus, (pw[W_PLUS] == NONE ? NONE31 : pw[W_PLUS]) |
                                                                ((pw[W_FLAGS] & FLAG_HAS_EXACT) ? 0x80000000u : 0u));
                    }
                    n_next += __popc(mp);
                    if (n_next > capF) overflow = true;
                }
            }
            __syncwarp();
            uint2* tmp = fr_cur;
            fr_cur = fr_next;
            fr_next = tmp;
            n_fr = n_next;
            pos = e + 1;
            if (last) break;
        }
    }

    // ---- warp totals
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_r += __shfl_xor_sync(FULL, acc_r, o);
        acc_p += __shfl_xor_sync(FULL, acc_p, o);
        acc_g += __shfl_xor_sync(FULL, acc_g, o);
    }
    if (overflow) {
        if (lane == 0) {
            if (!kBig) {
                const unsigned long long idx = atomicAdd(&p.counters[CTR_OVERFLOW], 1ull);
                p.overflow_list[idx] = t;
                p.span_begin[t] = 0;
                p.span_count[t] = SPAN_OVERFLOW;
                p.route_count[t] = 0;
            } else {
                atomicAdd(&p.counters[CTR_ERROR], 1ull);
                p.span_begin[t] = 0;
                p.span_count[t] = 0;
                p.route_count[t] = 0;
            }
        }
        return;
    }
    const int maxP = p.max_pfanout[tenant], maxG = p.max_gfanout[tenant];
    // ranks are < 2^31-1, so a cap of INT_MAX can never be exceeded
    const bool flag_p = maxP != 0x7FFFFFFF && acc_p > (uint64_t) (maxP < 0 ? 0 : maxP);
    const bool flag_g = maxG != 0x7FFFFFFF && acc_g > (uint64_t) (maxG < 0 ? 0 : maxG);
    const bool flagged = flag_p || flag_g;
    unsigned long long base = 0;
    if (n_rg > 0) {
        if (lane == 0) base = atomicAdd(&p.counters[CTR_RANGES], (unsigned long long) n_rg);
        base = __shfl_sync(FULL, base, 0);
        if (base + n_rg <= p.ranges_cap)
            for (uint32_t i = lane; i < n_rg; i += 32) p.ranges[base + i] = rg[i];
    }
    if (lane == 0) {
        p.span_begin[t] = (uint32_t) base;
        p.span_count[t] = n_rg | (flagged ? SPAN_FLAGGED : 0u);
        p.route_count[t] = acc_r;
        if (flagged) {
            const unsigned long long idx = atomicAdd(&p.counters[CTR_FLAGGED], 1ull);
            p.flagged_list[idx] = t;
        }
    }
}

template <bool kBig>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) match_topics_kernel(const MatchParams p) {
    __shared__ WarpSmem sm[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmem& ws = sm[wid];
    const int64_t gw = (int64_t) blockIdx.x * WARPS_PER_CTA + wid;
    const int64_t nw = (int64_t) gridDim.x * WARPS_PER_CTA;
    if (kBig) {
        uint2* basep = p.scratch + (uint64_t) gw * (2 * p.scratch_frontier_cap + p.scratch_ranges_cap);
        uint2* fr_a = basep;
        uint2* fr_b = basep + p.scratch_frontier_cap;
        uint2* rg = basep + 2 * p.scratch_frontier_cap;
        const uint32_t capF = (uint32_t) min((uint64_t) 0x3FFFFFFFull, p.scratch_frontier_cap);
        const uint32_t capR = (uint32_t) min((uint64_t) SPAN_COUNT_MASK, p.scratch_ranges_cap);
        for (int64_t it = gw; it < p.n_work; it += nw) match_one<true>(p, ws, p.work_list[it], lane, fr_a, fr_b, rg, capF, capR);
    } else {
        for (int64_t it = gw; it < p.n_topics; it += nw)
            match_one<false>(p, ws, (uint32_t) it, lane, ws.fr[0], ws.fr[1], ws.rg, FR_CAP, RG_CAP);
    }
}

// ------------------------------------------------------------------------------------------------ caps
// Fan-out caps of MatchedRoutes (bifromq-dist/bifromq-dist-worker/.../cache/MatchedRoutes.java:87-141): among a
// topic's matched routes taken in KV (rank) order, only the first maxPersistentFanout persistent normal routes
// (subBrokerId == 1) and the first maxGroupFanout group routes survive; every later one is dropped and reported.
// One CTA per flagged topic. A route's index among the matched persistent routes is
//   (# persistent routes in matched segments that start before its segment) + (# persistent before it in its segment),
// both read off exclusive prefix counts over the rank space, so no sort of the ranges is needed.
constexpr int CAPS_THREADS = 128;

struct SegIter {  // iterate the segments {first,count} behind a topic's ranges (resolving multi-segment ranges)
    const uint2* ranges;
    const uint32_t* segs;
    uint32_t n;
    template <typename F>
    __device__ __forceinline__ void for_range(uint32_t j, F&& f) const {
        const uint2 r = ranges[j];
        if (r.y & RANGE_MULTI) {
            const uint32_t nseg = segs[2 * (uint64_t) r.x];
            for (uint32_t s = 0; s < nseg; s++) f(segs[2 * ((uint64_t) r.x + 1 + s)], segs[2 * ((uint64_t) r.x + 1 + s) + 1]);
        } else {
            f(r.x, r.y);
        }
    }
};

__global__ void __launch_bounds__(CAPS_THREADS) caps_kernel(const CapsParams p) {
    const uint32_t t = p.flagged_list[blockIdx.x];
    const int tenant = p.topic_tenant[t];
    const uint64_t maxP = (uint64_t) max(p.max_pfanout[tenant], 0), maxG = (uint64_t) max(p.max_gfanout[tenant], 0);
    SegIter it{p.ranges + p.span_begin[t], p.segs, p.span_count[t] & SPAN_COUNT_MASK};
    __shared__ unsigned long long kept;
    if (threadIdx.x == 0) kept = 0;
    __syncthreads();
    unsigned long long my_kept = 0;
    for (uint32_t j = threadIdx.x; j < it.n; j += CAPS_THREADS) {
        it.for_range(j, [&](uint32_t first, uint32_t count) {
            // persistent / group routes in matched segments that start before this one
            uint64_t baseP = 0, baseG = 0;
            for (uint32_t q = 0; q < it.n; q++)
                it.for_range(q, [&](uint32_t f2, uint32_t c2) {
                    if (f2 < first) {
                        baseP += p.pfx_persistent[f2 + c2] - p.pfx_persistent[f2];
                        baseG += p.pfx_group[f2 + c2] - p.pfx_group[f2];
                    }
                });
            const uint64_t cntP = p.pfx_persistent[first + count] - p.pfx_persistent[first];
            const uint64_t cntG = p.pfx_group[first + count] - p.pfx_group[first];
            if (baseP + cntP <= maxP && baseG + cntG <= maxG) {
                my_kept += count;
                return;
            }
            for (uint32_t r = first; r < first + count; r++) {
                const uint8_t kind = p.rkind[r];
                uint32_t drop = 0;
                if (kind == 1 && baseP + (p.pfx_persistent[r] - p.pfx_persistent[first]) >= maxP) drop = 1;
                else if (kind == 2 && baseG + (p.pfx_group[r] - p.pfx_group[first]) >= maxG) drop = 2;
                if (drop) {
                    const unsigned long long idx = atomicAdd(&p.counters[CTR_THROTTLED], 1ull);
                    if (idx < p.throttled_cap) p.throttled[idx] = make_uint3(t, r, drop);
                } else {
                    my_kept++;
                }
            }
        });
    }
    atomicAdd(&kept, my_kept);
    __syncthreads();
    if (threadIdx.x == 0 && p.kept_count) p.kept_count[t] = (uint32_t) kept;
}

}  // namespace

int match_kernel_smem_bytes() { return (int) sizeof(WarpSmem) * WARPS_PER_CTA; }

void launch_match(const MatchParams& p, bool tier2, int n_warps_tier2, cudaStream_t stream) {
    if (tier2) {
        const int ctas = (n_warps_tier2 + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        match_topics_kernel<true><<<ctas, WARPS_PER_CTA * 32, 0, stream>>>(p);
        return;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, match_topics_kernel<false>, WARPS_PER_CTA * 32, 0);
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    // persistent grid: a whole number of waves (SM count x resident CTAs per SM), grid-stride over topics
    int64_t ctas = (int64_t) sms * ctas_per_sm;
    const int64_t need = (p.n_topics + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (need < ctas) ctas = need < 1 ? 1 : need;
    match_topics_kernel<false><<<(unsigned) ctas, WARPS_PER_CTA * 32, 0, stream>>>(p);
}

void launch_caps(const CapsParams& p, cudaStream_t stream) {
    if (p.n_flagged <= 0) return;
    caps_kernel<<<(unsigned) p.n_flagged, CAPS_THREADS, 0, stream>>>(p);
}

}  // namespace bfq
