// rmatch_kernels.cu — inverse match on sm_100a: a batch of topic FILTERS against an index of TOPICS.
//
// Replaces RetainTopicIndex.match / TopicIndex.match
//   (bifromq-retain/bifromq-retain-store/src/main/java/org/apache/bifromq/retain/store/index/RetainTopicIndex.java:36-138,
//    bifromq-dist/bifromq-dist-worker/src/main/java/org/apache/bifromq/dist/worker/TopicIndex.java:40-155,
//    traversal bifromq-util/src/main/java/org/apache/bifromq/util/index/TopicLevelTrie.java:190-249).
// Same machinery as the forward kernel with the roles swapped: ONE WARP PER FILTER walks the topic trie.
//
// Layout (HBM): the per-tenant topic tries are numbered in one global BFS, so
//   * the children of a node — and the children of any RUN of consecutive nodes of one depth — are one
//     contiguous id interval: a '+' level maps a frontier interval to ONE interval with two record loads,
//     it never explodes the frontier;
//   * topics get two ranks: their DFS (pre-order) rank, in which a whole subtree is a contiguous range
//     ("prefix/#" = one range per frontier node), and their BFS rank, in which the topics ending at a run of
//     consecutive nodes are contiguous (a final '+' = one range per frontier interval).
//   * exact levels use the same 64-byte (parent, token) hash slots as the forward index (trie_layout.h),
//     payload word W_PLUS holding the child's BFS id.
// Results are emitted as ranges {space|first, count}; a second kernel maps them to stable topic ids and
// applies the per-filter limit (RetainStoreCoProc.match stops after `limit` messages,
// bifromq-retain/bifromq-retain-store/.../RetainStoreCoProc.java:167-190).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "../../include/bfq_gpumatch.h"
#include "codec.h"
#include "trie_layout.h"
#include "hash_probe.cuh"
#include "match_kernels.cuh"

using namespace bfq;

namespace {

int32_t rfail(int32_t code, const std::string& msg) { return bfq::set_error(code, msg); }
#define RCUDA_TRY(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return rfail(BFQ_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// one topic-trie node, indexed by BFS id; rnodes has a sentinel entry at [n_nodes]
struct RNode {
    uint32_t child_begin;   // BFS id of the first child (children are consecutive)
    uint32_t child_count;
    uint32_t sub_begin;     // DFS-rank range of the topics in this subtree (own topic first)
    uint32_t sub_end;
    uint32_t own_prefix;    // BFS-rank of the first topic ending at a node with id >= this one
    uint32_t sys_begin;     // children whose name starts with '$' form the run [sys_begin, sys_begin+sys_count)
    uint32_t sys_count;
    uint32_t pad;
};
static_assert(sizeof(RNode) == 32, "RNode is one 32-byte sector");

constexpr uint32_t SPACE_BFS = 0x80000000u;   // tag in a range's `first` word: BFS-rank space (else DFS)
constexpr uint32_t VIRT_BASE = 0x80000000u;   // ids of the intermediate nodes of long-token chunk chains

enum : int { RC_RANGES = 0, RC_OVERFLOW = 1, RC_ERROR = 2, RC_COUNT = 4 };

struct RMatchParams {
    const RNode* nodes;
    const Slot* slots;
    const uint4* tags;
    uint32_t n_blocks;
    const uint8_t* filters;
    const int64_t* filter_off;
    const int32_t* filter_tenant;
    const int32_t* tenant_root;     // BFS id of the tenant's root or -1
    int64_t n_filters;
    const uint32_t* work_list;
    int64_t n_work;
    uint32_t* span_begin;
    uint32_t* span_count;
    unsigned long long* total;      // [n] matches before the limit
    uint2* ranges;
    uint64_t ranges_cap;
    uint32_t* overflow_list;
    unsigned long long* counters;
    uint2* scratch;
    uint64_t scratch_frontier_cap, scratch_ranges_cap;
};

constexpr int R_WARPS = 8;
constexpr int R_STAGE = 256;
constexpr uint32_t R_FR_CAP = 64, R_RG_CAP = 64;
constexpr unsigned RFULL = 0xFFFFFFFFu;
constexpr uint32_t SPAN_OVF = 0x40000000u;

struct RWarpSmem {
    uint8_t stage[R_STAGE];
    uint32_t keyw[8];
    uint2 fr[2][R_FR_CAP];   // frontier: intervals {first id, count}
    uint2 rg[R_RG_CAP];
};

__device__ __forceinline__ RNode load_node(const RNode* n) {
    const uint4* p = reinterpret_cast<const uint4*>(n);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    RNode r;
    r.child_begin = a.x; r.child_count = a.y; r.sub_begin = a.z; r.sub_end = a.w;
    r.own_prefix = b.x; r.sys_begin = b.y; r.sys_count = b.z; r.pad = b.w;
    return r;
}

// (parent, lenw, k) -> child id, or NONE; *nd = the child's record (without its '$' run), *own_next = own_prefix of the node
// behind it — both carried in the slot's payload half (see rebuild)
__device__ __forceinline__ uint32_t rprobe(const Slot* slots, const uint4* tags, uint32_t n_blocks, uint32_t parent, uint32_t lenw,
                                           const uint32_t (&k)[6], uint64_t tokh, RNode* nd, uint32_t* own_next) {
    uint32_t w[16], slot = 0;
    if (!probe(slots, tags, n_blocks, parent, lenw, k, tokh, w, slot)) return NONE;
    nd->child_begin = w[9]; nd->child_count = w[10]; nd->sub_begin = w[11]; nd->sub_end = w[12];
    nd->own_prefix = w[13]; nd->sys_begin = 0; nd->sys_count = 0; nd->pad = 0;
    *own_next = w[14];
    return w[W_PLUS];
}

template <bool kBig>
__device__ __forceinline__ void rmatch_one(const RMatchParams& p, RWarpSmem& ws, uint32_t f, int lane, uint2* fr_a, uint2* fr_b,
                                           uint2* rg, uint32_t capF, uint32_t capR) {
    const int64_t fb = p.filter_off[f];
    const int len = (int) (p.filter_off[f + 1] - fb);
    const uint8_t* src = p.filters + fb;
    const bool staged = len <= R_STAGE;
    __syncwarp();
    if (staged)
        for (int i = lane; i < len; i += 32) ws.stage[i] = src[i];
    __syncwarp();
    auto byte_at = [&](int i) -> uint32_t { return staged ? (uint32_t) ws.stage[i] : (uint32_t) src[i]; };

    const int root = p.tenant_root[p.filter_tenant[f]];
    uint32_t n_rg = 0;
    unsigned long long total = 0;   // lane-local, reduced at the end
    bool overflow = false;
    auto emit = [&](bool valid, uint32_t first, uint32_t count) {
        valid = valid && count > 0;
        const unsigned m = __ballot_sync(RFULL, valid);
        if (m == 0) return;
        if (valid) {
            const uint32_t idx = n_rg + __popc(m & ((1u << lane) - 1));
            if (idx < capR) rg[idx] = make_uint2(first, count);
            total += count;
        }
        n_rg += __popc(m);
        if (n_rg > capR) overflow = true;
    };
    // append intervals to the next frontier (one optional interval per lane)
    uint2* fr_cur = fr_a;
    uint2* fr_next = fr_b;
    uint32_t n_fr = 0, n_next = 0;
    auto push = [&](bool valid, uint32_t first, uint32_t count) {
        valid = valid && count > 0;
        const unsigned m = __ballot_sync(RFULL, valid);
        if (m == 0) return;
        if (valid) {
            const uint32_t idx = n_next + __popc(m & ((1u << lane) - 1));
            if (idx < capF) fr_next[idx] = make_uint2(first, count);
        }
        n_next += __popc(m);
        if (n_next > capF) overflow = true;
    };

    // Visits every NODE of every frontier interval, 32 nodes per round, one per lane: the intervals of a block of 32 are
    // flattened with a warp scan and each lane finds its (interval, offset) with a 5-step search over the scanned prefixes.
    // (Round 1 gave each lane one INTERVAL and walked it sequentially: after a '+' level the frontier is ONE interval of
    // hundreds of nodes, so one lane did hundreds of dependent probes while 31 idled.) fn(alive, node) is called by the whole
    // warp in lock step: it may use warp collectives.
    auto for_each_frontier_node = [&](auto&& fn) {
        for (uint32_t base = 0; base < n_fr && !overflow; base += 32) {
            const bool act = base + lane < n_fr;
            const uint2 iv = act ? fr_cur[base + lane] : make_uint2(0u, 0u);
            uint32_t inc = iv.y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(RFULL, inc, o);
                if (lane >= o) inc += y;
            }
            const uint32_t pre = inc - iv.y, tot = __shfl_sync(RFULL, inc, 31);
            for (uint32_t t0 = 0; t0 < tot && !overflow; t0 += 32) {
                const uint32_t g = t0 + lane;
                int lo = 0, hi = 32;
#pragma unroll
                for (int it = 0; it < 5; it++) {   // largest lane whose exclusive prefix is <= g (prefixes are non-decreasing)
                    const int mid = (lo + hi) >> 1;
                    const uint32_t pm = __shfl_sync(RFULL, pre, mid);
                    if (pm <= g) lo = mid;
                    else hi = mid;
                }
                const uint32_t ix = __shfl_sync(RFULL, iv.x, lo), ip = __shfl_sync(RFULL, pre, lo);
                fn(g < tot, ix + (g - ip));
            }
        }
    };

    if (root >= 0) {
        if (lane == 0) fr_cur[0] = make_uint2((uint32_t) root, 1u);
        n_fr = 1;
        int pos = 0;
        int level = 0;
        while (n_fr > 0 && !overflow) {
            int e = len;
            for (int b = pos; b < len; b += 32) {
                const int i = b + lane;
                const unsigned m = __ballot_sync(RFULL, i < len && byte_at(i) == '/');
                if (m) {
                    e = b + __ffs(m) - 1;
                    break;
                }
            }
            const bool last = e == len;
            const int tlen = e - pos;
            const uint32_t c0 = tlen >= 1 ? byte_at(pos) : 0u;
            const bool is_plus = tlen == 1 && c0 == '+';
            const bool is_hash = last && tlen == 1 && c0 == '#';
            // does "/#" follow this level as the final level?  (matchParent of the reference selectors)
            bool hash_next = false;
            if (!last && e + 2 == len) hash_next = byte_at(e + 1) == '#';
            n_next = 0;
            __syncwarp();
            if (is_hash) {
                // '#': every child subtree of every frontier node. At level 0 the '$' children are skipped.
                for (uint32_t base = 0; base < n_fr && !overflow; base += 32) {
                    const bool act = base + lane < n_fr;
                    const uint2 iv = act ? fr_cur[base + lane] : make_uint2(0u, 0u);
                    // an interval of frontier nodes: emit per node (subtrees of different parents are not adjacent in DFS rank)
                    // intervals are short here except after '+' levels; lanes walk their interval sequentially
                    for (uint32_t j = 0; __any_sync(RFULL, act && j < iv.y); j++) {
                        const bool a2 = act && j < iv.y;
                        RNode nd{};
                        if (a2) nd = load_node(p.nodes + iv.x + j);
                        if (level == 0) {
                            // children subtrees minus the '$' run: [first child .. sys) and (sys .. last child]
                            RNode s0{}, s1{};
                            const bool has_sys = a2 && nd.sys_count > 0;
                            if (has_sys) {
                                s0 = load_node(p.nodes + nd.sys_begin);
                                s1 = load_node(p.nodes + nd.sys_begin + nd.sys_count - 1);
                            }
                            const uint32_t own = a2 ? (load_node(p.nodes + iv.x + j + 1).own_prefix - nd.own_prefix) : 0u;
                            const uint32_t lo = nd.sub_begin + own;
                            emit(a2 && !has_sys, lo, nd.sub_end - lo);
                            emit(has_sys, lo, s0.sub_begin - lo);
                            emit(has_sys, s1.sub_end, nd.sub_end - s1.sub_end);
                        } else {
                            // reached through "x/#" handling below, never here: kept for completeness
                            emit(a2, nd.sub_begin, nd.sub_end - nd.sub_begin);
                        }
                    }
                }
                break;
            }
            if (is_plus) {
                for (uint32_t base = 0; base < n_fr && !overflow; base += 32) {
                    const bool act = base + lane < n_fr;
                    const uint2 iv = act ? fr_cur[base + lane] : make_uint2(0u, 0u);
                    RNode n0{}, n1{};
                    if (act) {
                        n0 = load_node(p.nodes + iv.x);
                        n1 = iv.y > 1 ? load_node(p.nodes + iv.x + iv.y - 1) : n0;
                    }
                    // children of the whole interval = [first child of the first node, last child of the last node]
                    uint32_t cb = n0.child_begin, ce = n1.child_begin + n1.child_count;
                    // level 0: the frontier is the single tenant root; skip its '$' children
                    const bool split = act && level == 0 && n0.sys_count > 0;
                    const uint32_t sb = n0.sys_begin, se = n0.sys_begin + n0.sys_count;
                    if (last) {
                        // MATCH_AND_STOP on every child: the topics ending exactly at those nodes = one BFS-rank range
                        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
                        if (act && ce > cb) {
                            o0 = load_node(p.nodes + cb).own_prefix;
                            o3 = load_node(p.nodes + ce).own_prefix;
                            if (split) {
                                o1 = load_node(p.nodes + sb).own_prefix;
                                o2 = load_node(p.nodes + se).own_prefix;
                            }
                        }
                        emit(act && !split, SPACE_BFS | o0, o3 - o0);
                        emit(split, SPACE_BFS | o0, o1 - o0);
                        emit(split, SPACE_BFS | o2, o3 - o2);
                    } else if (hash_next) {
                        // "+/#": handled below, per frontier NODE
                    } else {
                        push(act && !split, cb, ce - cb);
                        push(split, cb, sb - cb);
                        push(split, se, ce - se);
                    }
                }
                if (hash_next && !last) {
                    // "+/#": whole subtrees of all children; per frontier node one DFS range (minus its own topic)
                    for_each_frontier_node([&](bool a2, uint32_t id) {
                        RNode nd{};
                        uint32_t own = 0;
                        if (a2) {
                            nd = load_node(p.nodes + id);
                            own = load_node(p.nodes + id + 1).own_prefix - nd.own_prefix;
                        }
                        const uint32_t lo = nd.sub_begin + own;
                        const bool sp = a2 && level == 0 && nd.sys_count > 0;
                        RNode s0{}, s1{};
                        if (sp) {
                            s0 = load_node(p.nodes + nd.sys_begin);
                            s1 = load_node(p.nodes + nd.sys_begin + nd.sys_count - 1);
                        }
                        emit(a2 && !sp, lo, nd.sub_end - lo);
                        emit(sp, lo, s0.sub_begin - lo);
                        emit(sp, s1.sub_end, nd.sub_end - s1.sub_end);
                    });
                }
                if (last || hash_next) break;
            } else {
                // exact level: probe every node of every frontier interval
                const int nchunks = tlen <= (int) TOKEN_BYTES ? 1 : (tlen + (int) TOKEN_BYTES - 1) / (int) TOKEN_BYTES;
                for_each_frontier_node([&](bool alive, uint32_t node) {
                    RNode cnd{};
                    uint32_t own_next = 0;
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
                        for (int q = 0; q < 6; q++) k[q] = ws.keyw[q];
                        const uint64_t tokh = token_hash(lenw, k);
                        if (alive) {
                            node = rprobe(p.slots, p.tags, p.n_blocks, node, lenw, k, tokh, &cnd, &own_next);
                            alive = node != NONE;
                        }
                    }
                    if (last) {
                        emit(alive, SPACE_BFS | cnd.own_prefix, own_next - cnd.own_prefix);
                    } else if (hash_next) {
                        emit(alive, cnd.sub_begin, cnd.sub_end - cnd.sub_begin);   // "x/#": x itself and everything below
                    } else {
                        push(alive, node, 1u);
                    }
                });
                if (last || hash_next) break;
            }
            __syncwarp();
            uint2* tmp = fr_cur;
            fr_cur = fr_next;
            fr_next = tmp;
            n_fr = n_next;
            pos = e + 1;
            level++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(RFULL, total, o);
    __syncwarp();
    if (overflow) {
        if (lane == 0) {
            if (!kBig) {
                const unsigned long long idx = atomicAdd(&p.counters[RC_OVERFLOW], 1ull);
                p.overflow_list[idx] = f;
                p.span_count[f] = SPAN_OVF;
            } else {
                atomicAdd(&p.counters[RC_ERROR], 1ull);
                p.span_count[f] = 0;
            }
            p.span_begin[f] = 0;
            p.total[f] = 0;
        }
        return;
    }
    unsigned long long base = 0;
    if (n_rg > 0) {
        if (lane == 0) base = atomicAdd(&p.counters[RC_RANGES], (unsigned long long) n_rg);
        base = __shfl_sync(RFULL, base, 0);
        if (base + n_rg <= p.ranges_cap)
            for (uint32_t i = lane; i < n_rg; i += 32) p.ranges[base + i] = rg[i];
    }
    if (lane == 0) {
        p.span_begin[f] = (uint32_t) base;
        p.span_count[f] = n_rg;
        p.total[f] = total;
    }
}

template <bool kBig>
__global__ void __launch_bounds__(R_WARPS * 32, 5) rmatch_kernel(const RMatchParams p) {
    __shared__ RWarpSmem sm[R_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    RWarpSmem& ws = sm[wid];
    const int64_t gw = (int64_t) blockIdx.x * R_WARPS + wid, nw = (int64_t) gridDim.x * R_WARPS;
    if (kBig) {
        uint2* basep = p.scratch + (uint64_t) gw * (2 * p.scratch_frontier_cap + p.scratch_ranges_cap);
        for (int64_t it = gw; it < p.n_work; it += nw)
            rmatch_one<true>(p, ws, p.work_list[it], lane, basep, basep + p.scratch_frontier_cap, basep + 2 * p.scratch_frontier_cap,
                             (uint32_t) min((uint64_t) 0x3FFFFFFFull, p.scratch_frontier_cap),
                             (uint32_t) min((uint64_t) 0x3FFFFFFFull, p.scratch_ranges_cap));
    } else {
        // p.work_list here = the locality order of the batch (filters grouped by tenant and leading levels), or nullptr
        for (int64_t it = gw; it < p.n_filters; it += nw)
            rmatch_one<false>(p, ws, p.work_list ? p.work_list[it] : (uint32_t) it, lane, ws.fr[0], ws.fr[1], ws.rg, R_FR_CAP, R_RG_CAP);
    }
}

// kept[i] = min(total[i], limit[i])
__global__ void rkept_kernel(int64_t n, const unsigned long long* total, const int64_t* limit, unsigned long long* kept) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned long long t = total[i];
        const long long l = limit ? limit[i] : -1;
        kept[i] = l < 0 ? t : (t < (unsigned long long) l ? t : (unsigned long long) l);
    }
}

// one warp per filter: map its ranges to topic ids, first `kept` of them
__global__ void __launch_bounds__(256) rexpand_kernel(int64_t n, const uint32_t* span_begin, const uint32_t* span_count,
                                                      const uint2* ranges, const unsigned long long* offsets,
                                                      const unsigned long long* kept, const int64_t* dfs_to_id,
                                                      const int64_t* bfs_to_id, int64_t* ids) {
    const int lane = threadIdx.x & 31;
    const int64_t f = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= n) return;
    const unsigned long long out0 = offsets[f], want = kept[f];
    unsigned long long done = 0;
    const uint32_t b = span_begin[f], c = span_count[f];
    for (uint32_t j = 0; j < c && done < want; j++) {
        const uint2 r = ranges[b + j];
        const bool bfs = r.x & SPACE_BFS;
        const uint32_t first = r.x & ~SPACE_BFS;
        const unsigned long long take = min((unsigned long long) r.y, want - done);
        const int64_t* map = bfs ? bfs_to_id : dfs_to_id;
        for (unsigned long long x = lane; x < take; x += 32) ids[out0 + done + x] = map[first + x];
        done += take;
    }
}

template <typename T>
struct DBuf {
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
};

}  // namespace

struct bfq_rresult {
    std::vector<int64_t> offsets, ids, totals;
    double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct bfq_rindex {
    int device = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // [0,1] around all kernels of a match, [2,3] around rmatch_kernel
    unsigned long long* h_small = nullptr;   // pinned scalars
    // staging: (tenant, topic) -> id ; id -> (tenant, topic)
    std::map<std::pair<std::string, std::string>, int64_t> staged;
    std::vector<std::pair<std::string, std::string>> by_id;   // id -> strings (tombstones keep their slot)
    std::vector<char> alive;
    bool have_snapshot = false;
    // snapshot
    std::unordered_map<std::string, int32_t> tenant_root;
    uint32_t n_blocks = 0;
    int64_t n_nodes = 0, max_nodes_per_depth = 0, n_topics = 0;
    DBuf<RNode> d_nodes;
    DBuf<Slot> d_slots;
    DBuf<uint8_t> d_tags;
    DBuf<int64_t> d_dfs_to_id, d_bfs_to_id;
    // workspace
    DBuf<uint8_t> d_filters, d_scan_tmp;
    DBuf<int64_t> d_filter_off, d_limit, d_ids;
    DBuf<int32_t> d_filter_tenant, d_tenant_root;
    DBuf<uint32_t> d_span_begin, d_span_count, d_overflow;
    DBuf<unsigned long long> d_total, d_kept, d_offsets, d_counters;
    DBuf<uint2> d_ranges, d_scratch;
    // locality order of a batch of filters (match_kernels.cu: launch_order, without de-duplication)
    DBuf<uint32_t> d_ord_keys, d_ord_leader, d_order, d_hist;
    DBuf<unsigned long long> d_ord_ctr;
    int64_t launches = 0;

    ~bfq_rindex() {
        cudaSetDevice(device);
        d_nodes.release(); d_slots.release(); d_tags.release(); d_dfs_to_id.release(); d_bfs_to_id.release(); d_filters.release();
        d_scan_tmp.release(); d_filter_off.release(); d_limit.release(); d_ids.release(); d_filter_tenant.release();
        d_tenant_root.release(); d_span_begin.release(); d_span_count.release(); d_overflow.release(); d_total.release();
        d_kept.release(); d_offsets.release(); d_counters.release(); d_ranges.release(); d_scratch.release();
        d_ord_keys.release(); d_ord_leader.release(); d_order.release(); d_hist.release(); d_ord_ctr.release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (h_small) cudaFreeHost(h_small);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

struct HNode {  // host build node
    std::map<std::string, uint32_t> children;   // name -> host node index (sorted => '$' children form one run)
    int64_t own = -1;                           // topic id ending here
    uint32_t bfs = 0;
};

inline void make_tok(sv chunk, uint32_t* tok) {
    for (uint32_t k = 0; k < TOKEN_WORDS; k++) tok[k] = 0;
    for (size_t j = 0; j < chunk.size(); j++) tok[j >> 2] |= (uint32_t) (uint8_t) chunk[j] << (8 * (j & 3));
}

struct Edge {
    uint32_t parent, lenw, tok[TOKEN_WORDS], child;
};

int32_t rebuild(bfq_rindex* h) {
    std::vector<HNode> nodes;
    std::unordered_map<std::string, uint32_t> root_of;
    std::vector<uint32_t> roots;
    for (const auto& e : h->staged) {
        const std::string& tenant = e.first.first;
        auto it = root_of.find(tenant);
        uint32_t cur;
        if (it == root_of.end()) {
            nodes.emplace_back();
            cur = (uint32_t) nodes.size() - 1;
            root_of.emplace(tenant, cur);
            roots.push_back(cur);
        } else {
            cur = it->second;
        }
        for_each_level(sv(e.first.second), '/', [&](sv l) {
            auto c = nodes[cur].children.find(std::string(l));
            if (c == nodes[cur].children.end()) {
                nodes.emplace_back();
                uint32_t idx = (uint32_t) nodes.size() - 1;
                nodes[cur].children.emplace(std::string(l), idx);
                cur = idx;
            } else {
                cur = c->second;
            }
        });
        nodes[cur].own = e.second;
    }
    const size_t N = nodes.size();
    if (N >= 0x7FFFFFF0ull) return rfail(BFQ_E_RANGE, "topic index too large");
    // ---- BFS numbering (roots first, then level by level in parent order)
    std::vector<uint32_t> order;
    order.reserve(N);
    for (uint32_t r : roots) order.push_back(r);
    int64_t max_depth_nodes = (int64_t) roots.size();
    for (size_t lo = 0, hi = order.size(); lo < hi;) {
        for (size_t i = lo; i < hi; i++)
            for (const auto& c : nodes[order[i]].children) order.push_back(c.second);
        lo = hi;
        hi = order.size();
        max_depth_nodes = std::max<int64_t>(max_depth_nodes, (int64_t) (hi - lo));
    }
    for (size_t i = 0; i < N; i++) nodes[order[i]].bfs = (uint32_t) i;
    std::vector<RNode> rn(N + 1);
    std::vector<int64_t> bfs_to_id, dfs_to_id;
    // child intervals + BFS topic ranks
    {
        uint32_t next_child = (uint32_t) roots.size();
        for (size_t i = 0; i < N; i++) {
            const HNode& hn = nodes[order[i]];
            RNode& r = rn[i];
            r.child_begin = next_child;
            r.child_count = (uint32_t) hn.children.size();
            next_child += r.child_count;
            r.own_prefix = (uint32_t) bfs_to_id.size();
            if (hn.own >= 0) bfs_to_id.push_back(hn.own);
            r.sys_begin = 0;
            r.sys_count = 0;
            r.pad = 0;
            uint32_t k = 0;
            for (const auto& c : hn.children) {
                if (!c.first.empty() && c.first[0] == '$') {
                    if (r.sys_count == 0) r.sys_begin = r.child_begin + k;
                    r.sys_count++;
                }
                k++;
            }
        }
        rn[N] = RNode{next_child, 0, 0, 0, (uint32_t) bfs_to_id.size(), 0, 0, 0};
    }
    // DFS (pre-order) topic ranks, iterative
    {
        std::vector<std::pair<uint32_t, std::map<std::string, uint32_t>::const_iterator>> st;
        for (uint32_t r : roots) {
            st.clear();
            rn[nodes[r].bfs].sub_begin = (uint32_t) dfs_to_id.size();
            if (nodes[r].own >= 0) dfs_to_id.push_back(nodes[r].own);
            st.push_back({r, nodes[r].children.begin()});
            while (!st.empty()) {
                auto& top = st.back();
                if (top.second == nodes[top.first].children.end()) {
                    rn[nodes[top.first].bfs].sub_end = (uint32_t) dfs_to_id.size();
                    st.pop_back();
                    continue;
                }
                uint32_t c = top.second->second;
                ++top.second;
                rn[nodes[c].bfs].sub_begin = (uint32_t) dfs_to_id.size();
                if (nodes[c].own >= 0) dfs_to_id.push_back(nodes[c].own);
                st.push_back({c, nodes[c].children.begin()});
            }
        }
    }
    // ---- exact-edge hash table keyed by the parent's BFS id (long names: chains of virtual nodes)
    std::vector<Edge> edges;
    edges.reserve(N);
    uint32_t next_virtual = VIRT_BASE;
    std::map<std::tuple<uint32_t, uint32_t, std::string>, uint32_t> virt;
    for (size_t i = 0; i < N; i++) {
        const HNode& hn = nodes[order[i]];
        for (const auto& c : hn.children) {
            sv name(c.first);
            uint32_t parent = (uint32_t) i;
            size_t off = 0;
            uint32_t j = 0;
            while (name.size() - off > TOKEN_BYTES) {
                Edge e{};
                e.parent = parent;
                e.lenw = LEN_CONT | j;
                make_tok(name.substr(off, TOKEN_BYTES), e.tok);
                // identical chunk prefixes under the same parent share one virtual node
                auto vk = std::make_tuple(e.parent, e.lenw, std::string((const char*) e.tok, sizeof(e.tok)));
                auto vit = virt.find(vk);
                uint32_t found;
                if (vit == virt.end()) {
                    e.child = next_virtual++;
                    edges.push_back(e);
                    virt.emplace(std::move(vk), e.child);
                    found = e.child;
                } else {
                    found = vit->second;
                }
                parent = found;
                off += TOKEN_BYTES;
                j++;
            }
            Edge e{};
            e.parent = parent;
            e.lenw = (uint32_t) name.size();
            make_tok(name.substr(off), e.tok);
            e.child = nodes[c.second].bfs;
            edges.push_back(e);
        }
    }
    if (((uint64_t) edges.size() * 2 / BLOCK_USABLE + 64) * BLOCK_SLOTS >= 0x7FFFFFF0ull) return rfail(BFQ_E_RANGE, "topic index too large");
    EdgeTable table;
    table.init(edges.size());
    for (const Edge& e : edges) {
        const uint32_t s = table.place(e.parent, e.lenw, e.tok);
        uint32_t* w = table.slots[s].w;
        w[W_PLUS] = e.child;
        if (e.child < VIRT_BASE) {
            // the child's node record rides in the slot's payload half: an exact step is tag + slot, with no third dependent
            // access for the record (words 9..14: child_begin, child_count, sub_begin, sub_end, own_prefix, own_prefix of the
            // next node; the '$' run is only needed for tenant roots, which are never reached through a slot)
            const RNode& r = rn[e.child];
            w[9] = r.child_begin;
            w[10] = r.child_count;
            w[11] = r.sub_begin;
            w[12] = r.sub_end;
            w[13] = r.own_prefix;
            w[14] = rn[e.child + 1].own_prefix;
        }
    }
    SlotVec& slots = table.slots;
    // ---- upload
    RCUDA_TRY(cudaSetDevice(h->device));
    RCUDA_TRY(cudaStreamSynchronize(h->stream));
    RCUDA_TRY(h->d_nodes.reserve(rn.size()));
    RCUDA_TRY(h->d_slots.reserve(slots.size()));
    RCUDA_TRY(h->d_tags.reserve(table.tags.size()));
    RCUDA_TRY(h->d_dfs_to_id.reserve(std::max<size_t>(dfs_to_id.size(), 1)));
    RCUDA_TRY(h->d_bfs_to_id.reserve(std::max<size_t>(bfs_to_id.size(), 1)));
    RCUDA_TRY(cudaMemcpy(h->d_nodes.p, rn.data(), rn.size() * sizeof(RNode), cudaMemcpyHostToDevice));
    RCUDA_TRY(cudaMemcpy(h->d_slots.p, slots.data(), slots.size() * sizeof(Slot), cudaMemcpyHostToDevice));
    RCUDA_TRY(cudaMemcpy(h->d_tags.p, table.tags.data(), table.tags.size(), cudaMemcpyHostToDevice));
    if (!dfs_to_id.empty()) {
        RCUDA_TRY(cudaMemcpy(h->d_dfs_to_id.p, dfs_to_id.data(), dfs_to_id.size() * 8, cudaMemcpyHostToDevice));
        RCUDA_TRY(cudaMemcpy(h->d_bfs_to_id.p, bfs_to_id.data(), bfs_to_id.size() * 8, cudaMemcpyHostToDevice));
    }
    h->tenant_root.clear();
    for (const auto& e : root_of) h->tenant_root[e.first] = (int32_t) nodes[e.second].bfs;
    h->n_blocks = table.n_blocks;
    h->n_nodes = (int64_t) N;
    h->n_topics = (int64_t) dfs_to_id.size();
    h->max_nodes_per_depth = max_depth_nodes;
    h->have_snapshot = true;
    return BFQ_OK;
}

}  // namespace

extern "C" {

int32_t bfq_rindex_create(int32_t device_ordinal, bfq_rindex** out) {
    if (!out) return rfail(BFQ_E_INVALID, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return rfail(BFQ_E_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device_ordinal < 0 || device_ordinal >= count) return rfail(BFQ_E_INVALID, "device ordinal out of range");
    RCUDA_TRY(cudaSetDevice(device_ordinal));
    auto* h = new bfq_rindex();
    h->device = device_ordinal;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return rfail(BFQ_E_CUDA, "cudaStreamCreate failed");
    }
    *out = h;
    return BFQ_OK;
}
void bfq_rindex_destroy(bfq_rindex* h) { delete h; }

int32_t bfq_rindex_reset(bfq_rindex* h) {
    if (!h) return rfail(BFQ_E_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    h->staged.clear();
    h->by_id.clear();
    h->alive.clear();
    return BFQ_OK;
}

int32_t bfq_rindex_add(bfq_rindex* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                       const uint8_t* topics, const int64_t* topic_off, const int32_t* topic_tenant, int64_t n, int64_t* ids_out) {
    if (!h || n < 0 || n_tenants < 0) return rfail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    std::vector<std::string> ts((size_t) n_tenants);
    for (int32_t t = 0; t < n_tenants; t++) ts[(size_t) t].assign((const char*) tenants + tenant_off[t], (size_t) (tenant_off[t + 1] - tenant_off[t]));
    for (int64_t i = 0; i < n; i++) {
        if (topic_tenant[i] < 0 || topic_tenant[i] >= n_tenants) return rfail(BFQ_E_RANGE, "topic_tenant out of range");
        std::pair<std::string, std::string> key(ts[(size_t) topic_tenant[i]],
                                                std::string((const char*) topics + topic_off[i], (size_t) (topic_off[i + 1] - topic_off[i])));
        auto it = h->staged.find(key);
        int64_t id;
        if (it == h->staged.end()) {
            id = (int64_t) h->by_id.size();
            h->by_id.push_back(key);
            h->alive.push_back(1);
            h->staged.emplace(std::move(key), id);
        } else {
            id = it->second;
        }
        if (ids_out) ids_out[i] = id;
    }
    return BFQ_OK;
}

// The feed of RetainStoreCoProc.load() (RS/RetainStoreCoProc.java:279-296): raw retain-store KV keys from a range scan. The
// reference parses every VALUE (a TopicMessage proto) for the topic; the key carries it too (escaped, behind the level-hash bytes),
// so the index is fed from the keys alone. ids_out[i] < 0 marks a key that is not a retain key (skipped, as the reference logs
// and skips an unparsable entry).
int32_t bfq_rindex_load_keys(bfq_rindex* h, const uint8_t* keys, const int64_t* key_off, int64_t n, int64_t* ids_out) {
    if (!h || n < 0 || (n > 0 && (!keys || !key_off))) return rfail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    for (int64_t i = 0; i < n; i++) {
        sv tenant;
        std::string topic;
        if (key_off[i + 1] < key_off[i] || !decode_retain_key(sv((const char*) keys + key_off[i], (size_t) (key_off[i + 1] - key_off[i])), &tenant, &topic)) {
            if (ids_out) ids_out[i] = -1;
            continue;
        }
        std::pair<std::string, std::string> key(std::string(tenant), std::move(topic));
        auto it = h->staged.find(key);
        int64_t id;
        if (it == h->staged.end()) {
            id = (int64_t) h->by_id.size();
            h->by_id.push_back(key);
            h->alive.push_back(1);
            h->staged.emplace(std::move(key), id);
        } else {
            id = it->second;
        }
        if (ids_out) ids_out[i] = id;
    }
    return BFQ_OK;
}

// retainMessageKey(tenant, topic) of every id in a match result, in result order: the keys of the follow-up reader.get calls of
// RetainStoreCoProc.match (RS/RetainStoreCoProc.java:177-188), as one batch. Returns the blob length; copies if it fits.
int64_t bfq_rresult_retain_keys(bfq_rindex* h, const bfq_rresult* r, uint8_t* blob_out, int64_t blob_cap, int64_t* key_off_out) {
    if (!h || !r) return BFQ_E_INVALID;
    std::lock_guard<std::mutex> g(h->mu);
    int64_t at = 0;
    const int64_t n = (int64_t) r->ids.size();
    for (int64_t i = 0; i < n; i++) {
        const int64_t id = r->ids[(size_t) i];
        if (id < 0 || id >= (int64_t) h->by_id.size()) return BFQ_E_RANGE;
        const std::string k = make_retain_key(h->by_id[(size_t) id].first, h->by_id[(size_t) id].second);
        if (key_off_out) key_off_out[i] = at;
        if (blob_out && at + (int64_t) k.size() <= blob_cap) memcpy(blob_out + at, k.data(), k.size());
        at += (int64_t) k.size();
    }
    if (key_off_out) key_off_out[n] = at;
    return at;
}

int32_t bfq_rindex_remove(bfq_rindex* h, const uint8_t* tenant, int64_t tn, const uint8_t* topic, int64_t n) {
    if (!h) return rfail(BFQ_E_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    auto it = h->staged.find({std::string((const char*) tenant, (size_t) tn), std::string((const char*) topic, (size_t) n)});
    if (it != h->staged.end()) {
        h->alive[(size_t) it->second] = 0;
        h->staged.erase(it);
    }
    return BFQ_OK;
}

int32_t bfq_rindex_commit(bfq_rindex* h) {
    if (!h) return rfail(BFQ_E_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    return rebuild(h);
}

int32_t bfq_rindex_lookup(bfq_rindex* h, int64_t id, uint8_t* tenant_out, int64_t tenant_cap, int64_t* tenant_len,
                          uint8_t* topic_out, int64_t topic_cap, int64_t* topic_len) {
    if (!h) return rfail(BFQ_E_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    if (id < 0 || id >= (int64_t) h->by_id.size()) return rfail(BFQ_E_RANGE, "id out of range");
    const auto& e = h->by_id[(size_t) id];
    if (tenant_len) *tenant_len = (int64_t) e.first.size();
    if (topic_len) *topic_len = (int64_t) e.second.size();
    if (tenant_out && (int64_t) e.first.size() <= tenant_cap) memcpy(tenant_out, e.first.data(), e.first.size());
    if (topic_out && (int64_t) e.second.size() <= topic_cap) memcpy(topic_out, e.second.data(), e.second.size());
    return BFQ_OK;
}

int32_t bfq_rmatch(bfq_rindex* h, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                   const uint8_t* filters, const int64_t* filter_off, const int32_t* filter_tenant, int64_t n,
                   const int64_t* limit, bfq_rresult** out) {
    if (!h || !out || n < 0 || n_tenants < 0) return rfail(BFQ_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->have_snapshot) return rfail(BFQ_E_STATE, "bfq_rmatch before the first bfq_rindex_commit");
    for (int64_t i = 0; i < n; i++)
        if (filter_tenant[i] < 0 || filter_tenant[i] >= n_tenants) return rfail(BFQ_E_RANGE, "filter_tenant out of range");
    RCUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    auto t0 = std::chrono::steady_clock::now();
    const size_t nn = (size_t) std::max<int64_t>(n, 1), nt = (size_t) std::max(n_tenants, 1);
    std::vector<int32_t> troot(nt, -1);
    for (int32_t t = 0; t < n_tenants; t++) {
        auto it = h->tenant_root.find(std::string((const char*) tenants + tenant_off[t], (size_t) (tenant_off[t + 1] - tenant_off[t])));
        if (it != h->tenant_root.end()) troot[(size_t) t] = it->second;
    }
    const int64_t fbytes = n ? filter_off[n] : 0;
    RCUDA_TRY(h->d_filters.reserve((size_t) std::max<int64_t>(fbytes, 1)));
    RCUDA_TRY(h->d_filter_off.reserve(nn + 1));
    RCUDA_TRY(h->d_filter_tenant.reserve(nn));
    RCUDA_TRY(h->d_tenant_root.reserve(nt));
    RCUDA_TRY(h->d_limit.reserve(nn));
    RCUDA_TRY(h->d_span_begin.reserve(nn));
    RCUDA_TRY(h->d_span_count.reserve(nn));
    RCUDA_TRY(h->d_overflow.reserve(nn));
    RCUDA_TRY(h->d_total.reserve(nn));
    RCUDA_TRY(h->d_kept.reserve(nn));
    RCUDA_TRY(h->d_offsets.reserve(nn + 1));
    RCUDA_TRY(h->d_counters.reserve(RC_COUNT));
    if (h->d_ranges.cap == 0) RCUDA_TRY(h->d_ranges.reserve(std::max<size_t>(1 << 18, 8 * nn)));
    auto* res = new bfq_rresult();
    res->offsets.assign((size_t) n + 1, 0);
    res->totals.assign((size_t) n, 0);
    if (n == 0) {
        *out = res;
        return BFQ_OK;
    }
    RCUDA_TRY(cudaMemcpyAsync(h->d_filters.p, filters, (size_t) fbytes, cudaMemcpyHostToDevice, st));
    RCUDA_TRY(cudaMemcpyAsync(h->d_filter_off.p, filter_off, (size_t) (n + 1) * 8, cudaMemcpyHostToDevice, st));
    RCUDA_TRY(cudaMemcpyAsync(h->d_filter_tenant.p, filter_tenant, (size_t) n * 4, cudaMemcpyHostToDevice, st));
    RCUDA_TRY(cudaMemcpyAsync(h->d_tenant_root.p, troot.data(), nt * 4, cudaMemcpyHostToDevice, st));
    if (limit) RCUDA_TRY(cudaMemcpyAsync(h->d_limit.p, limit, (size_t) n * 8, cudaMemcpyHostToDevice, st));
    auto t1 = std::chrono::steady_clock::now();
    for (auto& e : h->ev)
        if (!e) RCUDA_TRY(cudaEventCreate(&e));
    RCUDA_TRY(cudaEventRecord(h->ev[0], st));   // the inputs are (enqueued to be) resident: device time of the kernels from here

    RMatchParams p{};
    p.nodes = h->d_nodes.p;
    p.slots = h->d_slots.p;
    p.tags = reinterpret_cast<const uint4*>(h->d_tags.p);
    p.n_blocks = h->n_blocks;
    p.filters = h->d_filters.p;
    p.filter_off = h->d_filter_off.p;
    p.filter_tenant = h->d_filter_tenant.p;
    p.tenant_root = h->d_tenant_root.p;
    p.n_filters = n;
    p.span_begin = h->d_span_begin.p;
    p.span_count = h->d_span_count.p;
    p.total = h->d_total.p;
    p.overflow_list = h->d_overflow.p;
    p.counters = h->d_counters.p;
    unsigned long long hc[RC_COUNT];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = 4;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, rmatch_kernel<false>, R_WARPS * 32, 0);
    ctas_per_sm = std::max(1, ctas_per_sm);
    // filters that share a tenant and leading levels walk the same part of the topic trie: group them so that the warps
    // running at the same time hit the same node records in L2 (the forward path's locality order, reused)
    const uint32_t* order = nullptr;
    if (n >= 4096) {
        const size_t buckets = order_hist_buckets(n, n_tenants);
        const size_t hist_words = (buckets + 2 * (buckets / 4096) + 64 + 63) / 64 * 64;
        RCUDA_TRY(h->d_ord_keys.reserve(nn));
        RCUDA_TRY(h->d_ord_leader.reserve(nn));
        RCUDA_TRY(h->d_order.reserve(nn));
        RCUDA_TRY(h->d_hist.reserve(hist_words));
        RCUDA_TRY(h->d_ord_ctr.reserve(CTR_COUNT));
        OrderParams q{};
        q.n_topics = n;
        q.topics = h->d_filters.p;
        q.topic_off = h->d_filter_off.p;
        q.topic_tenant = h->d_filter_tenant.p;
        q.n_tenants = n_tenants;
        q.keys = h->d_ord_keys.p;
        q.leader = h->d_ord_leader.p;
        q.order = h->d_order.p;
        q.hash_tab = nullptr;
        q.hash_mask = 0;
        q.hist = h->d_hist.p;
        q.blk_tot = q.hist + buckets;
        q.blk_pfx = q.blk_tot + buckets / 4096;
        q.ticket = q.blk_pfx + buckets / 4096;
        q.hist_bits = 0;
        while (((size_t) 1 << q.hist_bits) < buckets) q.hist_bits++;
        q.dedup = 0;
        q.counters = h->d_ord_ctr.p;
        RCUDA_TRY(cudaMemsetAsync(q.hist, 0, hist_words * sizeof(uint32_t), st));
        RCUDA_TRY(launch_order(q, st));
        h->launches += 3;
        order = q.order;
    }
    for (int attempt = 0; attempt < 8; attempt++) {
        p.ranges = h->d_ranges.p;
        p.ranges_cap = h->d_ranges.cap;
        p.work_list = order;
        p.n_work = 0;
        RCUDA_TRY(cudaMemsetAsync(h->d_counters.p, 0, sizeof(hc), st));
        int64_t ctas = std::min<int64_t>((n + R_WARPS - 1) / R_WARPS, (int64_t) sms * ctas_per_sm);
        RCUDA_TRY(cudaEventRecord(h->ev[2], st));
        rmatch_kernel<false><<<(unsigned) std::max<int64_t>(ctas, 1), R_WARPS * 32, 0, st>>>(p);
        RCUDA_TRY(cudaEventRecord(h->ev[3], st));
        h->launches++;
        RCUDA_TRY(cudaGetLastError());
        RCUDA_TRY(cudaMemcpyAsync(hc, h->d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        RCUDA_TRY(cudaStreamSynchronize(st));
        if (hc[RC_OVERFLOW] > 0) {
            const uint64_t capF = (uint64_t) h->max_nodes_per_depth + 2;
            const uint64_t capR = 3 * ((uint64_t) h->n_nodes + 2) + 2;
            const uint64_t per_warp = 2 * capF + capR;
            uint64_t warps = std::min<uint64_t>(hc[RC_OVERFLOW], std::max<uint64_t>(8, (1ull << 31) / (per_warp * sizeof(uint2))));
            warps = std::min<uint64_t>((warps + 7) / 8 * 8, (uint64_t) sms * 8);
            RCUDA_TRY(h->d_scratch.reserve((size_t) (warps * per_warp)));
            p.scratch = h->d_scratch.p;
            p.scratch_frontier_cap = capF;
            p.scratch_ranges_cap = capR;
            p.work_list = h->d_overflow.p;
            p.n_work = (int64_t) hc[RC_OVERFLOW];
            rmatch_kernel<true><<<(unsigned) (warps / 8), R_WARPS * 32, 0, st>>>(p);
            h->launches++;
            RCUDA_TRY(cudaGetLastError());
            RCUDA_TRY(cudaMemcpyAsync(hc, h->d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
            RCUDA_TRY(cudaStreamSynchronize(st));
            if (hc[RC_ERROR] != 0) {
                delete res;
                return rfail(BFQ_E_STATE, "tier-2 scratch exhausted");
            }
        }
        if (hc[RC_RANGES] <= h->d_ranges.cap) break;
        const size_t want = (size_t) (hc[RC_RANGES] + hc[RC_RANGES] / 4 + 1024);
        if (want >= 0xFFFFFFF0ull || attempt == 7) {
            delete res;
            return rfail(BFQ_E_RANGE, "too many matched ranges in one batch; split the batch");
        }
        RCUDA_TRY(h->d_ranges.reserve(want));
    }
    // kept = min(total, limit); exclusive scan; expand to ids
    rkept_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(n, h->d_total.p, limit ? h->d_limit.p : nullptr, h->d_kept.p);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_kept.p, h->d_offsets.p, (int) n, st);
    RCUDA_TRY(h->d_scan_tmp.reserve(tmp_bytes));
    cub::DeviceScan::ExclusiveSum(h->d_scan_tmp.p, tmp_bytes, h->d_kept.p, h->d_offsets.p, (int) n, st);
    h->launches += 2;
    // only the grand total is needed on the host before the expansion (to size the id buffer): two scalars, not the arrays
    if (!h->h_small) RCUDA_TRY(cudaMallocHost(&h->h_small, 4 * sizeof(unsigned long long)));
    RCUDA_TRY(cudaMemcpyAsync(h->h_small, h->d_offsets.p + (n - 1), 8, cudaMemcpyDeviceToHost, st));
    RCUDA_TRY(cudaMemcpyAsync(h->h_small + 1, h->d_kept.p + (n - 1), 8, cudaMemcpyDeviceToHost, st));
    RCUDA_TRY(cudaStreamSynchronize(st));
    const unsigned long long total_ids = h->h_small[0] + h->h_small[1];
    RCUDA_TRY(h->d_ids.reserve((size_t) std::max<unsigned long long>(total_ids, 1)));
    rexpand_kernel<<<(unsigned) ((n * 32 + 255) / 256), 256, 0, st>>>(n, h->d_span_begin.p, h->d_span_count.p, h->d_ranges.p,
                                                                      h->d_offsets.p, h->d_kept.p, h->d_dfs_to_id.p,
                                                                      h->d_bfs_to_id.p, h->d_ids.p);
    h->launches++;
    RCUDA_TRY(cudaGetLastError());
    RCUDA_TRY(cudaEventRecord(h->ev[1], st));
    auto t2 = std::chrono::steady_clock::now();
    // the result owns its arrays: offsets / totals / ids are read back straight into them (same 8-byte element types)
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "offsets are copied without conversion");
    res->ids.resize((size_t) total_ids);
    RCUDA_TRY(cudaMemcpyAsync(res->offsets.data(), h->d_offsets.p, (size_t) n * 8, cudaMemcpyDeviceToHost, st));
    RCUDA_TRY(cudaMemcpyAsync(res->totals.data(), h->d_total.p, (size_t) n * 8, cudaMemcpyDeviceToHost, st));
    if (total_ids) RCUDA_TRY(cudaMemcpyAsync(res->ids.data(), h->d_ids.p, (size_t) total_ids * 8, cudaMemcpyDeviceToHost, st));
    RCUDA_TRY(cudaStreamSynchronize(st));
    res->offsets[(size_t) n] = (int64_t) total_ids;
    auto t3 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    res->ms[0] = ms(t0, t1);
    res->ms[1] = ms(t1, t2);
    res->ms[2] = ms(t2, t3);
    res->ms[3] = ms(t0, t3);
    {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&b, h->ev[2], h->ev[3]);
        res->ms[4] = a;                       // device time from "inputs resident" to "ids expanded" (all kernels + the host's counter reads between them)
        res->ms[5] = b;                       // device time of rmatch_kernel (the last attempt)
        res->ms[6] = (double) hc[RC_RANGES];  // rank ranges emitted (8 bytes each): the kernel's output size
        res->ms[7] = (double) hc[RC_OVERFLOW];
    }
    *out = res;
    return BFQ_OK;
}

int64_t bfq_rresult_num_filters(const bfq_rresult* r) { return r ? (int64_t) r->totals.size() : 0; }
const int64_t* bfq_rresult_offsets(const bfq_rresult* r) { return r->offsets.data(); }
const int64_t* bfq_rresult_ids(const bfq_rresult* r, int64_t* n) {
    if (n) *n = (int64_t) r->ids.size();
    return r->ids.data();
}
const int64_t* bfq_rresult_total_matches(const bfq_rresult* r) { return r->totals.data(); }
int32_t bfq_rresult_timings(const bfq_rresult* r, double* ms, int32_t n) {
    if (!r || !ms) return rfail(BFQ_E_INVALID, "bad argument");
    for (int32_t i = 0; i < n && i < 8; i++) ms[i] = r->ms[i];
    return BFQ_OK;
}
void bfq_rresult_free(bfq_rresult* r) { delete r; }

}  // extern "C"
