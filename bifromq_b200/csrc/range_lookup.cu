// range_lookup.cu — batched range pruning on the dist-server side.
//
// Before a publish is sent to the dist-workers, TenantRangeLookupCache.lookup
// (bifromq-dist/bifromq-dist-server/src/main/java/org/apache/bifromq/dist/server/scheduler/TenantRangeLookupCache.java:70-106) decides
// which KV ranges of the tenant can hold a matching route: every range publishes a Fact {firstGlobalFilterLevels,
// lastGlobalFilterLevels} (the smallest and largest filter it stores, tenant id as level 0), and a range stays a candidate iff
// the topic's EXPANSION SET (every filter that matches the topic) has a member inside [first, last]. The reference builds a
// one-topic trie and runs its lazy expansion iterator: seek(first), then compares the filter found with `last` — per topic, per
// candidate, behind a Caffeine cache. Here the whole batch is answered by one kernel, one thread per (topic, candidate):
//
// The expansion set of a topic t_1/../t_n (global mode: level 0 is the tenant id and is never wildcard-matched,
// TopicTrieNode.java:146-152) is an implicit trie: after i matched levels the children are "#" (terminal; not under the tenant
// level when t_1 starts with '$'), "+" and t_{i+1} (the last two only while i < n; "+" not for a '$' first level), and the
// node with i == n is itself a filter. seek(B) = the least member >= B in level-wise String.compareTo order is a lower-bound
// walk of that trie: follow B while it is a path, remember the deepest level that has a greater sibling, and complete
// minimally ("#" as soon as it is the smallest child). No trie is materialised; a member is a bitmask of '+' choices plus
// where it ends. UTF-8 byte order equals UTF-16 code-unit order on BMP text (what the MQTT edge admits).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bfq_gpumatch.h"

namespace bfq {
int32_t set_error(int32_t code, const std::string& msg);
}

namespace {

constexpr int RL_MAX_LEVELS = 34;   // levels of a topic / of a bound this kernel handles (MaxTopicLevels default is 16)

struct Str {
    const uint8_t* p;
    int n;
};
__device__ __forceinline__ int cmp(Str a, Str b) {   // bytewise, shorter prefix first
    const int m = a.n < b.n ? a.n : b.n;
    for (int i = 0; i < m; i++)
        if (a.p[i] != b.p[i]) return a.p[i] < b.p[i] ? -1 : 1;
    return a.n == b.n ? 0 : (a.n < b.n ? -1 : 1);
}
__device__ __forceinline__ int split(const uint8_t* s, int len, uint8_t sep, int* start, int* end) {
    int n = 0, b = 0;
    for (int i = 0; i <= len; i++)
        if (i == len || s[i] == sep) {
            if (n < RL_MAX_LEVELS) {
                start[n] = b;
                end[n] = i;
            }
            n++;
            b = i + 1;
        }
    return n;
}

// a member of the expansion set: levels 1..depth choose '+' where the bit is set, else the topic's level; `hash` = it ends with
// a "#" level behind them (else it is the full-length filter: depth == n)
struct Member {
    uint64_t plus;
    int depth;
    bool hash;
};

struct Ctx {
    const uint8_t* topic;
    int ts[RL_MAX_LEVELS], te[RL_MAX_LEVELS], n;   // topic levels
    bool sys;                                       // t_1 starts with '$'
    __device__ Str t(int i) const { return Str{topic + ts[i - 1], te[i - 1] - ts[i - 1]}; }   // 1-based
    // children of the node behind i matched levels, smallest first; kinds: 0 = "#", 1 = "+", 2 = t_{i+1}
    __device__ int children(int i, int* kind) const {
        const bool wild = !(i == 0 && sys);
        int c = 0;
        if (i < n) {
            const uint8_t H = '#', P = '+';
            const Str h{&H, 1}, pl{&P, 1};
            const Str tx = t(i + 1);
            // order "#" < "+" always; place t_{i+1} among them
            const int ch = cmp(tx, h), cp = cmp(tx, pl);
            if (!wild) {
                kind[c++] = 2;
            } else if (ch < 0) {
                kind[c++] = 2; kind[c++] = 0; kind[c++] = 1;
            } else if (ch == 0) {            // the topic level is literally "#" (not a valid topic, but keep the order total)
                kind[c++] = 0; kind[c++] = 1;
            } else if (cp < 0) {
                kind[c++] = 0; kind[c++] = 2; kind[c++] = 1;
            } else if (cp == 0) {
                kind[c++] = 0; kind[c++] = 1;
            } else {
                kind[c++] = 0; kind[c++] = 1; kind[c++] = 2;
            }
        } else if (wild) {
            kind[c++] = 0;   // "#" matches the parent level
        }
        return c;
    }
    // smallest member in the subtree of the node behind i matched levels (path so far in m.plus)
    __device__ Member complete(Member m, int i) const {
        while (i < n) {
            int kind[3];
            const int c = children(i, kind);
            (void) c;
            if (kind[0] == 0) {
                m.depth = i;
                m.hash = true;
                return m;
            }
            if (kind[0] == 1) m.plus |= 1ull << i;
            i++;
        }
        m.depth = n;
        m.hash = false;
        return m;
    }
    __device__ Str level_of(const Member& m, int i, const uint8_t* H, const uint8_t* P) const {   // level i (1-based) of member m
        if (m.hash && i == m.depth + 1) return Str{H, 1};
        if ((m.plus >> (i - 1)) & 1ull) return Str{P, 1};
        return t(i);
    }
    __device__ int levels_of(const Member& m) const { return m.depth + (m.hash ? 1 : 0); }
};

// least member >= bound (bound levels b[0..k) WITHOUT the tenant level); false: none
__device__ bool seek(const Ctx& c, const uint8_t* bound, const int* bs, const int* be, int k, Member* out) {
    const uint8_t H = '#', P = '+';
    Member m{0ull, 0, false};
    int fb_depth = -1, fb_kind = 0;   // deepest level on the tight path with a child greater than the bound's level
    uint64_t fb_plus = 0;
    int i = 0;                         // matched levels so far (tight)
    while (true) {
        if (i >= k) {                  // the bound is exhausted: everything below this node is >= it
            *out = c.complete(m, i);
            return true;
        }
        const Str b{bound + bs[i], be[i] - bs[i]};
        int kind[3];
        const int nc = c.children(i, kind);
        int eq = -1, gt = -1;
        for (int j = 0; j < nc; j++) {
            const Str s = kind[j] == 0 ? Str{&H, 1} : kind[j] == 1 ? Str{&P, 1} : c.t(i + 1);
            const int r = cmp(s, b);
            if (r == 0) eq = kind[j];
            else if (r > 0 && gt < 0) gt = kind[j];
        }
        if (gt >= 0) {
            fb_depth = i;
            fb_kind = gt;
            fb_plus = m.plus;
        }
        if (eq == 0) {                 // "#": terminal. Equal to the bound iff the bound ends here too, else it is a proper prefix (<)
            if (i + 1 == k) {
                m.depth = i;
                m.hash = true;
                *out = m;
                return true;
            }
            break;
        }
        if (eq > 0) {
            if (eq == 1) m.plus |= 1ull << i;
            i++;
            if (i == c.n && i == k) {  // the full-length filter equals the bound
                m.depth = c.n;
                m.hash = false;
                *out = m;
                return true;
            }
            continue;
        }
        break;                         // no child equals the bound's level: leave the tight path
    }
    if (fb_depth < 0) return false;
    m.plus = fb_plus;
    if (fb_kind == 0) {
        m.depth = fb_depth;
        m.hash = true;
        *out = m;
        return true;
    }
    if (fb_kind == 1) m.plus |= 1ull << fb_depth;
    *out = c.complete(m, fb_depth + 1);
    return true;
}

// 0 = the range cannot hold a match, 1 = candidate, 2 = seek past the end (the reference stops looking at later candidates)
__global__ void __launch_bounds__(128) range_lookup_kernel(int64_t n_pairs, const int64_t* pair_off, int64_t n_topics, const uint8_t* topics,
                                                           const int64_t* topic_off, const int32_t* topic_tenant, const uint8_t* tenants,
                                                           const int64_t* tenant_off, const int64_t* cand_off, const uint8_t* first_blob,
                                                           const int64_t* first_off, const uint8_t* last_blob, const int64_t* last_off,
                                                           uint8_t* out) {
    const int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pairs) return;
    int64_t lo = 0, hi = n_topics;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (pair_off[mid] <= j) lo = mid;
        else hi = mid;
    }
    const int64_t ti = lo;
    const int tn = topic_tenant[ti];
    const int64_t cand = cand_off[tn] + (j - pair_off[ti]);
    Ctx c;
    c.topic = topics + topic_off[ti];
    const int tlen = (int) (topic_off[ti + 1] - topic_off[ti]);
    c.n = split(c.topic, tlen, '/', c.ts, c.te);
    c.sys = tlen > 0 && c.topic[0] == '$';
    const uint8_t* fb = first_blob + first_off[cand];
    const uint8_t* lb = last_blob + last_off[cand];
    int fs[RL_MAX_LEVELS], fe[RL_MAX_LEVELS], ls[RL_MAX_LEVELS], le[RL_MAX_LEVELS];
    const int fk = split(fb, (int) (first_off[cand + 1] - first_off[cand]), 0, fs, fe);
    const int lk = split(lb, (int) (last_off[cand + 1] - last_off[cand]), 0, ls, le);
    if (c.n > RL_MAX_LEVELS || fk > RL_MAX_LEVELS || lk > RL_MAX_LEVELS) {
        out[j] = 3;   // unsupported depth: reported to the caller as an error
        return;
    }
    const Str tenant{tenants + tenant_off[tn], (int) (tenant_off[tn + 1] - tenant_off[tn])};
    // level 0 is the tenant id: every member starts with it
    Member m;
    const int r0 = cmp(tenant, Str{fb + fs[0], fe[0] - fs[0]});
    if (r0 < 0) {
        out[j] = 2;   // the bound's tenant sorts behind this tenant: nothing >= first
        return;
    }
    bool found;
    if (r0 > 0) {
        m = c.complete(Member{0ull, 0, false}, 0);   // every member is greater: the smallest one
        found = true;
    } else {
        found = seek(c, fb, fs + 1, fe + 1, fk - 1, &m);
    }
    if (!found) {
        out[j] = 2;
        return;
    }
    // found == first, or found <= last (level-wise)
    const uint8_t H = '#', P = '+';
    const int ml = c.levels_of(m);
    bool equal_first = r0 == 0 && ml == fk - 1;
    for (int i = 1; equal_first && i <= ml; i++) equal_first = cmp(c.level_of(m, i, &H, &P), Str{fb + fs[i], fe[i] - fs[i]}) == 0;
    int r = cmp(tenant, Str{lb + ls[0], le[0] - ls[0]});
    for (int i = 1; r == 0; i++) {
        const bool me = i > ml, le_ = i > lk - 1;
        if (me || le_) {
            r = me && le_ ? 0 : (me ? -1 : 1);
            break;
        }
        r = cmp(c.level_of(m, i, &H, &P), Str{lb + ls[i], le[i] - ls[i]});
    }
    out[j] = (equal_first || r <= 0) ? 1 : 0;
}

int32_t rl_fail(int32_t code, const std::string& msg) { return bfq::set_error(code, msg); }
#define RL_CUDA(expr)                                                                               \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            for (void* q : allocs) cudaFree(q);                                                     \
            return rl_fail(BFQ_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
        }                                                                                           \
    } while (0)

}  // namespace

extern "C" int32_t bfq_range_lookup(int32_t device_ordinal, const uint8_t* tenants, const int64_t* tenant_off, int32_t n_tenants,
                                    const uint8_t* topics, const int64_t* topic_off, const int32_t* topic_tenant, int64_t n_topics,
                                    const int64_t* cand_off, const uint8_t* cand_flags, const uint8_t* first_blob, const int64_t* first_off,
                                    const uint8_t* last_blob, const int64_t* last_off, int64_t* keep_off_out, uint8_t* keep_out) {
    if (n_tenants < 0 || n_topics < 0 || !keep_off_out || (n_topics > 0 && (!topics || !topic_off || !topic_tenant || !keep_out)) ||
        (n_tenants > 0 && (!tenants || !tenant_off || !cand_off)))
        return rl_fail(BFQ_E_INVALID, "bad argument");
    const int64_t n_cand = n_tenants ? cand_off[n_tenants] : 0;
    if (n_cand > 0 && (!cand_flags || !first_off || !last_off)) return rl_fail(BFQ_E_INVALID, "NULL candidate arrays");
    // rows: topic i has one cell per candidate of its tenant
    keep_off_out[0] = 0;
    for (int64_t i = 0; i < n_topics; i++) {
        const int t = topic_tenant[i];
        if (t < 0 || t >= n_tenants) return rl_fail(BFQ_E_RANGE, "topic_tenant out of range");
        keep_off_out[i + 1] = keep_off_out[i] + (cand_off[t + 1] - cand_off[t]);
    }
    const int64_t n_pairs = keep_off_out[n_topics];
    if (n_pairs == 0) return BFQ_OK;
    std::vector<void*> allocs;
    RL_CUDA(cudaSetDevice(device_ordinal));
    auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return e;
        allocs.push_back(*dst);
        return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    void *d_pair_off, *d_topics, *d_topic_off, *d_tt, *d_tenants, *d_tenant_off, *d_cand_off, *d_first, *d_first_off, *d_last, *d_last_off, *d_out;
    RL_CUDA(up(keep_off_out, (size_t) (n_topics + 1) * 8, &d_pair_off));
    RL_CUDA(up(topics + topic_off[0], (size_t) (topic_off[n_topics] - topic_off[0]), &d_topics));
    std::vector<int64_t> toff((size_t) n_topics + 1);
    for (int64_t i = 0; i <= n_topics; i++) toff[(size_t) i] = topic_off[i] - topic_off[0];
    RL_CUDA(up(toff.data(), toff.size() * 8, &d_topic_off));
    RL_CUDA(up(topic_tenant, (size_t) n_topics * 4, &d_tt));
    RL_CUDA(up(tenants, (size_t) tenant_off[n_tenants], &d_tenants));
    RL_CUDA(up(tenant_off, (size_t) (n_tenants + 1) * 8, &d_tenant_off));
    RL_CUDA(up(cand_off, (size_t) (n_tenants + 1) * 8, &d_cand_off));
    RL_CUDA(up(first_blob, (size_t) first_off[n_cand], &d_first));
    RL_CUDA(up(first_off, (size_t) (n_cand + 1) * 8, &d_first_off));
    RL_CUDA(up(last_blob, (size_t) last_off[n_cand], &d_last));
    RL_CUDA(up(last_off, (size_t) (n_cand + 1) * 8, &d_last_off));
    RL_CUDA(cudaMalloc(&d_out, (size_t) n_pairs));
    allocs.push_back(d_out);
    range_lookup_kernel<<<(unsigned) ((n_pairs + 127) / 128), 128>>>(n_pairs, (const int64_t*) d_pair_off, n_topics, (const uint8_t*) d_topics,
                                                                    (const int64_t*) d_topic_off, (const int32_t*) d_tt, (const uint8_t*) d_tenants,
                                                                    (const int64_t*) d_tenant_off, (const int64_t*) d_cand_off,
                                                                    (const uint8_t*) d_first, (const int64_t*) d_first_off, (const uint8_t*) d_last,
                                                                    (const int64_t*) d_last_off, (uint8_t*) d_out);
    RL_CUDA(cudaGetLastError());
    RL_CUDA(cudaMemcpy(keep_out, d_out, (size_t) n_pairs, cudaMemcpyDeviceToHost));
    for (void* q : allocs) cudaFree(q);
    allocs.clear();
    // the reference's candidate loop (TenantRangeLookupCache.java:78-104): a range without a Fact is kept, one whose Fact lacks
    // first or last is empty (skipped), and the first range whose seek runs past the end ends the scan
    for (int64_t i = 0; i < n_topics; i++) {
        const int t = topic_tenant[i];
        bool stopped = false;
        for (int64_t k = 0; k < cand_off[t + 1] - cand_off[t]; k++) {
            uint8_t& cell = keep_out[keep_off_out[i] + k];
            const uint8_t fl = cand_flags[cand_off[t] + k];
            if (cell == 3) return rl_fail(BFQ_E_RANGE, "a topic or a range bound has more levels than bfq_range_lookup handles");
            if (stopped) {
                cell = 0;
            } else if (!(fl & 1)) {
                cell = 1;
            } else if ((fl & 6) != 6) {
                cell = 0;
            } else if (cell == 2) {
                cell = 0;
                stopped = true;
            }
        }
    }
    return BFQ_OK;
}
