// fanout.cu — fan-out expansion on the GPU: the step right behind the match.
//
// The reference walks every matched route of a message on the CPU: DeliverExecutorGroup.submit
// (bifromq-dist/bifromq-dist-worker/src/main/java/org/apache/bifromq/dist/worker/DeliverExecutorGroup.java:112-231) iterates the
// route set, resolves a shared subscription to ONE member (send(GroupMatching) :242-278: a uniformly random member for
// $share, a rendezvous hash of the publisher for $oshare), and DeliverExecutor.send (DeliverExecutor.java:89-93) turns each
// route into a DeliveryCall keyed by (subBrokerId, delivererKey), which the deliverer batches into one DeliveryPack list per
// deliverer. With the matched route ranks already on the device that is a group-by:
//   input   the device CSR of a completed match (surviving ranks per topic, caps applied: bfq_expand_device)
//   output  every (topic, route) pair of the batch grouped by DELIVERER id: pack_offsets[D + 1], pack_topic[], pack_rank[]
//           (+ pack_member[] for shared subscriptions: the index of the member the pair was resolved to)
// A deliverer id is a dense index over the distinct (subBrokerId, delivererKey) pairs of the index, interned on the host when a
// snapshot is first used for fan-out (bfq_fanout_deliverer gives the pair back). Unordered shared subscriptions pick member
// hash(topic position, route rank) mod n — the reference picks uniformly at random (ThreadLocalRandom), so any member is a valid
// outcome and the pick here is reproducible; ORDERED shared subscriptions need the publisher of each message
// (RendezvousHash over ClientInfo.hashCode(), :253-270) which a topic batch does not carry: their pairs are grouped under the
// reserved deliverer id BFQ_FANOUT_ORDERED_SHARE with every member left to the host.
//
// Kernels: a two-pass radix partition on the deliverer id. Pass 1 counts per (CTA tile, deliverer) in shared memory; a scan over
// the [deliverer][tile] count matrix gives every tile its write cursor per deliverer; pass 2 re-reads the tile and scatters.
// Per pair: one 8-byte rank read (streaming), one 4-byte deliverer-id read (the table is 4 bytes per route: L2 resident at
// 10M filters), 12 bytes written. HBM-streaming bound.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "../../include/bfq_gpumatch.h"
#include "codec.h"
#include "fanout.h"

namespace bfq {

namespace {

constexpr int FO_THREADS = 256;
constexpr int FO_TILE = 4096;          // pairs per tile
constexpr uint32_t FO_MAX_D = 8192;    // deliverer ids a tile counts in shared memory (32 KB); more -> BFQ_E_RANGE

__device__ __forceinline__ uint32_t fo_mix(uint32_t a, uint32_t b) {
    uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u + (a << 6) + (a >> 2));
    h ^= h >> 15;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    return h;
}

// deliverer of the pair (topic position t, rank r); *member = member index of a shared subscription or 0xFFFFFFFF
__device__ __forceinline__ uint32_t fo_deliverer(const FanoutParams& p, uint32_t t, int64_t r, uint32_t* member) {
    const uint32_t d = p.rdeliv[r];
    *member = 0xFFFFFFFFu;
    if (!(d & FO_GROUP_BIT)) return d;
    const uint32_t g = d & ~FO_GROUP_BIT;                 // index into the group table
    const uint32_t b = p.gmem_off[g], n = p.gmem_off[g + 1] - b;
    if (n == 0) return p.n_deliverers - 1;                // empty group: nothing to deliver (parked under the ordered-share id)
    if (p.gordered[g]) return p.n_deliverers - 1;         // BFQ_FANOUT_ORDERED_SHARE: the host picks per publisher
    const uint32_t m = fo_mix(t, (uint32_t) r) % n;
    *member = m;
    return p.gmem_deliv[b + m];
}

// topic of pair position j: binary search in offsets (monotone), amortised by doing it once per thread then walking
__device__ __forceinline__ uint32_t fo_topic_of(const int64_t* offsets, int64_t n_topics, int64_t j) {
    int64_t lo = 0, hi = n_topics;   // offsets[lo] <= j < offsets[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= j) lo = mid;
        else hi = mid;
    }
    return (uint32_t) lo;
}

__global__ void __launch_bounds__(FO_THREADS) fanout_count_kernel(const FanoutParams p) {
    extern __shared__ uint32_t hist[];
    for (uint32_t i = threadIdx.x; i < p.n_deliverers; i += FO_THREADS) hist[i] = 0;
    __syncthreads();
    // a thread owns FO_TILE / FO_THREADS consecutive pairs (one 128-byte line of ranks): one binary search for the first one's
    // topic, then the topic index only walks forward
    constexpr int PER = FO_TILE / FO_THREADS;
    const int64_t j0 = (int64_t) blockIdx.x * FO_TILE + (int64_t) threadIdx.x * PER;
    if (j0 < p.n_pairs) {
        uint32_t t = fo_topic_of(p.offsets, p.n_topics, j0);
        for (int q = 0; q < PER && j0 + q < p.n_pairs; q++) {
            const int64_t j = j0 + q;
            while (p.offsets[t + 1] <= j) t++;
            uint32_t member;
            atomicAdd(&hist[fo_deliverer(p, t, p.ranks[j], &member)], 1u);
        }
    }
    __syncthreads();
    // count matrix in [deliverer][tile] order: its exclusive scan is, for every deliverer, the cursor of every tile
    for (uint32_t i = threadIdx.x; i < p.n_deliverers; i += FO_THREADS) p.tile_counts[(uint64_t) i * gridDim.x + blockIdx.x] = hist[i];
}

__global__ void __launch_bounds__(FO_THREADS) fanout_scatter_kernel(const FanoutParams p) {
    extern __shared__ uint32_t cur[];
    for (uint32_t i = threadIdx.x; i < p.n_deliverers; i += FO_THREADS) cur[i] = 0;
    __syncthreads();
    constexpr int PER = FO_TILE / FO_THREADS;
    const int64_t j0 = (int64_t) blockIdx.x * FO_TILE + (int64_t) threadIdx.x * PER;
    if (j0 < p.n_pairs) {
        uint32_t t = fo_topic_of(p.offsets, p.n_topics, j0);
        for (int q = 0; q < PER && j0 + q < p.n_pairs; q++) {
            const int64_t j = j0 + q;
            while (p.offsets[t + 1] <= j) t++;
            uint32_t member;
            const int64_t r = p.ranks[j];
            const uint32_t d = fo_deliverer(p, t, r, &member);
            const uint32_t at = p.tile_base[(uint64_t) d * gridDim.x + blockIdx.x] + atomicAdd(&cur[d], 1u);
            p.pack_topic[at] = t;
            p.pack_rank[at] = (uint32_t) r;
            if (p.pack_member) p.pack_member[at] = member;
        }
    }
}

// pack_offsets[d] = tile_base[d][0]; pack_offsets[D] = n_pairs
__global__ void fanout_offsets_kernel(const FanoutParams p, uint32_t n_tiles) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < p.n_deliverers) p.pack_offsets[d] = (long long) p.tile_base[(uint64_t) d * n_tiles];
    if (d == p.n_deliverers) p.pack_offsets[d] = (long long) p.n_pairs;
}

}  // namespace

cudaError_t launch_fanout(const FanoutParams& p, void* d_tmp, size_t* tmp_bytes, cudaStream_t stream) {
    const uint32_t n_tiles = (uint32_t) std::max<int64_t>(1, (p.n_pairs + FO_TILE - 1) / FO_TILE);
    const size_t cells = (size_t) p.n_deliverers * n_tiles;
    if (!d_tmp) return cub::DeviceScan::ExclusiveSum(nullptr, *tmp_bytes, p.tile_counts, p.tile_base, (int) std::min<size_t>(cells, 0x7FFFFFFF), stream);
    const size_t smem = (size_t) p.n_deliverers * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(fanout_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        cudaFuncSetAttribute(fanout_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    }
    fanout_count_kernel<<<n_tiles, FO_THREADS, smem, stream>>>(p);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(d_tmp, *tmp_bytes, p.tile_counts, p.tile_base, (int) cells, stream);
    if (e != cudaSuccess) return e;
    fanout_scatter_kernel<<<n_tiles, FO_THREADS, smem, stream>>>(p);
    fanout_offsets_kernel<<<(p.n_deliverers + 1 + 255) / 256, 256, 0, stream>>>(p, n_tiles);
    return cudaGetLastError();
}

uint32_t fanout_max_deliverers() { return FO_MAX_D; }
int64_t fanout_tile() { return FO_TILE; }

// ------------------------------------------------------------------------------------------------ host: interning
namespace {
// "<decimal subBrokerId>\0<receiverId>\0<delivererKey>"  (DWS/KVSchemaUtil.java:56-58; parsed by cache/ReceiverCache.java:32-36)
bool split_receiver_url(sv url, int32_t* broker, sv* deliverer_key) {
    const size_t a = url.find('\0');
    if (a == sv::npos) return false;
    const size_t b = url.find('\0', a + 1);
    if (b == sv::npos) return false;
    int64_t v = 0;
    if (a == 0) return false;
    for (size_t i = 0; i < a; i++) {
        if (url[i] < '0' || url[i] > '9') return false;
        v = v * 10 + (url[i] - '0');
        if (v > 0x7FFFFFFF) return false;
    }
    *broker = (int32_t) v;
    *deliverer_key = url.substr(b + 1);
    return true;
}
// RouteGroup { map<string, uint64> members = 1; }  (bifromq-dist-worker-schema/src/main/proto/distservice/RouteGroup.proto:27-29):
// repeated field 1, each a nested message {1: string key, 2: varint value}. Calls f(receiverUrl) per member in wire order.
template <typename F>
bool for_each_group_member(sv b, F&& f) {
    size_t i = 0;
    auto varint = [&](uint64_t* out) {
        uint64_t v = 0;
        int shift = 0;
        while (i < b.size()) {
            const uint8_t c = (uint8_t) b[i++];
            v |= (uint64_t) (c & 0x7F) << shift;
            if (!(c & 0x80)) {
                *out = v;
                return true;
            }
            shift += 7;
            if (shift > 63) return false;
        }
        return false;
    };
    while (i < b.size()) {
        uint64_t tag, len;
        if (!varint(&tag)) return false;
        if (tag != ((1u << 3) | 2u)) return false;
        if (!varint(&len) || i + len > b.size()) return false;
        const size_t end = i + (size_t) len;
        sv key;
        while (i < end) {
            uint64_t t2;
            if (!varint(&t2)) return false;
            if (t2 == ((1u << 3) | 2u)) {
                uint64_t kl;
                if (!varint(&kl) || i + kl > end) return false;
                key = b.substr(i, (size_t) kl);
                i += (size_t) kl;
            } else if (t2 == (2u << 3)) {
                uint64_t v;
                if (!varint(&v)) return false;
            } else {
                return false;
            }
        }
        f(key);
    }
    return true;
}
}  // namespace

uint32_t DelivererTable::intern(int32_t broker, sv key) {
    std::string k = std::to_string(broker);
    k.push_back('\0');
    k.append(key);
    std::lock_guard<std::mutex> g(mu);
    auto it = ids.find(k);
    if (it != ids.end()) return it->second;
    const uint32_t id = (uint32_t) list.size();
    list.emplace_back(broker, std::string(key));
    ids.emplace(std::move(k), id);
    return id;
}

bool build_tenant_fan(const KVBlob& kv, DelivererTable* table, TenantFan* out, std::string* err) {
    const int64_t n = kv.n();
    out->rdeliv.assign((size_t) n, 0);
    out->gmem_off.assign(1, 0);
    out->gmem_deliv.clear();
    out->gordered.clear();
    for (int64_t r = 0; r < n; r++) {
        DecodedKey d;
        if (!decode_route_key(kv.key(r), &d)) {
            if (err) *err = "undecodable route key";
            return false;
        }
        if (d.kind != KIND_GROUP) {
            int32_t broker = 0;
            sv dk;
            if (!split_receiver_url(d.receiver, &broker, &dk)) {
                if (err) *err = "receiver url without subBrokerId / delivererKey";
                return false;
            }
            out->rdeliv[(size_t) r] = table->intern(broker, dk);
            continue;
        }
        out->rdeliv[(size_t) r] = FO_GROUP_BIT | (uint32_t) out->gordered.size();
        out->gordered.push_back(d.flag == FLAG_ORDERED ? 1 : 0);
        bool ok = true;
        const bool parsed = for_each_group_member(kv.val(r), [&](sv url) {
            int32_t broker = 0;
            sv dk;
            if (!split_receiver_url(url, &broker, &dk)) {
                ok = false;
                return;
            }
            out->gmem_deliv.push_back(table->intern(broker, dk));
        });
        if (!parsed || !ok) {
            if (err) *err = "undecodable RouteGroup value";
            return false;
        }
        out->gmem_off.push_back((uint32_t) out->gmem_deliv.size());
    }
    return true;
}

}  // namespace bfq
