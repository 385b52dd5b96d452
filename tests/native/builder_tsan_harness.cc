// builder_tsan_harness.cc — host-only harness for ThreadSanitizer runs of the index builder (threads inside a large tenant: plan,
// record emission, slot fill; workers across tenants). Not part of the pytest suite (a TSAN build takes a minute):
//   g++ -O1 -g -fsanitize=thread -std=c++17 -Ibifromq_b200/csrc tests/native/builder_tsan_harness.cc bifromq_b200/csrc/index_builder.cc \
//       bifromq_b200/csrc/codec.cc -lpthread -o /tmp/builder_tsan && /tmp/builder_tsan     (round 2: no reports)
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include "index_builder.h"
using namespace bfq;
int32_t bfq::set_error(int32_t code, const std::string&) { return code; }
int main() {
    std::mt19937 rng(11);
    std::map<std::string, std::string> kvs;
    const char* tenants[] = {"big", "small"};
    for (int t = 0; t < 2; t++) {
        const int nf = t == 0 ? 260000 : 2000;
        for (int i = 0; i < nf; i++) {
            std::string f;
            int depth = 2 + rng() % 4;
            for (int d = 0; d < depth; d++) {
                if (d) f.push_back('/');
                if (rng() % 9 == 0) f += "+";
                else f += "l" + std::to_string(rng() % (d == 0 ? 50 : 400));
            }
            if (rng() % 6 == 0) f += "/#";
            std::string url = make_receiver_url(rng() % 2, "r" + std::to_string(rng() % 1000), "d");
            kvs[make_route_key(tenants[t], f, url)] = std::string(8, '\1');
        }
    }
    std::vector<uint8_t> kb, vb; std::vector<int64_t> ko{0}, vo{0};
    for (auto& kv : kvs) { kb.insert(kb.end(), kv.first.begin(), kv.first.end()); vb.insert(vb.end(), kv.second.begin(), kv.second.end()); ko.push_back(kb.size()); vo.push_back(vb.size()); }
    Staging st; std::string err;
    if (!st.load(kb.data(), ko.data(), vb.data(), vo.data(), (int64_t) kvs.size(), &err)) { printf("load: %s\n", err.c_str()); return 1; }
    std::vector<const KVBlob*> parts;
    for (auto& kvp : st.tenants()) parts.push_back(kvp.second.base.get());
    FlatIndex flat;
    if (!build_flat_index_parts(parts, &flat, &err)) { printf("build: %s\n", err.c_str()); return 1; }
    printf("nodes %lld slots %u tenants %zu\n", (long long) flat.n_nodes, flat.n_slots, flat.tenants.size());
    size_t ti = 0;
    for (auto& kvp : st.tenants()) {
        const TenantMeta& m = flat.tenants[ti++];
        TenantImage img;
        if (!build_tenant_image(*kvp.second.base, sv(m.tenant), m.ordinal, m.lo, m.region_base, m.seg_base, m.pp_base, m.pg_base, &img, &err)) { printf("image: %s\n", err.c_str()); return 1; }
        if (m.big_edges == 0 && memcmp(img.slots.data(), flat.slots.data() + m.region_base, (size_t) m.csr_slots * sizeof(Slot)) != 0) { printf("image differs\n"); return 1; }
    }
    printf("tsan run ok\n");
    return 0;
}
