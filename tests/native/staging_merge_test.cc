// staging_merge_test.cc — host-only check of Staging (bifromq_b200/csrc/index_builder.cc): load + random upsert / erase rounds + merge_all must equal a
// std::map kept beside it (the delta merge copies base runs in bulk between the delta keys). Built and run by tests/test_host_cpu.py.
#include <cstdio>
#include <map>
#include <random>
#include <string>
#include "index_builder.h"
using namespace bfq;
static std::string mk(const std::string& tenant, const std::string& rest) {
    std::string k; k.push_back(0); k.push_back((char)(tenant.size() >> 8)); k.push_back((char) tenant.size()); k += tenant; k += rest; return k;
}
int main() {
    std::mt19937 rng(7);
    for (int round = 0; round < 200; round++) {
        Staging st;
        std::map<std::string, std::string> ref;
        // base
        std::vector<std::pair<std::string, std::string>> base;
        int nb = rng() % 400;
        for (int i = 0; i < nb; i++) {
            std::string t = (rng() % 3 == 0) ? "ta" : (rng() % 2 ? "tb" : "t");
            std::string rest; int L = 1 + rng() % 6; for (int j = 0; j < L; j++) rest.push_back('a' + rng() % 4);
            ref[mk(t, rest)] = std::string(1 + rng() % 5, 'v');
        }
        {
            std::vector<uint8_t> kb, vb; std::vector<int64_t> ko{0}, vo{0};
            for (auto& kv : ref) { kb.insert(kb.end(), kv.first.begin(), kv.first.end()); vb.insert(vb.end(), kv.second.begin(), kv.second.end()); ko.push_back(kb.size()); vo.push_back(vb.size()); }
            std::string err;
            if (!st.load(kb.data(), ko.data(), vb.data(), vo.data(), (int64_t) ref.size(), &err)) { printf("load failed %s\n", err.c_str()); return 1; }
        }
        for (int commit = 0; commit < 4; commit++) {
            int nd = rng() % 30;
            for (int i = 0; i < nd; i++) {
                std::string t = (rng() % 3 == 0) ? "ta" : (rng() % 2 ? "tb" : (rng() % 5 ? "t" : "tnew"));
                std::string rest; int L = 1 + rng() % 6; for (int j = 0; j < L; j++) rest.push_back('a' + rng() % 4);
                std::string k = mk(t, rest);
                if (rng() % 3 == 0) { st.erase(k); ref.erase(k); }
                else { std::string v(1 + rng() % 7, 'w'); st.upsert(k, v); ref[k] = v; }
            }
            st.merge_all();
            KVBlob all = st.concat();
            if ((size_t) all.n() != ref.size()) { printf("round %d: size %lld vs %zu\n", round, (long long) all.n(), ref.size()); return 1; }
            int64_t i = 0;
            for (auto& kv : ref) {
                if (all.key(i) != sv(kv.first) || all.val(i) != sv(kv.second)) { printf("round %d: mismatch at %lld\n", round, (long long) i); return 1; }
                i++;
            }
        }
    }
    printf("merge ok\n");
    return 0;
}
