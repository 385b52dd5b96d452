// oracle/matcher.cc — CPU ORACLE (test infrastructure only; see oracle.h).
// Three forward matchers that must agree:
//   (1) match_all_reference — literal restatement of DW/cache/TenantRouteMatcher.java:68-161
//       (topic trie -> expansion-set cursor -> merge-join with the sorted KV, probe-20-then-seek)
//       with the caps of DW/cache/MatchedRoutes.java:87-141;
//   (2) match_all_brute — the SURVEY.md §8a predicate applied to every (topic, route) pair;
//   (3) match_all_trie — a straightforward per-topic filter-trie walk (the strongest honest
//       CPU competitor; also yields the §8(d) algorithmic-byte counters V / P / R).
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <unordered_map>

#include "oracle.h"

namespace orc {

// ---------------------------------------------------------------- SortedKV
void SortedKV::freeze() {
    if (frozen_) return;
    order_.clear();
    order_.reserve(kv_.size());
    for (const auto& e : kv_) order_.push_back(&e);
    frozen_ = true;
}
int64_t SortedKV::lower_bound(const std::string& k) const {
    if (!frozen_) throw std::runtime_error("SortedKV not frozen");
    int64_t lo = 0, hi = (int64_t) order_.size();
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (order_[mid]->first < k) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------- MatchedRoutes (caps)
namespace {
struct MatchedRoutes {  // DW/cache/MatchedRoutes.java
    std::set<int64_t> allMatchings;                       // by KV rank (keys are unique => value equality == rank equality)
    std::map<std::string, int64_t> groupMatchings;        // mqttTopicFilter -> rank
    int persistentFanout = 0;
    int maxPersistentFanout, maxGroupFanout;
    int topicIdx;
    std::vector<ThrottleEvent>* events;

    void add_normal(int64_t rank, const Matching& m) {    // :87-108
        if (allMatchings.insert(rank).second) {
            if (m.subBrokerId == 1) {
                if (persistentFanout < maxPersistentFanout) {
                    persistentFanout++;
                } else {
                    allMatchings.erase(rank);
                    events->push_back({ThrottleEvent::Persistent, topicIdx, rank, maxPersistentFanout});
                }
            }
        }
    }
    void put_group(int64_t rank, const Matching& m) {     // :119-141
        auto it = groupMatchings.find(m.mqttTopicFilter);
        if (it == groupMatchings.end()) {
            groupMatchings[m.mqttTopicFilter] = rank;
            if ((int) groupMatchings.size() <= maxGroupFanout) {
                allMatchings.insert(rank);
            } else {
                groupMatchings.erase(m.mqttTopicFilter);
                events->push_back({ThrottleEvent::Group, topicIdx, rank, maxGroupFanout});
            }
        } else {
            allMatchings.erase(it->second);
            it->second = rank;
            allMatchings.insert(rank);
        }
    }
    void add(int64_t rank, const Matching& m) {
        if (m.type == Matching::Normal) add_normal(rank, m); else put_group(rank, m);
    }
};

std::vector<MatchedRoutes> new_results(size_t n, int maxP, int maxG, std::vector<ThrottleEvent>* ev) {
    std::vector<MatchedRoutes> r(n);
    for (size_t i = 0; i < n; i++) {
        r[i].maxPersistentFanout = maxP;
        r[i].maxGroupFanout = maxG;
        r[i].topicIdx = (int) i;
        r[i].events = ev;
    }
    return r;
}

void finish(std::vector<MatchedRoutes>& mr, MatchResult& out) {
    out.routes.resize(mr.size());
    out.persistentFanout.resize(mr.size());
    out.groupFanout.resize(mr.size());
    for (size_t i = 0; i < mr.size(); i++) {
        out.routes[i].assign(mr[i].allMatchings.begin(), mr[i].allMatchings.end());
        out.persistentFanout[i] = mr[i].persistentFanout;
        out.groupFanout[i] = (int) mr[i].groupMatchings.size();
    }
}
}  // namespace

// ---------------------------------------------------------------- (1) literal reference algorithm
MatchResult match_all_reference(const SortedKV& kv, const std::string& tenantId, const std::vector<std::string>& topics,
                                int maxPersistentFanout, int maxGroupFanout) {
    MatchResult out;
    auto matchedRoutes = new_results(topics.size(), maxPersistentFanout, maxGroupFanout, &out.events);
    TopicTrie topicTrie(false);
    for (size_t i = 0; i < topics.size(); i++) topicTrie.add_topic(parse(topics[i], false), (int) i);

    std::string tenantStartKey = tenant_begin_key(tenantId);
    bool openEnd = false;
    std::string tenantEndKey = upper_bound(tenantStartKey, &openEnd);
    TopicFilterIterator expansionSetItr(topicTrie);  // init(): seek(emptyList)
    std::map<Levels, std::vector<int>> matchedTopicFilters;
    int64_t itr = kv.lower_bound(tenantStartKey);  // itr.seek(tenantBoundary.getStartKey())
    out.stats.seeks++;
    int probe = 0;
    const int64_t n = kv.n();
    while (itr < n && (openEnd || kv.key(itr) < tenantEndKey)) {
        Matching matching = build_match_route(kv.key(itr), kv.value(itr));
        auto seen = matchedTopicFilters.find(matching.filterLevels);
        if (seen == matchedTopicFilters.end()) {
            const Levels& seekTopicFilter = matching.filterLevels;
            expansionSetItr.seek(seekTopicFilter);
            if (expansionSetItr.is_valid()) {
                Levels topicFilterToMatch = expansionSetItr.key();
                if (topicFilterToMatch == seekTopicFilter) {
                    std::vector<int> backingTopics = expansionSetItr.value();
                    for (int t : backingTopics) matchedRoutes[t].add(itr, matching);
                    matchedTopicFilters[seekTopicFilter] = backingTopics;
                    itr++;
                    out.stats.nexts++;
                    probe = 0;
                } else {
                    // next() is much cheaper than seek(): probe the following 20 entries first (:127-136)
                    if (probe++ < 20) {
                        itr++;
                        out.stats.nexts++;
                    } else {
                        int64_t target = kv.lower_bound(tenant_route_start_key(tenantId, topicFilterToMatch));
                        out.stats.seeks++;
                        // Guard (oracle only): keys of a filter with an EMPTY level after prefix F sort inside
                        // F's bucket range, so this seek can land at or before the cursor and the reference
                        // loop would never terminate. Force progress and count it (tests assert on it).
                        if (target <= itr) {
                            out.stats.backwardSeeks++;
                            target = itr + 1;
                        }
                        itr = target;
                    }
                }
            } else {
                break;  // no more topic filter to match
            }
        } else {
            int64_t rank = itr;
            itr++;
            out.stats.nexts++;
            for (int t : seen->second) matchedRoutes[t].add(rank, matching);
        }
    }
    finish(matchedRoutes, out);
    return out;
}

// ---------------------------------------------------------------- (2) brute force
// SURVEY.md §8a predicate (derived from DCP/TopicTrieNode.java:147-152 and the N/S/M node rules).
bool topic_matches_filter(const Levels& T, const Levels& F) {
    const size_t n = T.size(), m = F.size();
    const bool sys = !T.empty() && !T[0].empty() && T[0][0] == '$';
    for (size_t i = 0; i < m; i++) {
        if (F[i] == "#") {
            if (i == 0 && sys) return false;
            return true;  // i <= n always holds here: '#' also matches the parent level (i == n)
        }
        if (i >= n) return false;
        if (F[i] == "+") {
            if (i == 0 && sys) return false;
        } else if (F[i] != T[i]) {
            return false;
        }
    }
    return m == n;
}

MatchResult match_all_brute(const SortedKV& kv, const std::string& tenantId, const std::vector<std::string>& topics,
                            int maxPersistentFanout, int maxGroupFanout) {
    MatchResult out;
    auto matchedRoutes = new_results(topics.size(), maxPersistentFanout, maxGroupFanout, &out.events);
    std::vector<Levels> topicLevels;
    for (const auto& t : topics) topicLevels.push_back(parse(t, false));
    std::string start = tenant_begin_key(tenantId);
    bool openEnd = false;
    std::string end = upper_bound(start, &openEnd);
    for (int64_t r = kv.lower_bound(start); r < kv.n() && (openEnd || kv.key(r) < end); r++) {
        Matching m = build_match_route(kv.key(r), kv.value(r));
        for (size_t t = 0; t < topics.size(); t++)
            if (topic_matches_filter(topicLevels[t], m.filterLevels)) matchedRoutes[t].add(r, m);
    }
    finish(matchedRoutes, out);
    return out;
}

// ---------------------------------------------------------------- (3) filter-trie walk
class FilterTrie {
public:
    struct Node {
        std::unordered_map<std::string, int> children;  // exact children
        int plus = -1, hash = -1;
        std::vector<int64_t> routes;  // KV ranks, ascending
    };
    std::vector<Node> nodes;
    std::unordered_map<std::string, int> tenantRoot;
    std::vector<uint8_t> kind;  // per rank: 0 normal, 1 normal persistent (subBrokerId==1), 2 group
    std::vector<std::string> groupFilter;  // unused placeholder for future use
    int child(int n, const std::string& name, bool create) {
        if (name == "+") {
            if (nodes[n].plus < 0 && create) { nodes.emplace_back(); nodes[n].plus = (int) nodes.size() - 1; }
            return nodes[n].plus;
        }
        if (name == "#") {
            if (nodes[n].hash < 0 && create) { nodes.emplace_back(); nodes[n].hash = (int) nodes.size() - 1; }
            return nodes[n].hash;
        }
        auto it = nodes[n].children.find(name);
        if (it != nodes[n].children.end()) return it->second;
        if (!create) return -1;
        nodes.emplace_back();
        int id = (int) nodes.size() - 1;
        nodes[n].children[name] = id;
        return id;
    }
};

std::shared_ptr<FilterTrie> build_filter_trie(const SortedKV& kv) {
    auto t = std::make_shared<FilterTrie>();
    t->kind.resize(kv.n());
    for (int64_t r = 0; r < kv.n(); r++) {
        RouteDetail d = decode_route_key(kv.key(r));
        auto it = t->tenantRoot.find(d.tenantId);
        int node;
        if (it == t->tenantRoot.end()) {
            t->nodes.emplace_back();
            node = (int) t->nodes.size() - 1;
            t->tenantRoot[d.tenantId] = node;
        } else {
            node = it->second;
        }
        for (const auto& l : d.matcher.filterLevels) node = t->child(node, l, true);
        t->nodes[node].routes.push_back(r);
        if (d.matcher.type == RouteMatcher::Normal)
            t->kind[r] = parse_receiver(d.receiverUrl).subBrokerId == 1 ? 1 : 0;
        else
            t->kind[r] = 2;
    }
    return t;
}

MatchResult match_all_trie(const FilterTrie& trie, const SortedKV& kv, const std::string& tenantId,
                           const std::vector<std::string>& topics, int maxP, int maxG) {
    (void) kv;
    MatchResult out;
    const size_t nt = topics.size();
    out.routes.resize(nt);
    out.persistentFanout.assign(nt, 0);
    out.groupFanout.assign(nt, 0);
    auto rootIt = trie.tenantRoot.find(tenantId);
    if (rootIt == trie.tenantRoot.end()) return out;
    std::vector<int> frontier, next;
    std::vector<int64_t> matched;
    for (size_t ti = 0; ti < nt; ti++) {
        Levels T = parse(topics[ti], false);
        const bool sys = !T[0].empty() && T[0][0] == '$';
        frontier.assign(1, rootIt->second);
        matched.clear();
        out.stats.V++;  // tenant root
        auto emit = [&](int node) {
            const auto& r = trie.nodes[node].routes;
            if (!r.empty()) {
                out.stats.ranges++;
                matched.insert(matched.end(), r.begin(), r.end());
            }
        };
        for (size_t i = 0; i <= T.size(); i++) {
            next.clear();
            for (int n : frontier) {
                const auto& nd = trie.nodes[n];
                const bool wild_ok = !(i == 0 && sys);
                if (nd.hash >= 0 && wild_ok) {  // '#' matches the rest, including nothing (parent match)
                    out.stats.V++;
                    emit(nd.hash);
                }
                if (i == T.size()) {
                    emit(n);
                    continue;
                }
                if (nd.plus >= 0 && wild_ok) {
                    out.stats.V++;
                    next.push_back(nd.plus);
                }
                if (!nd.children.empty()) {
                    out.stats.P += (uint64_t) std::ceil(std::log2((double) nd.children.size())) + 1;
                    auto c = nd.children.find(T[i]);
                    if (c != nd.children.end()) {
                        out.stats.V++;
                        next.push_back(c->second);
                    }
                }
            }
            frontier.swap(next);
            if (frontier.empty()) break;
        }
        std::sort(matched.begin(), matched.end());
        out.stats.R += matched.size();
        // caps in KV order (MatchedRoutes.java:87-141)
        int p = 0, g = 0;
        for (int64_t r : matched) {
            uint8_t k = trie.kind[r];
            if (k == 1) {
                if (p < maxP) { p++; out.routes[ti].push_back(r); }
                else out.events.push_back({ThrottleEvent::Persistent, (int) ti, r, maxP});
            } else if (k == 2) {
                if (g < maxG) { g++; out.routes[ti].push_back(r); }
                else out.events.push_back({ThrottleEvent::Group, (int) ti, r, maxG});
            } else {
                out.routes[ti].push_back(r);
            }
        }
        out.persistentFanout[ti] = p;
        out.groupFanout[ti] = g;
    }
    return out;
}

}  // namespace orc
