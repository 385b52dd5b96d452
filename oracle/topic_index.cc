// oracle/topic_index.cc — CPU ORACLE (test infrastructure only; see oracle.h).
// Restates the inverse match (topic filter -> indexed topics):
//   traversal  U/index/TopicLevelTrie.java:190-249 (lookup + Action handling),
//   selectors  DW/TopicIndex.java:40-117 (TopicMatcher, sys level 0), :119-131 (TopicGetter),
//              RS/index/RetainTopicIndex.java:36-124 (RetainMatcher, tenantId is level 0 => sys level 1).
// The lock-free Ctrie machinery of the reference (CAS, TNode contraction) is concurrency
// plumbing and not part of the result; a plain tree is used here.
#include <algorithm>

#include "oracle.h"

namespace orc {

struct TopicLevelIndex::Node {
    struct Branch {
        std::set<int64_t> values;
        std::unique_ptr<Node> child;
    };
    std::map<std::string, Branch> branches;
};

TopicLevelIndex::TopicLevelIndex() : root_(new Node()) {}
TopicLevelIndex::~TopicLevelIndex() = default;

void TopicLevelIndex::add(const Levels& levels, int64_t value) {
    Node* n = root_.get();
    for (size_t i = 0; i < levels.size(); i++) {
        Node::Branch& b = n->branches[levels[i]];
        if (i + 1 == levels.size()) {
            b.values.insert(value);
        } else {
            if (!b.child) b.child.reset(new Node());
            n = b.child.get();
        }
    }
}

namespace {
bool remove_rec(TopicLevelIndex::Node* n, const Levels& levels, size_t i, int64_t value) {
    auto it = n->branches.find(levels[i]);
    if (it == n->branches.end()) return false;
    auto& b = it->second;
    if (i + 1 == levels.size()) {
        b.values.erase(value);
    } else if (b.child) {
        if (remove_rec(b.child.get(), levels, i + 1, value)) b.child.reset();
    }
    if (b.values.empty() && !b.child) n->branches.erase(it);
    return n->branches.empty();
}

enum Action { CONTINUE, MATCH_AND_CONTINUE, MATCH_AND_STOP };

void lookup(const TopicLevelIndex::Node* n, const Levels& F, int currentLevel, int sysLevel,
            std::vector<int64_t>& out, uint64_t* visited) {
    const int m = (int) F.size();
    auto apply = [&](const TopicLevelIndex::Node::Branch& b, Action a) {
        if (visited) (*visited)++;
        if (a == MATCH_AND_CONTINUE || a == MATCH_AND_STOP) out.insert(out.end(), b.values.begin(), b.values.end());
        if (a != MATCH_AND_STOP && b.child) lookup(b.child.get(), F, currentLevel + 1, sysLevel, out, visited);
    };
    auto is_sys = [&](const std::string& name) {
        return currentLevel == sysLevel && !name.empty() && name[0] == '$';
    };
    if (m == 0) {  // RetainTopicIndex.findAll: every branch MATCH_AND_CONTINUE (RetainTopicIndex.java:41-48)
        for (const auto& e : n->branches) apply(e.second, MATCH_AND_CONTINUE);
        return;
    }
    if (currentLevel < m - 1) {
        // not the last filter level
        const bool matchParent = currentLevel + 1 == m - 1 && F[currentLevel + 1] == "#";
        const std::string& lvl = F[currentLevel];
        if (lvl == "+") {
            for (const auto& e : n->branches) {
                if (is_sys(e.first)) continue;  // '+' skips SYS topics at the first user level
                apply(e.second, matchParent ? MATCH_AND_CONTINUE : CONTINUE);
            }
        } else {
            auto it = n->branches.find(lvl);
            if (it != n->branches.end()) apply(it->second, matchParent ? MATCH_AND_CONTINUE : CONTINUE);
        }
    } else if (currentLevel == m - 1) {
        const std::string& lvl = F[currentLevel];
        if (lvl == "+") {
            for (const auto& e : n->branches) {
                if (is_sys(e.first)) continue;
                apply(e.second, MATCH_AND_STOP);
            }
        } else if (lvl == "#") {
            for (const auto& e : n->branches) {
                if (is_sys(e.first)) continue;
                apply(e.second, MATCH_AND_CONTINUE);
            }
        } else {
            auto it = n->branches.find(lvl);
            if (it != n->branches.end()) apply(it->second, MATCH_AND_STOP);
        }
    } else {
        // below a '#': every descendant matches
        for (const auto& e : n->branches) apply(e.second, MATCH_AND_CONTINUE);
    }
}
}  // namespace

void TopicLevelIndex::remove(const Levels& levels, int64_t value) {
    if (!levels.empty()) remove_rec(root_.get(), levels, 0, value);
}

std::vector<int64_t> TopicLevelIndex::match(const Levels& F, int sysLevel, uint64_t* visited) const {
    std::vector<int64_t> out;
    lookup(root_.get(), F, 0, sysLevel, out, visited);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

std::vector<int64_t> TopicLevelIndex::get(const Levels& T) const {
    const Node* n = root_.get();
    for (size_t i = 0; i < T.size() && n; i++) {
        auto it = n->branches.find(T[i]);
        if (it == n->branches.end()) return {};
        if (i + 1 == T.size()) return std::vector<int64_t>(it->second.values.begin(), it->second.values.end());
        n = it->second.child.get();
    }
    return {};
}

std::vector<int64_t> TopicLevelIndex::find_all() const { return match({}, 1, nullptr); }

}  // namespace orc
