// oracle/expansion.cc — CPU ORACLE (test infrastructure only; see oracle.h).
// Restates the per-batch topic trie and the lazily enumerated "expansion set" cursor:
//   DCP/TopicTrieNode.java:135-161, DCP/TopicFilterIterator.java:62-122,224-300,
//   DCP/NTopicFilterTrieNode.java:118-153, DCP/STopicFilterTrieNode.java:117-148,
//   DCP/MTopicFilterTrieNode.java:105-135.
#include <algorithm>
#include <stdexcept>

#include "oracle.h"

namespace orc {

static const std::string NUL(1, '\0');

// ---------------------------------------------------------------- TopicTrie
TopicTrie::TopicTrie(bool global) : root(new TopicTrieNode()), isGlobal(global) {
    root->levelName = NUL;  // TopicTrieNode() -> this(NUL, false)   (TopicTrieNode.java:48-50)
    root->wildcardMatchable = false;
}

void TopicTrie::add_topic(const Levels& topicLevels, int value) {
    if (topicLevels.empty()) return;
    TopicTrieNode* node = root.get();
    for (size_t level = 0; level < topicLevels.size(); level++) {
        const std::string& levelName = topicLevels[level];
        // system topics ('$...') at the first user level are not reachable through wildcards (:147-152)
        bool startsWithSys = !levelName.empty() && levelName[0] == '$';
        bool wildcardMatchable = isGlobal ? (level > 1 || (level == 1 && !startsWithSys))
                                          : (level > 0 || !startsWithSys);
        auto it = node->children.find(levelName);
        if (it == node->children.end()) {
            auto child = std::make_unique<TopicTrieNode>();
            child->levelName = levelName;
            child->wildcardMatchable = wildcardMatchable;
            it = node->children.emplace(levelName, std::move(child)).first;
        }
        node = it->second.get();
        if (level == topicLevels.size() - 1) {
            node->topic = topicLevels;
            if (std::find(node->values.begin(), node->values.end(), value) == node->values.end())
                node->values.push_back(value);
        }
    }
}

// ---------------------------------------------------------------- virtual filter nodes
struct TopicFilterIterator::FNode {
    enum Kind { N, S, M } kind;
    FNode* parent = nullptr;
    std::string levelName;
    std::set<std::string, JavaLess> subLevelNames;
    std::map<std::string, std::vector<const TopicTrieNode*>, JavaLess> subTopicTrieNodes;
    std::vector<const TopicTrieNode*> subWildcardMatchable;
    std::vector<const TopicTrieNode*> backingTopics;
    bool hasCur = false;
    std::string cur;  // subLevelName (null <=> !hasCur)

    // N and S nodes share the same child bookkeeping (NTopicFilterTrieNode.init :118-153,
    // STopicFilterTrieNode.init :117-148)
    void init_ns(const std::vector<const TopicTrieNode*>& siblings) {
        for (const TopicTrieNode* sibling : siblings) {
            if (sibling->is_user_topic()) backingTopics.push_back(sibling);
            for (const auto& e : sibling->children) {
                const TopicTrieNode* sub = e.second.get();
                if (sub->wildcardMatchable) subWildcardMatchable.push_back(sub);
                subTopicTrieNodes[sub->levelName].push_back(sub);
                subLevelNames.insert(sub->levelName);
            }
        }
        if (!backingTopics.empty()) subLevelNames.insert("#");  // '#' matches the parent level
        if (!subWildcardMatchable.empty()) {
            subLevelNames.insert("#");
            subLevelNames.insert("+");
        }
        seek_child("");
    }
    static void collect(const TopicTrieNode* n, std::vector<const TopicTrieNode*>& out) {
        if (n->is_user_topic()) out.push_back(n);
        for (const auto& e : n->children) collect(e.second.get(), out);
    }
    // MTopicFilterTrieNode.init :105-116
    void init_m(const std::vector<const TopicTrieNode*>& siblings) {
        if (parent) backingTopics = parent->backingTopics;
        for (const TopicTrieNode* s : siblings) collect(s, backingTopics);
        std::sort(backingTopics.begin(), backingTopics.end());
        backingTopics.erase(std::unique(backingTopics.begin(), backingTopics.end()), backingTopics.end());
    }
    void seek_child(const std::string& name) {
        if (kind == M) return;
        if (!subLevelNames.empty()) {
            auto it = subLevelNames.lower_bound(name);  // ceiling
            hasCur = it != subLevelNames.end();
            if (hasCur) cur = *it;
        }
    }
    bool at_valid_child() const { return kind != M && hasCur; }
    void next_child() {
        if (kind == M || !hasCur) return;
        auto it = subLevelNames.upper_bound(cur);  // higher
        hasCur = it != subLevelNames.end();
        if (hasCur) cur = *it;
    }
    FNode* child_node() {
        if (!at_valid_child()) throw std::out_of_range("NoSuchElementException");
        FNode* c = new FNode();
        c->parent = this;
        if (cur == "#") {
            c->kind = M;
            c->levelName = "#";
            c->init_m(subWildcardMatchable);
        } else if (cur == "+") {
            c->kind = S;
            c->levelName = "+";
            c->init_ns(subWildcardMatchable);
        } else {
            c->kind = N;
            c->levelName = cur;
            c->init_ns(subTopicTrieNodes.at(cur));
        }
        return c;
    }
};

TopicFilterIterator::TopicFilterIterator(const TopicTrie& trie) : trie_(trie) { seek({}); }
TopicFilterIterator::~TopicFilterIterator() { clear(); }

void TopicFilterIterator::pop() {
    delete stack_.back();
    stack_.pop_back();
}
void TopicFilterIterator::clear() {
    while (!stack_.empty()) pop();
}
bool TopicFilterIterator::is_valid() const { return !stack_.empty(); }

// TopicFilterIterator.seek :62-122 — position at the least expansion filter >= filterLevels
void TopicFilterIterator::seek(const Levels& filterLevels) {
    clear();
    {
        FNode* root = new FNode();  // TopicFilterTrieNode.from(root): N node named NUL over {root}
        root->kind = FNode::N;
        root->levelName = NUL;
        root->init_ns({trie_.root.get()});
        stack_.push_back(root);
    }
    int i = -1;
    const int n = (int) filterLevels.size();
    bool out = false;
    while (!out && !stack_.empty() && i < n) {
        const std::string& levelNameToSeek = i == -1 ? NUL : filterLevels[i];
        i++;
        FNode* node = stack_.back();
        int cmp = java_compare(levelNameToSeek, node->levelName);
        if (cmp < 0) {
            break;
        } else if (cmp == 0) {
            if (i == n) break;
            node->seek_child(filterLevels[i]);
            if (node->at_valid_child()) {
                stack_.push_back(node->child_node());
            } else {
                // backtrace: replace the current node with its next sibling
                pop();
                if (stack_.empty()) break;
                while (!stack_.empty()) {
                    FNode* parent = stack_.back();
                    parent->next_child();
                    if (parent->at_valid_child()) {
                        stack_.push_back(parent->child_node());
                        out = true;
                        break;
                    } else {
                        pop();
                    }
                }
            }
        } else {
            // no least next topic filter exists in the expansion set
            clear();
        }
    }
    // descend to the least filter that has backing topics
    while (!stack_.empty()) {
        FNode* node = stack_.back();
        if (node->backingTopics.empty()) {
            stack_.push_back(node->child_node());
        } else {
            break;
        }
    }
}

// TopicFilterIterator.next :260-278
void TopicFilterIterator::next() {
    while (!stack_.empty()) {
        FNode* node = stack_.back();
        if (node->at_valid_child()) {
            FNode* sub = node->child_node();
            stack_.push_back(sub);
            if (!sub->backingTopics.empty()) break;
        } else {
            pop();
            if (!stack_.empty()) stack_.back()->next_child();
        }
    }
}

// TopicFilterIterator.key :280-288 with TopicFilterTrieNode.topicFilterPrefix (skips the NUL root)
Levels TopicFilterIterator::key() const {
    if (stack_.empty()) throw std::out_of_range("NoSuchElementException");
    Levels out;
    for (const FNode* f : stack_)
        if (f->levelName != NUL) out.push_back(f->levelName);
    return out;
}

std::vector<int> TopicFilterIterator::value() const {
    if (stack_.empty()) throw std::out_of_range("NoSuchElementException");
    std::vector<int> out;
    for (const TopicTrieNode* t : stack_.back()->backingTopics)
        out.insert(out.end(), t->values.begin(), t->values.end());
    std::sort(out.begin(), out.end());
    return out;
}

}  // namespace orc
