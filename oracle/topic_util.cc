// oracle/topic_util.cc — CPU ORACLE (test infrastructure only; see oracle.h).
// Restates U/TopicUtil.java, U/TopicConst.java and the JDK String behaviours the path relies on.
#include "oracle.h"

namespace orc {

namespace {
const char16_t NUL_CHAR = 0x0000;        // U/TopicConst.java:27
const char16_t DELIMITER_CHAR = u'/';    // :28
const char16_t SINGLE_WILDCARD_CHAR = u'+';  // :29
const char16_t MULTIPLE_WILDCARD_CHAR = u'#';  // :30
const std::string PREFIX_UNORDERED_SHARE = "$share/";   // TopicUtil.java:39
const std::string PREFIX_ORDERED_SHARE = "$oshare/";    // :40

bool starts_with(const std::string& s, const std::string& p) {
    return s.size() >= p.size() && s.compare(0, p.size(), p) == 0;
}
bool starts_with16(const std::u16string& s, const std::string& ascii) {
    if (s.size() < ascii.size()) return false;
    for (size_t i = 0; i < ascii.size(); i++)
        if (s[i] != (char16_t) (unsigned char) ascii[i]) return false;
    return true;
}
}  // namespace

std::u16string to_utf16(const std::string& s) {
    std::u16string out;
    out.reserve(s.size());
    size_t i = 0, n = s.size();
    while (i < n) {
        unsigned char c = (unsigned char) s[i];
        uint32_t cp;
        int extra;
        if (c < 0x80) { cp = c; extra = 0; }
        else if ((c >> 5) == 0x6) { cp = c & 0x1F; extra = 1; }
        else if ((c >> 4) == 0xE) { cp = c & 0x0F; extra = 2; }
        else if ((c >> 3) == 0x1E) { cp = c & 0x07; extra = 3; }
        else { cp = 0xFFFD; extra = 0; }
        i++;
        for (int k = 0; k < extra && i < n; k++, i++) cp = (cp << 6) | ((unsigned char) s[i] & 0x3F);
        if (cp >= 0x10000) {
            cp -= 0x10000;
            out.push_back((char16_t) (0xD800 + (cp >> 10)));
            out.push_back((char16_t) (0xDC00 + (cp & 0x3FF)));
        } else {
            out.push_back((char16_t) cp);
        }
    }
    return out;
}

int32_t java_hash(const std::string& utf8) {
    uint32_t h = 0;
    for (char16_t c : to_utf16(utf8)) h = 31u * h + (uint32_t) c;
    return (int32_t) h;
}

int java_compare(const std::string& a, const std::string& b) {
    // fast path: pure ASCII / BMP strings order identically in UTF-8 bytes and UTF-16 units
    // unless a 4-byte sequence (lead >= 0xF0) meets a 3-byte one with lead >= 0xEE.
    bool risky = false;
    for (unsigned char c : a) if (c >= 0xEE) { risky = true; break; }
    if (!risky) for (unsigned char c : b) if (c >= 0xEE) { risky = true; break; }
    if (!risky) {
        int c = a.compare(b);
        return c < 0 ? -1 : (c > 0 ? 1 : 0);
    }
    std::u16string ua = to_utf16(a), ub = to_utf16(b);
    size_t n = std::min(ua.size(), ub.size());
    for (size_t i = 0; i < n; i++)
        if (ua[i] != ub[i]) return ua[i] < ub[i] ? -1 : 1;
    return ua.size() < ub.size() ? -1 : (ua.size() > ub.size() ? 1 : 0);
}

bool JavaLevelsLess::operator()(const Levels& a, const Levels& b) const {
    size_t n = std::min(a.size(), b.size());
    for (size_t i = 0; i < n; i++) {
        int c = java_compare(a[i], b[i]);
        if (c != 0) return c < 0;
    }
    return a.size() < b.size();
}

// U/TopicUtil.java:206-225 — split on '/' (or NUL when escaped) keeping empty levels.
Levels parse(const std::string& topic, bool escaped) {
    Levels out;
    char splitter = escaped ? '\0' : '/';
    std::string cur;
    for (char c : topic) {
        if (c == splitter) {
            out.push_back(cur);
            cur.clear();
        } else {
            cur.push_back(c);
        }
    }
    out.push_back(cur);
    return out;
}

// U/TopicUtil.java:42-72
bool is_valid_topic(const std::string& topic8, int maxLevelLength, int maxLevel, int maxLength) {
    std::u16string topic = to_utf16(topic8);
    if (topic.empty() || (int) topic.size() > maxLength) return false;
    if (starts_with16(topic, PREFIX_ORDERED_SHARE) || starts_with16(topic, PREFIX_UNORDERED_SHARE)) return false;
    int topicLevelLength = 0;
    int level = 1;
    for (size_t i = 0; i < topic.size(); i++) {
        char16_t c = topic[i];
        if (c == DELIMITER_CHAR) {
            if (++level > maxLevel) return false;
            if (topicLevelLength > maxLevelLength) return false;
            topicLevelLength = 0;
        } else {
            if (c == NUL_CHAR || c == SINGLE_WILDCARD_CHAR || c == MULTIPLE_WILDCARD_CHAR) return false;
            topicLevelLength++;
        }
    }
    return topicLevelLength <= maxLevelLength;
}

// U/TopicUtil.java:74-163
bool is_valid_topic_filter(const std::string& tf8, int maxLevelLength, int maxLevel, int maxLength) {
    std::u16string tf = to_utf16(tf8);
    bool unordered = starts_with16(tf, PREFIX_UNORDERED_SHARE);
    bool ordered = starts_with16(tf, PREFIX_ORDERED_SHARE);
    if (unordered) maxLength += (int) PREFIX_UNORDERED_SHARE.size();
    if (ordered) maxLength += (int) PREFIX_ORDERED_SHARE.size();
    if (tf.empty() || (int) tf.size() > maxLength) return false;
    size_t i = 0;
    int topicLevelLength = 0;
    const size_t n = tf.size();
    if (ordered || unordered) {
        // validate share name
        size_t firstDelim = tf.find(DELIMITER_CHAR);
        for (i = firstDelim + 1; i < n; i++) {
            char16_t c = tf[i];
            if (c == DELIMITER_CHAR) break;
            if (c == MULTIPLE_WILDCARD_CHAR || c == SINGLE_WILDCARD_CHAR || c == NUL_CHAR) return false;
            topicLevelLength++;
        }
        if (topicLevelLength == 0) return false;
        if (i == n) return false;
        topicLevelLength = 0;
        i++;  // skip the separator in front of the real topic filter
    }
    size_t startIdx = i;
    int level = 1;
    for (; i < n; i++) {
        char16_t c = tf[i];
        if (c == DELIMITER_CHAR) {
            if (++level > maxLevel) return false;
            if (topicLevelLength > maxLevelLength) return false;
            topicLevelLength = 0;
        } else {
            if (c == NUL_CHAR) return false;
            if (c == MULTIPLE_WILDCARD_CHAR) {
                if (i != n - 1) return false;
                if (i != startIdx && tf[i - 1] != DELIMITER_CHAR) return false;
            }
            if (c == SINGLE_WILDCARD_CHAR) {
                if (i == startIdx) {
                    if (i != n - 1 && tf[i + 1] != DELIMITER_CHAR) return false;
                } else if (i == n - 1) {
                    if (tf[i - 1] != DELIMITER_CHAR) return false;
                } else {
                    if (tf[i - 1] != DELIMITER_CHAR || tf[i + 1] != DELIMITER_CHAR) return false;
                }
            }
            topicLevelLength++;
        }
    }
    if (level > maxLevel) return false;
    return topicLevelLength <= maxLevelLength;
}

bool is_multi_wildcard_topic_filter(const std::string& tf) { return !tf.empty() && tf.back() == '#'; }
bool is_wildcard_topic_filter(const std::string& tf) {
    return tf.find('+') != std::string::npos || is_multi_wildcard_topic_filter(tf);
}
bool is_unordered_shared(const std::string& tf) { return starts_with(tf, PREFIX_UNORDERED_SHARE); }
bool is_ordered_shared(const std::string& tf) { return starts_with(tf, PREFIX_ORDERED_SHARE); }
bool is_shared_subscription(const std::string& tf) { return is_ordered_shared(tf) || is_unordered_shared(tf); }

std::string escape(const std::string& tf) {
    std::string o = tf;
    for (char& c : o) if (c == '/') c = '\0';
    return o;
}
std::string unescape(const std::string& tf) {
    std::string o = tf;
    for (char& c : o) if (c == '\0') c = '/';
    return o;
}
std::string join(const Levels& levels, char sep) {
    std::string o;
    for (size_t i = 0; i < levels.size(); i++) {
        if (i) o.push_back(sep);
        o += levels[i];
    }
    return o;
}

// U/TopicUtil.java:252-272
RouteMatcher route_matcher_from(const std::string& topicFilter) {
    RouteMatcher m;
    m.mqttTopicFilter = topicFilter;
    if (!is_shared_subscription(topicFilter)) {
        m.type = RouteMatcher::Normal;
        m.filterLevels = parse(topicFilter, false);
        return m;
    }
    // note: the reference tests startsWith("$share") (no slash) to pick the prefix
    bool unordered = starts_with(topicFilter, "$share");
    const std::string sharePrefix = unordered ? "$share" : "$oshare";
    std::string rest = topicFilter.substr(sharePrefix.size() + 1);
    size_t sep = rest.find('/');
    m.group = rest.substr(0, sep);
    m.type = unordered ? RouteMatcher::UnorderedShare : RouteMatcher::OrderedShare;
    m.filterLevels = parse(rest.substr(sep + 1), false);
    return m;
}

}  // namespace orc
