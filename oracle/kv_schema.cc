// oracle/kv_schema.cc — CPU ORACLE (test infrastructure only; see oracle.h).
// Restates the dist-worker route key/value codec:
//   DWS/KVSchemaUtil.java:56-130, DWS/KVSchemaConstants.java:24-34,
//   DWS/cache/RouteDetailCache.java:53-117, DWS/cache/ReceiverCache.java:32-36,
//   U/BSUtil.java:29-69, base-kv BoundaryUtil.upperBound (:299-307).
#include <stdexcept>

#include "oracle.h"

namespace orc {

namespace {
const char SCHEMA_VER = 0x00;       // KVSchemaConstants.java:25
const char FLAG_NORMAL = 0x01;      // :27
const char FLAG_UNORDERED = 0x02;   // :28
const char FLAG_ORDERED = 0x03;     // :29
const char SEPARATOR_BYTE = 0x00;   // :30

std::string u16_be(size_t v) {
    std::string s(2, '\0');
    s[0] = (char) ((v >> 8) & 0xFF);
    s[1] = (char) (v & 0xFF);
    return s;
}
int16_t i16_from_be(const std::string& k, size_t pos) {
    return (int16_t) ((((unsigned char) k[pos]) << 8) | (unsigned char) k[pos + 1]);
}
// KVSchemaUtil.toReceiverBytes :122-125
std::string to_receiver_bytes(const std::string& receiver) { return receiver + u16_be(receiver.size()); }
}  // namespace

std::string to_receiver_url(int subBrokerId, const std::string& receiverId, const std::string& delivererKey) {
    std::string s = std::to_string(subBrokerId);
    s.push_back('\0');
    s += receiverId;
    s.push_back('\0');
    s += delivererKey;
    return s;
}

std::string tenant_begin_key(const std::string& tenantId) {
    std::string k(1, SCHEMA_VER);
    k += u16_be(tenantId.size());
    k += tenantId;
    return k;
}

std::string tenant_route_start_key(const std::string& tenantId, const Levels& filterLevels) {
    std::string k = tenant_begin_key(tenantId);
    for (const auto& l : filterLevels) {
        k += l;
        k.push_back(SEPARATOR_BYTE);
    }
    k.push_back(SEPARATOR_BYTE);
    return k;
}

uint8_t bucket(const std::string& receiver) {
    int32_t hash = java_hash(receiver);
    uint32_t u = (uint32_t) hash;
    return (uint8_t) ((u ^ (u >> 16)) & 0xFF);  // hash ^ (hash >>> 16), MAX_RECEIVER_BUCKETS = 0xFF
}

std::string to_normal_route_key(const std::string& tenantId, const RouteMatcher& m, const std::string& receiverUrl) {
    std::string k = tenant_route_start_key(tenantId, m.filterLevels);
    k.push_back((char) bucket(receiverUrl));
    k.push_back(FLAG_NORMAL);
    k += to_receiver_bytes(receiverUrl);
    return k;
}

std::string to_group_route_key(const std::string& tenantId, const RouteMatcher& m) {
    std::string k = tenant_route_start_key(tenantId, m.filterLevels);
    k.push_back((char) bucket(m.group));
    k.push_back(m.type == RouteMatcher::OrderedShare ? FLAG_ORDERED : FLAG_UNORDERED);
    k += to_receiver_bytes(m.group);
    return k;
}

std::string upper_bound(const std::string& key, bool* open_end) {
    int idx = (int) key.size() - 1;
    while (idx >= 0 && (unsigned char) key[idx] == 0xFF) idx--;
    if (idx < 0) {
        if (open_end) *open_end = true;
        return std::string();
    }
    if (open_end) *open_end = false;
    std::string up = key.substr(0, idx + 1);
    up[idx] = (char) ((unsigned char) up[idx] + 1);
    return up;
}

std::string u64_be(uint64_t v) {
    std::string s(8, '\0');
    for (int i = 0; i < 8; i++) s[i] = (char) ((v >> (56 - 8 * i)) & 0xFF);
    return s;
}
uint64_t u64_from_be(const std::string& b) {
    if (b.size() != 8) throw std::runtime_error("incarnation value must be 8 bytes");
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | (unsigned char) b[i];
    return v;
}

Receiver parse_receiver(const std::string& url) {
    // String.split(NUL): parts[0] = subBrokerId (decimal int), parts[1] = receiverId, parts[2] = delivererKey
    Receiver r;
    size_t a = url.find('\0');
    size_t b = url.find('\0', a + 1);
    r.subBrokerId = std::stoi(url.substr(0, a));
    r.receiverId = url.substr(a + 1, b - a - 1);
    r.delivererKey = b == std::string::npos ? "" : url.substr(b + 1);
    return r;
}

std::string parse_tenant_id(const std::string& key) {
    int16_t len = i16_from_be(key, 1);
    return key.substr(3, len);
}
uint8_t parse_flag(const std::string& key) {
    int16_t rlen = i16_from_be(key, key.size() - 2);
    size_t receiverStart = key.size() - 2 - rlen;
    return (uint8_t) key[receiverStart - 1];
}

// RouteDetailCache.get :53-109 —
// <VER><LENGTH_PREFIX_TENANT_ID><ESCAPED_TOPIC_FILTER><SEP><BUCKET_BYTE><FLAG_BYTE><LENGTH_SUFFIX_RECEIVER_BYTES>
RouteDetail decode_route_key(const std::string& k) {
    RouteDetail d;
    int16_t tenantIdLen = i16_from_be(k, 1);
    size_t tenantIdStartIdx = 1 + 2;
    size_t escapedTopicFilterStartIdx = tenantIdStartIdx + tenantIdLen;
    int receiverBytesLen = i16_from_be(k, k.size() - 2);
    size_t receiverBytesStartIdx = k.size() - 2 - receiverBytesLen;
    size_t receiverBytesEndIdx = k.size() - 2;
    size_t flagByteIdx = receiverBytesStartIdx - 1;
    size_t separatorBytesIdx = flagByteIdx - 1 - 2;  // 2 bytes separator
    std::string receiverInfo = k.substr(receiverBytesStartIdx, receiverBytesEndIdx - receiverBytesStartIdx);
    char flag = k[flagByteIdx];
    d.tenantId = k.substr(tenantIdStartIdx, tenantIdLen);
    std::string escapedTopicFilter = k.substr(escapedTopicFilterStartIdx, separatorBytesIdx - escapedTopicFilterStartIdx);
    d.matcher.filterLevels = parse(escapedTopicFilter, true);
    switch (flag) {
        case FLAG_NORMAL:
            d.matcher.type = RouteMatcher::Normal;
            d.matcher.mqttTopicFilter = unescape(escapedTopicFilter);
            d.receiverUrl = receiverInfo;
            break;
        case FLAG_UNORDERED:
            d.matcher.type = RouteMatcher::UnorderedShare;
            d.matcher.group = receiverInfo;
            d.matcher.mqttTopicFilter = "$share/" + receiverInfo + "/" + unescape(escapedTopicFilter);
            break;
        case FLAG_ORDERED:
            d.matcher.type = RouteMatcher::OrderedShare;
            d.matcher.group = receiverInfo;
            d.matcher.mqttTopicFilter = "$oshare/" + receiverInfo + "/" + unescape(escapedTopicFilter);
            break;
        default:
            throw std::runtime_error("Unknown route type: " + std::to_string((int) flag));
    }
    return d;
}

// minimal protobuf codec for RouteGroup { map<string, uint64> members = 1; }
namespace {
void put_varint(std::string& o, uint64_t v) {
    while (v >= 0x80) {
        o.push_back((char) ((v & 0x7F) | 0x80));
        v >>= 7;
    }
    o.push_back((char) v);
}
uint64_t get_varint(const std::string& b, size_t& p) {
    uint64_t v = 0;
    int shift = 0;
    while (p < b.size()) {
        unsigned char c = (unsigned char) b[p++];
        v |= (uint64_t) (c & 0x7F) << shift;
        if (!(c & 0x80)) break;
        shift += 7;
    }
    return v;
}
}  // namespace

std::string encode_route_group(const std::vector<std::pair<std::string, uint64_t>>& members) {
    std::string out;
    for (const auto& kv : members) {
        std::string entry;
        entry.push_back(0x0A);  // field 1 (key), wire type 2
        put_varint(entry, kv.first.size());
        entry += kv.first;
        entry.push_back(0x10);  // field 2 (value), wire type 0
        put_varint(entry, kv.second);
        out.push_back(0x0A);  // field 1 (members entry), wire type 2
        put_varint(out, entry.size());
        out += entry;
    }
    return out;
}

std::map<std::string, uint64_t> decode_route_group(const std::string& b) {
    std::map<std::string, uint64_t> members;
    size_t p = 0;
    while (p < b.size()) {
        uint64_t tag = get_varint(b, p);
        if ((tag >> 3) != 1 || (tag & 7) != 2) throw std::runtime_error("Unable to parse matching record");
        uint64_t len = get_varint(b, p);
        size_t end = p + len;
        std::string key;
        uint64_t val = 0;
        while (p < end) {
            uint64_t t = get_varint(b, p);
            if (t == 0x0A) {
                uint64_t kl = get_varint(b, p);
                key = b.substr(p, kl);
                p += kl;
            } else if (t == 0x10) {
                val = get_varint(b, p);
            } else {
                throw std::runtime_error("Unable to parse matching record");
            }
        }
        members[key] = val;
    }
    return members;
}

Matching build_match_route(const std::string& key, const std::string& value) {
    RouteDetail d = decode_route_key(key);
    Matching m;
    m.tenantId = d.tenantId;
    m.mqttTopicFilter = d.matcher.mqttTopicFilter;
    m.filterLevels = d.matcher.filterLevels;
    if (d.matcher.type == RouteMatcher::Normal) {
        m.type = Matching::Normal;
        m.receiverUrl = d.receiverUrl;
        m.incarnation = u64_from_be(value);
        m.subBrokerId = parse_receiver(d.receiverUrl).subBrokerId;
    } else {
        m.type = Matching::Group;
        m.members = decode_route_group(value);
    }
    return m;
}

}  // namespace orc

// ------------------------------------------------------------------ retain store schema
namespace orc {
// LevelHash.hashToByte (bifromq-retain/bifromq-retain-store-schema/.../schema/LevelHash.java:41-48): FNV-1a over data.charAt(i),
// int arithmetic, low byte
uint8_t level_hash_byte(const std::string& level) {
    const std::u16string u = to_utf16(level);
    uint32_t hash = 0x811c9dc5u;
    for (char16_t c : u) {
        hash ^= (uint32_t) c;
        hash *= 0x01000193u;
    }
    return (uint8_t) (hash & 0xff);
}
// LevelHash.hash :33-39
std::string level_hash(const std::vector<std::string>& levels) {
    std::string out;
    for (const auto& l : levels) out.push_back((char) level_hash_byte(l));
    return out;
}
// KVSchemaUtil.retainMessageKey :44-50: tenantBeginKey ++ toByteString((short) levels) ++ LevelHash.hash(levels) ++ escape(topic)
std::string retain_message_key(const std::string& tenantId, const std::string& topic) {
    const std::vector<std::string> levels = parse(topic, false);
    std::string k = tenant_begin_key(tenantId);
    k.push_back((char) ((levels.size() >> 8) & 0xff));
    k.push_back((char) (levels.size() & 0xff));
    k += level_hash(levels);
    std::string esc = topic;   // TopicUtil.escape :189-192
    for (char& c : esc)
        if (c == '/') c = '\0';
    return k + esc;
}
// KVSchemaUtil.filterPrefix :52-62
std::vector<std::string> retain_filter_prefix(const std::vector<std::string>& filterLevels) {
    int firstWildcard = -1;
    for (size_t i = 0; i < filterLevels.size(); i++)
        if (filterLevels[i] == "+") {
            firstWildcard = (int) i;
            break;
        }
    if (firstWildcard == -1) {
        if (!filterLevels.empty() && filterLevels.back() == "#")
            return std::vector<std::string>(filterLevels.begin(), filterLevels.end() - 1);
        return filterLevels;
    }
    return std::vector<std::string>(filterLevels.begin(), filterLevels.begin() + firstWildcard);
}
// KVSchemaUtil.retainKeyPrefix :64-68
std::string retain_key_prefix(const std::string& tenantId, int levels, const std::vector<std::string>& filterPrefix) {
    std::string k = tenant_begin_key(tenantId);
    k.push_back((char) ((levels >> 8) & 0xff));
    k.push_back((char) (levels & 0xff));
    return k + level_hash(filterPrefix);
}
}  // namespace orc
