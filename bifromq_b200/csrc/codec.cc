// codec.cc — see codec.h. Byte-level restatement of the reference's route key codec and topic validators.
#include "codec.h"

#include <cstdlib>

namespace bfq {

namespace {
inline uint16_t be16(const char* p) { return (uint16_t) (((uint8_t) p[0] << 8) | (uint8_t) p[1]); }
inline void put_be16(std::string& s, size_t v) {
    s.push_back((char) ((v >> 8) & 0xFF));
    s.push_back((char) (v & 0xFF));
}
// number of UTF-16 code units contributed by the UTF-8 byte c (lead bytes only)
inline int utf16_units(uint8_t c) {
    if ((c & 0xC0) == 0x80) return 0;   // continuation
    return c >= 0xF0 ? 2 : 1;           // 4-byte sequence => surrogate pair
}
inline bool has_prefix(sv s, sv p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
}  // namespace

// <VER 0x00><u16 BE tenant len><tenant><level NUL>*<NUL><bucket><flag><receiver><u16 BE receiver len>
bool decode_route_key(sv k, DecodedKey* out) {
    if (k.size() < 1 + 2 + 2 + 2 + 2 || k[0] != 0x00) return false;
    const size_t tenant_len = be16(k.data() + 1);
    const size_t receiver_len = be16(k.data() + k.size() - 2);
    const size_t filter_start = 3 + tenant_len;
    if (k.size() < filter_start + 2 + 2 + receiver_len + 2) return false;
    const size_t receiver_start = k.size() - 2 - receiver_len;
    const size_t flag_idx = receiver_start - 1;
    const size_t bucket_idx = flag_idx - 1;
    if (bucket_idx < filter_start + 2) return false;
    const size_t sep_idx = bucket_idx - 2;   // the NUL closing the last level, followed by the extra NUL
    if (k[sep_idx] != 0 || k[sep_idx + 1] != 0) return false;
    out->tenant = k.substr(3, tenant_len);
    out->escaped_filter = k.substr(filter_start, sep_idx - filter_start);
    out->receiver = k.substr(receiver_start, receiver_len);
    out->flag = (uint8_t) k[flag_idx];
    out->bucket = (uint8_t) k[bucket_idx];
    if (out->flag == FLAG_NORMAL) {
        // receiverUrl = decimal(subBrokerId) NUL receiverId NUL delivererKey; persistent session broker id == 1
        sv r = out->receiver;
        size_t nul = r.find('\0');
        sv id = r.substr(0, nul);
        out->kind = (id.size() == 1 && id[0] == '1') ? KIND_PERSISTENT : KIND_NORMAL;
        const bool one_digit = id.size() == 1 && id[0] >= '0' && id[0] <= '9';   // the usual spelling: decided above
        if (out->kind == KIND_NORMAL && !id.empty() && !one_digit) {
            // tolerate "+1" / "01" spellings Integer.parseInt would accept
            char* end = nullptr;
            std::string tmp(id);
            long v = std::strtol(tmp.c_str(), &end, 10);
            if (end && *end == 0 && v == 1) out->kind = KIND_PERSISTENT;
        }
    } else if (out->flag == FLAG_UNORDERED || out->flag == FLAG_ORDERED) {
        out->kind = KIND_GROUP;
    } else {
        return false;
    }
    return true;
}

int32_t java_string_hash(sv s) {
    uint32_t h = 0;
    const size_t n = s.size();
    for (size_t i = 0; i < n;) {
        uint8_t c = (uint8_t) s[i];
        uint32_t cp;
        int len;
        if (c < 0x80) { cp = c; len = 1; }
        else if (c < 0xE0) { cp = c & 0x1F; len = 2; }
        else if (c < 0xF0) { cp = c & 0x0F; len = 3; }
        else { cp = c & 0x07; len = 4; }
        for (int j = 1; j < len && i + j < n; j++) cp = (cp << 6) | ((uint8_t) s[i + j] & 0x3F);
        i += len;
        if (cp >= 0x10000) {
            cp -= 0x10000;
            h = 31u * h + (0xD800u + (cp >> 10));
            h = 31u * h + (0xDC00u + (cp & 0x3FF));
        } else {
            h = 31u * h + cp;
        }
    }
    return (int32_t) h;
}

uint8_t level_hash_byte(sv s) {
    uint32_t h = 0x811c9dc5u;
    const size_t n = s.size();
    auto step = [&](uint32_t unit) {
        h ^= unit;
        h *= 0x01000193u;
    };
    for (size_t i = 0; i < n;) {   // UTF-8 -> UTF-16 code units, as String.charAt sees them
        uint8_t c = (uint8_t) s[i];
        uint32_t cp;
        int len;
        if (c < 0x80) { cp = c; len = 1; }
        else if (c < 0xE0) { cp = c & 0x1F; len = 2; }
        else if (c < 0xF0) { cp = c & 0x0F; len = 3; }
        else { cp = c & 0x07; len = 4; }
        for (int j = 1; j < len && i + j < n; j++) cp = (cp << 6) | ((uint8_t) s[i + j] & 0x3F);
        i += len;
        if (cp >= 0x10000) {
            cp -= 0x10000;
            step(0xD800u + (cp >> 10));
            step(0xDC00u + (cp & 0x3FF));
        } else {
            step(cp);
        }
    }
    return (uint8_t) (h & 0xFF);
}

std::string make_retain_key(sv tenant, sv topic) {
    std::string k = make_tenant_begin_key(tenant);
    size_t levels = 0;
    std::string hashes;
    for_each_level(topic, '/', [&](sv l) {
        levels++;
        hashes.push_back((char) level_hash_byte(l));
    });
    put_be16(k, levels);
    k.append(hashes);
    for (char c : topic) k.push_back(c == '/' ? '\0' : c);
    return k;
}

std::string make_retain_key_prefix(sv tenant, sv tf) {
    std::vector<sv> lv;
    for_each_level(tf, '/', [&](sv l) { lv.push_back(l); });
    const bool multi = !lv.empty() && lv.back().size() == 1 && lv.back()[0] == '#';
    const size_t levels = multi ? lv.size() - 1 : lv.size();
    size_t cut = lv.size();   // filterPrefix: up to the first '+', else without a final '#'
    for (size_t i = 0; i < lv.size(); i++)
        if (lv[i].size() == 1 && lv[i][0] == '+') {
            cut = i;
            break;
        }
    if (cut == lv.size() && multi) cut = lv.size() - 1;
    std::string k = make_tenant_begin_key(tenant);
    put_be16(k, levels);
    for (size_t i = 0; i < cut; i++) k.push_back((char) level_hash_byte(lv[i]));
    return k;
}

bool decode_retain_key(sv key, sv* tenant, std::string* topic) {
    if (key.size() < 5 || key[0] != 0) return false;
    const size_t tl = ((size_t) (uint8_t) key[1] << 8) | (uint8_t) key[2];
    if (key.size() < 3 + tl + 2) return false;
    *tenant = key.substr(3, tl);
    const size_t levels = ((size_t) (uint8_t) key[3 + tl] << 8) | (uint8_t) key[4 + tl];
    const size_t at = 5 + tl + levels;
    if (levels == 0 || key.size() < at) return false;
    const sv esc = key.substr(at);
    // the escaped topic must have exactly `levels` levels and the stored hash bytes must be theirs
    size_t n = 0;
    bool ok = true;
    for_each_level(esc, '\0', [&](sv l) {
        if (n < levels && (uint8_t) key[5 + tl + n] != level_hash_byte(l)) ok = false;
        n++;
    });
    if (!ok || n != levels) return false;
    topic->assign(esc);
    for (char& c : *topic)
        if (c == '\0') c = '/';
    return true;
}

uint8_t receiver_bucket(sv receiver) {
    uint32_t h = (uint32_t) java_string_hash(receiver);
    return (uint8_t) ((h ^ (h >> 16)) & 0xFF);
}

std::string make_receiver_url(int32_t sub_broker_id, sv receiver_id, sv deliverer_key) {
    std::string s = std::to_string(sub_broker_id);
    s.push_back('\0');
    s.append(receiver_id);
    s.push_back('\0');
    s.append(deliverer_key);
    return s;
}

std::string make_tenant_begin_key(sv tenant) {
    std::string k(1, '\0');
    put_be16(k, tenant.size());
    k.append(tenant);
    return k;
}

std::string make_route_key(sv tenant, sv tf, sv receiver_url) {
    uint8_t flag = FLAG_NORMAL;
    sv receiver = receiver_url;
    sv filter = tf;
    if (has_prefix(tf, "$share/") || has_prefix(tf, "$oshare/")) {
        flag = tf[1] == 's' ? FLAG_UNORDERED : FLAG_ORDERED;
        sv rest = tf.substr(flag == FLAG_UNORDERED ? 7 : 8);
        size_t sep = rest.find('/');
        receiver = rest.substr(0, sep);                       // group name
        filter = sep == sv::npos ? sv() : rest.substr(sep + 1);
    }
    std::string k = make_tenant_begin_key(tenant);
    for (char c : filter) k.push_back(c == '/' ? '\0' : c);   // levels, each closed by NUL ...
    k.push_back('\0');
    k.push_back('\0');                                        // ... plus the extra separator
    k.push_back((char) receiver_bucket(receiver));
    k.push_back((char) flag);
    k.append(receiver);
    put_be16(k, receiver.size());
    return k;
}

std::string prefix_upper_bound(sv key, bool* open_end) {
    size_t n = key.size();
    while (n > 0 && (uint8_t) key[n - 1] == 0xFF) n--;
    *open_end = n == 0;
    std::string up(key.substr(0, n));
    if (n) up[n - 1] = (char) ((uint8_t) up[n - 1] + 1);
    return up;
}

// TopicUtil.isValidTopic :42-72 — lengths are counted in UTF-16 code units like java.lang.String
bool is_valid_topic(sv topic, int max_level_length, int max_level, int max_length) {
    if (topic.empty()) return false;
    if (has_prefix(topic, "$oshare/") || has_prefix(topic, "$share/")) return false;
    int total = 0, level_len = 0, level = 1;
    for (char ch : topic) {
        uint8_t c = (uint8_t) ch;
        total += utf16_units(c);
        if (c == '/') {
            if (++level > max_level) return false;
            if (level_len > max_level_length) return false;
            level_len = 0;
        } else {
            if (c == 0 || c == '+' || c == '#') return false;
            level_len += utf16_units(c);
        }
    }
    if (total > max_length) return false;
    return level_len <= max_level_length;
}

// TopicUtil.isValidTopicFilter :74-163
bool is_valid_topic_filter(sv tf, int max_level_length, int max_level, int max_length) {
    const bool unordered = has_prefix(tf, "$share/"), ordered = has_prefix(tf, "$oshare/");
    if (unordered) max_length += 7;
    if (ordered) max_length += 8;
    if (tf.empty()) return false;
    int total = 0;
    for (char ch : tf) total += utf16_units((uint8_t) ch);
    if (total > max_length) return false;
    size_t i = 0;
    const size_t n = tf.size();
    if (unordered || ordered) {
        int name_len = 0;
        for (i = tf.find('/') + 1; i < n; i++) {
            uint8_t c = (uint8_t) tf[i];
            if (c == '/') break;
            if (c == '#' || c == '+' || c == 0) return false;
            name_len += utf16_units(c);
        }
        if (name_len == 0 || i == n) return false;
        i++;
    }
    const size_t start = i;
    int level = 1, level_len = 0;
    for (; i < n; i++) {
        uint8_t c = (uint8_t) tf[i];
        if (c == '/') {
            if (++level > max_level) return false;
            if (level_len > max_level_length) return false;
            level_len = 0;
            continue;
        }
        if (c == 0) return false;
        const bool first = i == start, last = i == n - 1;
        if (c == '#') {
            if (!last) return false;
            if (!first && tf[i - 1] != '/') return false;
        } else if (c == '+') {
            if (!first && tf[i - 1] != '/') return false;
            if (!last && tf[i + 1] != '/') return false;
        }
        level_len += utf16_units(c);
    }
    if (level > max_level) return false;
    return level_len <= max_level_length;
}

}  // namespace bfq
