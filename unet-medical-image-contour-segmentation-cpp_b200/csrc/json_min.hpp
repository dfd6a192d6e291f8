// json_min.hpp -- minimal JSON reader (config, weight-blob header, size sidecar) and the byte-exact
// writers for the two documents the reference emits with nlohmann::json 3.12.0
// (/root/reference/include/nlohmann/json.hpp -- third-party, not copied here):
//   * the size sidecar   `jf << j << std::endl`            (src/preprocess.cpp:126-134)  compact dump
//   * the LabelMe result `f << std::setw(4) << j << endl`  (src/mask2polygon.cpp:74-108) indent 4
// nlohmann's default object is std::map-backed, so keys come out alphabetically; both writers
// hard-code that order instead of building a DOM.
#pragma once
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ms {
namespace json {

struct Value {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;

    bool has(const std::string& k) const { return type == Obj && obj.count(k); }
    const Value& at(const std::string& k) const {
        auto it = obj.find(k);
        if (type != Obj || it == obj.end()) throw std::runtime_error("json: missing key '" + k + "'");
        return it->second;
    }
    double number(const std::string& k, double dflt) const { return has(k) && at(k).type == Num ? at(k).num : dflt; }
    int64_t integer(const std::string& k, int64_t dflt) const { return (int64_t)number(k, (double)dflt); }
    bool boolean(const std::string& k, bool dflt) const { return has(k) && at(k).type == Bool ? at(k).b : dflt; }
    std::string string(const std::string& k, const std::string& dflt) const {
        return has(k) && at(k).type == Str ? at(k).str : dflt;
    }
};

class Parser {
  public:
    explicit Parser(const std::string& s) : s_(s) {}
    Value parse() {
        Value v = value();
        ws();
        if (i_ != s_.size()) err("trailing characters");
        return v;
    }

  private:
    const std::string& s_;
    size_t i_ = 0;
    [[noreturn]] void err(const char* m) const { throw std::runtime_error(std::string("json: ") + m + " at offset " + std::to_string(i_)); }
    void ws() { while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_; }
    bool eat(char c) { ws(); if (i_ < s_.size() && s_[i_] == c) { ++i_; return true; } return false; }
    Value value() {
        ws();
        if (i_ >= s_.size()) err("unexpected end");
        char c = s_[i_];
        Value v;
        if (c == '{') {
            ++i_;
            v.type = Value::Obj;
            if (eat('}')) return v;
            do {
                ws();
                Value k = value();
                if (k.type != Value::Str) err("object key must be a string");
                if (!eat(':')) err("expected ':'");
                v.obj[k.str] = value();
            } while (eat(','));
            if (!eat('}')) err("expected '}'");
        } else if (c == '[') {
            ++i_;
            v.type = Value::Arr;
            if (eat(']')) return v;
            do v.arr.push_back(value()); while (eat(','));
            if (!eat(']')) err("expected ']'");
        } else if (c == '"') {
            ++i_;
            v.type = Value::Str;
            while (i_ < s_.size() && s_[i_] != '"') {
                char ch = s_[i_++];
                if (ch == '\\') {
                    if (i_ >= s_.size()) err("bad escape");
                    char e = s_[i_++];
                    switch (e) {
                        case 'n': v.str += '\n'; break;
                        case 't': v.str += '\t'; break;
                        case 'r': v.str += '\r'; break;
                        case 'b': v.str += '\b'; break;
                        case 'f': v.str += '\f'; break;
                        case 'u': {
                            if (i_ + 4 > s_.size()) err("bad \\u escape");
                            unsigned cp = (unsigned)std::strtoul(s_.substr(i_, 4).c_str(), nullptr, 16);
                            i_ += 4;
                            if (cp < 0x80) v.str += (char)cp;
                            else if (cp < 0x800) { v.str += (char)(0xC0 | (cp >> 6)); v.str += (char)(0x80 | (cp & 0x3F)); }
                            else { v.str += (char)(0xE0 | (cp >> 12)); v.str += (char)(0x80 | ((cp >> 6) & 0x3F)); v.str += (char)(0x80 | (cp & 0x3F)); }
                            break;
                        }
                        default: v.str += e;
                    }
                } else {
                    v.str += ch;
                }
            }
            if (i_ >= s_.size()) err("unterminated string");
            ++i_;
        } else if (s_.compare(i_, 4, "true") == 0) { v.type = Value::Bool; v.b = true; i_ += 4;
        } else if (s_.compare(i_, 5, "false") == 0) { v.type = Value::Bool; v.b = false; i_ += 5;
        } else if (s_.compare(i_, 4, "null") == 0) { v.type = Value::Null; i_ += 4;
        } else {
            char* end = nullptr;
            v.num = std::strtod(s_.c_str() + i_, &end);
            if (end == s_.c_str() + i_) err("unexpected character");
            v.type = Value::Num;
            i_ = (size_t)(end - s_.c_str());
        }
        return v;
    }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

// nlohmann escapes ", \, control characters; non-ASCII bytes pass through (ensure_ascii = false).
inline void append_escaped(std::string& out, const std::string& s) {
    static const char* hex = "0123456789abcdef";
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) { out += "\\u00"; out += hex[c >> 4]; out += hex[c & 15]; }
                else out += (char)c;
        }
    }
}

// src/preprocess.cpp:126-134
inline std::string sidecar_text(const std::string& filename, int w, int h, int scaled_w, int scaled_h) {
    std::string o = "{\"";
    append_escaped(o, filename);
    o += "\":{\"original_height\":" + std::to_string(h) + ",\"original_width\":" + std::to_string(w) +
         ",\"scaled_height\":" + std::to_string(scaled_h) + ",\"scaled_width\":" + std::to_string(scaled_w) + "}}\n";
    return o;
}

// src/mask2polygon.cpp:74-108.  contour c = xy[2*cstart[c] .. 2*cstart[c+1]).  Byte-exact with nlohmann's dump(4) of the
// reference's document; formatted straight into one buffer (a point is ~88 bytes of fixed text around two integers, so
// the whole cost is two integer conversions and two memcpys per point: ~1 GB/s per core).
inline void append_int(std::string& o, int v) {
    char buf[16];
    char* e = buf + sizeof buf;
    char* p = e;
    unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
    do {
        *--p = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *--p = '-';
    o.append(p, (size_t)(e - p));
}
inline void labelme_append(std::string& o, const int32_t* xy, const int32_t* cstart, int n_contours, const std::string& base_name,
                           int orig_w, int orig_h) {
    static const char kPointOpen[] = "                [\n                    ";
    static const char kPointMid[] = ",\n                    ";
    static const char kPointClose[] = "\n                ]";
    o.reserve(o.size() + 256 + (size_t)n_contours * 256 + (size_t)(n_contours ? cstart[n_contours] - cstart[0] : 0) * 96);
    o += "{\n    \"flags\": {},\n    \"imageData\": null,\n    \"imageHeight\": ";
    append_int(o, orig_h);
    o += ",\n    \"imagePath\": \"";
    append_escaped(o, base_name + ".raw");
    o += "\",\n    \"imageWidth\": ";
    append_int(o, orig_w);
    o += ",\n    \"shapes\": ";
    if (n_contours == 0) {
        o += "[]";
    } else {
        o += "[\n";
        for (int c = 0; c < n_contours; ++c) {
            o += "        {\n            \"description\": \"\",\n            \"flags\": {},\n            \"group_id\": null,\n"
                 "            \"label\": 1,\n            \"labelIndex\": 0,\n            \"mask\": null,\n            \"points\": ";
            const int a = cstart[c], b = cstart[c + 1];
            if (a == b) {
                o += "null";
            } else {
                o += "[\n";
                for (int i = a; i < b; ++i) {
                    o.append(kPointOpen, sizeof kPointOpen - 1);
                    append_int(o, xy[2 * i]);
                    o.append(kPointMid, sizeof kPointMid - 1);
                    append_int(o, xy[2 * i + 1]);
                    o.append(kPointClose, sizeof kPointClose - 1);
                    if (i + 1 < b) o += ",\n";
                    else o += '\n';
                }
                o += "            ]";
            }
            o += ",\n            \"shape_type\": \"polygon\"\n        }";
            o += c + 1 < n_contours ? ",\n" : "\n";
        }
        o += "    ]";
    }
    o += ",\n    \"version\": \"1.0.2.812\"\n}\n";
}
inline std::string labelme_text(const int32_t* xy, const int32_t* cstart, int n_contours, const std::string& base_name,
                                int orig_w, int orig_h) {
    std::string o;
    labelme_append(o, xy, cstart, n_contours, base_name, orig_w, orig_h);
    return o;
}

}  // namespace json
}  // namespace ms
