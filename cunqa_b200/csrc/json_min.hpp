// json_min.hpp -- minimal JSON value / parser / writer for the wire format
// (/root/reference/src/quantum_task.cpp:18-36 parses the same messages with nlohmann::json; the
// engine keeps no third-party dependency, so it carries this ~200-line reader instead).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cb200 {

class Json {
public:
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ, RAW };  // RAW: pre-rendered JSON text (big count tables are written directly)
    Type type = NUL;
    bool b = false;
    double num = 0.0;
    bool is_int = false;
    long long inum = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // insertion order preserved

    Json() = default;
    static Json object() { Json j; j.type = OBJ; return j; }
    static Json array() { Json j; j.type = ARR; return j; }
    static Json number(double v) { Json j; j.type = NUM; j.num = v; return j; }
    static Json integer(long long v) { Json j; j.type = NUM; j.num = (double)v; j.is_int = true; j.inum = v; return j; }
    static Json string(const std::string &s) { Json j; j.type = STR; j.str = s; return j; }
    static Json boolean(bool v) { Json j; j.type = BOOL; j.b = v; return j; }
    static Json raw(std::string text) { Json j; j.type = RAW; j.str = std::move(text); return j; }

    bool is_object() const { return type == OBJ; }
    bool is_array() const { return type == ARR; }
    bool is_number() const { return type == NUM; }
    bool is_string() const { return type == STR; }

    const Json *find(const std::string &k) const
    {
        if (type != OBJ) return nullptr;
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    Json *find(const std::string &k)
    {
        if (type != OBJ) return nullptr;
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    bool contains(const std::string &k) const { return find(k) != nullptr; }
    const Json &at(const std::string &k) const
    {
        const Json *p = find(k);
        if (!p) throw std::runtime_error("key '" + k + "' not found");
        return *p;
    }
    Json &set(const std::string &k, Json v)
    {
        if (type != OBJ) { type = OBJ; }
        for (auto &kv : obj) if (kv.first == k) { kv.second = std::move(v); return kv.second; }
        obj.emplace_back(k, std::move(v));
        return obj.back().second;
    }
    const Json &at(size_t i) const
    {
        if (type != ARR || i >= arr.size()) throw std::runtime_error("array index out of range");
        return arr[i];
    }
    size_t size() const { return type == ARR ? arr.size() : (type == OBJ ? obj.size() : 0); }
    double as_double() const
    {
        if (type == NUM) return num;
        if (type == BOOL) return b ? 1.0 : 0.0;
        throw std::runtime_error("number expected");
    }
    long long as_int() const
    {
        if (type == NUM) return is_int ? inum : (long long)std::llround(num);
        if (type == BOOL) return b ? 1 : 0;
        throw std::runtime_error("integer expected");
    }
    bool as_bool() const
    {
        if (type == BOOL) return b;
        if (type == NUM) return num != 0.0;
        throw std::runtime_error("boolean expected");
    }
    const std::string &as_string() const
    {
        if (type != STR) throw std::runtime_error("string expected");
        return str;
    }

    static Json parse(const std::string &text)
    {
        size_t pos = 0;
        Json v = parse_value(text, pos);
        skip_ws(text, pos);
        if (pos != text.size()) throw std::runtime_error("trailing characters in JSON");
        return v;
    }

    std::string dump() const
    {
        std::string out;
        dump_to(out);
        return out;
    }

private:
    static void skip_ws(const std::string &s, size_t &p)
    {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++;
    }
    static Json parse_value(const std::string &s, size_t &p)
    {
        skip_ws(s, p);
        if (p >= s.size()) throw std::runtime_error("unexpected end of JSON");
        const char c = s[p];
        if (c == '{') {
            Json j = object();
            p++;
            skip_ws(s, p);
            if (p < s.size() && s[p] == '}') { p++; return j; }
            while (true) {
                skip_ws(s, p);
                if (p >= s.size() || s[p] != '"') throw std::runtime_error("object key expected");
                std::string k = parse_string(s, p);
                skip_ws(s, p);
                if (p >= s.size() || s[p] != ':') throw std::runtime_error("':' expected");
                p++;
                j.obj.emplace_back(std::move(k), parse_value(s, p));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                throw std::runtime_error("',' or '}' expected");
            }
            return j;
        }
        if (c == '[') {
            Json j = array();
            p++;
            skip_ws(s, p);
            if (p < s.size() && s[p] == ']') { p++; return j; }
            while (true) {
                j.arr.push_back(parse_value(s, p));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                throw std::runtime_error("',' or ']' expected");
            }
            return j;
        }
        if (c == '"') return string(parse_string(s, p));
        if (s.compare(p, 4, "true") == 0) { p += 4; return boolean(true); }
        if (s.compare(p, 5, "false") == 0) { p += 5; return boolean(false); }
        if (s.compare(p, 4, "null") == 0) { p += 4; return Json(); }
        // number
        const char *start = s.c_str() + p;
        char *end = nullptr;
        const double v = std::strtod(start, &end);
        if (end == start) throw std::runtime_error(std::string("unexpected character '") + c + "' in JSON");
        bool integral = true;
        for (const char *q = start; q < end; q++)
            if (*q == '.' || *q == 'e' || *q == 'E') integral = false;
        p += (size_t)(end - start);
        if (integral) {
            Json j = integer(std::strtoll(start, nullptr, 10));
            j.num = v;
            return j;
        }
        return number(v);
    }
    static std::string parse_string(const std::string &s, size_t &p)
    {
        std::string out;
        p++;  // opening quote
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                p++;
                switch (s[p]) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {
                    if (p + 4 < s.size()) {
                        const unsigned cp = (unsigned)std::strtoul(s.substr(p + 1, 4).c_str(), nullptr, 16);
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        p += 4;
                    }
                    break;
                }
                default: out += s[p];
                }
                p++;
            } else {
                out += s[p++];
            }
        }
        if (p >= s.size()) throw std::runtime_error("unterminated string in JSON");
        p++;  // closing quote
        return out;
    }
    static void dump_string(const std::string &s, std::string &out)
    {
        out += '"';
        for (char c : s) {
            switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\t': out += "\\t"; break;
            case '\r': out += "\\r"; break;
            default:
                if ((unsigned char)c < 0x20) { char buf[8]; std::snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
                else out += c;
            }
        }
        out += '"';
    }
    void dump_to(std::string &out) const
    {
        switch (type) {
        case NUL: out += "null"; break;
        case BOOL: out += b ? "true" : "false"; break;
        case NUM: {
            char buf[40];
            if (is_int) std::snprintf(buf, sizeof buf, "%lld", inum);
            else if (std::isfinite(num)) std::snprintf(buf, sizeof buf, "%.17g", num);
            else std::snprintf(buf, sizeof buf, "null");
            out += buf;
            break;
        }
        case STR: dump_string(str, out); break;
        case RAW: out += str; break;
        case ARR:
            out += '[';
            for (size_t i = 0; i < arr.size(); i++) { if (i) out += ','; arr[i].dump_to(out); }
            out += ']';
            break;
        case OBJ:
            out += '{';
            for (size_t i = 0; i < obj.size(); i++) {
                if (i) out += ',';
                dump_string(obj[i].first, out);
                out += ':';
                obj[i].second.dump_to(out);
            }
            out += '}';
            break;
        }
    }
};

}  // namespace cb200
