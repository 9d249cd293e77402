// TEST INFRASTRUCTURE (oracle/): minimal JSON value + parser/printer used only by the
// oracle's text interface. Not part of the product path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pjson {

struct Value;
using Object = std::vector<std::pair<std::string, Value>>;  // insertion-ordered, like a JS object
using Array = std::vector<Value>;

struct Value {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::shared_ptr<Array> arr;
    std::shared_ptr<Object> obj;

    Value() {}
    static Value boolean(bool v) { Value x; x.kind = Bool; x.b = v; return x; }
    static Value number(double v) { Value x; x.kind = Num; x.num = v; return x; }
    static Value string(const std::string& s) { Value x; x.kind = Str; x.str = s; return x; }
    static Value array() { Value x; x.kind = Arr; x.arr = std::make_shared<Array>(); return x; }
    static Value object() { Value x; x.kind = Obj; x.obj = std::make_shared<Object>(); return x; }

    bool isNull() const { return kind == Null; }
    bool isStr() const { return kind == Str; }
    bool isObj() const { return kind == Obj; }
    bool isArr() const { return kind == Arr; }
    bool isNum() const { return kind == Num; }

    const Value* get(const std::string& key) const {
        if (kind != Obj) return nullptr;
        for (auto& kv : *obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    void set(const std::string& key, const Value& v) {
        for (auto& kv : *obj) if (kv.first == key) { kv.second = v; return; }
        obj->push_back({key, v});
    }
    void push(const Value& v) { arr->push_back(v); }
    size_t size() const { return kind == Arr ? arr->size() : kind == Obj ? obj->size() : 0; }
    const Value& at(size_t i) const { return (*arr)[i]; }
};

inline void escape_to(const std::string& s, std::string& out) {
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
                else out.push_back((char)c);
        }
    }
    out.push_back('"');
}

inline void dump_to(const Value& v, std::string& out) {
    switch (v.kind) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Num: {
            char buf[40];
            if (v.num == (double)(long long)v.num) snprintf(buf, sizeof buf, "%lld", (long long)v.num);
            else snprintf(buf, sizeof buf, "%.17g", v.num);
            out += buf; break;
        }
        case Value::Str: escape_to(v.str, out); break;
        case Value::Arr: {
            out.push_back('[');
            bool first = true;
            for (auto& e : *v.arr) { if (!first) out.push_back(','); first = false; dump_to(e, out); }
            out.push_back(']'); break;
        }
        case Value::Obj: {
            out.push_back('{');
            bool first = true;
            for (auto& kv : *v.obj) {
                if (!first) out.push_back(','); first = false;
                escape_to(kv.first, out); out.push_back(':'); dump_to(kv.second, out);
            }
            out.push_back('}'); break;
        }
    }
}
inline std::string dump(const Value& v) { std::string s; dump_to(v, s); return s; }

// Canonical dump: object keys sorted (deep-equality friendly string form).
inline void canon_to(const Value& v, std::string& out) {
    if (v.kind == Value::Obj) {
        std::map<std::string, const Value*> m;
        for (auto& kv : *v.obj) m[kv.first] = &kv.second;
        out.push_back('{');
        bool first = true;
        for (auto& kv : m) {
            if (!first) out.push_back(','); first = false;
            escape_to(kv.first, out); out.push_back(':'); canon_to(*kv.second, out);
        }
        out.push_back('}');
    } else if (v.kind == Value::Arr) {
        out.push_back('[');
        bool first = true;
        for (auto& e : *v.arr) { if (!first) out.push_back(','); first = false; canon_to(e, out); }
        out.push_back(']');
    } else dump_to(v, out);
}
inline std::string canon(const Value& v) { std::string s; canon_to(v, s); return s; }

struct Parser {
    const char* p; const char* end;
    explicit Parser(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
    [[noreturn]] void fail(const char* msg) { throw std::runtime_error(std::string("JSON parse error: ") + msg); }
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    static void put_utf8(unsigned cp, std::string& out) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    unsigned hex4() {
        if (end - p < 4) fail("bad \\u");
        unsigned v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p++; v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad hex");
        }
        return v;
    }
    std::string str() {
        if (*p != '"') fail("expected string");
        ++p; std::string out;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                ++p; if (p >= end) fail("bad escape");
                char c = *p++;
                switch (c) {
                    case 'n': out.push_back('\n'); break; case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break; case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break; case '/': out.push_back('/'); break;
                    case '\\': out.push_back('\\'); break; case '"': out.push_back('"'); break;
                    case 'u': {
                        unsigned cp = hex4();
                        if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2; unsigned lo = hex4();
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        put_utf8(cp, out); break;
                    }
                    default: fail("bad escape char");
                }
            } else out.push_back(*p++);
        }
        if (p >= end) fail("unterminated string");
        ++p; return out;
    }
    Value value() {
        ws(); if (p >= end) fail("unexpected end");
        char c = *p;
        if (c == '{') {
            ++p; Value v = Value::object(); ws();
            if (*p == '}') { ++p; return v; }
            for (;;) {
                ws(); std::string k = str(); ws();
                if (*p != ':') fail("expected :"); ++p;
                Value e = value(); v.obj->push_back({k, e}); ws();
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; return v; }
                fail("expected , or }");
            }
        }
        if (c == '[') {
            ++p; Value v = Value::array(); ws();
            if (*p == ']') { ++p; return v; }
            for (;;) {
                v.arr->push_back(value()); ws();
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; return v; }
                fail("expected , or ]");
            }
        }
        if (c == '"') return Value::string(str());
        if (c == 't' && end - p >= 4 && std::string(p, 4) == "true") { p += 4; return Value::boolean(true); }
        if (c == 'f' && end - p >= 5 && std::string(p, 5) == "false") { p += 5; return Value::boolean(false); }
        if (c == 'n' && end - p >= 4 && std::string(p, 4) == "null") { p += 4; return Value(); }
        char* e2 = nullptr;
        double d = strtod(p, &e2);
        if (e2 == p) fail("bad value");
        p = e2; return Value::number(d);
    }
};
inline Value parse(const std::string& s) { Parser ps(s); Value v = ps.value(); ps.ws(); return v; }

}  // namespace pjson
