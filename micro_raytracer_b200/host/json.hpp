// json.hpp — the small JSON reader/writer of the native host (stands in for serde_json, which
// the reference uses for every description: src/parser.rs:16-186, src/cli.rs:84, src/http.rs:115).
// Objects keep insertion order so `-v` dumps read like the input.  Header-only, no dependencies.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mrt_host {

struct Error : std::runtime_error {  // ≙ the String of Result<_, String>, parser.rs:12-14
    using std::runtime_error::runtime_error;
};

class Json {
public:
    enum Type { Null, Bool, Num, Str, Arr, Obj };
    using Member = std::pair<std::string, Json>;

    Json() = default;
    static Json boolean(bool b) { Json j; j.t_ = Bool; j.b_ = b; return j; }
    static Json number(double v) { Json j; j.t_ = Num; j.n_ = v; return j; }
    static Json string(std::string s) { Json j; j.t_ = Str; j.s_ = std::move(s); return j; }
    static Json array() { Json j; j.t_ = Arr; return j; }
    static Json object() { Json j; j.t_ = Obj; return j; }
    static Json numbers(const std::vector<double>& v) { Json j = array(); for (double x : v) j.push(number(x)); return j; }

    Type type() const { return t_; }
    bool is_null() const { return t_ == Null; }
    bool is_num() const { return t_ == Num; }
    bool is_str() const { return t_ == Str; }
    bool is_arr() const { return t_ == Arr; }
    bool is_obj() const { return t_ == Obj; }
    bool as_bool() const { need(Bool, "a boolean"); return b_; }
    double as_num() const { need(Num, "a number"); return n_; }
    const std::string& as_str() const { need(Str, "a string"); return s_; }
    const std::vector<Json>& items() const { need(Arr, "an array"); return a_; }
    std::vector<Json>& items() { need(Arr, "an array"); return a_; }
    const std::vector<Member>& members() const { need(Obj, "an object"); return o_; }
    size_t size() const { return t_ == Arr ? a_.size() : t_ == Obj ? o_.size() : 0; }

    // object access: find() returns nullptr for a missing key or an explicit null (serde Option<T>)
    const Json* find(const std::string& k) const {
        if (t_ != Obj) return nullptr;
        for (const auto& m : o_) if (m.first == k) return m.second.is_null() ? nullptr : &m.second;
        return nullptr;
    }
    bool has(const std::string& k) const { return find(k) != nullptr; }
    Json& set(const std::string& k, Json v) {
        need(Obj, "an object");
        for (auto& m : o_) if (m.first == k) { m.second = std::move(v); return m.second; }
        o_.emplace_back(k, std::move(v));
        return o_.back().second;
    }
    Json& operator[](const std::string& k) {  // creates a null member
        need(Obj, "an object");
        for (auto& m : o_) if (m.first == k) return m.second;
        o_.emplace_back(k, Json());
        return o_.back().second;
    }
    void push(Json v) { need(Arr, "an array"); a_.push_back(std::move(v)); }

    // ---- parse
    static Json parse(const std::string& text) {
        Parser p{text.data(), text.data() + text.size(), text.data()};
        Json j = p.value(0);
        p.ws();
        if (p.c != p.e) p.fail("trailing characters");
        return j;
    }

    // ---- serialise (indent < 0: compact like serde_json::to_string, else to_string_pretty)
    std::string dump(int indent = -1) const { std::string out; write(out, indent, 0); return out; }

private:
    Type t_ = Null;
    bool b_ = false;
    double n_ = 0.0;
    std::string s_;
    std::vector<Json> a_;
    std::vector<Member> o_;

    void need(Type t, const char* what) const {
        if (t_ != t) throw Error(std::string("invalid type: expected ") + what);
    }

    struct Parser {
        const char* b; const char* e; const char* c;
        [[noreturn]] void fail(const char* msg) const {
            size_t line = 1, col = 1;
            for (const char* p = b; p < c; p++) { if (*p == '\n') { line++; col = 1; } else col++; }
            throw Error(std::string(msg) + " at line " + std::to_string(line) + " column " + std::to_string(col));
        }
        void ws() { while (c < e && (*c == ' ' || *c == '\t' || *c == '\n' || *c == '\r')) c++; }
        bool lit(const char* s) {
            const char* p = c;
            while (*s) { if (p >= e || *p != *s) return false; p++; s++; }
            c = p; return true;
        }
        static void utf8(std::string& o, uint32_t cp) {
            if (cp < 0x80) o += (char)cp;
            else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
            else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
            else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        }
        uint32_t hex4() {
            if (e - c < 4) fail("EOF while parsing a string");
            uint32_t v = 0;
            for (int i = 0; i < 4; i++) {
                char ch = *c++;
                v <<= 4;
                if (ch >= '0' && ch <= '9') v |= (uint32_t)(ch - '0');
                else if (ch >= 'a' && ch <= 'f') v |= (uint32_t)(ch - 'a' + 10);
                else if (ch >= 'A' && ch <= 'F') v |= (uint32_t)(ch - 'A' + 10);
                else fail("invalid escape");
            }
            return v;
        }
        std::string str() {
            std::string o;
            c++;  // opening quote
            for (;;) {
                if (c >= e) fail("EOF while parsing a string");
                char ch = *c++;
                if (ch == '"') return o;
                if ((unsigned char)ch < 0x20) fail("control character (\\u0000-\\u001F) found while parsing a string");
                if (ch != '\\') { o += ch; continue; }
                if (c >= e) fail("EOF while parsing a string");
                ch = *c++;
                switch (ch) {
                    case '"': o += '"'; break;   case '\\': o += '\\'; break; case '/': o += '/'; break;
                    case 'b': o += '\b'; break;  case 'f': o += '\f'; break;  case 'n': o += '\n'; break;
                    case 'r': o += '\r'; break;  case 't': o += '\t'; break;
                    case 'u': {
                        uint32_t cp = hex4();
                        if (cp >= 0xD800 && cp < 0xDC00 && e - c >= 6 && c[0] == '\\' && c[1] == 'u') {
                            c += 2;
                            uint32_t lo = hex4();
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        utf8(o, cp);
                        break;
                    }
                    default: fail("invalid escape");
                }
            }
        }
        Json value(int depth) {
            if (depth > 128) fail("recursion limit exceeded");
            ws();
            if (c >= e) fail("EOF while parsing a value");
            switch (*c) {
                case '{': {
                    c++;
                    Json j = Json::object();
                    ws();
                    if (c < e && *c == '}') { c++; return j; }
                    for (;;) {
                        ws();
                        if (c >= e || *c != '"') fail("key must be a string");
                        std::string k = str();
                        ws();
                        if (c >= e || *c != ':') fail("expected `:`");
                        c++;
                        j.set(k, value(depth + 1));
                        ws();
                        if (c < e && *c == ',') { c++; continue; }
                        if (c < e && *c == '}') { c++; return j; }
                        fail("expected `,` or `}`");
                    }
                }
                case '[': {
                    c++;
                    Json j = Json::array();
                    ws();
                    if (c < e && *c == ']') { c++; return j; }
                    for (;;) {
                        j.push(value(depth + 1));
                        ws();
                        if (c < e && *c == ',') { c++; continue; }
                        if (c < e && *c == ']') { c++; return j; }
                        fail("expected `,` or `]`");
                    }
                }
                case '"': return Json::string(str());
                case 't': if (lit("true")) return Json::boolean(true); fail("expected ident");
                case 'f': if (lit("false")) return Json::boolean(false); fail("expected ident");
                case 'n': if (lit("null")) return Json(); fail("expected ident");
                default: {
                    const char* s = c;
                    if (c < e && *c == '-') c++;
                    if (c >= e || *c < '0' || *c > '9') fail("expected value");
                    while (c < e && ((*c >= '0' && *c <= '9') || *c == '.' || *c == 'e' || *c == 'E' || *c == '+' || *c == '-')) c++;
                    std::string num(s, c);
                    char* endp = nullptr;
                    double v = std::strtod(num.c_str(), &endp);
                    if (endp != num.c_str() + num.size()) fail("invalid number");
                    return Json::number(v);
                }
            }
        }
    };

    static void write_str(std::string& o, const std::string& s) {
        o += '"';
        for (char ch : s) {
            switch (ch) {
                case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break;
                case '\r': o += "\\r"; break; case '\t': o += "\\t"; break; case '\b': o += "\\b"; break; case '\f': o += "\\f"; break;
                default:
                    if ((unsigned char)ch < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
                    else o += ch;
            }
        }
        o += '"';
    }
    static void write_num(std::string& o, double v) {
        if (!std::isfinite(v)) { o += "null"; return; }  // as serde_json
        if (v == std::floor(v) && std::fabs(v) < 1e15) {
            char b[32]; std::snprintf(b, sizeof b, "%.1f", v); o += b; return;  // 1.0, -0.0: floats keep their point
        }
        char b[40];
        for (int prec = 1; prec <= 17; prec++) {  // shortest representation that round-trips
            std::snprintf(b, sizeof b, "%.*g", prec, v);
            if (std::strtod(b, nullptr) == v) break;
        }
        o += b;
    }
    void write(std::string& o, int indent, int level) const {
        auto nl = [&](int lv) { if (indent >= 0) { o += '\n'; o.append((size_t)(indent * lv), ' '); } };
        switch (t_) {
            case Null: o += "null"; break;
            case Bool: o += b_ ? "true" : "false"; break;
            case Num: write_num(o, n_); break;
            case Str: write_str(o, s_); break;
            case Arr:
                o += '[';
                for (size_t i = 0; i < a_.size(); i++) { if (i) o += ','; nl(level + 1); a_[i].write(o, indent, level + 1); }
                if (!a_.empty()) nl(level);
                o += ']';
                break;
            case Obj:
                o += '{';
                for (size_t i = 0; i < o_.size(); i++) {
                    if (i) o += ',';
                    nl(level + 1);
                    write_str(o, o_[i].first);
                    o += indent >= 0 ? ": " : ":";
                    o_[i].second.write(o, indent, level + 1);
                }
                if (!o_.empty()) nl(level);
                o += '}';
                break;
        }
    }
};

}  // namespace mrt_host
