// json_lite.h — tiny JSON reader for model `config.json` files (the reference parses these only on
// the Go side, server/main.go:605-674; the C++ ModelRepository ignored their content,
// model_repository.cpp:136-145 — reading them here is a documented addition).
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {
namespace json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;
    const Value* Get(const std::string& k) const {
        auto it = obj.find(k);
        return it == obj.end() ? nullptr : &it->second;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    Value Parse() {
        Value v = ParseValue();
        Ws();
        if (p_ != s_.size()) throw std::runtime_error("json: trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_ = 0;
    void Ws() { while (p_ < s_.size() && isspace((unsigned char)s_[p_])) ++p_; }
    char Peek() { Ws(); if (p_ >= s_.size()) throw std::runtime_error("json: unexpected end"); return s_[p_]; }
    void Expect(char c) { if (Peek() != c) throw std::runtime_error(std::string("json: expected '") + c + "'"); ++p_; }
    Value ParseValue() {
        char c = Peek();
        Value v;
        if (c == '{') {
            v.kind = Value::Object; ++p_;
            if (Peek() == '}') { ++p_; return v; }
            while (true) {
                std::string k = ParseString();
                Expect(':');
                v.obj[k] = ParseValue();
                if (Peek() == ',') { ++p_; continue; }
                Expect('}');
                return v;
            }
        }
        if (c == '[') {
            v.kind = Value::Array; ++p_;
            if (Peek() == ']') { ++p_; return v; }
            while (true) {
                v.arr.push_back(ParseValue());
                if (Peek() == ',') { ++p_; continue; }
                Expect(']');
                return v;
            }
        }
        if (c == '"') { v.kind = Value::String; v.str = ParseString(); return v; }
        if (s_.compare(p_, 4, "true") == 0) { p_ += 4; v.kind = Value::Bool; v.b = true; return v; }
        if (s_.compare(p_, 5, "false") == 0) { p_ += 5; v.kind = Value::Bool; return v; }
        if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return v; }
        char* end = nullptr;
        v.num = strtod(s_.c_str() + p_, &end);
        if (end == s_.c_str() + p_) throw std::runtime_error("json: bad token");
        p_ = (size_t)(end - s_.c_str());
        v.kind = Value::Number;
        return v;
    }
    std::string ParseString() {
        Expect('"');
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c == '\\' && p_ < s_.size()) {
                char e = s_[p_++];
                switch (e) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break;
                    case 'u': p_ += 4; out.push_back('?'); break;
                    default: out.push_back(e);
                }
            } else {
                out.push_back(c);
            }
        }
        if (p_ >= s_.size()) throw std::runtime_error("json: unterminated string");
        ++p_;
        return out;
    }
};

inline Value ParseString(const std::string& s) { return Parser(s).Parse(); }

}  // namespace json
}  // namespace b200
