// A small JSON reader for the factories' `create(const char* jsonParameters)` entry points (the reference uses
// nlohmann/json 3.11.2, src/CMakeLists.txt:29-36; this library has no third-party dependency).  Parses objects,
// arrays, strings (with the common escapes), numbers, true/false/null into a tree; throws std::invalid_argument on
// malformed input (mapped to Status_ParseError / Status_InvalidArgument at the library edge).
#pragma once

#include <cctype>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace gs {

class Json {
 public:
  enum Kind { Null, Bool, Number, String, Array, Object };
  Kind kind = Null;
  bool boolean = false;
  double number = 0.0;
  std::string text;
  std::vector<Json> items;
  std::vector<std::pair<std::string, Json>> members;  // insertion order kept

  static Json parse(const char* src) {
    if (src == nullptr) throw std::invalid_argument("JSON text is null");
    const char* p = src;
    Json v = parseValue(p);
    skip(p);
    if (*p != '\0') throw std::invalid_argument("trailing characters after JSON value");
    return v;
  }

  bool contains(const std::string& key) const { return find(key) != nullptr; }
  const Json& at(const std::string& key) const {
    const Json* v = find(key);
    if (v == nullptr) throw std::invalid_argument("missing JSON key \"" + key + "\"");
    return *v;
  }
  const std::string& str() const {
    if (kind != String) throw std::invalid_argument("JSON value is not a string");
    return text;
  }
  double num() const {
    if (kind != Number) throw std::invalid_argument("JSON value is not a number");
    return number;
  }
  const std::vector<Json>& array() const {
    if (kind != Array) throw std::invalid_argument("JSON value is not an array");
    return items;
  }
  std::string dump() const {
    switch (kind) {
      case Null: return "null";
      case Bool: return boolean ? "true" : "false";
      case Number: {
        char buf[40];
        snprintf(buf, sizeof(buf), "%.17g", number);
        return buf;
      }
      case String: return "\"" + text + "\"";
      case Array: {
        std::string s = "[";
        for (size_t i = 0; i < items.size(); i++) s += (i ? "," : "") + items[i].dump();
        return s + "]";
      }
      default: {
        std::string s = "{";
        for (size_t i = 0; i < members.size(); i++) s += (i ? ",\"" : "\"") + members[i].first + "\":" + members[i].second.dump();
        return s + "}";
      }
    }
  }

 private:
  const Json* find(const std::string& key) const {
    if (kind != Object) return nullptr;
    for (const auto& m : members)
      if (m.first == key) return &m.second;
    return nullptr;
  }
  static void skip(const char*& p) {
    while (*p && isspace(static_cast<unsigned char>(*p))) p++;
  }
  static std::string parseString(const char*& p) {
    std::string out;
    p++;  // opening quote
    while (*p && *p != '"') {
      if (*p == '\\') {
        p++;
        switch (*p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case '\0': throw std::invalid_argument("unterminated escape in JSON string");
          default: out += *p; break;  // \" \\ \/ (and \uXXXX is passed through unescaped)
        }
        p++;
      } else {
        out += *p++;
      }
    }
    if (*p != '"') throw std::invalid_argument("unterminated JSON string");
    p++;
    return out;
  }
  static Json parseValue(const char*& p) {
    skip(p);
    Json v;
    if (*p == '{') {
      v.kind = Object;
      p++;
      skip(p);
      if (*p == '}') { p++; return v; }
      for (;;) {
        skip(p);
        if (*p != '"') throw std::invalid_argument("expected a string key in JSON object");
        std::string key = parseString(p);
        skip(p);
        if (*p != ':') throw std::invalid_argument("expected ':' in JSON object");
        p++;
        v.members.emplace_back(std::move(key), parseValue(p));
        skip(p);
        if (*p == ',') { p++; continue; }
        if (*p == '}') { p++; return v; }
        throw std::invalid_argument("expected ',' or '}' in JSON object");
      }
    }
    if (*p == '[') {
      v.kind = Array;
      p++;
      skip(p);
      if (*p == ']') { p++; return v; }
      for (;;) {
        v.items.push_back(parseValue(p));
        skip(p);
        if (*p == ',') { p++; continue; }
        if (*p == ']') { p++; return v; }
        throw std::invalid_argument("expected ',' or ']' in JSON array");
      }
    }
    if (*p == '"') {
      v.kind = String;
      v.text = parseString(p);
      return v;
    }
    if (!strncmp(p, "true", 4)) { v.kind = Bool; v.boolean = true; p += 4; return v; }
    if (!strncmp(p, "false", 5)) { v.kind = Bool; p += 5; return v; }
    if (!strncmp(p, "null", 4)) { p += 4; return v; }
    char* end = nullptr;
    v.number = strtod(p, &end);
    if (end == p) throw std::invalid_argument("unexpected character in JSON text");
    v.kind = Number;
    p = end;
    return v;
  }
};

}  // namespace gs
