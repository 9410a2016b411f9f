// Minimal JSON reader for the reference's scene files (objects, arrays, numbers, strings, true/false/null;
// tolerant of the tabs / trailing whitespace the shipped scenes contain).  The reference parses scenes with
// nlohmann::json through Utilities::SceneLoader (SPlisHSPlasH/Utilities/SceneLoader.cpp:22-40); only the
// value model the loader needs is provided here.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dfrhost {

struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept (RigidBodies order = boundary model index)

  const Json *find(const std::string &key) const {
    if (kind != Object) return nullptr;
    for (const auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool has(const std::string &key) const { return find(key) != nullptr; }
  // SceneLoader::readValue semantics (SceneLoader.cpp:516-530): missing or null -> leave the target untouched
  bool read(const std::string &key, double &v) const {
    const Json *j = find(key);
    if (!j) return false;
    if (j->kind == Number) { v = j->num; return true; }
    if (j->kind == Bool) { v = j->b ? 1.0 : 0.0; return true; }
    return false;
  }
  bool read(const std::string &key, int &v) const {
    double d;
    if (!read(key, d)) return false;
    v = (int)d;
    return true;
  }
  bool read(const std::string &key, bool &v) const {
    const Json *j = find(key);
    if (!j) return false;
    if (j->kind == Bool) { v = j->b; return true; }
    if (j->kind == Number) { v = j->num != 0.0; return true; }  // "isDynamic": 1 in the shipped scenes
    return false;
  }
  bool read(const std::string &key, std::string &v) const {
    const Json *j = find(key);
    if (!j || j->kind != String) return false;
    v = j->str;
    return true;
  }
  template <size_t N>
  bool read_vec(const std::string &key, double (&v)[N]) const {
    const Json *j = find(key);
    if (!j || j->kind != Array || j->arr.size() < N) return false;
    for (size_t i = 0; i < N; i++) v[i] = j->arr[i].num;
    return true;
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string &text) : s(text) {}
  Json parse() {
    Json v = value();
    ws();
    if (p != s.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string &s;
  size_t p = 0;
  [[noreturn]] void fail(const char *what) const {
    throw std::runtime_error(std::string("scene JSON: ") + what + " at offset " + std::to_string(p));
  }
  void ws() {
    while (p < s.size()) {
      if (std::isspace((unsigned char)s[p])) { p++; continue; }
      if (s[p] == '/' && p + 1 < s.size() && s[p + 1] == '/') {  // comment lines appear in some scenes
        while (p < s.size() && s[p] != '\n') p++;
        continue;
      }
      break;
    }
  }
  Json value() {
    ws();
    if (p >= s.size()) fail("unexpected end");
    const char c = s[p];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') { Json j; j.kind = Json::String; j.str = string(); return j; }
    if (s.compare(p, 4, "true") == 0) { p += 4; Json j; j.kind = Json::Bool; j.b = true; return j; }
    if (s.compare(p, 5, "false") == 0) { p += 5; Json j; j.kind = Json::Bool; j.b = false; return j; }
    if (s.compare(p, 4, "null") == 0) { p += 4; return Json(); }
    char *end = nullptr;
    const double d = std::strtod(s.c_str() + p, &end);
    if (end == s.c_str() + p) fail("bad value");
    p = (size_t)(end - s.c_str());
    Json j; j.kind = Json::Number; j.num = d;
    return j;
  }
  std::string string() {
    std::string out;
    p++;  // opening quote
    while (p < s.size() && s[p] != '"') {
      if (s[p] == '\\' && p + 1 < s.size()) {
        p++;
        switch (s[p]) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          default: out += s[p];
        }
        p++;
      } else
        out += s[p++];
    }
    if (p >= s.size()) fail("unterminated string");
    p++;
    return out;
  }
  Json array() {
    Json j; j.kind = Json::Array;
    p++;
    ws();
    if (p < s.size() && s[p] == ']') { p++; return j; }
    for (;;) {
      j.arr.push_back(value());
      ws();
      if (p < s.size() && s[p] == ',') { p++; ws(); if (p < s.size() && s[p] == ']') { p++; return j; } continue; }
      if (p < s.size() && s[p] == ']') { p++; return j; }
      fail("expected , or ]");
    }
  }
  Json object() {
    Json j; j.kind = Json::Object;
    p++;
    ws();
    if (p < s.size() && s[p] == '}') { p++; return j; }
    for (;;) {
      ws();
      if (p >= s.size() || s[p] != '"') fail("expected key");
      std::string k = string();
      ws();
      if (p >= s.size() || s[p] != ':') fail("expected :");
      p++;
      j.obj.emplace_back(std::move(k), value());
      ws();
      if (p < s.size() && s[p] == ',') { p++; ws(); if (p < s.size() && s[p] == '}') { p++; return j; } continue; }
      if (p < s.size() && s[p] == '}') { p++; return j; }
      fail("expected , or }");
    }
  }
};

}  // namespace dfrhost
