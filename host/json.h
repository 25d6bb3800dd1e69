// Minimal JSON DOM (RFC 8259) for the Asuna scene format.  The reference parses scenes with
// nlohmann::json (src/ext/json.hpp, third-party); this is an independent ~200-line reader that covers
// what scene files use: objects, arrays, strings with escapes, numbers, booleans, null.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace asuna_host {

class Json {
 public:
  enum Type { Null, Bool, Number, String, Array, Object };
  Type type = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept (ids are assigned by order)

  bool is_object() const { return type == Object; }
  bool is_array() const { return type == Array; }
  bool is_string() const { return type == String; }
  bool is_number() const { return type == Number; }
  bool contains(const std::string& k) const { return find(k) != nullptr; }
  const Json* find(const std::string& k) const {
    if (type != Object) return nullptr;
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Json& at(const std::string& k) const {
    const Json* j = find(k);
    if (!j) throw std::runtime_error("missing key [\"" + k + "\"]");
    return *j;
  }
  const Json& operator[](size_t i) const {
    if (type != Array || i >= arr.size()) throw std::runtime_error("json: array index out of range");
    return arr[i];
  }
  size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }
  double as_number() const {
    if (type == Number) return num;
    if (type == Bool) return b ? 1.0 : 0.0;
    throw std::runtime_error("json: expected a number");
  }
  float as_float() const { return (float)as_number(); }
  int as_int() const { return (int)as_number(); }
  bool as_bool() const {
    if (type == Bool) return b;
    if (type == Number) return num != 0.0;
    throw std::runtime_error("json: expected a boolean");
  }
  const std::string& as_string() const {
    if (type != String) throw std::runtime_error("json: expected a string");
    return str;
  }
  std::vector<float> as_floats() const {
    if (type != Array) throw std::runtime_error("json: expected an array of numbers");
    std::vector<float> v;
    for (auto& e : arr) v.push_back(e.as_float());
    return v;
  }
  // value of key `k` or `def`
  double number_or(const std::string& k, double def) const {
    const Json* j = find(k);
    return j ? j->as_number() : def;
  }
  bool bool_or(const std::string& k, bool def) const {
    const Json* j = find(k);
    return j ? j->as_bool() : def;
  }

  static Json parse(const std::string& text) {
    Parser p{text, 0};
    Json j = p.value();
    p.ws();
    if (p.i != text.size()) p.fail("trailing characters");
    return j;
  }

 private:
  struct Parser {
    const std::string& s;
    size_t i;
    [[noreturn]] void fail(const std::string& m) const {
      size_t line = 1;
      for (size_t k = 0; k < i && k < s.size(); k++)
        if (s[k] == '\n') line++;
      throw std::runtime_error("json parse error at line " + std::to_string(line) + ": " + m);
    }
    void ws() {
      while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++;
    }
    bool lit(const char* w) {
      size_t n = 0;
      while (w[n]) n++;
      if (s.compare(i, n, w) == 0) {
        i += n;
        return true;
      }
      return false;
    }
    static void utf8(std::string& out, unsigned cp) {
      if (cp < 0x80) out += (char)cp;
      else if (cp < 0x800) out += (char)(0xC0 | (cp >> 6)), out += (char)(0x80 | (cp & 0x3F));
      else if (cp < 0x10000)
        out += (char)(0xE0 | (cp >> 12)), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
      else
        out += (char)(0xF0 | (cp >> 18)), out += (char)(0x80 | ((cp >> 12) & 0x3F)), out += (char)(0x80 | ((cp >> 6) & 0x3F)),
            out += (char)(0x80 | (cp & 0x3F));
    }
    unsigned hex4() {
      if (i + 4 > s.size()) fail("bad \\u escape");
      unsigned v = 0;
      for (int k = 0; k < 4; k++) {
        char c = s[i++];
        v <<= 4;
        if (c >= '0' && c <= '9') v |= c - '0';
        else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
        else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
        else fail("bad \\u escape");
      }
      return v;
    }
    std::string string() {
      if (s[i] != '"') fail("expected string");
      i++;
      std::string out;
      while (true) {
        if (i >= s.size()) fail("unterminated string");
        char c = s[i++];
        if (c == '"') break;
        if (c == '\\') {
          if (i >= s.size()) fail("unterminated escape");
          char e = s[i++];
          switch (e) {
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case '/': out += '/'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'n': out += '\n'; break;
            case 'r': out += '\r'; break;
            case 't': out += '\t'; break;
            case 'u': {
              unsigned cp = hex4();
              if (cp >= 0xD800 && cp < 0xDC00 && i + 1 < s.size() && s[i] == '\\' && s[i + 1] == 'u') {
                i += 2;
                unsigned lo = hex4();
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              }
              utf8(out, cp);
              break;
            }
            default: fail("bad escape");
          }
        } else
          out += c;
      }
      return out;
    }
    Json value() {
      ws();
      if (i >= s.size()) fail("unexpected end of input");
      Json j;
      char c = s[i];
      if (c == '{') {
        j.type = Object;
        i++;
        ws();
        if (i < s.size() && s[i] == '}') {
          i++;
          return j;
        }
        while (true) {
          ws();
          std::string k = string();
          ws();
          if (i >= s.size() || s[i] != ':') fail("expected ':'");
          i++;
          j.obj.emplace_back(std::move(k), value());
          ws();
          if (i < s.size() && s[i] == ',') {
            i++;
            continue;
          }
          if (i < s.size() && s[i] == '}') {
            i++;
            break;
          }
          fail("expected ',' or '}'");
        }
      } else if (c == '[') {
        j.type = Array;
        i++;
        ws();
        if (i < s.size() && s[i] == ']') {
          i++;
          return j;
        }
        while (true) {
          j.arr.push_back(value());
          ws();
          if (i < s.size() && s[i] == ',') {
            i++;
            continue;
          }
          if (i < s.size() && s[i] == ']') {
            i++;
            break;
          }
          fail("expected ',' or ']'");
        }
      } else if (c == '"') {
        j.type = String;
        j.str = string();
      } else if (lit("true")) {
        j.type = Bool, j.b = true;
      } else if (lit("false")) {
        j.type = Bool, j.b = false;
      } else if (lit("null")) {
        j.type = Null;
      } else {
        const char* start = s.c_str() + i;
        char* end = nullptr;
        double v = std::strtod(start, &end);
        if (end == start) fail("unexpected character");
        i += (size_t)(end - start);
        j.type = Number, j.num = v;
      }
      return j;
    }
  };
};

}  // namespace asuna_host
