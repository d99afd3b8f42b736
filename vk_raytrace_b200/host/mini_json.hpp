// mini_json.hpp -- a small recursive-descent JSON reader (RFC 8259) for the glTF importer.
// The reference parses glTF with tinygltf (src/scene.cpp:123-155), which is not vendored; nothing else in this
// repository needs JSON, so a 150-line reader is preferable to a dependency.
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace mj {

struct Value;
using Array  = std::vector<Value>;
using Object = std::map<std::string, Value>;

struct Value
{
  enum Type { Null, Bool, Number, String, Arr, Obj } type = Null;
  bool                    b = false;
  double                  num = 0.0;
  std::string             str;
  std::shared_ptr<Array>  arr;
  std::shared_ptr<Object> obj;

  bool isNull() const { return type == Null; }
  bool isObject() const { return type == Obj; }
  bool isArray() const { return type == Arr; }
  bool isNumber() const { return type == Number; }
  bool isString() const { return type == String; }
  bool has(const std::string& k) const { return type == Obj && obj->count(k) != 0; }
  const Value& operator[](const std::string& k) const
  {
    static const Value null;
    if(type != Obj)
      return null;
    auto it = obj->find(k);
    return it == obj->end() ? null : it->second;
  }
  const Value& operator[](size_t i) const
  {
    static const Value null;
    return (type == Arr && i < arr->size()) ? (*arr)[i] : null;
  }
  size_t size() const { return type == Arr ? arr->size() : (type == Obj ? obj->size() : 0); }
  double number(double def = 0.0) const { return type == Number ? num : def; }
  int    integer(int def = -1) const { return type == Number ? (int)std::llround(num) : def; }
  bool   boolean(bool def = false) const { return type == Bool ? b : def; }
  const std::string& string() const { return str; }
};

class Parser
{
public:
  explicit Parser(const std::string& text) : s(text) {}
  bool parse(Value& out, std::string& err)
  {
    try
    {
      skip();
      out = value();
      skip();
      if(p != s.size())
        fail("trailing characters");
      return true;
    }
    catch(const std::string& e)
    {
      err = e + " at byte " + std::to_string(p);
      return false;
    }
  }

private:
  const std::string& s;
  size_t             p = 0;
  [[noreturn]] void fail(const char* m) { throw std::string(m); }
  void skip() { while(p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++; }
  bool match(const char* lit)
  {
    size_t n = std::char_traits<char>::length(lit);
    if(s.compare(p, n, lit) == 0) { p += n; return true; }
    return false;
  }
  Value value()
  {
    if(p >= s.size()) fail("unexpected end");
    Value v;
    const char c = s[p];
    if(c == '{')
    {
      v.type = Value::Obj;
      v.obj  = std::make_shared<Object>();
      p++; skip();
      if(p < s.size() && s[p] == '}') { p++; return v; }
      for(;;)
      {
        skip();
        if(p >= s.size() || s[p] != '"') fail("expected key");
        std::string k = str();
        skip();
        if(p >= s.size() || s[p] != ':') fail("expected ':'");
        p++; skip();
        (*v.obj)[k] = value();
        skip();
        if(p < s.size() && s[p] == ',') { p++; continue; }
        if(p < s.size() && s[p] == '}') { p++; break; }
        fail("expected ',' or '}'");
      }
    }
    else if(c == '[')
    {
      v.type = Value::Arr;
      v.arr  = std::make_shared<Array>();
      p++; skip();
      if(p < s.size() && s[p] == ']') { p++; return v; }
      for(;;)
      {
        skip();
        v.arr->push_back(value());
        skip();
        if(p < s.size() && s[p] == ',') { p++; continue; }
        if(p < s.size() && s[p] == ']') { p++; break; }
        fail("expected ',' or ']'");
      }
    }
    else if(c == '"')
    {
      v.type = Value::String;
      v.str  = str();
    }
    else if(match("true")) { v.type = Value::Bool; v.b = true; }
    else if(match("false")) { v.type = Value::Bool; v.b = false; }
    else if(match("null")) { v.type = Value::Null; }
    else
    {
      char*  end = nullptr;
      double d   = std::strtod(s.c_str() + p, &end);
      if(end == s.c_str() + p) fail("bad token");
      p      = (size_t)(end - s.c_str());
      v.type = Value::Number;
      v.num  = d;
    }
    return v;
  }
  std::string str()
  {
    std::string out;
    p++;  // opening quote
    while(p < s.size() && s[p] != '"')
    {
      char c = s[p++];
      if(c == '\\')
      {
        if(p >= s.size()) fail("bad escape");
        char e = s[p++];
        switch(e)
        {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u':
          {
            if(p + 4 > s.size()) fail("bad \\u");
            unsigned cp = (unsigned)std::strtoul(s.substr(p, 4).c_str(), nullptr, 16);
            p += 4;
            if(cp < 0x80) out += (char)cp;
            else if(cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e; break;  // \" \\ \/
        }
      }
      else
        out += c;
    }
    if(p >= s.size()) fail("unterminated string");
    p++;
    return out;
  }
};

inline bool parse(const std::string& text, Value& out, std::string& err)
{
  Parser pr(text);
  return pr.parse(out, err);
}

}  // namespace mj
