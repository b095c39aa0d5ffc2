#include "input.h"
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

struct json_parser {
  const std::string &s;
  size_t p = 0;
  explicit json_parser(const std::string &text) : s(text) {}
  [[noreturn]] void fail(const std::string &m) const { throw invalidType("JSON parse error at offset " + std::to_string(p) + ": " + m); }
  void ws()
  {
    while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
  }
  json_t value()
  {
    ws();
    if (p >= s.size()) fail("unexpected end");
    json_t v;
    const char c = s[p];
    if (c == '{')
      {
        v.kind = json_t::Object;
        ++p;
        ws();
        if (s[p] == '}') { ++p; return v; }
        for (;;)
          {
            ws();
            if (s[p] != '"') fail("expected key");
            std::string k = string();
            ws();
            if (s[p] != ':') fail("expected ':'");
            ++p;
            v.obj[k] = value();
            ws();
            if (s[p] == ',') { ++p; continue; }
            if (s[p] == '}') { ++p; break; }
            fail("expected ',' or '}'");
          }
      }
    else if (c == '[')
      {
        v.kind = json_t::Array;
        ++p;
        ws();
        if (s[p] == ']') { ++p; return v; }
        for (;;)
          {
            v.arr.push_back(value());
            ws();
            if (s[p] == ',') { ++p; continue; }
            if (s[p] == ']') { ++p; break; }
            fail("expected ',' or ']'");
          }
      }
    else if (c == '"')
      {
        v.kind = json_t::String;
        v.str = string();
      }
    else if (s.compare(p, 4, "true") == 0) { v.kind = json_t::Bool; v.b = true; p += 4; }
    else if (s.compare(p, 5, "false") == 0) { v.kind = json_t::Bool; v.b = false; p += 5; }
    else if (s.compare(p, 4, "null") == 0) { p += 4; }
    else
      {
        const char *b = s.c_str() + p;
        char *e = nullptr;
        v.num = std::strtod(b, &e);
        if (e == b) fail("unexpected character");
        v.kind = json_t::Number;
        v.isInteger = true;
        for (const char *q = b; q < e; ++q)
          if (*q == '.' || *q == 'e' || *q == 'E') v.isInteger = false;
        p += (size_t)(e - b);
      }
    return v;
  }
  std::string string()
  {
    std::string out;
    ++p;
    while (p < s.size() && s[p] != '"')
      {
        if (s[p] == '\\' && p + 1 < s.size())
          {
            ++p;
            switch (s[p])
              {
              case 'n': out += '\n'; break;
              case 't': out += '\t'; break;
              case 'r': out += '\r'; break;
              case 'b': out += '\b'; break;
              case 'f': out += '\f'; break;
              case 'u': out += '?'; p += 4; break;
              default: out += s[p];
              }
            ++p;
          }
        else out += s[p++];
      }
    if (p >= s.size()) fail("unterminated string");
    ++p;
    return out;
  }
};

json_t json_t::parse(const std::string &text)
{
  json_parser P(text);
  json_t v = P.value();
  P.ws();
  if (P.p != text.size()) P.fail("trailing characters");
  v.resolveLinks();
  return v;
}

json_t json_t::loadFromFile(const std::string &filename)
{
  std::ifstream f(filename);
  if (!f) throw invalidType("cannot open " + filename);
  std::stringstream ss;
  ss << f.rdbuf();
  return parse(ss.str());
}

void json_t::resolveLinks()
{ /* string values "@path" are replaced by the JSON file at path (dmc/input.cpp:7-34) */
  if (kind == String && str.size() > 1 && str[0] == '@')
    {
      *this = loadFromFile(str.substr(1));
      return;
    }
  for (auto &e : arr) e.resolveLinks();
  for (auto &kv : obj) kv.second.resolveLinks();
}

const json_t &json_t::operator[](const std::string &key) const
{
  if (kind != Object) throw invalidType("not an object while accessing key [" + key + "] in: \n" + dump());
  auto it = obj.find(key);
  if (it == obj.end()) throw missingKey(key, dump());
  return it->second;
}

const json_t &json_t::operator[](size_t i) const
{
  if (kind != Array) throw invalidType("not an array while accessing index " + std::to_string(i) + " in: \n" + dump());
  if (i >= arr.size()) throw missingKey((int)i, (int)arr.size());
  return arr[i];
}

std::vector<const json_t *> json_t::children() const
{
  std::vector<const json_t *> out;
  for (const auto &e : arr) out.push_back(&e);
  for (const auto &kv : obj) out.push_back(&kv.second);
  return out;
}

std::string json_t::dump() const
{
  std::ostringstream o;
  switch (kind)
    {
    case Null: o << "null"; break;
    case Bool: o << (b ? "true" : "false"); break;
    case Number:
      if (isInteger) o << (long long)num;
      else { o.precision(17); o << num; }
      break;
    case String: o << '"' << str << '"'; break;
    case Array:
      o << '[';
      for (size_t i = 0; i < arr.size(); ++i) o << (i ? "," : "") << arr[i].dump();
      o << ']';
      break;
    case Object:
      {
        o << '{';
        bool first = true;
        for (const auto &kv : obj)
          {
            o << (first ? "" : ",") << '"' << kv.first << "\":" << kv.second.dump();
            first = false;
          }
        o << '}';
      }
    }
  return o.str();
}

static void need(bool ok, const json_t &j, const char *want)
{
  if (!ok) throw invalidType(std::string("Wrong type, expected ") + want + " in: \n" + j.dump());
}
template <> double json_t::get<double>() const { need(kind == Number, *this, "number"); return num; }
template <> int json_t::get<int>() const { need(kind == Number, *this, "integer"); return (int)num; }
template <> long json_t::get<long>() const { need(kind == Number, *this, "integer"); return (long)num; }
template <> size_t json_t::get<size_t>() const { need(kind == Number, *this, "integer"); return (size_t)num; }
template <> bool json_t::get<bool>() const { need(kind == Bool, *this, "bool"); return b; }
template <> std::string json_t::get<std::string>() const { need(kind == String, *this, "string"); return str; }
template <> std::vector<double> json_t::get<std::vector<double> >() const
{
  need(kind == Array, *this, "array");
  std::vector<double> v;
  for (const auto &e : arr) v.push_back(e.get<double>());
  return v;
}
template <> std::vector<int> json_t::get<std::vector<int> >() const
{
  need(kind == Array, *this, "array");
  std::vector<int> v;
  for (const auto &e : arr) v.push_back(e.get<int>());
  return v;
}
