// input.h — the host driver's JSON document type.
// Mirrors the reference's `json_t = input` wrapper (dmc/input.h:97-330): operator[] with
// access-path error messages (missingKey / invalidType), get<T>(), find()/contains(), iteration in
// key-sorted order (nlohmann's std::map objects — the factories' id scheme depends on it,
// dmc/abstractFactory.cpp:3-20) and "@file" links (dmc/input.cpp:7-34).  Self-contained
// recursive-descent parser: no third-party code.
#pragma once
#include <exception>
#include <map>
#include <memory>
#include <string>
#include <vector>

class missingKey : public std::exception {
 public:
  explicit missingKey(const std::string &k) : message("Missing key: " + k) {}
  missingKey(const std::string &k, const std::string &parent) : message("Missing key [" + k + "] in json object: \n" + parent) {}
  missingKey(int index, int maxLen) : message("Missing Index [" + std::to_string(index) + "] . Max Range is : " + std::to_string(maxLen)) {}
  const char *what() const throw() override { return message.c_str(); }

 private:
  std::string message;
};

class invalidType : public std::exception {
 public:
  explicit invalidType(const std::string &m) : message(m) {}
  const char *what() const throw() override { return message.c_str(); }

 private:
  std::string message;
};

class json_t {
 public:
  enum kind_t { Null, Bool, Number, String, Array, Object };
  json_t() : kind(Null) {}
  static json_t parse(const std::string &text);
  static json_t loadFromFile(const std::string &filename);

  bool is_null() const { return kind == Null; }
  bool is_object() const { return kind == Object; }
  bool is_array() const { return kind == Array; }
  bool is_structured() const { return kind == Object || kind == Array; }
  bool is_string() const { return kind == String; }
  bool is_number() const { return kind == Number; }
  size_t size() const { return kind == Array ? arr.size() : kind == Object ? obj.size() : (kind == Null ? 0 : 1); }

  bool contains(const std::string &key) const { return kind == Object && obj.count(key) > 0; }
  const json_t &operator[](const std::string &key) const;
  const json_t &operator[](size_t i) const;
  // structured children in iteration order: array order, or key-sorted for objects
  std::vector<const json_t *> children() const;
  const std::map<std::string, json_t> &items() const { return obj; }
  const std::vector<json_t> &elements() const { return arr; }

  template <class T> T get() const;
  template <class T> T value(const std::string &key, T fallback) const { return contains(key) ? (*this)[key].get<T>() : fallback; }
  std::string dump() const;
  void resolveLinks();

 private:
  friend struct json_parser;
  kind_t kind;
  bool b = false;
  double num = 0;
  bool isInteger = false;
  std::string str;
  std::vector<json_t> arr;
  std::map<std::string, json_t> obj;
};

template <> double json_t::get<double>() const;
template <> int json_t::get<int>() const;
template <> long json_t::get<long>() const;
template <> size_t json_t::get<size_t>() const;
template <> bool json_t::get<bool>() const;
template <> std::string json_t::get<std::string>() const;
template <> std::vector<double> json_t::get<std::vector<double> >() const;
template <> std::vector<int> json_t::get<std::vector<int> >() const;
