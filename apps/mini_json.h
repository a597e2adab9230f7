// mini_json.h — the small subset of JSON the MyTRIM input files use, with // and /* */ comments
// (the reference reads them through jsoncpp, which is not a dependency of this project).
#ifndef MYTRIM_B200_MINI_JSON_H
#define MYTRIM_B200_MINI_JSON_H

#include <cctype>
#include <cstdlib>
#include <istream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json
{

class Value
{
public:
  enum Kind { Null, Bool, Number, String, Array, Object };

  Value() : _kind(Null), _num(0), _bool(false) {}

  bool isNull() const { return _kind == Null; }
  bool isBool() const { return _kind == Bool; }
  bool isNumeric() const { return _kind == Number; }
  bool isString() const { return _kind == String; }
  bool isArray() const { return _kind == Array; }
  bool isObject() const { return _kind == Object; }

  double asDouble() const { return _num; }
  int asInt() const { return (int)_num; }
  long long asInt64() const { return (long long)_num; }
  bool asBool() const { return _bool; }
  const std::string & asString() const { return _str; }
  size_t size() const { return _kind == Array ? _items.size() : _members.size(); }

  // missing members / out-of-range items read as Null, like jsoncpp
  const Value & operator[](const std::string & key) const
  {
    auto it = _members.find(key);
    return it == _members.end() ? null() : it->second;
  }
  const Value & operator[](const char * key) const { return (*this)[std::string(key)]; }
  const Value & operator[](size_t i) const { return i < _items.size() ? _items[i] : null(); }
  const Value & operator[](int i) const { return (*this)[(size_t)i]; }

  static Value parse(const std::string & text)
  {
    Parser p{text, 0};
    Value v = p.value();
    p.skip();
    if (p.pos != text.size())
      p.fail("trailing characters");
    return v;
  }

  static Value parse(std::istream & in)
  {
    return parse(std::string(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>()));
  }

private:
  static const Value & null()
  {
    static const Value v;
    return v;
  }

  struct Parser
  {
    const std::string & s;
    size_t pos;

    [[noreturn]] void fail(const std::string & what)
    {
      size_t line = 1;
      for (size_t i = 0; i < pos && i < s.size(); ++i)
        if (s[i] == '\n')
          ++line;
      throw std::runtime_error("JSON error at line " + std::to_string(line) + ": " + what);
    }

    void skip()
    {
      for (;;)
      {
        while (pos < s.size() && std::isspace((unsigned char)s[pos]))
          ++pos;
        if (pos + 1 < s.size() && s[pos] == '/' && s[pos + 1] == '/')
        {
          while (pos < s.size() && s[pos] != '\n')
            ++pos;
        }
        else if (pos + 1 < s.size() && s[pos] == '/' && s[pos + 1] == '*')
        {
          const size_t end = s.find("*/", pos + 2);
          if (end == std::string::npos)
            fail("unterminated comment");
          pos = end + 2;
        }
        else
          return;
      }
    }

    std::string string()
    {
      std::string out;
      ++pos; // opening quote
      while (pos < s.size() && s[pos] != '"')
      {
        char c = s[pos++];
        if (c == '\\' && pos < s.size())
        {
          const char e = s[pos++];
          switch (e)
          {
            case 'n': c = '\n'; break;
            case 't': c = '\t'; break;
            case 'r': c = '\r'; break;
            case 'b': c = '\b'; break;
            case 'f': c = '\f'; break;
            default: c = e;
          }
        }
        out.push_back(c);
      }
      if (pos >= s.size())
        fail("unterminated string");
      ++pos;
      return out;
    }

    bool at(char c) const { return pos < s.size() && s[pos] == c; }

    Value value()
    {
      skip();
      if (pos >= s.size())
        fail("unexpected end of input");
      Value v;
      const char c = s[pos];
      if (c == '{')
      {
        v._kind = Object;
        ++pos;
        skip();
        if (at('}'))
        {
          ++pos;
          return v;
        }
        for (;;)
        {
          skip();
          if (!at('"'))
            fail("expected a member name");
          const std::string key = string();
          skip();
          if (!at(':'))
            fail("expected ':'");
          ++pos;
          v._members[key] = value();
          skip();
          if (at(','))
          {
            ++pos;
            continue;
          }
          if (at('}'))
          {
            ++pos;
            return v;
          }
          fail("expected ',' or '}'");
        }
      }
      if (c == '[')
      {
        v._kind = Array;
        ++pos;
        skip();
        if (at(']'))
        {
          ++pos;
          return v;
        }
        for (;;)
        {
          v._items.push_back(value());
          skip();
          if (at(','))
          {
            ++pos;
            continue;
          }
          if (at(']'))
          {
            ++pos;
            return v;
          }
          fail("expected ',' or ']'");
        }
      }
      if (c == '"')
      {
        v._kind = String;
        v._str = string();
        return v;
      }
      if (s.compare(pos, 4, "true") == 0 || s.compare(pos, 5, "false") == 0)
      {
        v._kind = Bool;
        v._bool = s[pos] == 't';
        pos += v._bool ? 4 : 5;
        return v;
      }
      if (s.compare(pos, 4, "null") == 0)
      {
        pos += 4;
        return v;
      }
      char * end = nullptr;
      const double d = std::strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos)
        fail("unexpected character");
      pos = (size_t)(end - s.c_str());
      v._kind = Number;
      v._num = d;
      return v;
    }
  };

  Kind _kind;
  double _num;
  bool _bool;
  std::string _str;
  std::vector<Value> _items;
  std::map<std::string, Value> _members;
};

} // namespace mini_json
#endif
