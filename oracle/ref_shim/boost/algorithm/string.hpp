#pragma once
#include <algorithm>
#include <cctype>
#include <string>
namespace boost {
namespace algorithm {
inline void to_lower(std::string &s) { for (char &c : s) c = (char)std::tolower((unsigned char)c); }
inline void to_upper(std::string &s) { for (char &c : s) c = (char)std::toupper((unsigned char)c); }
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
inline std::string to_upper_copy(std::string s) { to_upper(s); return s; }
inline bool starts_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
inline bool ends_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }
inline void trim(std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    s = s.substr(a, b - a);
}
inline std::string trim_copy(std::string s) { trim(s); return s; }
inline bool iequals(const std::string &a, const std::string &b) { return to_lower_copy(a) == to_lower_copy(b); }
}
using algorithm::to_lower; using algorithm::to_upper; using algorithm::to_lower_copy; using algorithm::to_upper_copy;
using algorithm::starts_with; using algorithm::ends_with; using algorithm::trim; using algorithm::trim_copy; using algorithm::iequals;
}
