#pragma once
#include <unordered_set>
namespace boost { template <class K, class H = std::hash<K>, class E = std::equal_to<K> > using unordered_set = std::unordered_set<K, H, E>; }
