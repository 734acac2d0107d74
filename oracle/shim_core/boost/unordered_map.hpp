/* Stand-in header (test infrastructure only, see oracle/shim_core/README). */
#pragma once
#include <unordered_map>
namespace boost { template <typename K, typename V, typename H = std::hash<K>, typename E = std::equal_to<K> > using unordered_map = std::unordered_map<K, V, H, E>; }
