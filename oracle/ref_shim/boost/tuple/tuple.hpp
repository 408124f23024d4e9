#pragma once
#include <tuple>
namespace boost {
template <class... T> using tuple = std::tuple<T...>;
using std::make_tuple; using std::tie;
namespace tuples { using std::get; }
using std::get;
}
