#pragma once
#include <functional>
namespace boost {
namespace placeholders { using namespace std::placeholders; }
using std::bind; using std::ref; using std::cref;
}
