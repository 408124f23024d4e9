#pragma once
#include <functional>
namespace boost {
template <class S> class function : public std::function<S> {
public:
    using std::function<S>::function;
    function() {}
    bool empty() const { return !static_cast<const std::function<S> &>(*this); }
    void clear() { std::function<S>::operator=(nullptr); }
};
}
