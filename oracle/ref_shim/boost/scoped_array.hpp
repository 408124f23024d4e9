#pragma once
#include <cstddef>
namespace boost {
template <class T> class scoped_array {
    T *p_;
    scoped_array(const scoped_array &);
    scoped_array &operator=(const scoped_array &);
public:
    explicit scoped_array(T *p = 0) : p_(p) {}
    ~scoped_array() { delete[] p_; }
    void reset(T *p = 0) { if (p != p_) { delete[] p_; p_ = p; } }
    T &operator[](std::ptrdiff_t i) const { return p_[i]; }
    T *get() const { return p_; }
};
}
