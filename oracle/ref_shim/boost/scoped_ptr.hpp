#pragma once
namespace boost {
template <class T> class scoped_ptr {
    T *p_;
    scoped_ptr(const scoped_ptr &);
    scoped_ptr &operator=(const scoped_ptr &);
public:
    explicit scoped_ptr(T *p = 0) : p_(p) {}
    ~scoped_ptr() { delete p_; }
    void reset(T *p = 0) { if (p != p_) { delete p_; p_ = p; } }
    T &operator*() const { return *p_; }
    T *operator->() const { return p_; }
    T *get() const { return p_; }
    explicit operator bool() const { return p_ != 0; }
    bool operator!() const { return p_ == 0; }
};
}
