#pragma once
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <thread>
namespace boost {
namespace posix_time {
typedef std::chrono::system_clock::time_point ptime;
inline std::chrono::milliseconds milliseconds(long long ms) { return std::chrono::milliseconds(ms); }
}
inline posix_time::ptime get_system_time() { return std::chrono::system_clock::now(); }
template <class M> class scoped_lock_shim : public std::unique_lock<M> {
public:
    explicit scoped_lock_shim(M &m) : std::unique_lock<M>(m) {}
};
class mutex : public std::mutex { public: typedef scoped_lock_shim<mutex> scoped_lock; };
class recursive_mutex : public std::recursive_mutex { public: typedef scoped_lock_shim<recursive_mutex> scoped_lock; };
class timed_mutex : public std::timed_mutex { public: typedef scoped_lock_shim<timed_mutex> scoped_lock; };
class recursive_timed_mutex : public std::recursive_timed_mutex { public: typedef scoped_lock_shim<recursive_timed_mutex> scoped_lock; };
template <class M> using lock_guard = std::lock_guard<M>;
template <class M> using unique_lock = std::unique_lock<M>;
class condition_variable_any : public std::condition_variable_any {
public:
    template <class L> bool timed_wait(L &l, const posix_time::ptime &t) { return wait_until(l, t) == std::cv_status::no_timeout; }
};
struct thread_interrupted {};
struct thread_resource_error : public std::runtime_error { thread_resource_error() : std::runtime_error("thread_resource_error") {} };
class thread : public std::thread {
public:
    thread() {}
    template <class F, class... A> explicit thread(F &&f, A &&...a) : std::thread(std::forward<F>(f), std::forward<A>(a)...) {}
    thread(thread &&o) : std::thread(std::move(static_cast<std::thread &>(o))) {}
    thread &operator=(thread &&o) { std::thread::operator=(std::move(static_cast<std::thread &>(o))); return *this; }
    static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
};
namespace this_thread {
inline void yield() { std::this_thread::yield(); }
template <class D> inline void sleep(const D &d) { std::this_thread::sleep_for(d); }
}
}
