#pragma once
#include <filesystem>
#include <fstream>
#include <string>
namespace boost { namespace filesystem {
namespace sfs = std::filesystem;
// a real class (the reference forward-declares `class path` in core/fwd.h), otherwise std::filesystem::path
class path : public sfs::path {
public:
    path() {}
    path(const sfs::path &p) : sfs::path(p) {}
    path(const std::string &s) : sfs::path(s) {}
    path(const char *s) : sfs::path(s) {}
    path parent_path() const { return sfs::path::parent_path(); }
    path filename() const { return sfs::path::filename(); }
    path extension() const { return sfs::path::extension(); }
    path stem() const { return sfs::path::stem(); }
    path leaf() const { return sfs::path::filename(); }
    path branch_path() const { return sfs::path::parent_path(); }
    std::string file_string() const { return string(); }
    path &replace_extension(const path &e = path()) { sfs::path::replace_extension(e); return *this; }
    path &operator/=(const path &o) { sfs::path::operator/=(o); return *this; }
    bool is_complete() const { return is_absolute(); }
};
inline path operator/(const path &a, const path &b) { return path(static_cast<const sfs::path &>(a) / static_cast<const sfs::path &>(b)); }
inline bool exists(const path &p) { std::error_code ec; return sfs::exists(p, ec); }
inline bool is_directory(const path &p) { std::error_code ec; return sfs::is_directory(p, ec); }
inline bool is_regular_file(const path &p) { std::error_code ec; return sfs::is_regular_file(p, ec); }
inline path canonical(const path &p) { return sfs::canonical(p); }
inline path absolute(const path &p) { return sfs::absolute(p); }
inline path complete(const path &p) { return sfs::absolute(p); }
inline path current_path() { return sfs::current_path(); }
inline unsigned long long file_size(const path &p) { return sfs::file_size(p); }
inline bool remove(const path &p) { std::error_code ec; return sfs::remove(p, ec); }
inline void resize_file(const path &p, unsigned long long n) { sfs::resize_file(p, n); }
inline bool create_directory(const path &p) { std::error_code ec; return sfs::create_directory(p, ec); }
inline bool create_directories(const path &p) { std::error_code ec; return sfs::create_directories(p, ec); }
inline void rename(const path &a, const path &b) { sfs::rename(a, b); }
class directory_entry_shim {
    sfs::directory_entry e_;
public:
    directory_entry_shim() {}
    directory_entry_shim(const sfs::directory_entry &e) : e_(e) {}
    filesystem::path path() const { return e_.path(); }
};
class directory_iterator {
    sfs::directory_iterator it_;
    mutable directory_entry_shim cur_;
public:
    directory_iterator() {}
    explicit directory_iterator(const path &p) : it_(p) {}
    directory_iterator &operator++() { ++it_; return *this; }
    const directory_entry_shim &operator*() const { cur_ = directory_entry_shim(*it_); return cur_; }
    const directory_entry_shim *operator->() const { cur_ = directory_entry_shim(*it_); return &cur_; }
    bool operator!=(const directory_iterator &o) const { return it_ != o.it_; }
    bool operator==(const directory_iterator &o) const { return it_ == o.it_; }
};
class ifstream : public std::ifstream {
public:
    ifstream() {}
    explicit ifstream(const path &p, std::ios_base::openmode m = std::ios_base::in) : std::ifstream(p.string(), m) {}
    void open(const path &p, std::ios_base::openmode m = std::ios_base::in) { std::ifstream::open(p.string(), m); }
};
class ofstream : public std::ofstream {
public:
    ofstream() {}
    explicit ofstream(const path &p, std::ios_base::openmode m = std::ios_base::out) : std::ofstream(p.string(), m) {}
    void open(const path &p, std::ios_base::openmode m = std::ios_base::out) { std::ofstream::open(p.string(), m); }
};
class fstream : public std::fstream {
public:
    fstream() {}
    explicit fstream(const path &p, std::ios_base::openmode m = std::ios_base::in | std::ios_base::out) : std::fstream(p.string(), m) {}
    void open(const path &p, std::ios_base::openmode m = std::ios_base::in | std::ios_base::out) { std::fstream::open(p.string(), m); }
};
} }
