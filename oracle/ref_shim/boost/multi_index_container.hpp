#pragma once
// The one shape the reference uses (core/tls.cpp): elements in insertion order + a unique hashed key over a member.
#include <list>
#include <unordered_map>
namespace boost { namespace multi_index {
template <class T> struct tag { typedef T type; };
template <class C, class M, M C::*P> struct member { typedef M result_type; static const M &key(const C &c) { return c.*P; } };
template <class Tag, class Key> struct hashed_unique { typedef Key key; };
template <class Tag> struct sequenced {};
template <class A, class B> struct indexed_by { typedef typename A::key key; };
template <class T, class Spec> class multi_index_container {
    typedef typename Spec::key key_from;
    typedef typename key_from::result_type key_type;
public:
    typedef std::list<T> sequence;
    typedef typename sequence::iterator iterator;
    sequence seq;
    std::unordered_map<key_type, iterator> idx;
    iterator end() { return seq.end(); }
    iterator find(const key_type &k) { auto i = idx.find(k); return i == idx.end() ? seq.end() : i->second; }
    void erase(iterator it) { idx.erase(key_from::key(*it)); seq.erase(it); }
    std::pair<iterator, bool> insert(const T &v) {
        auto i = idx.find(key_from::key(v));
        if (i != idx.end()) return std::make_pair(i->second, false);
        seq.push_back(v);
        iterator it = --seq.end();
        idx[key_from::key(v)] = it;
        return std::make_pair(it, true);
    }
};
template <class C, class Tag> struct index { typedef typename C::sequence type; };
template <class Tag, class T, class S> inline typename multi_index_container<T, S>::sequence &get(multi_index_container<T, S> &c) { return c.seq; }
} }
