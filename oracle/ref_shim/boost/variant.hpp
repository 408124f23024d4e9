#pragma once
#include <variant>
namespace boost {
template <class... T> class variant : public std::variant<T...> {
public:
    using std::variant<T...>::variant;
    using std::variant<T...>::operator=;
    variant() {}
    const std::variant<T...> &base() const { return *this; }
    std::variant<T...> &base() { return *this; }
};
template <class R = void> struct static_visitor { typedef R result_type; };
template <class U, class... T> inline U &get(variant<T...> &v) { return std::get<U>(v.base()); }
template <class U, class... T> inline const U &get(const variant<T...> &v) { return std::get<U>(v.base()); }
template <class U, class... T> inline U *get(variant<T...> *v) { return std::get_if<U>(&v->base()); }
template <class U, class... T> inline const U *get(const variant<T...> *v) { return std::get_if<U>(&v->base()); }
template <class V, class... T> inline typename V::result_type apply_visitor(const V &vis, const variant<T...> &v) {
    return std::visit([&](const auto &x) -> typename V::result_type { return vis(x); }, v.base());
}
template <class V, class... T> inline typename V::result_type apply_visitor(V &vis, const variant<T...> &v) {
    return std::visit([&](const auto &x) -> typename V::result_type { return vis(x); }, v.base());
}
typedef std::bad_variant_access bad_get;
}
