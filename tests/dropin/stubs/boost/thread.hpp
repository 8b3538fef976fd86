#pragma once
#include <functional>
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
template <class S> using function = std::function<S>;
template <class... A> auto bind(A &&...a) -> decltype(std::bind(std::forward<A>(a)...)) { return std::bind(std::forward<A>(a)...); }
}
using std::placeholders::_1;
using std::placeholders::_2;
