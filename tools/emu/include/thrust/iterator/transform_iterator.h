// minimal stand-in for thrust::transform_iterator (tools/emu)
#pragma once
#include <cstddef>
#include <iterator>
#include <type_traits>
namespace thrust {
template <class F, class It> struct transform_iterator {
    using value_type = std::decay_t<decltype(std::declval<F>()(*std::declval<It>()))>;
    using difference_type = std::ptrdiff_t;
    using reference = value_type;
    using pointer = const value_type *;
    using iterator_category = std::random_access_iterator_tag;
    It it;
    F f;
    transform_iterator(It it_, F f_) : it(it_), f(f_) {}
    value_type operator*() const { return f(*it); }
    value_type operator[](difference_type i) const { return f(it[i]); }
    transform_iterator operator+(difference_type i) const { return transform_iterator(it + i, f); }
};
template <class It, class F> transform_iterator<F, It> make_transform_iterator(It it, F f) { return transform_iterator<F, It>(it, f); }
}  // namespace thrust
