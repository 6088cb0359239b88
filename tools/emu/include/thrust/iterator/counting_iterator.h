// minimal stand-in for thrust::counting_iterator (tools/emu)
#pragma once
#include <cstddef>
#include <iterator>
namespace thrust {
template <class T> struct counting_iterator {
    using value_type = T;
    using difference_type = std::ptrdiff_t;
    using reference = T;
    using pointer = const T *;
    using iterator_category = std::random_access_iterator_tag;
    T v;
    explicit counting_iterator(T v_ = T()) : v(v_) {}
    T operator*() const { return v; }
    T operator[](difference_type i) const { return (T)(v + i); }
    counting_iterator operator+(difference_type i) const { return counting_iterator((T)(v + i)); }
    counting_iterator &operator++() { ++v; return *this; }
};
}  // namespace thrust
