// minimal stand-ins for the four CUB device-wide primitives libb200trie.so uses (tools/emu): sequential, same results.
#pragma once
#include <algorithm>
#include <cstdint>
#include <iterator>
#include <numeric>
#include <vector>
#include "../../cuda_emu.h"
namespace cub {
struct DeviceRadixSort {
    // stable LSD semantics: order by bits [begin_bit, end_bit) of the key, ties keep input order
    template <class K, class V, class N>
    static cudaError_t SortPairs(void *temp, size_t &temp_bytes, const K *keys_in, K *keys_out, const V *vals_in, V *vals_out, N n,
                                 int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!temp) { temp_bytes = 1; return cudaSuccess; }
        size_t m = (size_t)n;
        std::vector<size_t> order(m);
        std::iota(order.begin(), order.end(), 0);
        const int bits = end_bit - begin_bit;
        auto field = [&](K k) -> uint64_t {
            uint64_t v = (uint64_t)k >> begin_bit;
            return bits >= 64 ? v : (v & ((1ull << bits) - 1));
        };
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return field(keys_in[a]) < field(keys_in[b]); });
        std::vector<K> ko(m);
        std::vector<V> vo(m);
        for (size_t i = 0; i < m; i++) { ko[i] = keys_in[order[i]]; vo[i] = vals_in[order[i]]; }
        std::copy(ko.begin(), ko.end(), keys_out);
        std::copy(vo.begin(), vo.end(), vals_out);
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <class In, class Out, class N>
    static cudaError_t ExclusiveSum(void *temp, size_t &temp_bytes, In in, Out out, N n, cudaStream_t = nullptr) {
        if (!temp) { temp_bytes = 1; return cudaSuccess; }
        using T = typename std::iterator_traits<Out>::value_type;
        T acc = T();
        for (size_t i = 0; i < (size_t)n; i++) { T v = (T)in[i]; out[i] = acc; acc = (T)(acc + v); }
        return cudaSuccess;
    }
    template <class In, class Out, class N>
    static cudaError_t InclusiveSum(void *temp, size_t &temp_bytes, In in, Out out, N n, cudaStream_t = nullptr) {
        if (!temp) { temp_bytes = 1; return cudaSuccess; }
        using T = typename std::iterator_traits<Out>::value_type;
        T acc = T();
        for (size_t i = 0; i < (size_t)n; i++) { acc = (T)(acc + (T)in[i]); out[i] = acc; }
        return cudaSuccess;
    }
};
struct DeviceSelect {
    template <class In, class Flags, class Out, class Cnt, class N>
    static cudaError_t Flagged(void *temp, size_t &temp_bytes, In in, Flags flags, Out out, Cnt n_selected, N n, cudaStream_t = nullptr) {
        if (!temp) { temp_bytes = 1; return cudaSuccess; }
        size_t k = 0;
        for (size_t i = 0; i < (size_t)n; i++)
            if (flags[i]) out[k++] = in[i];
        *n_selected = (typename std::remove_reference<decltype(*n_selected)>::type)k;
        return cudaSuccess;
    }
};
}  // namespace cub
