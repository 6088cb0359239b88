// hash_sort.cu — "hash, then sort by digest": what AccountHashingStage / StorageHashingStage hand to the ETL
// collector (crates/stages/stages/src/stages/hashing_account.rs:192-230, crates/etl/src/lib.rs:31-60), done on
// the device: keccak of every key, then a radix sort of the 32-byte digests.
//
// Digests are keccak outputs, so their leading 64 bits are distinct with overwhelming probability: sort
// (prefix64, index) pairs with one 64-bit radix sort, then verify strict order of the full 32-byte keys.
// Only if the verification finds an unordered neighbour pair (adversarial / equal-prefix input) the keys are
// re-sorted by a stable LSD over all four 64-bit words.
#include <cub/cub.cuh>

#include <algorithm>

#include "engine.h"
#include "kernels.h"

using namespace b200;

namespace {

__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}

// keys64[i] = big-endian word `w` (0 = most significant) of the digest at perm[i] (or i when perm is null)
__global__ void extract_word_kernel(const uint64_t *__restrict__ digests, const uint32_t *__restrict__ perm, int w,
                                    uint64_t n, uint64_t *__restrict__ keys64, uint32_t *__restrict__ idx_out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t src = perm ? perm[i] : i;
    keys64[i] = bswap64(digests[4 * src + w]);
    if (idx_out) idx_out[i] = (uint32_t)src;
}

__global__ void gather32_kernel(const uint4 *__restrict__ digests, const uint32_t *__restrict__ perm, uint64_t n,
                                uint4 *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = perm[i];
    out[2 * i] = digests[2 * s];
    out[2 * i + 1] = digests[2 * s + 1];
}

// flag = 1 if any neighbour pair is not strictly ascending
__global__ void check_sorted_kernel(const uint64_t *__restrict__ sorted, uint64_t n, int *__restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 || i >= n) return;
    bool less = false, decided = false;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint64_t a = bswap64(sorted[4 * (i - 1) + w]), b = bswap64(sorted[4 * i + w]);
        if (!decided && a != b) {
            decided = true;
            less = a < b;
        }
    }
    if (!less) *flag = 1;  // equal keys also land here; harmless (the fallback is stable)
}

inline unsigned nblk(uint64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

// d_digests: [n][32] (input, unsorted) -> d_sorted [n][32], d_perm [n]
int32_t sort_digests_on_device(b200_ctx *c, const void *d_digests, uint64_t n, void *d_sorted, uint32_t *d_perm,
                               DevBuf &keys_a, DevBuf &keys_b, DevBuf &idx_a, DevBuf &flag) {
    if (n == 0) return B200_OK;
    if (n >= (1ull << 32)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^32-1 keys per sort");
    cudaStream_t st = c->stream;
    TRY(ensure(c, keys_a, n * 8));
    TRY(ensure(c, keys_b, n * 8));
    TRY(ensure(c, idx_a, n * 4));
    TRY(ensure(c, flag, 16));
    uint64_t *ka = static_cast<uint64_t *>(keys_a.p), *kb = static_cast<uint64_t *>(keys_b.p);
    uint32_t *ia = static_cast<uint32_t *>(idx_a.p);
    const uint64_t *dig = static_cast<const uint64_t *>(d_digests);
    size_t temp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, temp, ka, kb, ia, d_perm, (int64_t)n, 0, 64, st));
    TRY(ensure(c, c->cub_temp, temp));
    extract_word_kernel<<<nblk(n), 256, 0, st>>>(dig, nullptr, 0, n, ka, ia);
    CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, temp, ka, kb, ia, d_perm, (int64_t)n, 0, 64, st));
    gather32_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const uint4 *>(d_digests), d_perm, n,
                                             static_cast<uint4 *>(d_sorted));
    CU(cudaMemsetAsync(flag.p, 0, 4, st));
    check_sorted_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const uint64_t *>(d_sorted), n, static_cast<int *>(flag.p));
    c->launches += 4;
    int *h_flag = reinterpret_cast<int *>(static_cast<uint8_t *>(c->pinned_small) + 3072);
    CU(cudaMemcpyAsync(h_flag, flag.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (*h_flag == 0) return B200_OK;
    // fallback: stable LSD over the four words, least significant first
    uint32_t *cur = nullptr;  // identity
    uint32_t *bufs[2] = {ia, d_perm};
    int which = 0;
    for (int w = 3; w >= 0; w--) {
        // keys of word w in the current order, then a stable sort carrying the source index
        uint32_t *idx_in = bufs[which], *idx_out = bufs[which ^ 1];
        extract_word_kernel<<<nblk(n), 256, 0, st>>>(dig, cur, w, n, ka, idx_in);
        CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, temp, ka, kb, idx_in, idx_out, (int64_t)n, 0, 64, st));
        cur = idx_out;
        which ^= 1;
        c->launches += 2;
    }
    if (cur != d_perm) CU(cudaMemcpyAsync(d_perm, cur, n * 4, cudaMemcpyDeviceToDevice, st));
    gather32_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const uint4 *>(d_digests), d_perm, n,
                                             static_cast<uint4 *>(d_sorted));
    c->launches++;
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------- composite keys (a3)
// StorageHashingStage (crates/stages/stages/src/stages/hashing_storage.rs:106-178) sorts by the 64-byte key
// keccak(address) || keccak(slot).  Element i belongs to address addr_index[i]; the address digest is computed
// once per address (the reference caches it across consecutive entries, :129-134).
namespace {

// big-endian word w (0..7) of the composite key of element src
__device__ __forceinline__ uint64_t composite_word(const uint64_t *__restrict__ ha, const uint32_t *__restrict__ addr_index,
                                                   const uint64_t *__restrict__ hs, uint64_t src, int w) {
    return bswap64(w < 4 ? ha[4 * (uint64_t)addr_index[src] + w] : hs[4 * src + (w - 4)]);
}

__global__ void extract_composite_kernel(const uint64_t *__restrict__ ha, const uint32_t *__restrict__ addr_index,
                                         const uint64_t *__restrict__ hs, const uint32_t *__restrict__ perm, int w, uint64_t n,
                                         uint64_t *__restrict__ keys64, uint32_t *__restrict__ idx_out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t src = perm ? perm[i] : i;
    keys64[i] = composite_word(ha, addr_index, hs, src, w);
    idx_out[i] = (uint32_t)src;
}

__global__ void gather_composite_kernel(const uint4 *__restrict__ ha, const uint32_t *__restrict__ addr_index,
                                        const uint4 *__restrict__ hs, const uint32_t *__restrict__ perm, uint64_t n,
                                        uint4 *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = perm[i], a = addr_index[s];
    out[4 * i] = ha[2 * a];
    out[4 * i + 1] = ha[2 * a + 1];
    out[4 * i + 2] = hs[2 * s];
    out[4 * i + 3] = hs[2 * s + 1];
}

// neighbours must be strictly ascending (a duplicate (address, slot) pair is reported separately)
__global__ void check_sorted_composite_kernel(const uint64_t *__restrict__ sorted, uint64_t n, int *__restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 || i >= n) return;
    bool less = false, decided = false;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        uint64_t a = bswap64(sorted[8 * (i - 1) + w]), b = bswap64(sorted[8 * i + w]);
        if (!decided && a != b) {
            decided = true;
            less = a < b;
        }
    }
    if (!decided) atomicMax(flag, 2);       // equal composite keys
    else if (!less) atomicMax(flag, 1);     // out of order: the prefix passes were not enough
}

__global__ void check_index_kernel(const uint32_t *__restrict__ addr_index, uint64_t n, uint32_t n_addr, int *__restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && addr_index[i] >= n_addr) *flag = 3;
}

}  // namespace

// d_ha [n_addr][32], d_hs [n][32] digests; -> d_sorted [n][64], d_perm [n]
int32_t sort_composite_on_device(b200_ctx *c, const void *d_ha, uint32_t n_addr, const uint32_t *d_addr_index,
                                 const void *d_hs, uint64_t n, void *d_sorted, uint32_t *d_perm, DevBuf &keys_a,
                                 DevBuf &keys_b, DevBuf &idx_a, DevBuf &flag) {
    if (n == 0) return B200_OK;
    if (n >= (1ull << 32)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^32-1 entries per sort");
    cudaStream_t st = c->stream;
    TRY(ensure(c, keys_a, n * 8));
    TRY(ensure(c, keys_b, n * 8));
    TRY(ensure(c, idx_a, n * 4));
    TRY(ensure(c, flag, 16));
    uint64_t *ka = static_cast<uint64_t *>(keys_a.p), *kb = static_cast<uint64_t *>(keys_b.p);
    uint32_t *ia = static_cast<uint32_t *>(idx_a.p);
    const uint64_t *ha = static_cast<const uint64_t *>(d_ha), *hs = static_cast<const uint64_t *>(d_hs);
    size_t temp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, temp, ka, kb, ia, d_perm, (int64_t)n, 0, 64, st));
    TRY(ensure(c, c->cub_temp, temp));
    CU(cudaMemsetAsync(flag.p, 0, 4, st));
    check_index_kernel<<<nblk(n), 256, 0, st>>>(d_addr_index, n, n_addr, static_cast<int *>(flag.p));
    int *h_flag = reinterpret_cast<int *>(static_cast<uint8_t *>(c->pinned_small) + 3072);
    CU(cudaMemcpyAsync(h_flag, flag.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (*h_flag == 3) return fail(c, B200_ERR_INVALID_ARG, "addr_index entry out of range");
    auto lsd = [&](const int *words, int n_words) -> int32_t {
        uint32_t *cur = nullptr;
        uint32_t *bufs[2] = {ia, d_perm};
        int which = 0;
        for (int k = 0; k < n_words; k++) {
            uint32_t *idx_in = bufs[which], *idx_out = bufs[which ^ 1];
            extract_composite_kernel<<<nblk(n), 256, 0, st>>>(ha, d_addr_index, hs, cur, words[k], n, ka, idx_in);
            CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, temp, ka, kb, idx_in, idx_out, (int64_t)n, 0, 64, st));
            cur = idx_out;
            which ^= 1;
            c->launches += 2;
        }
        if (cur != d_perm) CU(cudaMemcpyAsync(d_perm, cur, n * 4, cudaMemcpyDeviceToDevice, st));
        gather_composite_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const uint4 *>(d_ha), d_addr_index,
                                                        static_cast<const uint4 *>(d_hs), d_perm, n,
                                                        static_cast<uint4 *>(d_sorted));
        CU(cudaMemsetAsync(flag.p, 0, 4, st));
        check_sorted_composite_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const uint64_t *>(d_sorted), n,
                                                              static_cast<int *>(flag.p));
        c->launches += 2;
        CU(cudaMemcpyAsync(h_flag, flag.p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return B200_OK;
    };
    const int fast[2] = {4, 0};  // slot-digest prefix, then (stable) address-digest prefix
    TRY(lsd(fast, 2));
    if (*h_flag == 1) {
        const int full[8] = {7, 6, 5, 4, 3, 2, 1, 0};
        TRY(lsd(full, 8));
    }
    if (*h_flag == 2) return fail(c, B200_ERR_UNSORTED, "duplicate (address, slot) pair");
    return B200_OK;
}
