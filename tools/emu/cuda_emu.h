// cuda_emu.h — a small CPU emulation of the CUDA features libb200trie.so uses.  DEVELOPMENT / TEST TOOLING ONLY.
//
// Purpose: run the *unmodified* CUDA sources of reth_b200/csrc on a machine without a GPU so that kernel logic can be
// checked bit-exact against the oracle (and under ASan/UBSan) before GPU minutes are spent.  It is never built into or
// loaded by the product: reth_b200/_lib.py only ever loads reth_b200/libb200trie.so, and that library has no CPU
// path.  tools/emu/translate.py rewrites `k<<<g,b,s,st>>>(args)` into EMU_LAUNCH(...) and `__shared__` into static
// storage; everything else is provided here by macros and inline functions.
//
// Execution model: one OS thread.  Blocks of a grid run one after the other; the threads of a block are ucontext
// fibers scheduled round-robin.  __syncthreads / __syncwarp / shuffles / warp reductions yield until every
// participating thread has arrived.  Atomics are plain operations.  Streams are synchronous.  What this does NOT
// model: real concurrency between blocks, memory-model races, performance.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <utility>

// ------------------------------------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __constant__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// ------------------------------------------------------------------------------------------------ vector types
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct alignas(8) ushort4 { unsigned short x, y, z, w; };
struct alignas(4) ushort2 { unsigned short x, y; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline ushort4 make_ushort4(unsigned short x, unsigned short y, unsigned short z, unsigned short w) { return ushort4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ------------------------------------------------------------------------------------------------ scheduler state
namespace emu {
struct State {
    uint3 tid{0, 0, 0}, bid{0, 0, 0};
    dim3 bdim, gdim;
};
extern State g;
void yield();                          // give the other fibers of the block a turn
void sync_block();                     // __syncthreads
void sync_warp(uint32_t mask);         // __syncwarp / barrier part of the warp collectives
uint64_t warp_exchange(uint32_t mask, uint64_t mine, int src_lane, bool *src_valid);  // generic shuffle
uint64_t warp_reduce_or(uint32_t mask, uint64_t mine);
void run_grid(unsigned grid, unsigned block, const std::function<void()> &thread_body);
void *dynamic_smem(size_t bytes);
extern size_t stat_launches, stat_threads;
}  // namespace emu

#define threadIdx (emu::g.tid)
#define blockIdx (emu::g.bid)
#define blockDim (emu::g.bdim)
#define gridDim (emu::g.gdim)
static const int warpSize = 32;

// ------------------------------------------------------------------------------------------------ intrinsics
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *(const volatile T *)p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
static inline uint4 __ldcg(const uint4 *p) { return *p; }
static inline uint2 __ldcg(const uint2 *p) { return *p; }

static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;  // sign replication mode
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (sh & 31)) >> 32);
}
static inline uint32_t __funnelshift_lc(uint32_t lo, uint32_t hi, uint32_t sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    sh = sh > 32 ? 32 : sh;
    return sh == 32 ? lo : (uint32_t)((v << sh) >> 32);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh & 31));
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    sh = sh > 32 ? 32 : sh;
    return (uint32_t)(v >> sh);
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline uint32_t __brev(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu::sync_warp(mask); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline uint32_t __activemask() { return 0xffffffffu; }

namespace emu {
template <class T> static inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> static inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
}  // namespace emu
template <class T> static inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
    int lane = emu::g.tid.x & 31;
    int base = lane & ~(width - 1);
    bool ok;
    uint64_t r = emu::warp_exchange(mask, emu::to_bits(v), base + (src & (width - 1)), &ok);
    return ok ? emu::from_bits<T>(r) : v;
}
template <class T> static inline T __shfl_xor_sync(uint32_t mask, T v, int lane_mask, int width = 32) {
    int lane = emu::g.tid.x & 31;
    int src = lane ^ lane_mask;
    bool ok;
    bool in_range = (src & ~(width - 1)) == (lane & ~(width - 1)) || width == 32;
    uint64_t r = emu::warp_exchange(mask, emu::to_bits(v), src & 31, &ok);
    return (ok && in_range) ? emu::from_bits<T>(r) : v;
}
template <class T> static inline T __shfl_up_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
    int lane = emu::g.tid.x & 31;
    int src = lane - (int)delta;
    bool ok;
    bool in_range = src >= (lane & ~(width - 1));
    uint64_t r = emu::warp_exchange(mask, emu::to_bits(v), in_range ? src : lane, &ok);
    return (ok && in_range) ? emu::from_bits<T>(r) : v;
}
template <class T> static inline T __shfl_down_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
    int lane = emu::g.tid.x & 31;
    int src = lane + (int)delta;
    bool ok;
    bool in_range = src < (lane & ~(width - 1)) + width;
    uint64_t r = emu::warp_exchange(mask, emu::to_bits(v), in_range ? src : lane, &ok);
    return (ok && in_range) ? emu::from_bits<T>(r) : v;
}
static inline uint32_t __reduce_or_sync(uint32_t mask, uint32_t v) { return (uint32_t)emu::warp_reduce_or(mask, v); }
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    return (uint32_t)emu::warp_reduce_or(mask, pred ? (1ull << (emu::g.tid.x & 31)) : 0);
}
static inline int __any_sync(uint32_t mask, int pred) { return emu::warp_reduce_or(mask, pred ? 1 : 0) != 0; }
static inline int __all_sync(uint32_t mask, int pred) { return emu::warp_reduce_or(mask, pred ? 0 : 1) == 0; }

// atomics: one OS thread, so plain read-modify-write
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicSub(T *p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T *p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// min / max as the CUDA headers provide them for mixed integer types
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { return a < b ? a : b; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------ runtime API
typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct emuStream *cudaStream_t;
typedef struct emuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum : unsigned { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
#define cudaStreamLegacy ((cudaStream_t)0x1)

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDevice(int *);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int dev);
cudaError_t cudaMalloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *);
cudaError_t cudaMallocHost(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMallocHost((void **)p, bytes); }
cudaError_t cudaFreeHost(void *);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *, unsigned, int);
cudaError_t cudaDeviceGetStreamPriorityRange(int *, int *);
cudaError_t cudaDeviceGetPCIBusId(char *, int, int);
cudaError_t cudaStreamCreate(cudaStream_t *);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) {
    *n = 1;
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------ launches
// translate.py turns  kernel<<<grid, block, smem, stream>>>(args...)  into  EMU_LAUNCH(kernel, grid, block, smem, stream, args...)
namespace emu {
template <class K, class... A> static inline void launch(K kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t, A... args) {
    dynamic_smem(smem);
    run_grid(grid, block, [&] { kernel(args...); });
}
}  // namespace emu
#define EMU_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(kernel, (unsigned)(grid), (unsigned)(block), (size_t)(smem), stream, ##__VA_ARGS__)
namespace emu { void *smem_base(); }
