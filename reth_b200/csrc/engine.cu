// engine.cu — context, HBM scratch arena, build orchestration and the extern "C" B200_API boundary of
// libb200trie.so (include/b200trie.h).  No CPU fallback exists: every entry point needs a CUDA device.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "engine.h"
#include "kernels.h"
#include "trie_kernels.h"

using namespace b200;

#define B200_VERSION_STR "reth_b200 0.1.0 (sm_100a)"

static thread_local int g_create_status = 0;

// levels with at most this many nodes run one warp per node (shuffle-based Keccak): beyond ~3-4k nodes the 14x higher instruction count of the shuffle formulation outweighs its ~5x shorter latency
static constexpr uint32_t WARP_LEVEL_MAX = 4096;

// layout of the `small` device buffer (uint32 units)
enum : int { SM_BUCKET_OFF = 0, SM_LEVEL_LO = 80, SM_NNODES = 160, SM_ERR = 164, SM_NSTORED = 168, SM_COUNTERS = 176 /* 4 x u64 */, SM_HIST = 256 /* 256 x u32 */, SM_WORDS = 512 };

static uint32_t *small_u32(b200_ctx *c) { return static_cast<uint32_t *>(c->small.p); }

extern "C" B200_API int32_t b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" B200_API int32_t b200_create_status(void) { return g_create_status; }

extern "C" B200_API b200_ctx *b200_create(int32_t device_ordinal) {
    int n = b200_device_count();
    if (n <= 0 || device_ordinal < 0 || device_ordinal >= n) {
        g_create_status = n <= 0 ? B200_ERR_NO_DEVICE : B200_ERR_INVALID_ARG;
        return nullptr;
    }
    b200_ctx *c = new b200_ctx();
    c->device = device_ordinal;
    auto bail = [&](cudaError_t e) -> b200_ctx * {
        (void)e;
        g_create_status = B200_ERR_CUDA;
        delete c;
        return nullptr;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return bail(e);
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    c->stream = c->own_stream;
    for (int i = 0; i < 3; i++)
        if ((e = cudaStreamCreateWithFlags(&c->copy_streams[i], cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail(e);
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail(e);
    if ((e = cudaMallocHost(&c->pinned_small, 4096)) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&c->small.p, SM_WORDS * 4)) != cudaSuccess) return bail(e);
    c->small.cap = SM_WORDS * 4;
    c->dev_bytes += c->small.cap;
    if ((e = cudaMemset(c->small.p, 0, SM_WORDS * 4)) != cudaSuccess) return bail(e);
    g_create_status = B200_OK;
    return c;
}

extern "C" B200_API void b200_destroy(b200_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&c->Lp, &c->nibs, &c->leaf_ref, &c->leaf_meta, &c->S, &c->E, &c->iota, &c->depth_sorted,
                      &c->gap_sorted, &c->bound_rank, &c->head, &c->node_start, &c->node_ref, &c->node_meta,
                      &c->node_l, &c->node_r, &c->node_masks, &c->cub_temp, &c->small, &c->sroots, &c->buckets,
                      &c->upd_flags, &c->upd_nh, &c->upd_ids, &c->upd_prefix, &c->in_a, &c->in_b, &c->in_c,
                      &c->in_d, &c->in_e, &c->out_a, &c->chunk_in[0], &c->chunk_in[1], &c->chunk_in[2], &c->chunk_out[0],
                      &c->chunk_out[1], &c->chunk_out[2], &c->sort_ka, &c->sort_kb, &c->sort_ia, &c->sort_flag, &c->sort_perm,
                      &c->sort_out, &c->node_key, &c->node_key2, &c->node_ids, &c->node_order};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->pinned_small) cudaFreeHost(c->pinned_small);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->chunk_events) cudaEventDestroy(e);
    for (int i = 0; i < 3; i++)
        if (c->copy_streams[i]) cudaStreamDestroy(c->copy_streams[i]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" B200_API const char *b200_last_error(const b200_ctx *c) { return c ? c->err.c_str() : "null context"; }
extern "C" B200_API const char *b200_version(void) { return B200_VERSION_STR; }
extern "C" B200_API uint64_t b200_device_bytes(const b200_ctx *c) { return c ? c->dev_bytes : 0; }

extern "C" B200_API int32_t b200_set_stream(b200_ctx *c, void *cuda_stream) {
    if (!c) return B200_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(c->mu);
    // 0 is the CUDA legacy default stream; "no stream given" is spelled b200_set_stream(ctx, (void*)-1)
    c->stream = cuda_stream == reinterpret_cast<void *>(-1) ? c->own_stream
                : cuda_stream == nullptr                    ? cudaStreamLegacy
                                                            : static_cast<cudaStream_t>(cuda_stream);
    return B200_OK;
}

extern "C" B200_API void *b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" B200_API void b200_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

static int32_t map_dev_error(b200_ctx *c, int code) {
    switch (code) {
        case B200_DEVERR_NONE: return B200_OK;
        case B200_DEVERR_UNSORTED: return fail(c, B200_ERR_UNSORTED, "keys are not strictly ascending inside a trie");
        case B200_DEVERR_ZERO_VALUE: return fail(c, B200_ERR_ZERO_VALUE, "storage slot with zero value (zero means deleted)");
        case B200_DEVERR_INLINE_HASH_CHILD:
            return fail(c, B200_ERR_INLINE_HASH_CHILD, "inline (<32 byte) branch child under a hash_mask bit");
        case B200_DEVERR_BAD_OFFSETS: return fail(c, B200_ERR_INVALID_ARG, "seg_offsets must start at 0, end at n and be monotone");
        case B200_DEVERR_NOT_FOUND: return fail(c, B200_ERR_NOT_FOUND, "key not found in the resident trie");
        default: return fail(c, B200_ERR_CUDA, "unknown device error %d", code);
    }
}

// waits for the stream, folds the timing / counters of the last build into stats, returns the sticky status
static int32_t sync_and_status(b200_ctx *c) {
    CU(cudaStreamSynchronize(c->stream));
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    CU(cudaMemcpyAsync(ps, small_u32(c) + SM_ERR, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(ps + 8, small_u32(c) + SM_COUNTERS, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(ps + 4, small_u32(c) + SM_NSTORED, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->stats_pending) {
        if (c->stats_wavefront) c->stats.branches_added = ps[4];
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->stats.device_ms = ms;
        else cudaGetLastError();
        const unsigned long long *cnt = reinterpret_cast<const unsigned long long *>(ps + 8);
        c->stats.hashed_nodes = cnt[CNT_HASHED];
        c->stats.extension_nodes = cnt[CNT_EXT];
        c->stats_pending = false;
    }
    return map_dev_error(c, (int)ps[0]);
}

static int32_t reset_build_state(b200_ctx *c) {
    CU(cudaMemsetAsync(small_u32(c) + SM_ERR, 0, 4, c->stream));
    CU(cudaMemsetAsync(small_u32(c) + SM_COUNTERS, 0, 32, c->stream));
    c->stats = b200_stats{};
    c->stats_wavefront = false;
    CU(cudaEventRecord(c->ev0, c->stream));
    return B200_OK;
}
static int32_t finish_build_state(b200_ctx *c) {
    CU(cudaEventRecord(c->ev1, c->stream));
    c->stats_pending = true;
    return B200_OK;
}

extern "C" B200_API int32_t b200_sync(b200_ctx *c) {
    if (!c) return B200_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    return sync_and_status(c);
}
extern "C" B200_API int32_t b200_dev_status(b200_ctx *c) { return b200_sync(c); }
extern "C" B200_API int32_t b200_last_stats(b200_ctx *c, b200_stats *out) {
    if (!c || !out) return B200_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    int32_t r = sync_and_status(c);
    *out = c->stats;
    return r;
}
// number of kernel launches issued through this context (bench.py reports it as gpu_launches)
extern "C" B200_API uint64_t b200_launch_count(const b200_ctx *c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------------------------------------ keccak
extern "C" B200_API int32_t b200_keccak256_fixed_dev(b200_ctx *c, const void *d_in, uint32_t msg_len, uint32_t stride,
                                            uint64_t n, void *d_out32) {
    if (!c || (n && (!d_in || !d_out32)) || stride < msg_len) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(launch_keccak256_fixed(d_in, msg_len, stride, n, d_out32, c->stream, &c->launches));
    return B200_OK;
}

extern "C" B200_API int32_t b200_keccak256_var_dev(b200_ctx *c, const void *d_data, const void *d_offsets, uint64_t n,
                                          void *d_out32) {
    if (!c || (n && (!d_offsets || !d_out32))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(launch_keccak256_var(d_data, d_offsets, n, d_out32, c->stream, &c->launches));
    return B200_OK;
}

// Host buffers: chunked over three streams so that the H2D copy of chunk k+2, the hashing of chunk k+1 and the D2H
// copy of chunk k overlap (both DMA directions stay busy; fully asynchronous when the caller's buffers are
// page-locked, see b200_host_alloc).
extern "C" B200_API int32_t b200_keccak256_fixed(b200_ctx *c, const uint8_t *in, uint32_t msg_len, uint32_t stride, uint64_t n,
                                        uint8_t *out32) {
    if (!c || (n && (!in || !out32)) || stride < msg_len) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return B200_OK;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    static const uint64_t CHUNK = [] {  // messages per chunk (B200_KECCAK_CHUNK overrides, for tuning)
        const char *e = getenv("B200_KECCAK_CHUNK");
        uint64_t v = e ? strtoull(e, nullptr, 10) : 0;
        return v >= 1024 ? v : (1ull << 19);
    }();
    uint64_t chunk = n < CHUNK ? n : CHUNK;
    for (int i = 0; i < 3; i++) {
        TRY(ensure(c, c->chunk_in[i], chunk * stride));
        TRY(ensure(c, c->chunk_out[i], chunk * 32));
    }
    int slot = 0;
    for (uint64_t lo = 0; lo < n; lo += chunk, slot = (slot + 1) % 3) {
        uint64_t m = n - lo < chunk ? n - lo : chunk;
        cudaStream_t st = c->copy_streams[slot];
        size_t in_bytes = (m - 1) * (size_t)stride + msg_len;
        CU(cudaMemcpyAsync(c->chunk_in[slot].p, in + lo * stride, in_bytes, cudaMemcpyHostToDevice, st));
        CU(launch_keccak256_fixed(c->chunk_in[slot].p, msg_len, stride, m, c->chunk_out[slot].p, st, &c->launches));
        CU(cudaMemcpyAsync(out32 + lo * 32, c->chunk_out[slot].p, m * 32, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < 3; i++) CU(cudaStreamSynchronize(c->copy_streams[i]));
    return B200_OK;
}

extern "C" B200_API int32_t b200_keccak256_var(b200_ctx *c, const uint8_t *data, const uint64_t *offsets, uint64_t n,
                                      uint8_t *out32) {
    if (!c || (n && (!offsets || !out32))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return B200_OK;
    for (uint64_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(c, B200_ERR_INVALID_ARG, "offsets must be monotone");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    uint64_t base = offsets[0], total = offsets[n] - base;
    if (total && !data) return fail(c, B200_ERR_INVALID_ARG, "data is null");
    ENSURE(in_a, total ? total : 1);
    ENSURE(in_b, (n + 1) * 8);
    ENSURE(out_a, n * 32);
    std::vector<uint64_t> rel;
    const uint64_t *offs = offsets;
    if (base) {
        rel.resize(n + 1);
        for (uint64_t i = 0; i <= n; i++) rel[i] = offsets[i] - base;
        offs = rel.data();
    }
    if (total) CU(cudaMemcpyAsync(c->in_a.p, data + base, total, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_b.p, offs, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // `rel` must outlive the copy
    CU(launch_keccak256_var(c->in_a.p, c->in_b.p, n, c->out_a.p, c->stream, &c->launches));
    CU(cudaMemcpyAsync(out32, c->out_a.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ hash + sort
int32_t sort_digests_on_device(b200_ctx *c, const void *d_digests, uint64_t n, void *d_sorted, uint32_t *d_perm,
                               DevBuf &keys_a, DevBuf &keys_b, DevBuf &idx_a, DevBuf &flag);

// Device-resident: d_in -> d_sorted32 (n x 32), d_perm (n x u32).  Synchronises once (tie check).
extern "C" B200_API int32_t b200_hash_sort_keys_dev(b200_ctx *c, const void *d_in, uint32_t msg_len, uint32_t stride,
                                           uint64_t n, void *d_sorted32, void *d_perm) {
    if (!c || (n && (!d_in || !d_sorted32 || !d_perm)) || stride < msg_len)
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ENSURE(out_a, (n ? n : 1) * 32);
    CU(launch_keccak256_fixed(d_in, msg_len, stride, n, c->out_a.p, c->stream, &c->launches));
    return sort_digests_on_device(c, c->out_a.p, n, d_sorted32, static_cast<uint32_t *>(d_perm), c->sort_ka,
                                  c->sort_kb, c->sort_ia, c->sort_flag);
}

// Sorts 32-byte keys that are already digests (no hashing): the ETL-replacement half on its own.
extern "C" B200_API int32_t b200_sort_keys32_dev(b200_ctx *c, const void *d_keys32, uint64_t n, void *d_sorted32,
                                        void *d_perm) {
    if (!c || (n && (!d_keys32 || !d_sorted32 || !d_perm))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    return sort_digests_on_device(c, d_keys32, n, d_sorted32, static_cast<uint32_t *>(d_perm), c->sort_ka,
                                  c->sort_kb, c->sort_ia, c->sort_flag);
}

extern "C" B200_API int32_t b200_hash_sort_keys(b200_ctx *c, const uint8_t *in, uint32_t msg_len, uint32_t stride, uint64_t n,
                                       uint8_t *out_sorted32, uint32_t *out_perm) {
    if (!c || (n && (!in || !out_sorted32 || !out_perm)) || stride < msg_len)
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return B200_OK;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    size_t in_bytes = (n - 1) * (size_t)stride + msg_len;
    ENSURE(in_a, in_bytes);
    ENSURE(out_a, n * 32);
    ENSURE(sort_out, n * 32);
    ENSURE(sort_perm, n * 4);
    CU(cudaMemcpyAsync(c->in_a.p, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
    CU(launch_keccak256_fixed(c->in_a.p, msg_len, stride, n, c->out_a.p, c->stream, &c->launches));
    TRY(sort_digests_on_device(c, c->out_a.p, n, c->sort_out.p, static_cast<uint32_t *>(c->sort_perm.p), c->sort_ka,
                               c->sort_kb, c->sort_ia, c->sort_flag));
    CU(cudaMemcpyAsync(out_sorted32, c->sort_out.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(out_perm, c->sort_perm.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

int32_t sort_composite_on_device(b200_ctx *c, const void *d_ha, uint32_t n_addr, const uint32_t *d_addr_index,
                                 const void *d_hs, uint64_t n, void *d_sorted, uint32_t *d_perm, DevBuf &keys_a,
                                 DevBuf &keys_b, DevBuf &idx_a, DevBuf &flag);

// StorageHashingStage full pass: hash n_addr addresses once, n slot keys, sort entries by keccak(address) || keccak(slot).
extern "C" B200_API int32_t b200_hash_sort_storage(b200_ctx *c, const uint8_t *addresses20, uint32_t n_addr,
                                                   const uint32_t *addr_index, const uint8_t *slots32, uint64_t n,
                                                   uint8_t *out_sorted64, uint32_t *out_perm) {
    if (!c || (n && (!addresses20 || !addr_index || !slots32 || !out_sorted64 || !out_perm)) || (n && !n_addr))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return B200_OK;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ENSURE(in_a, (size_t)n_addr * 20);
    ENSURE(in_b, n * 32);
    ENSURE(in_c, n * 4);
    ENSURE(in_d, (size_t)n_addr * 32);  // address digests
    ENSURE(out_a, n * 32);              // slot digests
    ENSURE(sort_out, n * 64);
    ENSURE(sort_perm, n * 4);
    CU(cudaMemcpyAsync(c->in_a.p, addresses20, (size_t)n_addr * 20, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_b.p, slots32, n * 32, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_c.p, addr_index, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(launch_keccak256_fixed(c->in_a.p, 20, 20, n_addr, c->in_d.p, c->stream, &c->launches));
    CU(launch_keccak256_fixed(c->in_b.p, 32, 32, n, c->out_a.p, c->stream, &c->launches));
    TRY(sort_composite_on_device(c, c->in_d.p, n_addr, static_cast<const uint32_t *>(c->in_c.p), c->out_a.p, n,
                                 c->sort_out.p, static_cast<uint32_t *>(c->sort_perm.p), c->sort_ka, c->sort_kb,
                                 c->sort_ia, c->sort_flag));
    CU(cudaMemcpyAsync(out_sorted64, c->sort_out.p, n * 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(out_perm, c->sort_perm.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ forest build
struct IsBoundary {
    __host__ __device__ uint32_t operator()(uint8_t v) const { return v == 0xFF ? 1u : 0u; }
};

struct Built {
    ForestDev f{};
    uint32_t n_nodes = 0;
    uint32_t levels = 0;
    uint32_t level_count[64] = {};  // branch nodes per depth
};

// Builds every trie of a forest over d_keys (n leaves).  d_seg_offsets == nullptr: one trie.
// account: leaves are accounts (d_values = b200_account[n], d_sroots = storage roots or null); else storage
// slots (d_values = U256 BE [n][32]).
static int32_t build_forest(b200_ctx *c, const uint8_t *d_keys, uint64_t n, const uint64_t *d_seg_offsets,
                            uint64_t n_segs, bool account, const uint8_t *d_values, const uint8_t *d_sroots,
                            bool retain_updates, Built &out) {
    if (n >= (1ull << 31)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^31-1 leaves per build");
    cudaStream_t st = c->stream;
    ForestDev &f = out.f;
    f.n = n;
    f.keys = d_keys;
    f.err = reinterpret_cast<int *>(small_u32(c) + SM_ERR);
    f.counters = reinterpret_cast<unsigned long long *>(small_u32(c) + SM_COUNTERS);
    f.retain_updates = retain_updates ? 1 : 0;
    out.n_nodes = 0;
    out.levels = 0;
    if (n == 0) return B200_OK;

    ENSURE(Lp, n + 1);
    ENSURE(nibs, n + 1);
    ENSURE(leaf_ref, n * 32);
    ENSURE(leaf_meta, n);
    ENSURE(S, n * 4);
    ENSURE(E, n * 4);
    f.Lp = static_cast<uint8_t *>(c->Lp.p);
    f.nibs = static_cast<uint8_t *>(c->nibs.p);
    f.leaf_ref = static_cast<uint8_t *>(c->leaf_ref.p);
    f.leaf_meta = static_cast<uint8_t *>(c->leaf_meta.p);
    f.S = static_cast<uint32_t *>(c->S.p);
    f.E = static_cast<uint32_t *>(c->E.p);

    CU(cudaMemsetAsync(f.Lp, 0, n + 1, st));
    if (d_seg_offsets) {
        CU(launch_mark_boundaries(d_seg_offsets, n_segs, n, f.Lp, f.err, st));
        c->launches++;
    }
    CU(launch_lcp(d_keys, n, f.Lp, f.nibs, f.err, st));
    CU(launch_leaves(f, account, d_values, d_sroots, st));
    c->launches += 2;
    if (n < 2) return B200_OK;

    // ---- gaps sorted by depth (stable: position order inside a depth) -> branch nodes in CSR form
    const uint64_t G = n - 1;
    ENSURE(iota, G * 4);
    ENSURE(depth_sorted, G);
    ENSURE(gap_sorted, G * 4);
    ENSURE(head, G);
    ENSURE(node_start, (G + 1) * 4);
    uint32_t *bucket_off = small_u32(c) + SM_BUCKET_OFF;
    uint32_t *level_lo = small_u32(c) + SM_LEVEL_LO;
    uint32_t *n_nodes_p = small_u32(c) + SM_NNODES;
    uint8_t *depth_sorted = static_cast<uint8_t *>(c->depth_sorted.p);
    uint32_t *gap_sorted = static_cast<uint32_t *>(c->gap_sorted.p);
    uint8_t *head = static_cast<uint8_t *>(c->head.p);
    uint32_t *node_start = static_cast<uint32_t *>(c->node_start.p);

    CU(launch_iota(static_cast<uint32_t *>(c->iota.p), G, 1, st));
    c->launches++;
    size_t t_sort = 0, t_sel = 0, t_scan = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, t_sort, f.Lp + 1, depth_sorted, static_cast<uint32_t *>(c->iota.p),
                                       gap_sorted, (int64_t)G, 0, 8, st));
    thrust::counting_iterator<uint32_t> counting(0);
    CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, head, node_start, n_nodes_p, (int64_t)G, st));
    auto bflags = thrust::make_transform_iterator(static_cast<const uint8_t *>(f.Lp), IsBoundary());
    uint32_t *bound_rank = nullptr;
    if (d_seg_offsets) {
        ENSURE(bound_rank, (n + 1) * 4);
        bound_rank = static_cast<uint32_t *>(c->bound_rank.p);
        CU(cub::DeviceScan::InclusiveSum(nullptr, t_scan, bflags, bound_rank, (int64_t)(n + 1), st));
    }
    size_t t_max = std::max(t_sort, std::max(t_sel, t_scan));
    ENSURE(cub_temp, t_max);
    CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, t_sort, f.Lp + 1, depth_sorted,
                                       static_cast<uint32_t *>(c->iota.p), gap_sorted, (int64_t)G, 0, 8, st));
    CU(launch_bucket_offsets(depth_sorted, G, bucket_off, st));
    if (d_seg_offsets)
        CU(cub::DeviceScan::InclusiveSum(c->cub_temp.p, t_scan, bflags, bound_rank, (int64_t)(n + 1), st));
    CU(cudaMemsetAsync(head, 0, G, st));
    CU(launch_head_flags(d_keys, depth_sorted, gap_sorted, bound_rank, bucket_off + 64, G, head, st));
    CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, head, node_start, n_nodes_p, (int64_t)G, st));
    CU(launch_level_ranges(node_start, n_nodes_p, bucket_off, level_lo, st));
    // (depth, child-count class) of every node + histogram, still without knowing the node count on the host
    ENSURE(node_key, G);
    ENSURE(node_ids, G * 4);
    uint8_t *nk = static_cast<uint8_t *>(c->node_key.p);
    uint32_t *nids = static_cast<uint32_t *>(c->node_ids.p);
    uint32_t *hist = small_u32(c) + SM_HIST;
    CU(cudaMemsetAsync(hist, 0, 256 * 4, st));
    CU(launch_node_class_keys(node_start, depth_sorted, n_nodes_p, G, nk, nids, hist, st));
    c->launches += 9;
    uint32_t *h_level = static_cast<uint32_t *>(c->pinned_small) + 64;
    uint32_t *h_hist = static_cast<uint32_t *>(c->pinned_small) + 256;
    CU(cudaMemcpyAsync(h_level, level_lo, 66 * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_hist, hist, 256 * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));  // the only host round trip of a build: 322 integers
    const uint32_t B = h_level[65];
    out.n_nodes = B;
    f.gap_sorted = gap_sorted;
    f.node_start = node_start;
    if (B == 0) return B200_OK;  // every trie has at most one leaf

    ENSURE(node_ref, (size_t)B * 32);
    ENSURE(node_meta, B);
    ENSURE(node_l, (size_t)B * 4);
    ENSURE(node_r, (size_t)B * 4);
    ENSURE(node_masks, (size_t)B * 8);
    f.node_ref = static_cast<uint8_t *>(c->node_ref.p);
    f.node_meta = static_cast<uint8_t *>(c->node_meta.p);
    f.node_l = static_cast<uint32_t *>(c->node_l.p);
    f.node_r = static_cast<uint32_t *>(c->node_r.p);
    f.node_masks = static_cast<ushort4 *>(c->node_masks.p);

    // ---- node visiting order: (depth descending, child-count class); ids stay what they are
    ENSURE(node_key2, B);
    ENSURE(node_order, (size_t)B * 4);
    uint8_t *nk2 = static_cast<uint8_t *>(c->node_key2.p);
    uint32_t *norder = static_cast<uint32_t *>(c->node_order.p);
    size_t t_ns = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, t_ns, nk, nk2, nids, norder, (int64_t)B, 0, 8, st));
    ENSURE(cub_temp, t_ns);
    CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, t_ns, nk, nk2, nids, norder, (int64_t)B, 0, 8, st));
    c->launches += 1;

    // ---- deepest level first; the per-level frontier stays in HBM.  Big levels get one launch per child-count
    // class (strip size and unrolling fit the class), small ones a single launch.
    uint32_t pos = 0;
    for (int d = 63; d >= 0; d--) {
        const uint32_t *hc = h_hist + 4 * (63 - d);
        uint32_t cnt = hc[0] + hc[1] + hc[2] + hc[3];
        if (!cnt) continue;
        out.levels++;
        out.level_count[d] = cnt;
        if (cnt <= WARP_LEVEL_MAX) {  // about one wave of warps: latency-bound, one warp per node
            CU(launch_branch_level(f, norder, pos, pos + cnt, d, -1, st));
            c->launches++;
            pos += cnt;
        } else {
            for (int cls = 0; cls < 4; cls++) {
                if (!hc[cls]) continue;
                // a sparsely populated class of a big level is latency-bound too: one warp per node
                CU(launch_branch_level(f, norder, pos, pos + hc[cls], d, hc[cls] <= WARP_LEVEL_MAX / 4 ? -1 : cls, st));
                c->launches++;
                pos += hc[cls];
            }
        }
    }
    if (pos != B) return fail(c, B200_ERR_CUDA, "internal: level histogram (%u) != node count (%u)", pos, B);
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ updates
struct UpdatesOwner {
    void *host = nullptr;  // one page-locked block holding every array
};

extern "C" B200_API void b200_updates_release(b200_updates *u) {
    if (!u) return;
    if (u->_owner) {
        UpdatesOwner *o = static_cast<UpdatesOwner *>(u->_owner);
        if (o->host) cudaFreeHost(o->host);
        delete o;
    }
    memset(u, 0, sizeof *u);
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Gathers the records of `n_stored` stored nodes (ids on the device) into `u` (host, page-locked).
static int32_t gather_and_copy(b200_ctx *c, const ForestDev &f, const uint32_t *d_stored_ids, uint32_t n_stored,
                               uint32_t n_hashes, const uint32_t *d_prefix_by_node, const uint32_t *d_prefix_by_record,
                               const uint64_t *d_seg_offsets, uint64_t n_segs, b200_updates *u, UpdatesOwner *owner) {
    cudaStream_t st = c->stream;
    // one device block + one pinned host block, same layout
    size_t o_tid = 0;
    size_t o_plen = align_up(o_tid + (size_t)n_stored * 4, 16);
    size_t o_path = align_up(o_plen + n_stored, 16);
    size_t o_sm = align_up(o_path + (size_t)n_stored * 32, 16);
    size_t o_tm = align_up(o_sm + (size_t)n_stored * 2, 16);
    size_t o_hm = align_up(o_tm + (size_t)n_stored * 2, 16);
    size_t o_ho32 = align_up(o_hm + (size_t)n_stored * 2, 16);
    size_t o_hash = align_up(o_ho32 + (size_t)n_stored * 4, 16);
    size_t o_ho64 = align_up(o_hash + (size_t)n_hashes * 32, 16);
    size_t dev_total = o_ho64;
    size_t host_total = o_ho64 + ((size_t)n_stored + 1) * 8;
    CU(cudaMallocHost(&owner->host, host_total ? host_total : 16));
    uint8_t *h = static_cast<uint8_t *>(owner->host);
    u->n_nodes = n_stored;
    u->trie_id = reinterpret_cast<uint32_t *>(h + o_tid);
    u->path_len = h + o_plen;
    u->path_packed = h + o_path;
    u->state_mask = reinterpret_cast<uint16_t *>(h + o_sm);
    u->tree_mask = reinterpret_cast<uint16_t *>(h + o_tm);
    u->hash_mask = reinterpret_cast<uint16_t *>(h + o_hm);
    u->hashes = h + o_hash;
    u->hash_offset = reinterpret_cast<uint64_t *>(h + o_ho64);
    if (n_stored) {
        ENSURE(out_a, dev_total);
        uint8_t *d = static_cast<uint8_t *>(c->out_a.p);
        UpdatesDev ud;
        ud.trie_id = reinterpret_cast<uint32_t *>(d + o_tid);
        ud.path_len = d + o_plen;
        ud.path_packed = d + o_path;
        ud.state_mask = reinterpret_cast<uint16_t *>(d + o_sm);
        ud.tree_mask = reinterpret_cast<uint16_t *>(d + o_tm);
        ud.hash_mask = reinterpret_cast<uint16_t *>(d + o_hm);
        ud.hash_offset = reinterpret_cast<uint32_t *>(d + o_ho32);
        ud.hashes = d + o_hash;
        CU(launch_gather_updates(f, d_stored_ids, n_stored, d_prefix_by_node, d_prefix_by_record, d_seg_offsets, n_segs,
                                 ud, st));
        c->launches++;
        CU(cudaMemcpyAsync(h, d, dev_total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const uint32_t *ho32 = reinterpret_cast<const uint32_t *>(h + o_ho32);
        for (uint32_t i = 0; i < n_stored; i++) u->hash_offset[i] = ho32[i];
    }
    u->hash_offset[n_stored] = n_hashes;
    return B200_OK;
}

// Collects the stored BranchNodeCompact records of a finished build into `u` (host, page-locked).
static int32_t collect_updates(b200_ctx *c, const Built &b, const uint64_t *d_seg_offsets, uint64_t n_segs,
                               b200_updates *u) {
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    cudaStream_t st = c->stream;
    const uint32_t B = b.n_nodes;
    uint32_t n_stored = 0, n_hashes = 0;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    if (B) {
        ENSURE(upd_flags, B);
        ENSURE(upd_nh, (size_t)B * 4);
        ENSURE(upd_ids, (size_t)B * 4);
        ENSURE(upd_prefix, (size_t)(B + 1) * 4);
        uint8_t *flags = static_cast<uint8_t *>(c->upd_flags.p);
        uint32_t *nh = static_cast<uint32_t *>(c->upd_nh.p);
        uint32_t *ids = static_cast<uint32_t *>(c->upd_ids.p);
        uint32_t *prefix = static_cast<uint32_t *>(c->upd_prefix.p);
        uint32_t *n_stored_p = small_u32(c) + SM_NSTORED;
        CU(launch_stored_flags(b.f, B, flags, nh, st));
        size_t t_sel = 0, t_scan = 0;
        thrust::counting_iterator<uint32_t> counting(0);
        CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, flags, ids, n_stored_p, (int64_t)B, st));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, nh, prefix, (int64_t)B, st));
        ENSURE(cub_temp, std::max(t_sel, t_scan));
        CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, flags, ids, n_stored_p, (int64_t)B, st));
        CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t_scan, nh, prefix, (int64_t)B, st));
        c->launches += 3;
        CU(cudaMemcpyAsync(ps + 200, n_stored_p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 201, prefix + (B - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 202, nh + (B - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        n_stored = ps[200];
        n_hashes = ps[201] + ps[202];
    }
    return gather_and_copy(c, b.f, static_cast<uint32_t *>(c->upd_ids.p), n_stored, n_hashes,
                           static_cast<uint32_t *>(c->upd_prefix.p), nullptr, d_seg_offsets, n_segs, u, owner);
}

// Same for a subset of nodes given by id (the dirty nodes of an incremental update), in list order.
static int32_t collect_updates_subset(b200_ctx *c, const ForestDev &f, const uint32_t *d_ids, uint32_t count,
                                      b200_updates *u) {
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    cudaStream_t st = c->stream;
    uint32_t n_stored = 0, n_hashes = 0;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    if (count) {
        ENSURE(upd_flags, count);
        ENSURE(upd_nh, (size_t)count * 4);
        ENSURE(upd_ids, (size_t)count * 4 * 3);  // selected positions | picked ids | picked prefixes
        ENSURE(upd_prefix, (size_t)(count + 1) * 4);
        uint8_t *flags = static_cast<uint8_t *>(c->upd_flags.p);
        uint32_t *nh = static_cast<uint32_t *>(c->upd_nh.p);
        uint32_t *sel = static_cast<uint32_t *>(c->upd_ids.p), *pick_ids = sel + count, *pick_prefix = sel + 2 * (size_t)count;
        uint32_t *prefix = static_cast<uint32_t *>(c->upd_prefix.p);
        uint32_t *n_stored_p = small_u32(c) + SM_NSTORED;
        CU(launch_stored_flags_subset(f, d_ids, count, flags, nh, st));
        size_t t_sel = 0, t_scan = 0;
        thrust::counting_iterator<uint32_t> counting(0);
        CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, flags, sel, n_stored_p, (int64_t)count, st));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, nh, prefix, (int64_t)count, st));
        ENSURE(cub_temp, std::max(t_sel, t_scan));
        CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, flags, sel, n_stored_p, (int64_t)count, st));
        CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t_scan, nh, prefix, (int64_t)count, st));
        c->launches += 3;
        CU(cudaMemcpyAsync(ps + 200, n_stored_p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 201, prefix + (count - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 202, nh + (count - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        n_stored = ps[200];
        n_hashes = ps[201] + ps[202];
        CU(launch_pick_subset(d_ids, prefix, sel, n_stored, pick_ids, pick_prefix, st));
        c->launches++;
        return gather_and_copy(c, f, pick_ids, n_stored, n_hashes, nullptr, pick_prefix, nullptr, 0, u, owner);
    }
    return gather_and_copy(c, f, nullptr, 0, 0, nullptr, nullptr, nullptr, 0, u, owner);
}

// ------------------------------------------------------------------------------------------------ device-level drivers
static int32_t storage_roots_on_device(b200_ctx *c, const uint8_t *d_keys, const uint8_t *d_vals,
                                       const uint64_t *d_offs, uint64_t n_accounts, uint64_t n_slots,
                                       uint8_t *d_roots, bool retain, Built &b) {
    TRY(build_forest(c, d_keys, n_slots, d_offs, n_accounts, false, d_vals, nullptr, retain, b));
    CU(launch_segment_roots(b.f, d_offs, n_accounts, d_roots, c->stream));
    c->launches++;
    c->stats.leaves_added += n_slots;
    c->stats.branches_added += b.n_nodes;
    c->stats.levels += b.levels;
    return B200_OK;
}

static int32_t account_root_on_device(b200_ctx *c, const uint8_t *d_keys, const uint8_t *d_accts,
                                      const uint8_t *d_sroots, uint64_t n, uint8_t *d_root, bool retain, Built &b) {
    TRY(build_forest(c, d_keys, n, nullptr, 0, true, d_accts, d_sroots, retain, b));
    CU(launch_segment_roots(b.f, nullptr, 1, d_root, c->stream));
    c->launches++;
    c->stats.leaves_added += n;
    c->stats.branches_added += b.n_nodes;
    c->stats.levels += b.levels;
    return B200_OK;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" B200_API int32_t b200_storage_roots_dev(b200_ctx *c, const void *d_slot_keys32, const void *d_values32_be,
                                          const void *d_seg_offsets, uint64_t n_accounts, uint64_t n_slots,
                                          void *d_roots32) {
    if (!c || !d_seg_offsets || (n_accounts && !d_roots32) || (n_slots && (!d_slot_keys32 || !d_values32_be)))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (!aligned16(d_slot_keys32) || !aligned16(d_values32_be) || !aligned16(d_roots32))
        return fail(c, B200_ERR_INVALID_ARG, "device buffers must be 16-byte aligned");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(reset_build_state(c));
    Built b;
    TRY(storage_roots_on_device(c, static_cast<const uint8_t *>(d_slot_keys32),
                                static_cast<const uint8_t *>(d_values32_be),
                                static_cast<const uint64_t *>(d_seg_offsets), n_accounts, n_slots,
                                static_cast<uint8_t *>(d_roots32), false, b));
    return finish_build_state(c);
}

extern "C" B200_API int32_t b200_state_root_dev(b200_ctx *c, const void *d_acct_keys32, const void *d_accts,
                                       const void *d_storage_roots32, uint64_t n, void *d_root32) {
    if (!c || !d_root32 || (n && (!d_acct_keys32 || !d_accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (!aligned16(d_acct_keys32) || !aligned16(d_root32) || !aligned16(d_storage_roots32) ||
        (reinterpret_cast<uintptr_t>(d_accts) & 7))
        return fail(c, B200_ERR_INVALID_ARG, "device buffers must be 16-byte aligned (accounts: 8)");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(reset_build_state(c));
    Built b;
    TRY(account_root_on_device(c, static_cast<const uint8_t *>(d_acct_keys32), static_cast<const uint8_t *>(d_accts),
                               static_cast<const uint8_t *>(d_storage_roots32), n, static_cast<uint8_t *>(d_root32),
                               false, b));
    return finish_build_state(c);
}

extern "C" B200_API int32_t b200_state_root_full_dev(b200_ctx *c, const void *d_acct_keys32, const void *d_accts,
                                            uint64_t n_accounts, const void *d_slot_keys32, const void *d_values32_be,
                                            const void *d_seg_offsets, uint64_t n_slots, void *d_root32) {
    if (!c || !d_root32 || !d_seg_offsets || (n_accounts && (!d_acct_keys32 || !d_accts)) ||
        (n_slots && (!d_slot_keys32 || !d_values32_be)))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(reset_build_state(c));
    ENSURE(sroots, (n_accounts ? n_accounts : 1) * 32);
    Built bs, ba;
    TRY(storage_roots_on_device(c, static_cast<const uint8_t *>(d_slot_keys32),
                                static_cast<const uint8_t *>(d_values32_be),
                                static_cast<const uint64_t *>(d_seg_offsets), n_accounts, n_slots,
                                static_cast<uint8_t *>(c->sroots.p), false, bs));
    TRY(account_root_on_device(c, static_cast<const uint8_t *>(d_acct_keys32), static_cast<const uint8_t *>(d_accts),
                               static_cast<const uint8_t *>(c->sroots.p), n_accounts,
                               static_cast<uint8_t *>(d_root32), false, ba));
    return finish_build_state(c);
}

// ------------------------------------------------------------------------------------------------ host-pointer drivers
static int32_t h2d(b200_ctx *c, DevBuf &b, const void *src, size_t bytes) {
    TRY(ensure(c, b, bytes ? bytes : 16));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return B200_OK;
}

static int32_t check_offsets_host(b200_ctx *c, const uint64_t *offs, uint64_t n_segs) {
    if (offs[0] != 0) return fail(c, B200_ERR_INVALID_ARG, "seg_offsets[0] must be 0");
    for (uint64_t i = 0; i < n_segs; i++)
        if (offs[i + 1] < offs[i]) return fail(c, B200_ERR_INVALID_ARG, "seg_offsets must be monotone");
    return B200_OK;
}

extern "C" B200_API int32_t b200_storage_roots(b200_ctx *c, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                      const uint64_t *seg_offsets, uint64_t n_accounts, uint8_t *roots32,
                                      b200_updates *opt_updates, b200_stats *opt_stats) {
    if (!c || !seg_offsets || (n_accounts && !roots32)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    TRY(check_offsets_host(c, seg_offsets, n_accounts));
    uint64_t n_slots = seg_offsets[n_accounts];
    if (n_slots && (!slot_keys32 || !values32_be)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (opt_updates) memset(opt_updates, 0, sizeof *opt_updates);
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(h2d(c, c->in_a, slot_keys32, n_slots * 32));
    TRY(h2d(c, c->in_b, values32_be, n_slots * 32));
    TRY(h2d(c, c->in_c, seg_offsets, (n_accounts + 1) * 8));
    ENSURE(sroots, (n_accounts ? n_accounts : 1) * 32);
    TRY(reset_build_state(c));
    Built b;
    TRY(storage_roots_on_device(c, static_cast<const uint8_t *>(c->in_a.p), static_cast<const uint8_t *>(c->in_b.p),
                                static_cast<const uint64_t *>(c->in_c.p), n_accounts, n_slots,
                                static_cast<uint8_t *>(c->sroots.p), opt_updates != nullptr, b));
    TRY(finish_build_state(c));
    if (n_accounts) CU(cudaMemcpyAsync(roots32, c->sroots.p, n_accounts * 32, cudaMemcpyDeviceToHost, c->stream));
    int32_t r = sync_and_status(c);
    if (r == B200_OK && opt_updates)
        r = collect_updates(c, b, static_cast<const uint64_t *>(c->in_c.p), n_accounts, opt_updates);
    if (r != B200_OK && opt_updates) b200_updates_release(opt_updates);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

extern "C" B200_API int32_t b200_state_root(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                   const uint8_t *storage_roots32, uint64_t n, uint8_t root32[32],
                                   b200_updates *opt_updates, b200_stats *opt_stats) {
    if (!c || !root32 || (n && (!acct_keys32 || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (opt_updates) memset(opt_updates, 0, sizeof *opt_updates);
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(h2d(c, c->in_a, acct_keys32, n * 32));
    TRY(h2d(c, c->in_b, accts, n * sizeof(b200_account)));
    if (storage_roots32) TRY(h2d(c, c->in_c, storage_roots32, n * 32));
    ENSURE(in_e, 32);
    TRY(reset_build_state(c));
    Built b;
    TRY(account_root_on_device(c, static_cast<const uint8_t *>(c->in_a.p), static_cast<const uint8_t *>(c->in_b.p),
                               storage_roots32 ? static_cast<const uint8_t *>(c->in_c.p) : nullptr, n,
                               static_cast<uint8_t *>(c->in_e.p), opt_updates != nullptr, b));
    TRY(finish_build_state(c));
    CU(cudaMemcpyAsync(root32, c->in_e.p, 32, cudaMemcpyDeviceToHost, c->stream));
    int32_t r = sync_and_status(c);
    if (r == B200_OK && opt_updates) r = collect_updates(c, b, nullptr, 0, opt_updates);
    if (r != B200_OK && opt_updates) b200_updates_release(opt_updates);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

// Host-pointer full state root without retained updates: the storage forest is cut into account ranges so that the
// H2D copy of range k+1 (copy stream) overlaps the build of range k (compute stream).  PCIe moves ~1.1 GB for the C3
// workload (≈20 ms) against ≈9 ms of hashing: the transfer is the critical path and the hashing hides under it.
static int32_t state_root_full_pipelined(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                         uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                         const uint64_t *seg_offsets, uint64_t n_slots, uint8_t root32[32]) {
    // chunk boundaries: ~n_slots/12 slots each, at least 1M
    const uint64_t target = std::max<uint64_t>(n_slots / 12, 1ull << 20);
    std::vector<uint64_t> cut{0};
    for (uint64_t a = 1; a <= n_accounts; a++)
        if (a == n_accounts || seg_offsets[a] - seg_offsets[cut.back()] >= target) cut.push_back(a);
    const size_t n_chunks = cut.size() - 1;
    // per-chunk offsets rebased to 0
    std::vector<uint64_t> rel(n_accounts + n_chunks);
    std::vector<uint64_t> rel_start(n_chunks);
    {
        uint64_t w = 0;
        for (size_t k = 0; k < n_chunks; k++) {
            rel_start[k] = w;
            uint64_t s0 = seg_offsets[cut[k]];
            for (uint64_t a = cut[k]; a <= cut[k + 1]; a++) rel[w++] = seg_offsets[a] - s0;
        }
    }
    ENSURE(in_a, n_slots * 32);
    ENSURE(in_b, n_slots * 32);
    ENSURE(in_c, rel.size() * 8);
    ENSURE(in_d, n_accounts * 32);
    ENSURE(in_e, n_accounts * sizeof(b200_account));
    ENSURE(sroots, n_accounts * 32 + 32);
    while (c->chunk_events.size() < n_chunks + 1) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_events.push_back(e);
    }
    cudaStream_t cs = c->copy_streams[0];
    CU(cudaStreamSynchronize(c->stream));  // scratch of an earlier call may still be in use
    CU(cudaMemcpyAsync(c->in_c.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, cs));
    uint8_t *d_keys = static_cast<uint8_t *>(c->in_a.p), *d_vals = static_cast<uint8_t *>(c->in_b.p);
    for (size_t k = 0; k < n_chunks; k++) {
        uint64_t s0 = seg_offsets[cut[k]], s1 = seg_offsets[cut[k + 1]];
        if (s1 > s0) {
            CU(cudaMemcpyAsync(d_keys + 32 * s0, slot_keys32 + 32 * s0, (s1 - s0) * 32, cudaMemcpyHostToDevice, cs));
            CU(cudaMemcpyAsync(d_vals + 32 * s0, values32_be + 32 * s0, (s1 - s0) * 32, cudaMemcpyHostToDevice, cs));
        }
        CU(cudaEventRecord(c->chunk_events[k], cs));
    }
    CU(cudaMemcpyAsync(c->in_d.p, acct_keys32, n_accounts * 32, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(c->in_e.p, accts, n_accounts * sizeof(b200_account), cudaMemcpyHostToDevice, cs));
    CU(cudaEventRecord(c->chunk_events[n_chunks], cs));

    TRY(reset_build_state(c));
    uint8_t *d_sroots = static_cast<uint8_t *>(c->sroots.p);
    uint8_t *d_root = d_sroots + n_accounts * 32;
    const uint64_t *d_rel = static_cast<const uint64_t *>(c->in_c.p);
    int32_t r = B200_OK;
    for (size_t k = 0; k < n_chunks && r == B200_OK; k++) {
        uint64_t a0 = cut[k], a1 = cut[k + 1], s0 = seg_offsets[a0], s1 = seg_offsets[a1];
        CU(cudaStreamWaitEvent(c->stream, c->chunk_events[k], 0));
        Built b;
        r = build_forest(c, d_keys + 32 * s0, s1 - s0, d_rel + rel_start[k], a1 - a0, false, d_vals + 32 * s0, nullptr,
                         false, b);
        if (r != B200_OK) break;
        CU(launch_segment_roots(b.f, d_rel + rel_start[k], a1 - a0, d_sroots + 32 * a0, c->stream));
        c->launches++;
        c->stats.leaves_added += s1 - s0;
        c->stats.branches_added += b.n_nodes;
        c->stats.levels += b.levels;
    }
    if (r != B200_OK) {
        cudaStreamSynchronize(cs);  // do not leave copies in flight into buffers a later call may resize
        return r;
    }
    CU(cudaStreamWaitEvent(c->stream, c->chunk_events[n_chunks], 0));
    Built ba;
    TRY(account_root_on_device(c, static_cast<const uint8_t *>(c->in_d.p), static_cast<const uint8_t *>(c->in_e.p),
                               d_sroots, n_accounts, d_root, false, ba));
    TRY(finish_build_state(c));
    CU(cudaMemcpyAsync(root32, d_root, 32, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_status(c);
}

extern "C" B200_API int32_t b200_state_root_full(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                        uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                        const uint64_t *seg_offsets, uint8_t root32[32],
                                        b200_updates *opt_account_updates, b200_updates *opt_storage_updates,
                                        b200_stats *opt_stats) {
    if (!c || !root32 || !seg_offsets || (n_accounts && (!acct_keys32 || !accts)))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    TRY(check_offsets_host(c, seg_offsets, n_accounts));
    uint64_t n_slots = seg_offsets[n_accounts];
    if (n_slots && (!slot_keys32 || !values32_be)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (opt_account_updates) memset(opt_account_updates, 0, sizeof *opt_account_updates);
    if (opt_storage_updates) memset(opt_storage_updates, 0, sizeof *opt_storage_updates);
    const bool retain = opt_account_updates || opt_storage_updates;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    if (!retain && n_slots >= (2ull << 20) && n_accounts >= 16) {
        int32_t pr = state_root_full_pipelined(c, acct_keys32, accts, n_accounts, slot_keys32, values32_be, seg_offsets,
                                               n_slots, root32);
        if (opt_stats) *opt_stats = c->stats;
        return pr;
    }
    TRY(h2d(c, c->in_a, slot_keys32, n_slots * 32));
    TRY(h2d(c, c->in_b, values32_be, n_slots * 32));
    TRY(h2d(c, c->in_c, seg_offsets, (n_accounts + 1) * 8));
    TRY(h2d(c, c->in_d, acct_keys32, n_accounts * 32));
    TRY(h2d(c, c->in_e, accts, n_accounts * sizeof(b200_account)));
    ENSURE(sroots, (n_accounts ? n_accounts : 1) * 32 + 32);
    uint8_t *d_root = static_cast<uint8_t *>(c->sroots.p) + (n_accounts ? n_accounts : 1) * 32;
    TRY(reset_build_state(c));
    Built bs, ba;
    int32_t r = storage_roots_on_device(c, static_cast<const uint8_t *>(c->in_a.p),
                                        static_cast<const uint8_t *>(c->in_b.p),
                                        static_cast<const uint64_t *>(c->in_c.p), n_accounts, n_slots,
                                        static_cast<uint8_t *>(c->sroots.p), retain, bs);
    // the storage forest's scratch is reused by the account build: gather its updates first
    if (r == B200_OK && opt_storage_updates) {
        r = sync_and_status(c);
        if (r == B200_OK)
            r = collect_updates(c, bs, static_cast<const uint64_t *>(c->in_c.p), n_accounts, opt_storage_updates);
    }
    if (r == B200_OK)
        r = account_root_on_device(c, static_cast<const uint8_t *>(c->in_d.p), static_cast<const uint8_t *>(c->in_e.p),
                                   static_cast<const uint8_t *>(c->sroots.p), n_accounts, d_root, retain, ba);
    if (r == B200_OK) r = finish_build_state(c);
    if (r == B200_OK) {
        cudaError_t e = cudaMemcpyAsync(root32, d_root, 32, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) r = fail(c, B200_ERR_CUDA, "root copy: %s", cudaGetErrorString(e));
    }
    if (r == B200_OK) r = sync_and_status(c);
    if (r == B200_OK && opt_account_updates) r = collect_updates(c, ba, nullptr, 0, opt_account_updates);
    if (r != B200_OK) {
        if (opt_account_updates) b200_updates_release(opt_account_updates);
        if (opt_storage_updates) b200_updates_release(opt_storage_updates);
    }
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

// ------------------------------------------------------------------------------------------------ multi-GPU frontier
static int32_t frontier_on_device(b200_ctx *c, const uint8_t *d_akeys, const uint8_t *d_accts, uint64_t n_accounts,
                                  const uint8_t *d_skeys, const uint8_t *d_svals, const uint64_t *d_offs,
                                  uint64_t n_slots, FrontierEntryDev *d_out) {
    ENSURE(sroots, (n_accounts ? n_accounts : 1) * 32);
    ENSURE(buckets, 17 * 8);
    Built bs, ba;
    TRY(storage_roots_on_device(c, d_skeys, d_svals, d_offs, n_accounts, n_slots, static_cast<uint8_t *>(c->sroots.p),
                                false, bs));
    uint64_t *d_buckets = static_cast<uint64_t *>(c->buckets.p);
    CU(launch_nibble_buckets(d_akeys, n_accounts, d_buckets, c->stream));
    // every top-nibble bucket is built as a trie of its own (16 segments)
    TRY(build_forest(c, d_akeys, n_accounts, d_buckets, 16, true, d_accts, static_cast<const uint8_t *>(c->sroots.p),
                     false, ba));
    CU(launch_frontier(ba.f, d_buckets, d_accts, static_cast<const uint8_t *>(c->sroots.p), d_out, c->stream));
    c->launches += 2;
    c->stats.leaves_added += n_accounts;
    c->stats.branches_added += ba.n_nodes;
    c->stats.levels += ba.levels;
    return B200_OK;
}

extern "C" B200_API int32_t b200_subtrie_frontier_dev(b200_ctx *c, const void *d_acct_keys32, const void *d_accts,
                                             uint64_t n_accounts, const void *d_slot_keys32,
                                             const void *d_values32_be, const void *d_seg_offsets, uint64_t n_slots,
                                             void *d_frontier) {
    if (!c || !d_frontier || !d_seg_offsets || (n_accounts && (!d_acct_keys32 || !d_accts)))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(reset_build_state(c));
    TRY(frontier_on_device(c, static_cast<const uint8_t *>(d_acct_keys32), static_cast<const uint8_t *>(d_accts),
                           n_accounts, static_cast<const uint8_t *>(d_slot_keys32),
                           static_cast<const uint8_t *>(d_values32_be), static_cast<const uint64_t *>(d_seg_offsets),
                           n_slots, static_cast<FrontierEntryDev *>(d_frontier)));
    return finish_build_state(c);
}

extern "C" B200_API int32_t b200_subtrie_frontier(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                         uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                         const uint64_t *seg_offsets, b200_frontier_entry frontier[16],
                                         b200_stats *opt_stats) {
    if (!c || !frontier || !seg_offsets || (n_accounts && (!acct_keys32 || !accts)))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    TRY(check_offsets_host(c, seg_offsets, n_accounts));
    uint64_t n_slots = seg_offsets[n_accounts];
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(h2d(c, c->in_a, slot_keys32, n_slots * 32));
    TRY(h2d(c, c->in_b, values32_be, n_slots * 32));
    TRY(h2d(c, c->in_c, seg_offsets, (n_accounts + 1) * 8));
    TRY(h2d(c, c->in_d, acct_keys32, n_accounts * 32));
    TRY(h2d(c, c->in_e, accts, n_accounts * sizeof(b200_account)));
    ENSURE(out_a, 16 * sizeof(FrontierEntryDev));
    TRY(reset_build_state(c));
    TRY(frontier_on_device(c, static_cast<const uint8_t *>(c->in_d.p), static_cast<const uint8_t *>(c->in_e.p),
                           n_accounts, static_cast<const uint8_t *>(c->in_a.p),
                           static_cast<const uint8_t *>(c->in_b.p), static_cast<const uint64_t *>(c->in_c.p), n_slots,
                           static_cast<FrontierEntryDev *>(c->out_a.p)));
    TRY(finish_build_state(c));
    CU(cudaMemcpyAsync(frontier, c->out_a.p, 16 * sizeof(FrontierEntryDev), cudaMemcpyDeviceToHost, c->stream));
    int32_t r = sync_and_status(c);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

extern "C" B200_API int32_t b200_root_from_frontier(b200_ctx *c, const b200_frontier_entry frontier[16], uint8_t root32[32]) {
    if (!c || !frontier || !root32) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    for (int i = 0; i < 16; i++)
        if (frontier[i].as_child_len > 33 || (frontier[i].as_root_len != 0 && frontier[i].as_root_len != 32))
            return fail(c, B200_ERR_INVALID_ARG, "malformed frontier entry %d", i);
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ENSURE(out_a, 16 * sizeof(FrontierEntryDev) + 64);
    uint8_t *d = static_cast<uint8_t *>(c->out_a.p);
    uint8_t *d_root = d + align_up(16 * sizeof(FrontierEntryDev), 16);
    CU(cudaMemcpyAsync(d, frontier, 16 * sizeof(FrontierEntryDev), cudaMemcpyHostToDevice, c->stream));
    CU(launch_root_from_frontier(reinterpret_cast<const FrontierEntryDev *>(d), d_root, c->stream));
    c->launches++;
    CU(cudaMemcpyAsync(root32, d_root, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

// device-resident variant used by the multi-GPU host after the NCCL all-gather
extern "C" B200_API int32_t b200_root_from_frontier_dev(b200_ctx *c, const void *d_frontier, void *d_root32) {
    if (!c || !d_frontier || !d_root32) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(launch_root_from_frontier(static_cast<const FrontierEntryDev *>(d_frontier), static_cast<uint8_t *>(d_root32),
                                 c->stream));
    c->launches++;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ resident trie (C5)
// The account trie of a whole state kept in HBM — keys, accounts, storage roots and the node-hash frontier of every
// level — so that a block's dirty accounts are committed by re-hashing only their root paths.  This is what reth
// gets from stored branch nodes + prefix sets (crates/trie/trie/src/walker.rs:172-202, node_iter.rs:205-300):
// untouched subtries are not revisited.  Scope: value changes of existing accounts (balance / nonce / code hash /
// storage root); inserting or deleting a key changes the trie shape and is reported as B200_ERR_NOT_FOUND so the
// caller rebuilds.
struct b200_trie {
    b200_ctx *c = nullptr;
    uint64_t n = 0;
    uint32_t B = 0;
    ForestDev f{};
    bool has_sroots = false;
    uint64_t bytes = 0;
    uint32_t level_count[64] = {};
    DevBuf keys, accts, sroots, Lp, nibs, leaf_ref, leaf_meta, S, E, gap_sorted, node_start, node_ref, node_meta, node_l,
        node_r, node_masks, leaf_parent, node_parent, dirty, dirty_ids, dirty_key, dirty_key2, dirty_order, idx, in_keys,
        in_accts, in_sroots, root;
};

static void steal(b200_trie *t, DevBuf &dst, DevBuf &src) {
    dst = src;
    src = DevBuf{};
    t->c->dev_bytes -= dst.cap;
    t->bytes += dst.cap;
}
static int32_t trie_alloc(b200_trie *t, DevBuf &b, size_t bytes) {
    b200_ctx *c = t->c;
    if (bytes <= b.cap) return B200_OK;
    if (b.p) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(b.p));
        t->bytes -= b.cap;
        b = DevBuf{};
    }
    size_t want = bytes + 256;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    t->bytes += want;
    return B200_OK;
}

extern "C" B200_API void b200_trie_destroy(b200_trie *t) {
    if (!t) return;
    cudaSetDevice(t->c->device);
    cudaStreamSynchronize(t->c->stream);
    DevBuf *bufs[] = {&t->keys, &t->accts, &t->sroots, &t->Lp, &t->nibs, &t->leaf_ref, &t->leaf_meta, &t->S, &t->E,
                      &t->gap_sorted, &t->node_start, &t->node_ref, &t->node_meta, &t->node_l, &t->node_r, &t->node_masks,
                      &t->leaf_parent, &t->node_parent, &t->dirty, &t->dirty_ids, &t->dirty_key, &t->dirty_key2,
                      &t->dirty_order, &t->idx, &t->in_keys, &t->in_accts, &t->in_sroots, &t->root};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    delete t;
}
extern "C" B200_API uint64_t b200_trie_device_bytes(const b200_trie *t) { return t ? t->bytes : 0; }
extern "C" B200_API uint64_t b200_trie_leaves(const b200_trie *t) { return t ? t->n : 0; }

// builds from device-resident inputs that the trie already owns (t->keys / accts / sroots)
static int32_t trie_build_owned(b200_trie *t) {
    b200_ctx *c = t->c;
    TRY(reset_build_state(c));
    Built b;
    TRY(trie_alloc(t, t->root, 64));
    TRY(account_root_on_device(c, static_cast<const uint8_t *>(t->keys.p), static_cast<const uint8_t *>(t->accts.p),
                               t->has_sroots ? static_cast<const uint8_t *>(t->sroots.p) : nullptr, t->n,
                               static_cast<uint8_t *>(t->root.p), true, b));
    TRY(finish_build_state(c));
    TRY(sync_and_status(c));
    t->f = b.f;
    t->B = b.n_nodes;
    memcpy(t->level_count, b.level_count, sizeof t->level_count);
    // the build's arrays become the trie's: same pointers, new owner; the context re-allocates on its next build
    steal(t, t->Lp, c->Lp);
    steal(t, t->nibs, c->nibs);
    steal(t, t->leaf_ref, c->leaf_ref);
    steal(t, t->leaf_meta, c->leaf_meta);
    steal(t, t->S, c->S);
    steal(t, t->E, c->E);
    if (t->n >= 2) {
        steal(t, t->gap_sorted, c->gap_sorted);
        steal(t, t->node_start, c->node_start);
    }
    if (t->B) {
        steal(t, t->node_ref, c->node_ref);
        steal(t, t->node_meta, c->node_meta);
        steal(t, t->node_l, c->node_l);
        steal(t, t->node_r, c->node_r);
        steal(t, t->node_masks, c->node_masks);
    }
    TRY(trie_alloc(t, t->leaf_parent, (t->n ? t->n : 1) * 4));
    TRY(trie_alloc(t, t->node_parent, ((size_t)t->B + 1) * 4));
    TRY(trie_alloc(t, t->dirty, ((size_t)t->B + 1) * 4));
    CU(cudaMemsetAsync(t->leaf_parent.p, 0xFF, (t->n ? t->n : 1) * 4, c->stream));
    CU(cudaMemsetAsync(t->node_parent.p, 0xFF, ((size_t)t->B + 1) * 4, c->stream));
    CU(cudaMemsetAsync(t->dirty.p, 0, ((size_t)t->B + 1) * 4, c->stream));
    CU(launch_parent_links(t->f, t->B, static_cast<uint32_t *>(t->leaf_parent.p),
                           static_cast<uint32_t *>(t->node_parent.p), c->stream));
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

static int32_t trie_create_common(b200_ctx *c, const void *keys, const void *accts, const void *sroots, uint64_t n,
                                  cudaMemcpyKind kind, b200_trie **out, void *root_out) {
    if (!c || !out || (n && (!keys || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    *out = nullptr;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    b200_trie *t = new b200_trie();
    t->c = c;
    t->n = n;
    t->has_sroots = sroots != nullptr;
    int32_t r = B200_OK;
    auto put = [&](DevBuf &b, const void *src, size_t bytes) -> int32_t {
        TRY(trie_alloc(t, b, bytes ? bytes : 16));
        if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, kind, c->stream));
        return B200_OK;
    };
    r = put(t->keys, keys, n * 32);
    if (r == B200_OK) r = put(t->accts, accts, n * 72);
    if (r == B200_OK && sroots) r = put(t->sroots, sroots, n * 32);
    if (r == B200_OK) r = trie_build_owned(t);
    if (r == B200_OK && root_out) {
        cudaError_t e = cudaMemcpyAsync(root_out, t->root.p, 32,
                                        kind == cudaMemcpyHostToDevice ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                        c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) r = fail(c, B200_ERR_CUDA, "root copy: %s", cudaGetErrorString(e));
    }
    if (r != B200_OK) {
        b200_trie_destroy(t);  // does not take the context lock
        return r;
    }
    *out = t;
    return B200_OK;
}

extern "C" B200_API int32_t b200_trie_create(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                             const uint8_t *storage_roots32, uint64_t n, b200_trie **out,
                                             uint8_t root32[32]) {
    return trie_create_common(c, acct_keys32, accts, storage_roots32, n, cudaMemcpyHostToDevice, out, root32);
}
extern "C" B200_API int32_t b200_trie_create_dev(b200_ctx *c, const void *d_acct_keys32, const void *d_accts,
                                                 const void *d_storage_roots32, uint64_t n, b200_trie **out,
                                                 void *d_root32) {
    return trie_create_common(c, d_acct_keys32, d_accts, d_storage_roots32, n, cudaMemcpyDeviceToDevice, out, d_root32);
}

// dirty inputs already on the device; result root in t->root.  Three launches, no host round trip:
// locate (binary search) -> mark_pending (count dirty children per ancestor) -> wavefront (leaf + root-path re-hash).
static int32_t trie_update_on_device(b200_trie *t, const uint8_t *d_keys, const uint8_t *d_accts, const uint8_t *d_sroots,
                                     uint64_t m) {
    b200_ctx *c = t->c;
    cudaStream_t st = c->stream;
    TRY(reset_build_state(c));
    if (m == 0 || t->n == 0) {
        if (m && t->n == 0) return fail(c, B200_ERR_NOT_FOUND, "the resident trie is empty");
        return finish_build_state(c);
    }
    if (d_sroots && !t->has_sroots) return fail(c, B200_ERR_INVALID_ARG, "trie was created without storage roots");
    ForestDev f = t->f;
    f.retain_updates = 1;
    TRY(trie_alloc(t, t->idx, m * 4));
    uint64_t max_dirty = std::min<uint64_t>((uint64_t)t->B, m * 64) + 1;  // at most 64 ancestors per dirty leaf
    TRY(trie_alloc(t, t->dirty_ids, max_dirty * 4));
    uint32_t *idx = static_cast<uint32_t *>(t->idx.p);
    uint32_t *count_p = small_u32(c) + SM_NSTORED;
    CU(cudaMemsetAsync(count_p, 0, 4, st));
    CU(launch_locate(static_cast<const uint8_t *>(t->keys.p), t->n, d_keys, m, idx, f.err, st));
    CU(launch_mark_pending(f, idx, m, static_cast<uint32_t *>(t->leaf_parent.p), static_cast<uint32_t *>(t->node_parent.p),
                           static_cast<uint32_t *>(t->dirty.p), st));
    // Populous deep levels (more dirty nodes than a wave of warps can absorb cheaply) are climbed by one thread per
    // leaf with the register-resident sponge; the sparse levels above by one warp per node (shuffle sponge).
    int split = 65;  // 65: everything warp-cooperative
    if (m > WARP_LEVEL_MAX)
        for (int d = 0; d < 64; d++)
            if (std::min<uint64_t>(m, t->level_count[d]) > WARP_LEVEL_MAX) {
                split = d;
                break;
            }
    uint8_t *accts = static_cast<uint8_t *>(t->accts.p);
    uint8_t *sroots = t->has_sroots ? static_cast<uint8_t *>(t->sroots.p) : nullptr;
    uint32_t *lp = static_cast<uint32_t *>(t->leaf_parent.p), *np = static_cast<uint32_t *>(t->node_parent.p);
    uint32_t *pending = static_cast<uint32_t *>(t->dirty.p), *dlist = static_cast<uint32_t *>(t->dirty_ids.p);
    if (split == 65) {
        CU(launch_wavefront(f, accts, sroots, d_accts, d_sroots, idx, m, lp, np, pending, dlist, count_p,
                            static_cast<uint8_t *>(t->root.p), st));
    } else {
        TRY(trie_alloc(t, t->dirty_order, m * 4));  // hand-over list: at most one entry per dirty leaf
        uint32_t *hcount = small_u32(c) + SM_NNODES + 1;
        CU(cudaMemsetAsync(hcount, 0, 4, st));
        CU(launch_wavefront_two_stage(f, accts, sroots, d_accts, d_sroots, idx, m, lp, np, pending, dlist, count_p,
                                      static_cast<uint32_t *>(t->dirty_order.p), hcount, m,
                                      static_cast<uint8_t *>(t->root.p), split, st));
        c->launches++;
    }
    c->launches += 3;
    c->stats.leaves_added += m;
    c->stats_wavefront = true;
    return finish_build_state(c);
}

// number of re-hashed branch nodes of the last update (after a sync)
static int32_t trie_read_dirty_count(b200_trie *t, uint32_t *out) {
    b200_ctx *c = t->c;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    CU(cudaMemcpyAsync(ps + 200, small_u32(c) + SM_NSTORED, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = ps[200];
    return B200_OK;
}

extern "C" B200_API int32_t b200_trie_update_dev(b200_trie *t, const void *d_dirty_keys32, const void *d_new_accts,
                                                 const void *d_new_storage_roots32, uint64_t m, void *d_root32) {
    if (!t || (m && (!d_dirty_keys32 || !d_new_accts))) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(trie_update_on_device(t, static_cast<const uint8_t *>(d_dirty_keys32), static_cast<const uint8_t *>(d_new_accts),
                              static_cast<const uint8_t *>(d_new_storage_roots32), m));
    if (d_root32) CU(cudaMemcpyAsync(d_root32, t->root.p, 32, cudaMemcpyDeviceToDevice, c->stream));
    return B200_OK;  // asynchronous: B200_ERR_NOT_FOUND etc. surface at the next b200_sync / b200_dev_status
}

extern "C" B200_API int32_t b200_trie_update(b200_trie *t, const uint8_t *dirty_keys32, const b200_account *new_accts,
                                             const uint8_t *new_storage_roots32, uint64_t m, uint8_t root32[32],
                                             b200_updates *opt_updates, b200_stats *opt_stats) {
    if (!t || !root32 || (m && (!dirty_keys32 || !new_accts)))
        return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    if (opt_updates) memset(opt_updates, 0, sizeof *opt_updates);
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(trie_alloc(t, t->in_keys, (m ? m : 1) * 32));
    TRY(trie_alloc(t, t->in_accts, (m ? m : 1) * 72));
    if (m) {
        CU(cudaMemcpyAsync(t->in_keys.p, dirty_keys32, m * 32, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(t->in_accts.p, new_accts, m * 72, cudaMemcpyHostToDevice, c->stream));
    }
    if (new_storage_roots32 && m) {
        TRY(trie_alloc(t, t->in_sroots, m * 32));
        CU(cudaMemcpyAsync(t->in_sroots.p, new_storage_roots32, m * 32, cudaMemcpyHostToDevice, c->stream));
    }
    uint32_t D = 0;
    int32_t r = trie_update_on_device(t, static_cast<const uint8_t *>(t->in_keys.p),
                                      static_cast<const uint8_t *>(t->in_accts.p),
                                      new_storage_roots32 ? static_cast<const uint8_t *>(t->in_sroots.p) : nullptr, m);
    if (r == B200_OK) {
        cudaError_t e = cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) r = fail(c, B200_ERR_CUDA, "root copy: %s", cudaGetErrorString(e));
    }
    if (r == B200_OK) r = sync_and_status(c);
    if (r == B200_OK) r = trie_read_dirty_count(t, &D);
    if (r == B200_OK) {
        c->stats.branches_added = D;
        if (opt_updates) r = collect_updates_subset(c, t->f, static_cast<const uint32_t *>(t->dirty_ids.p), D, opt_updates);
    }
    if (r != B200_OK && opt_updates) b200_updates_release(opt_updates);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

static void trie_free(b200_trie *t, DevBuf &b) {
    if (b.p) {
        cudaFree(b.p);
        t->bytes -= b.cap;
        b = DevBuf{};
    }
}

// General commit of a sorted dirty set (HashedPostStateSorted semantics: present = upsert, absent = delete).  If every
// entry is a value change of an existing account the dirty paths are re-hashed in place; otherwise the keys are
// merged on the device (two scans + two scatters) and the trie is rebuilt from the merged arrays — the state never
// travels back to the host.  *out_rebuilt tells which one happened: after a rebuild opt_updates holds the COMPLETE
// node set of the new trie (the caller clears AccountsTrie first, like MerkleStage's rebuild path, merkle.rs:237-238).
extern "C" B200_API int32_t b200_trie_apply(b200_trie *t, const uint8_t *keys32, const b200_account *accts,
                                            const uint8_t *present, const uint8_t *storage_roots32, uint64_t m,
                                            uint8_t root32[32], int32_t *out_rebuilt, b200_updates *opt_updates,
                                            b200_stats *opt_stats) {
    if (!t || !root32 || (m && (!keys32 || !accts))) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    if (opt_updates) memset(opt_updates, 0, sizeof *opt_updates);
    if (out_rebuilt) *out_rebuilt = 0;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    if (storage_roots32 && !t->has_sroots) return fail(c, B200_ERR_INVALID_ARG, "trie was created without storage roots");
    TRY(trie_alloc(t, t->in_keys, (m ? m : 1) * 32));
    TRY(trie_alloc(t, t->in_accts, (m ? m : 1) * 72));
    TRY(trie_alloc(t, t->idx, (m ? m : 1) * 4));
    TRY(trie_alloc(t, t->dirty_key, (m ? m : 1) * 2));  // kind[m] | present[m]
    uint8_t *d_kind = static_cast<uint8_t *>(t->dirty_key.p), *d_present = d_kind + (m ? m : 1);
    if (m) {
        CU(cudaMemcpyAsync(t->in_keys.p, keys32, m * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(t->in_accts.p, accts, m * 72, cudaMemcpyHostToDevice, st));
        if (present) CU(cudaMemcpyAsync(d_present, present, m, cudaMemcpyHostToDevice, st));
        if (storage_roots32) {
            TRY(trie_alloc(t, t->in_sroots, m * 32));
            CU(cudaMemcpyAsync(t->in_sroots.p, storage_roots32, m * 32, cudaMemcpyHostToDevice, st));
        }
    }
    const uint8_t *d_keys = static_cast<const uint8_t *>(t->in_keys.p), *d_accts = static_cast<const uint8_t *>(t->in_accts.p);
    const uint8_t *d_sr = storage_roots32 ? static_cast<const uint8_t *>(t->in_sroots.p) : nullptr;
    TRY(reset_build_state(c));
    uint32_t *counts = small_u32(c) + SM_HIST;  // [0] inserts [1] deletes [2] value updates
    CU(cudaMemsetAsync(counts, 0, 16, st));
    uint32_t *lb = static_cast<uint32_t *>(t->idx.p);
    CU(launch_locate_classify(static_cast<const uint8_t *>(t->keys.p), t->n, d_keys, present ? d_present : nullptr, m, lb, d_kind,
                              counts, reinterpret_cast<int *>(small_u32(c) + SM_ERR), st));
    c->launches++;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    CU(cudaMemcpyAsync(ps + 300, counts, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ps, small_u32(c) + SM_ERR, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ps[0] != B200_DEVERR_NONE) return map_dev_error(c, (int)ps[0]);
    const uint64_t n_ins = ps[300], n_del = ps[301], n_upd = ps[302];
    int32_t r = B200_OK;
    if (n_ins == 0 && n_del == 0 && n_upd == m) {
        // ---- value changes only: wavefront re-hash of the dirty paths
        r = trie_update_on_device(t, d_keys, d_accts, d_sr, m);
        if (r == B200_OK) {
            cudaError_t e = cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) r = fail(c, B200_ERR_CUDA, "root copy: %s", cudaGetErrorString(e));
        }
        if (r == B200_OK) r = sync_and_status(c);
        uint32_t D = 0;
        if (r == B200_OK) r = trie_read_dirty_count(t, &D);
        if (r == B200_OK) {
            c->stats.branches_added = D;
            if (opt_updates) r = collect_updates_subset(c, t->f, static_cast<const uint32_t *>(t->dirty_ids.p), D, opt_updates);
        }
    } else {
        // ---- shape changes: merge on the device, rebuild
        const uint64_t n = t->n, n2 = n + n_ins - n_del;
        if (n2 >= (1ull << 31)) return fail(c, B200_ERR_INVALID_ARG, "merged trie exceeds 2^31-1 leaves");
        DevBuf marks{}, nk{}, na{}, ns{};
        auto cleanup = [&]() {
            trie_free(t, marks);
            trie_free(t, nk);
            trie_free(t, na);
            trie_free(t, ns);
        };
        // marks: ins_at[n+1] | del[n+1] | ins_incl[n+1] | del_excl[n+1] | ins_flag[m] | ins_rank[m]
        size_t w = n + 1;
        r = trie_alloc(t, marks, (4 * w + 2 * (m ? m : 1)) * 4);
        if (r == B200_OK) r = trie_alloc(t, nk, (n2 ? n2 : 1) * 32);
        if (r == B200_OK) r = trie_alloc(t, na, (n2 ? n2 : 1) * 72);
        if (r == B200_OK && t->has_sroots) r = trie_alloc(t, ns, (n2 ? n2 : 1) * 32);
        if (r != B200_OK) {
            cleanup();
            return r;
        }
        uint32_t *ins_at = static_cast<uint32_t *>(marks.p), *del = ins_at + w, *ins_incl = del + w, *del_excl = ins_incl + w,
                 *ins_flag = del_excl + w, *ins_rank = ins_flag + (m ? m : 1);
        auto run = [&]() -> int32_t {
            CU(cudaMemsetAsync(ins_at, 0, 2 * w * 4, st));
            CU(launch_merge_marks(lb, d_kind, m, ins_at, del, ins_flag, st));
            size_t t1 = 0, t2 = 0, t3 = 0;
            CU(cub::DeviceScan::InclusiveSum(nullptr, t1, ins_at, ins_incl, (int64_t)w, st));
            CU(cub::DeviceScan::ExclusiveSum(nullptr, t2, del, del_excl, (int64_t)w, st));
            CU(cub::DeviceScan::ExclusiveSum(nullptr, t3, ins_flag, ins_rank, (int64_t)(m ? m : 1), st));
            ENSURE(cub_temp, std::max(t1, std::max(t2, t3)));
            CU(cub::DeviceScan::InclusiveSum(c->cub_temp.p, t1, ins_at, ins_incl, (int64_t)w, st));
            CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t2, del, del_excl, (int64_t)w, st));
            if (m) CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t3, ins_flag, ins_rank, (int64_t)m, st));
            CU(launch_merge_scatter(static_cast<const uint8_t *>(t->keys.p), static_cast<const uint8_t *>(t->accts.p),
                                    t->has_sroots ? static_cast<const uint8_t *>(t->sroots.p) : nullptr, n, ins_incl, del_excl, del,
                                    d_keys, d_accts, d_sr, lb, d_kind, ins_rank, m, static_cast<uint8_t *>(nk.p),
                                    static_cast<uint8_t *>(na.p), t->has_sroots ? static_cast<uint8_t *>(ns.p) : nullptr, st));
            c->launches += 6;
            CU(cudaStreamSynchronize(st));
            return B200_OK;
        };
        r = run();
        if (r != B200_OK) {
            cleanup();
            return r;
        }
        // the merged arrays become the trie's inputs; the old structure is dropped and rebuilt
        trie_free(t, marks);
        std::swap(t->keys, nk);
        std::swap(t->accts, na);
        if (t->has_sroots) std::swap(t->sroots, ns);
        cleanup();
        DevBuf *old[] = {&t->Lp, &t->nibs, &t->leaf_ref, &t->leaf_meta, &t->S, &t->E, &t->gap_sorted, &t->node_start,
                         &t->node_ref, &t->node_meta, &t->node_l, &t->node_r, &t->node_masks, &t->leaf_parent,
                         &t->node_parent, &t->dirty, &t->dirty_ids, &t->dirty_order};
        for (DevBuf *b : old) trie_free(t, *b);
        t->n = n2;
        r = trie_build_owned(t);
        if (r == B200_OK) {
            cudaError_t e = cudaMemcpy(root32, t->root.p, 32, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) r = fail(c, B200_ERR_CUDA, "root copy: %s", cudaGetErrorString(e));
        }
        if (out_rebuilt) *out_rebuilt = 1;
        if (r == B200_OK && opt_updates) {
            Built b;
            b.f = t->f;
            b.n_nodes = t->B;
            r = collect_updates(c, b, nullptr, 0, opt_updates);
        }
    }
    if (r != B200_OK && opt_updates) b200_updates_release(opt_updates);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

extern "C" B200_API int32_t b200_trie_root(b200_trie *t, uint8_t root32[32]) {
    if (!t || !root32) return B200_ERR_INVALID_ARG;
    b200_ctx *c = t->c;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}
