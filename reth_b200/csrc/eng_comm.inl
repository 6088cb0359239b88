// eng_comm.inl — the multi-GPU exchange steps behind the C ABI: a communicator handle (NCCL over NVLink / NVSwitch), the
// sharded state root (frontier -> all-gather -> root in one call) and the hash-partition step of the hashing stages.
// Part of the single translation unit engine.cu (textually included, in this order).
//
// a rank owns whole top-nibble buckets of the hashed address space with the storage tries of its accounts.
// The path has exactly two exchange steps: (1) one all-gather of the 16-entry subtrie frontier (68 bytes per entry) before
// the root — ParallelStateRoot's channel of storage roots, crates/trie/parallel/src/root.rs:101-197, turned sideways — and
// (2) when the input arrives unhashed and unpartitioned (AccountHashingStage / StorageHashingStage at N > 1,
// hashing_account.rs:176-238, hashing_storage.rs:106-178), one bucketed all-to-all of (digest, row) by owner rank before
// the sort.  NCCL is loaded at run time (dlopen of libnccl.so.2: the library has no link-time dependency on it and shares
// the copy a host process — e.g. torch — already has).  One process (or thread) per GPU; the unique id travels over
// whatever channel the host has.
#if !defined(B200_NO_NCCL) && defined(__has_include)
#if __has_include(<nccl.h>)
#define B200_HAVE_NCCL 1
#endif
#endif

#ifdef B200_HAVE_NCCL
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *override_path = getenv("B200_NCCL_LIB");
        const char *names[] = {override_path, "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        auto sym = [&](const char *n) { return dlsym(api.lib, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}
}  // namespace

struct b200_comm {
    b200_ctx *c = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    DevBuf mine, all, merged;            // frontier exchange
    DevBuf p_dig, p_owner, p_perm, p_owner2, p_iota, p_send_d, p_send_v, p_recv_d, p_recv_v, p_counts, p_allcounts, p_sortperm;
};

#define NC(call)                                                                                                   \
    do {                                                                                                           \
        ncclResult_t r__ = (call);                                                                                 \
        if (r__ != ncclSuccess) return fail(c, B200_ERR_CUDA, "%s: %s", #call, nccl_api().GetErrorString(r__));  \
    } while (0)

static_assert(sizeof(ncclUniqueId) == B200_COMM_ID_BYTES, "ncclUniqueId size");

extern "C" B200_API int32_t b200_comm_unique_id(uint8_t id[B200_COMM_ID_BYTES]) {
    if (!id || !nccl_api().ok) return B200_ERR_CUDA;
    ncclUniqueId u;
    if (nccl_api().GetUniqueId(&u) != ncclSuccess) return B200_ERR_CUDA;
    memcpy(id, &u, sizeof u);
    return B200_OK;
}

extern "C" B200_API int32_t b200_comm_create(b200_ctx *c, const uint8_t id[B200_COMM_ID_BYTES], int32_t n_ranks, int32_t rank,
                                             b200_comm **out) {
    if (!c || !id || !out || n_ranks < 1 || n_ranks > 16 || rank < 0 || rank >= n_ranks)
        return fail(c, B200_ERR_INVALID_ARG, "bad argument (1..16 ranks: one per top-nibble bucket at most)");
    if (!nccl_api().ok) return fail(c, B200_ERR_CUDA, "NCCL not available (libnccl.so.2 could not be loaded; B200_NCCL_LIB overrides)");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    b200_comm *m = new b200_comm();
    m->c = c;
    m->world = n_ranks;
    m->rank = rank;
    ncclResult_t r = nccl_api().CommInitRank(&m->comm, n_ranks, u, rank);
    if (r != ncclSuccess) {
        delete m;
        return fail(c, B200_ERR_CUDA, "ncclCommInitRank: %s", nccl_api().GetErrorString(r));
    }
    *out = m;
    return B200_OK;
}

extern "C" B200_API void b200_comm_destroy(b200_comm *m) {
    if (!m) return;
    b200_ctx *c = m->c;
    {
        std::lock_guard<std::mutex> g(c->mu);
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (m->comm) nccl_api().CommDestroy(m->comm);
        for (DevBuf *b : {&m->mine, &m->all, &m->merged, &m->p_dig, &m->p_owner, &m->p_perm, &m->p_owner2, &m->p_iota, &m->p_send_d,
                          &m->p_send_v, &m->p_recv_d, &m->p_recv_v, &m->p_counts, &m->p_allcounts, &m->p_sortperm})
            if (b->p) {
                cudaFree(b->p);
                c->dev_bytes -= b->cap;
            }
    }
    delete m;
}
extern "C" B200_API int32_t b200_comm_rank(const b200_comm *m) { return m ? m->rank : -1; }
extern "C" B200_API int32_t b200_comm_size(const b200_comm *m) { return m ? m->world : 0; }

// frontier of this rank's buckets -> all-gather -> merge -> root, all on the ctx stream; d_root32 receives the state root on
// every rank
static int32_t sharded_root_on_device(b200_comm *m, const uint8_t *d_akeys, const uint8_t *d_accts, uint64_t n_accounts,
                                      const uint8_t *d_skeys, const uint8_t *d_svals, const uint64_t *d_offs, uint64_t n_slots,
                                      uint8_t *d_root32) {
    b200_ctx *c = m->c;
    const size_t fb = 16 * sizeof(FrontierEntryDev);
    TRY(ensure(c, m->mine, fb));
    TRY(ensure(c, m->all, fb * m->world));
    TRY(ensure(c, m->merged, fb));
    TRY(frontier_on_device(c, d_akeys, d_accts, n_accounts, d_skeys, d_svals, d_offs, n_slots,
                           static_cast<FrontierEntryDev *>(m->mine.p)));
    NC(nccl_api().AllGather(m->mine.p, m->all.p, fb, ncclChar, m->comm, c->stream));
    CU(launch_merge_frontiers(static_cast<const FrontierEntryDev *>(m->all.p), m->world, static_cast<FrontierEntryDev *>(m->merged.p),
                              reinterpret_cast<int *>(small_u32(c) + SM_ERR), c->stream));
    CU(launch_root_from_frontier(static_cast<const FrontierEntryDev *>(m->merged.p), d_root32, c->stream));
    c->launches += 2;
    return B200_OK;
}

extern "C" B200_API int32_t b200_state_root_sharded_dev(b200_comm *m, const void *d_acct_keys32, const void *d_accts,
                                                        uint64_t n_accounts, const void *d_slot_keys32, const void *d_values32_be,
                                                        const void *d_seg_offsets, uint64_t n_slots, void *d_root32) {
    if (!m) return B200_ERR_INVALID_ARG;
    b200_ctx *c = m->c;
    if (!d_root32 || !d_seg_offsets || (n_accounts && (!d_acct_keys32 || !d_accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(reset_build_state(c));
    TRY(sharded_root_on_device(m, static_cast<const uint8_t *>(d_acct_keys32), static_cast<const uint8_t *>(d_accts), n_accounts,
                               static_cast<const uint8_t *>(d_slot_keys32), static_cast<const uint8_t *>(d_values32_be),
                               static_cast<const uint64_t *>(d_seg_offsets), n_slots, static_cast<uint8_t *>(d_root32)));
    return finish_build_state(c);
}

extern "C" B200_API int32_t b200_state_root_sharded(b200_comm *m, const uint8_t *acct_keys32, const b200_account *accts,
                                                    uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                                    const uint64_t *seg_offsets, uint8_t root32[32], b200_stats *opt_stats) {
    if (!m) return B200_ERR_INVALID_ARG;
    b200_ctx *c = m->c;
    if (!root32 || !seg_offsets || (n_accounts && (!acct_keys32 || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    TRY(check_offsets_host(c, seg_offsets, n_accounts));
    const uint64_t n_slots = seg_offsets[n_accounts];
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(h2d(c, c->in_a, slot_keys32, n_slots * 32));
    TRY(h2d(c, c->in_b, values32_be, n_slots * 32));
    TRY(h2d(c, c->in_c, seg_offsets, (n_accounts + 1) * 8));
    TRY(h2d(c, c->in_d, acct_keys32, n_accounts * 32));
    TRY(h2d(c, c->in_e, accts, n_accounts * sizeof(b200_account)));
    ENSURE(out_a, 64);
    TRY(reset_build_state(c));
    TRY(sharded_root_on_device(m, static_cast<const uint8_t *>(c->in_d.p), static_cast<const uint8_t *>(c->in_e.p), n_accounts,
                               static_cast<const uint8_t *>(c->in_a.p), static_cast<const uint8_t *>(c->in_b.p),
                               static_cast<const uint64_t *>(c->in_c.p), n_slots, static_cast<uint8_t *>(c->out_a.p)));
    TRY(finish_build_state(c));
    CU(cudaMemcpyAsync(root32, c->out_a.p, 32, cudaMemcpyDeviceToHost, c->stream));
    int32_t r = sync_and_status(c);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

// The sharded dynamic state (b200_dstate_create_sharded): after every rank has applied its part of a block, one all-gather of
// the resident frontiers and the root — the per-block exchange of the live path at N > 1.
extern "C" B200_API int32_t b200_dstate_root_sharded(b200_comm *m, b200_dstate *t, uint8_t root32[32]) {
    if (!m || !t || !root32) return B200_ERR_INVALID_ARG;
    b200_ctx *c = m->c;
    if (t->c != c) return fail(c, B200_ERR_INVALID_ARG, "the state and the communicator belong to different contexts");
    if (!t->sharded) return fail(c, B200_ERR_INVALID_ARG, "not a sharded state (b200_dstate_create_sharded)");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t fb = 16 * sizeof(FrontierEntryDev);
    TRY(ensure(c, m->all, fb * m->world));
    TRY(ensure(c, m->merged, fb + 64));
    NC(nccl_api().AllGather(t->frontier.p, m->all.p, fb, ncclChar, m->comm, c->stream));
    CU(launch_merge_frontiers(static_cast<const FrontierEntryDev *>(m->all.p), m->world, static_cast<FrontierEntryDev *>(m->merged.p),
                              reinterpret_cast<int *>(small_u32(c) + SM_ERR), c->stream));
    uint8_t *d_root = static_cast<uint8_t *>(m->merged.p) + align_up(fb, 16);
    CU(launch_root_from_frontier(static_cast<const FrontierEntryDev *>(m->merged.p), d_root, c->stream));
    c->launches += 2;
    CU(cudaMemcpyAsync(root32, d_root, 32, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_status(c);
}

// keccak of this rank's n messages, every (digest, row) sent to the rank that owns the digest's top nibble, what arrives
// sorted by digest.  d_sorted_keys32 / d_sorted_values must hold `capacity` rows; *n_out = rows this rank now owns.
extern "C" B200_API int32_t b200_hash_partition_dev(b200_comm *m, const void *d_in, uint32_t msg_len, uint32_t stride, uint64_t n,
                                                    const void *d_values, uint32_t value_bytes, uint64_t capacity,
                                                    void *d_sorted_keys32, void *d_sorted_values, uint64_t *n_out) {
    if (!m) return B200_ERR_INVALID_ARG;
    b200_ctx *c = m->c;
    if (!n_out || (n && !d_in) || (value_bytes && n && !d_values) || !d_sorted_keys32 || (value_bytes && !d_sorted_values) ||
        (msg_len != 20 && msg_len != 32) || stride < msg_len || n >= (1ull << 31) || value_bytes > 256)
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int W = m->world;
    const uint64_t n1 = n ? n : 1;
    TRY(ensure(c, m->p_dig, n1 * 32));
    TRY(ensure(c, m->p_owner, n1));
    TRY(ensure(c, m->p_owner2, n1));
    TRY(ensure(c, m->p_iota, n1 * 4));
    TRY(ensure(c, m->p_perm, n1 * 4));
    TRY(ensure(c, m->p_send_d, n1 * 32));
    TRY(ensure(c, m->p_send_v, n1 * (value_bytes ? value_bytes : 1)));
    TRY(ensure(c, m->p_counts, 16 * 8));
    TRY(ensure(c, m->p_allcounts, 16 * 8 * (size_t)W));
    // 1. hash, owner rank of every digest, rows in destination order (stable: the order inside a destination is the input order)
    phase_mark(c, "start");
    CU(cudaMemsetAsync(m->p_counts.p, 0, 16 * 8, st));
    if (n) {
        CU(launch_keccak256_fixed(d_in, msg_len, stride, n, m->p_dig.p, st, &c->launches));
        phase_mark(c, "p:hash");
        CU(launch_partition_owner(static_cast<const uint8_t *>(m->p_dig.p), n, W, static_cast<uint8_t *>(m->p_owner.p),
                                  static_cast<unsigned long long *>(m->p_counts.p), st));
        CU(launch_iota(static_cast<uint32_t *>(m->p_iota.p), n, 0, st));
        size_t t = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, t, static_cast<uint8_t *>(m->p_owner.p), static_cast<uint8_t *>(m->p_owner2.p),
                                           static_cast<uint32_t *>(m->p_iota.p), static_cast<uint32_t *>(m->p_perm.p), (int64_t)n, 0, 4, st));
        ENSURE(cub_temp, t);
        CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, t, static_cast<uint8_t *>(m->p_owner.p), static_cast<uint8_t *>(m->p_owner2.p),
                                           static_cast<uint32_t *>(m->p_iota.p), static_cast<uint32_t *>(m->p_perm.p), (int64_t)n, 0, 4, st));
        CU(launch_partition_gather(static_cast<const uint8_t *>(m->p_dig.p), static_cast<const uint8_t *>(d_values), value_bytes,
                                   static_cast<const uint32_t *>(m->p_perm.p), n, static_cast<uint8_t *>(m->p_send_d.p),
                                   static_cast<uint8_t *>(m->p_send_v.p), st));
        c->launches += 4;
        phase_mark(c, "p:owner+order+gather");
    }
    // 2. everybody learns everybody's per-destination counts (16 x u64 per rank)
    NC(nccl_api().AllGather(m->p_counts.p, m->p_allcounts.p, 16 * 8, ncclChar, m->comm, st));
    unsigned long long *hc = reinterpret_cast<unsigned long long *>(static_cast<uint8_t *>(c->pinned_small) + 2048);
    CU(cudaMemcpyAsync(hc, m->p_allcounts.p, 16 * 8 * (size_t)W, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint64_t send_off[17] = {0}, recv_off[17] = {0};
    for (int r = 0; r < W; r++) {
        send_off[r + 1] = send_off[r] + hc[16 * m->rank + r];
        recv_off[r + 1] = recv_off[r] + hc[16 * r + m->rank];
    }
    phase_mark(c, "p:counts");
    const uint64_t n_recv = recv_off[W];
    *n_out = n_recv;
    if (n_recv > capacity) return fail(c, B200_ERR_INVALID_ARG, "hash_partition: %llu rows arrive, capacity is %llu",
                                        (unsigned long long)n_recv, (unsigned long long)capacity);
    const uint64_t r1 = n_recv ? n_recv : 1;
    TRY(ensure(c, m->p_recv_d, r1 * 32));
    TRY(ensure(c, m->p_recv_v, r1 * (value_bytes ? value_bytes : 1)));
    TRY(ensure(c, m->p_sortperm, r1 * 4));
    // 3. the all-to-all: one grouped send / recv pair per peer (NVLink / NVSwitch: every pair at full bandwidth)
    NC(nccl_api().GroupStart());
    for (int r = 0; r < W; r++) {
        const uint64_t sc = send_off[r + 1] - send_off[r], rc = recv_off[r + 1] - recv_off[r];
        if (sc) {
            NC(nccl_api().Send(static_cast<uint8_t *>(m->p_send_d.p) + 32 * send_off[r], sc * 32, ncclChar, r, m->comm, st));
            if (value_bytes)
                NC(nccl_api().Send(static_cast<uint8_t *>(m->p_send_v.p) + (uint64_t)value_bytes * send_off[r], sc * value_bytes, ncclChar,
                                   r, m->comm, st));
        }
        if (rc) {
            NC(nccl_api().Recv(static_cast<uint8_t *>(m->p_recv_d.p) + 32 * recv_off[r], rc * 32, ncclChar, r, m->comm, st));
            if (value_bytes)
                NC(nccl_api().Recv(static_cast<uint8_t *>(m->p_recv_v.p) + (uint64_t)value_bytes * recv_off[r], rc * value_bytes, ncclChar,
                                   r, m->comm, st));
        }
    }
    NC(nccl_api().GroupEnd());
    phase_mark(c, "p:exchange");
    // 4. sort what arrived by digest (the ETL collector's job), rows follow their keys
    if (n_recv) {
        TRY(sort_digests_on_device(c, m->p_recv_d.p, n_recv, d_sorted_keys32, static_cast<uint32_t *>(m->p_sortperm.p), c->sort_ka,
                                   c->sort_kb, c->sort_ia, c->sort_flag, true));
        CU(launch_gather_values(static_cast<const uint8_t *>(m->p_recv_v.p), value_bytes, static_cast<const uint32_t *>(m->p_sortperm.p),
                                n_recv, static_cast<uint8_t *>(d_sorted_values), st));
        c->launches++;
    }
    phase_mark(c, "p:sort");
    CU(cudaStreamSynchronize(st));
    if (c->phase_timing) phase_report(c);
    return B200_OK;
}

#else  // no NCCL at build time (the CPU emulation build): the entry points exist and say so
struct b200_comm {
    int unused;
};
static int32_t no_nccl(b200_ctx *c) { return fail(c, B200_ERR_CUDA, "built without NCCL"); }
extern "C" B200_API int32_t b200_comm_unique_id(uint8_t *) { return B200_ERR_CUDA; }
extern "C" B200_API int32_t b200_comm_create(b200_ctx *c, const uint8_t *, int32_t, int32_t, b200_comm **) { return no_nccl(c); }
extern "C" B200_API void b200_comm_destroy(b200_comm *) {}
extern "C" B200_API int32_t b200_comm_rank(const b200_comm *) { return -1; }
extern "C" B200_API int32_t b200_comm_size(const b200_comm *) { return 0; }
extern "C" B200_API int32_t b200_state_root_sharded_dev(b200_comm *, const void *, const void *, uint64_t, const void *, const void *,
                                                        const void *, uint64_t, void *) { return B200_ERR_CUDA; }
extern "C" B200_API int32_t b200_state_root_sharded(b200_comm *, const uint8_t *, const b200_account *, uint64_t, const uint8_t *,
                                                    const uint8_t *, const uint64_t *, uint8_t *, b200_stats *) { return B200_ERR_CUDA; }
extern "C" B200_API int32_t b200_hash_partition_dev(b200_comm *, const void *, uint32_t, uint32_t, uint64_t, const void *, uint32_t,
                                                    uint64_t, void *, void *, uint64_t *) { return B200_ERR_CUDA; }
extern "C" B200_API int32_t b200_dstate_root_sharded(b200_comm *, b200_dstate *, uint8_t *) { return B200_ERR_CUDA; }
#endif
