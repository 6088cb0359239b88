// eng_keccak.inl — key hashing entry points (fixed / variable length, hash+sort for the hashing stages).
// Part of the single translation unit engine.cu (textually included, in this order).

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

static uint64_t keccak_host_chunk() {  // messages per chunk of the host-pointer paths (B200_KECCAK_CHUNK overrides: tuning, tests)
    static const uint64_t CHUNK = [] {
        const char *e = getenv("B200_KECCAK_CHUNK");
        uint64_t v = e ? strtoull(e, nullptr, 10) : 0;
        return v >= 1024 ? v : (1ull << 19);
    }();
    return CHUNK;
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
    const uint64_t CHUNK = keccak_host_chunk();
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
                               DevBuf &keys_a, DevBuf &keys_b, DevBuf &idx_a, DevBuf &flag, bool allow_equal = false);

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

// Host messages -> device digests (n x 32, d_digests) with the H2D copy of chunk k+1 under the hashing of chunk k: the
// hashing stages' inputs never sit in HBM as a whole, and the copy engine is busy from the first byte on.  The copy streams
// start after whatever is queued on c->stream and c->stream continues after the last chunk.
static int32_t hash_from_host(b200_ctx *c, const uint8_t *in, uint32_t msg_len, uint32_t stride, uint64_t n,
                              uint8_t *d_digests) {
    if (n == 0) return B200_OK;
    const uint64_t chunk = n < keccak_host_chunk() ? n : keccak_host_chunk();
    for (int i = 0; i < 3; i++) TRY(ensure(c, c->chunk_in[i], chunk * stride));
    while (c->chunk_events.size() < 4) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_events.push_back(e);
    }
    CU(cudaEventRecord(c->chunk_events[3], c->stream));
    for (int i = 0; i < 3; i++) CU(cudaStreamWaitEvent(c->copy_streams[i], c->chunk_events[3], 0));
    int slot = 0;
    for (uint64_t lo = 0; lo < n; lo += chunk, slot = (slot + 1) % 3) {
        uint64_t m = n - lo < chunk ? n - lo : chunk;
        cudaStream_t st = c->copy_streams[slot];
        CU(cudaMemcpyAsync(c->chunk_in[slot].p, in + lo * stride, (m - 1) * (size_t)stride + msg_len, cudaMemcpyHostToDevice,
                           st));
        CU(launch_keccak256_fixed(c->chunk_in[slot].p, msg_len, stride, m, d_digests + lo * 32, st, &c->launches));
    }
    for (int i = 0; i < 3; i++) {
        CU(cudaEventRecord(c->chunk_events[i], c->copy_streams[i]));
        CU(cudaStreamWaitEvent(c->stream, c->chunk_events[i], 0));
    }
    return B200_OK;
}

extern "C" B200_API int32_t b200_hash_sort_keys(b200_ctx *c, const uint8_t *in, uint32_t msg_len, uint32_t stride, uint64_t n,
                                       uint8_t *out_sorted32, uint32_t *out_perm) {
    if (!c || (n && (!in || !out_sorted32 || !out_perm)) || stride < msg_len)
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return B200_OK;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ENSURE(out_a, n * 32);
    ENSURE(sort_out, n * 32);
    ENSURE(sort_perm, n * 4);
    TRY(hash_from_host(c, in, msg_len, stride, n, static_cast<uint8_t *>(c->out_a.p)));
    TRY(sort_digests_on_device(c, c->out_a.p, n, c->sort_out.p, static_cast<uint32_t *>(c->sort_perm.p), c->sort_ka,
                               c->sort_kb, c->sort_ia, c->sort_flag));
    CU(cudaMemcpyAsync(out_sorted32, c->sort_out.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(out_perm, c->sort_perm.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

int32_t sort_composite_on_device(b200_ctx *c, const void *d_ha, uint32_t n_addr, const uint32_t *d_addr_index,
                                 const void *d_hs, uint64_t n, void *d_sorted, uint32_t *d_perm, DevBuf &keys_a,
                                 DevBuf &keys_b, DevBuf &idx_a, DevBuf &flag, bool allow_equal = false);

// The same entirely on the device (inputs and outputs in HBM).
extern "C" B200_API int32_t b200_hash_sort_storage_dev(b200_ctx *c, const void *d_addresses20, uint32_t n_addr,
                                                       const void *d_addr_index, const void *d_slots32, uint64_t n,
                                                       void *d_sorted64, void *d_perm) {
    if (!c || (n && (!d_addresses20 || !d_addr_index || !d_slots32 || !d_sorted64 || !d_perm)) || (n && !n_addr))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n == 0) return B200_OK;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ENSURE(in_d, (size_t)n_addr * 32);  // address digests
    ENSURE(out_a, n * 32);              // slot digests
    CU(launch_keccak256_fixed(d_addresses20, 20, 20, n_addr, c->in_d.p, c->stream, &c->launches));
    CU(launch_keccak256_fixed(d_slots32, 32, 32, n, c->out_a.p, c->stream, &c->launches));
    return sort_composite_on_device(c, c->in_d.p, n_addr, static_cast<const uint32_t *>(d_addr_index), c->out_a.p, n, d_sorted64,
                                    static_cast<uint32_t *>(d_perm), c->sort_ka, c->sort_kb, c->sort_ia, c->sort_flag);
}

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
    ENSURE(in_c, n * 4);
    ENSURE(in_d, (size_t)n_addr * 32);  // address digests
    ENSURE(out_a, n * 32);              // slot digests
    ENSURE(sort_out, n * 64);
    ENSURE(sort_perm, n * 4);
    CU(cudaMemcpyAsync(c->in_a.p, addresses20, (size_t)n_addr * 20, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_c.p, addr_index, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(launch_keccak256_fixed(c->in_a.p, 20, 20, n_addr, c->in_d.p, c->stream, &c->launches));
    TRY(hash_from_host(c, slots32, 32, 32, n, static_cast<uint8_t *>(c->out_a.p)));
    TRY(sort_composite_on_device(c, c->in_d.p, n_addr, static_cast<const uint32_t *>(c->in_c.p), c->out_a.p, n,
                                 c->sort_out.p, static_cast<uint32_t *>(c->sort_perm.p), c->sort_ka, c->sort_kb,
                                 c->sort_ia, c->sort_flag));
    CU(cudaMemcpyAsync(out_sorted64, c->sort_out.p, n * 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(out_perm, c->sort_perm.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}
