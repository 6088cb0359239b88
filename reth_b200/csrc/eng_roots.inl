// eng_roots.inl — storage_roots / state_root / state_root_full entry points (device and host pointers, pipelined H2D).
// Part of the single translation unit engine.cu (textually included, in this order).

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

// opt_*_rows: the stored nodes as finished table rows (encoded on the device) instead of / next to the records
static int32_t state_root_full_impl(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                    uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                    const uint64_t *seg_offsets, uint8_t root32[32], b200_updates *opt_account_updates,
                                    b200_updates *opt_storage_updates, int32_t key_format, b200_rows *opt_account_rows,
                                    b200_rows *opt_storage_rows, b200_stats *opt_stats) {
    if (!c || !root32 || !seg_offsets || (n_accounts && (!acct_keys32 || !accts)))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (opt_account_rows) memset(opt_account_rows, 0, sizeof *opt_account_rows);
    if (opt_storage_rows) memset(opt_storage_rows, 0, sizeof *opt_storage_rows);
    TRY(check_offsets_host(c, seg_offsets, n_accounts));
    uint64_t n_slots = seg_offsets[n_accounts];
    if (n_slots && (!slot_keys32 || !values32_be)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (opt_account_updates) memset(opt_account_updates, 0, sizeof *opt_account_updates);
    if (opt_storage_updates) memset(opt_storage_updates, 0, sizeof *opt_storage_updates);
    const bool retain = opt_account_updates || opt_storage_updates || opt_account_rows || opt_storage_rows;
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
    if (r == B200_OK && (opt_storage_updates || opt_storage_rows)) {
        r = sync_and_status(c);
        if (r == B200_OK && opt_storage_updates)
            r = collect_updates(c, bs, static_cast<const uint64_t *>(c->in_c.p), n_accounts, opt_storage_updates);
        if (r == B200_OK && opt_storage_rows)
            r = collect_rows(c, bs, static_cast<const uint64_t *>(c->in_c.p), n_accounts,
                             static_cast<const uint8_t *>(c->in_d.p), key_format, true, opt_storage_rows);
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
    if (r == B200_OK && opt_account_rows) r = collect_rows(c, ba, nullptr, 0, nullptr, key_format, false, opt_account_rows);
    if (r != B200_OK) {
        if (opt_account_updates) b200_updates_release(opt_account_updates);
        if (opt_storage_updates) b200_updates_release(opt_storage_updates);
        if (opt_account_rows) b200_rows_release(opt_account_rows);
        if (opt_storage_rows) b200_rows_release(opt_storage_rows);
    }
    if (opt_stats) *opt_stats = c->stats;
    return r;
}

extern "C" B200_API int32_t b200_state_root_full(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                        uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                        const uint64_t *seg_offsets, uint8_t root32[32],
                                        b200_updates *opt_account_updates, b200_updates *opt_storage_updates,
                                        b200_stats *opt_stats) {
    return state_root_full_impl(c, acct_keys32, accts, n_accounts, slot_keys32, values32_be, seg_offsets, root32,
                                opt_account_updates, opt_storage_updates, B200_KEYS_LEGACY, nullptr, nullptr, opt_stats);
}

extern "C" B200_API int32_t b200_state_root_full_rows(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                                      uint64_t n_accounts, const uint8_t *slot_keys32,
                                                      const uint8_t *values32_be, const uint64_t *seg_offsets,
                                                      int32_t key_format, uint8_t root32[32], b200_rows *account_rows,
                                                      b200_rows *storage_rows, b200_stats *opt_stats) {
    if (!account_rows || !storage_rows || (key_format != B200_KEYS_LEGACY && key_format != B200_KEYS_PACKED))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    return state_root_full_impl(c, acct_keys32, accts, n_accounts, slot_keys32, values32_be, seg_offsets, root32, nullptr,
                                nullptr, key_format, account_rows, storage_rows, opt_stats);
}
