// eng_frontier.inl — multi-GPU: subtrie frontier and root-from-frontier entry points.
// Part of the single translation unit engine.cu (textually included, in this order).

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
