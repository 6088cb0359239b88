// eng_items.inl — b200_root_from_items: the incremental root when the trie is NOT resident on the device.
// Part of the single translation unit engine.cu (textually included, in this order).
//
// reth computes an incremental root by walking the stored trie nodes next to the changed keys: TrieWalker descends only
// where the prefix set says something changed (PrefixSet::contains, crates/trie/common/src/prefix_set.rs:205-231;
// walker.rs:161-202 update_skip_node / advance), TrieNodeIter (node_iter.rs:200-304) turns that into a stream of
// `Branch(path, stored hash, children_are_in_trie)` for every subtree it may skip and `Leaf(key, value)` for everything else,
// and HashBuilder folds the stream (trie.rs:247-309 for accounts, :659-698 for storage).  The walk is cursor work over the
// database and stays with the host (reth's own walker, or the mirror in reth_b200/walker.py); the fold — every RLP and
// every keccak — is this call: the items of one trie, or of a forest of storage tries, in key order.
static int32_t items_on_device(b200_ctx *c, const uint8_t *d_keys, const uint8_t *d_nibs, const uint8_t *d_flags,
                               const uint8_t *d_values, const uint8_t *d_sroots, const uint64_t *d_offs, uint64_t n_segs,
                               uint64_t n, bool account, uint8_t *d_roots, bool retain, Built &b) {
    ItemLeavesDev it{d_nibs, d_flags, account ? 1 : 0};
    TRY(build_forest(c, d_keys, n, d_offs, d_offs ? n_segs : 0, account, d_values, d_sroots, retain, b, nullptr, &it));
    CU(launch_segment_roots(b.f, d_offs, d_offs ? n_segs : 1, d_roots, c->stream));
    c->launches++;
    c->stats.leaves_added += n;
    c->stats.branches_added += b.n_nodes;
    c->stats.levels += b.levels;
    return B200_OK;
}

extern "C" B200_API int32_t b200_root_from_items(b200_ctx *c, const uint8_t *keys32, const uint8_t *key_nibbles,
                                                 const uint8_t *item_flags, const uint8_t *values, const uint8_t *storage_roots32,
                                                 const uint64_t *seg_offsets, uint64_t n_segs, uint64_t n_items, int32_t account,
                                                 uint8_t *roots32, b200_updates *opt_updates, b200_stats *opt_stats) {
    if (!c || !roots32 || (n_items && (!keys32 || !key_nibbles || !item_flags || !values)))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (account && seg_offsets) return fail(c, B200_ERR_INVALID_ARG, "the account trie is a single trie (seg_offsets must be NULL)");
    if (opt_updates) memset(opt_updates, 0, sizeof *opt_updates);
    if (seg_offsets) {
        TRY(check_offsets_host(c, seg_offsets, n_segs));
        if (seg_offsets[n_segs] != n_items) return fail(c, B200_ERR_INVALID_ARG, "seg_offsets must end at n_items");
    }
    for (uint64_t i = 0; i < n_items; i++)
        if (key_nibbles[i] > 64) return fail(c, B200_ERR_INVALID_ARG, "key_nibbles[%llu] > 64", (unsigned long long)i);
    const uint64_t stride = account ? sizeof(b200_account) : 32, tries = seg_offsets ? n_segs : 1;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(h2d(c, c->in_a, keys32, n_items * 32));
    TRY(h2d(c, c->in_b, values, n_items * stride));
    if (seg_offsets) TRY(h2d(c, c->in_c, seg_offsets, (n_segs + 1) * 8));
    TRY(h2d(c, c->in_d, key_nibbles, n_items));
    TRY(h2d(c, c->in_e, item_flags, n_items));
    DevBuf &srb = c->sort_out;  // (free here: no sort runs inside a build)
    if (account && storage_roots32) TRY(h2d(c, srb, storage_roots32, n_items * 32));
    ENSURE(sroots, (tries ? tries : 1) * 32);
    TRY(reset_build_state(c));
    Built b;
    TRY(items_on_device(c, static_cast<const uint8_t *>(c->in_a.p), static_cast<const uint8_t *>(c->in_d.p),
                        static_cast<const uint8_t *>(c->in_e.p), static_cast<const uint8_t *>(c->in_b.p),
                        account && storage_roots32 ? static_cast<const uint8_t *>(srb.p) : nullptr,
                        seg_offsets ? static_cast<const uint64_t *>(c->in_c.p) : nullptr, n_segs, n_items, account != 0,
                        static_cast<uint8_t *>(c->sroots.p), opt_updates != nullptr, b));
    TRY(finish_build_state(c));
    if (tries) CU(cudaMemcpyAsync(roots32, c->sroots.p, tries * 32, cudaMemcpyDeviceToHost, c->stream));
    int32_t r = sync_and_status(c);
    if (r == B200_OK && opt_updates)
        r = collect_updates(c, b, seg_offsets ? static_cast<const uint64_t *>(c->in_c.p) : nullptr, seg_offsets ? n_segs : 0, opt_updates);
    if (r != B200_OK && opt_updates) b200_updates_release(opt_updates);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}
