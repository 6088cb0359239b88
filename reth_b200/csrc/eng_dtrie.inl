// eng_dtrie.inl — b200_dtrie_*: the account trie as one dynamic arena (eng_darena.inl).
// Part of the single translation unit engine.cu (textually included, in this order).

// ------------------------------------------------------------------------------------------------ b200_dtrie: accounts only
struct b200_dtrie {
    b200_ctx *c = nullptr;
    uint64_t bytes = 0;
    DArena a;
    DevBuf root, in_keys, in_accts, in_sroots, in_present;
};

static void dbuf_free(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b = DevBuf{};
}

extern "C" B200_API void b200_dtrie_destroy(b200_dtrie *t) {
    if (!t) return;
    cudaSetDevice(t->c->device);
    cudaStreamSynchronize(t->c->stream);
    da_free(&t->a);
    dbuf_free(t->root);
    dbuf_free(t->in_keys);
    dbuf_free(t->in_accts);
    dbuf_free(t->in_sroots);
    dbuf_free(t->in_present);
    delete t;
}
extern "C" B200_API uint64_t b200_dtrie_device_bytes(const b200_dtrie *t) { return t ? t->bytes : 0; }
extern "C" B200_API uint64_t b200_dtrie_leaves(const b200_dtrie *t) { return t ? t->a.n_leaves : 0; }
extern "C" B200_API uint64_t b200_dtrie_nodes(const b200_dtrie *t) { return t ? t->a.node_alloc : 0; }

extern "C" B200_API int32_t b200_dtrie_root(b200_dtrie *t, uint8_t root32[32]) {
    if (!t || !root32) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

// Built with the level-synchronous builder (every digest comes from there), then converted: ids carry over.
static int32_t dtrie_create_common(b200_ctx *c, const void *acct_keys32, const void *accts, const void *storage_roots32,
                                   uint64_t n, cudaMemcpyKind kind, b200_dtrie **out, void *root32) {
    if (!c || !out || (n && (!acct_keys32 || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    *out = nullptr;
    std::lock_guard<std::mutex> g(c->mu);
    b200_trie *src = nullptr;
    TRY(trie_create_locked(c, acct_keys32, accts, storage_roots32, n, kind, &src, nullptr));
    cudaStream_t st = c->stream;
    b200_dtrie *t = new b200_dtrie();
    t->c = c;
    t->a.c = c;
    t->a.bytes = &t->bytes;
    t->a.account = true;
    t->a.has_sroots = storage_roots32 != nullptr;
    auto body = [&]() -> int32_t {
        TRY(da_resize(&t->a, t->root, 64, 0, -1));
        t->a.top_out = static_cast<uint8_t *>(t->root.p);
        t->a.top_stride = 0;
        TRY(da_from_build(&t->a, src, 1));
        CU(cudaMemcpyAsync(t->root.p, src->root.p, 32, cudaMemcpyDeviceToDevice, st));
        if (root32)
            CU(cudaMemcpyAsync(root32, t->root.p, 32,
                               kind == cudaMemcpyHostToDevice ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
        CU(cudaStreamSynchronize(st));
        return B200_OK;
    };
    int32_t r = body();
    b200_trie_destroy(src);
    if (r != B200_OK) {
        b200_dtrie_destroy(t);
        return r;
    }
    *out = t;
    return B200_OK;
}

extern "C" B200_API int32_t b200_dtrie_create(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                              const uint8_t *storage_roots32, uint64_t n, b200_dtrie **out,
                                              uint8_t root32[32]) {
    return dtrie_create_common(c, acct_keys32, accts, storage_roots32, n, cudaMemcpyHostToDevice, out, root32);
}
// inputs (and the optional root output) in device memory
extern "C" B200_API int32_t b200_dtrie_create_dev(b200_ctx *c, const void *d_acct_keys32, const void *d_accts,
                                                  const void *d_storage_roots32, uint64_t n, b200_dtrie **out,
                                                  void *d_root32) {
    return dtrie_create_common(c, d_acct_keys32, d_accts, d_storage_roots32, n, cudaMemcpyDeviceToDevice, out, d_root32);
}

static int32_t h2d_into(DArena *a, DevBuf &b, const void *src, size_t bytes, cudaMemcpyKind kind = cudaMemcpyHostToDevice) {
    b200_ctx *c = a->c;
    TRY(da_scratch(a, b, bytes));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, kind, c->stream));
    return B200_OK;
}

// A block's HashedPostStateSorted-shaped dirty set (keys ascending; present[i] = 0 deletes, NULL = all upserts) applied
// in place.  opt_updated receives the re-hashed stored nodes, opt_removed the paths of stored nodes that ceased to
// exist — together reth's TrieUpdates{account_nodes, removed_nodes} for the block.
extern "C" B200_API int32_t b200_dtrie_apply(b200_dtrie *t, const uint8_t *keys32, const b200_account *accts,
                                             const uint8_t *present, const uint8_t *storage_roots32, uint64_t m,
                                             uint8_t root32[32], b200_updates *opt_updated, b200_updates *opt_removed,
                                             b200_stats *opt_stats) {
    if (!t || !root32 || (m && (!keys32 || !accts))) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    if (opt_updated) memset(opt_updated, 0, sizeof *opt_updated);
    if (opt_removed) memset(opt_removed, 0, sizeof *opt_removed);
    if (m >= (1ull << 28)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^28-1 dirty keys per apply");
    std::lock_guard<std::mutex> lock(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    DArena *a = &t->a;
    if (storage_roots32 && !a->has_sroots) return fail(c, B200_ERR_INVALID_ARG, "trie was created without storage roots");
    TRY(reset_build_state(c));
    a->n_built = a->n_removed = 0;
    if (m) {
        TRY(da_prepare(a, m, 1));
        TRY(h2d_into(a, t->in_keys, keys32, m * 32));
        TRY(h2d_into(a, t->in_accts, accts, m * 72));
        if (present) TRY(h2d_into(a, t->in_present, present, m));
        if (storage_roots32) TRY(h2d_into(a, t->in_sroots, storage_roots32, m * 32));
        TRY(da_restructure(a, nullptr, static_cast<const uint8_t *>(t->in_keys.p), static_cast<const uint8_t *>(t->in_accts.p),
                           present ? static_cast<const uint8_t *>(t->in_present.p) : nullptr,
                           storage_roots32 ? static_cast<const uint8_t *>(t->in_sroots.p) : nullptr, m));
        TRY(da_rehash(a, m));
        c->stats.leaves_added += m;
    }
    TRY(finish_build_state(c));
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    if (m) TRY(da_pull_counters(a, ps + 256));
    CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, st));
    TRY(sync_and_status(c));
    if (m) da_take_counters(a, ps + 256);
    c->stats.branches_added = a->n_built;
    if (opt_updated || opt_removed) {
        b200_updates tmp{};
        b200_updates *upd = opt_updated ? opt_updated : &tmp;
        int32_t r = da_collect_updates(a, upd);
        if (r == B200_OK && opt_removed) r = da_collect_removed(a, upd, opt_removed);
        if (!opt_updated) b200_updates_release(&tmp);
        if (r != B200_OK) {
            if (opt_updated) b200_updates_release(opt_updated);
            if (opt_removed) b200_updates_release(opt_removed);
            return r;
        }
    }
    if (opt_stats) *opt_stats = c->stats;
    return B200_OK;
}

