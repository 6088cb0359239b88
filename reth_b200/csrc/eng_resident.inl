// eng_resident.inl — resident trie (b200_trie_*): create, in-place update (wavefront), general apply (merge + rebuild).
// Part of the single translation unit engine.cu (textually included, in this order).

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
    // a storage FOREST instead of the account trie (used to seed the dynamic state, eng_dtrie.inl): `accts` then holds
    // the 32-byte slot values, seg_offsets the n_segs+1 segment bounds, seg_roots receives the n_segs storage roots
    bool forest = false;
    bool forest_account = false;  // forest of ACCOUNT tries (top-nibble buckets of a sharded state): accts / sroots as usual
    uint64_t n_segs = 0;
    DevBuf seg_offsets, seg_roots;
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
                      &t->dirty_order, &t->idx, &t->in_keys, &t->in_accts, &t->in_sroots, &t->root, &t->seg_offsets,
                      &t->seg_roots};
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
    if (t->forest) {
        TRY(trie_alloc(t, t->seg_roots, (t->n_segs ? t->n_segs : 1) * 32));
        if (t->forest_account) {
            TRY(build_forest(c, static_cast<const uint8_t *>(t->keys.p), t->n, static_cast<const uint64_t *>(t->seg_offsets.p),
                             t->n_segs, true, static_cast<const uint8_t *>(t->accts.p),
                             t->has_sroots ? static_cast<const uint8_t *>(t->sroots.p) : nullptr, true, b));
            CU(launch_segment_roots(b.f, static_cast<const uint64_t *>(t->seg_offsets.p), t->n_segs,
                                    static_cast<uint8_t *>(t->seg_roots.p), c->stream));
            c->launches++;
        } else {
            TRY(storage_roots_on_device(c, static_cast<const uint8_t *>(t->keys.p), static_cast<const uint8_t *>(t->accts.p),
                                        static_cast<const uint64_t *>(t->seg_offsets.p), t->n_segs, t->n,
                                        static_cast<uint8_t *>(t->seg_roots.p), true, b));
        }
    } else {
        TRY(account_root_on_device(c, static_cast<const uint8_t *>(t->keys.p), static_cast<const uint8_t *>(t->accts.p),
                                   t->has_sroots ? static_cast<const uint8_t *>(t->sroots.p) : nullptr, t->n,
                                   static_cast<uint8_t *>(t->root.p), true, b));
    }
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

// caller holds the context lock
static int32_t trie_create_locked(b200_ctx *c, const void *keys, const void *accts, const void *sroots, uint64_t n,
                                  cudaMemcpyKind kind, b200_trie **out, void *root_out) {
    *out = nullptr;
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

// A storage forest kept resident the same way (seed of the dynamic state): slot keys / values / segment table on the host
// or the device (`kind`), roots of all segments in (*out)->seg_roots.  Caller holds the context lock.
static int32_t forest_create_locked(b200_ctx *c, const void *slot_keys, const void *values, const void *seg_offsets,
                                    uint64_t n_segs, uint64_t n_slots, cudaMemcpyKind kind, b200_trie **out) {
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    b200_trie *t = new b200_trie();
    t->c = c;
    t->n = n_slots;
    t->forest = true;
    t->n_segs = n_segs;
    int32_t r = B200_OK;
    auto put = [&](DevBuf &b, const void *src, size_t bytes) -> int32_t {
        TRY(trie_alloc(t, b, bytes ? bytes : 16));
        if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, kind, c->stream));
        return B200_OK;
    };
    r = put(t->keys, slot_keys, n_slots * 32);
    if (r == B200_OK) r = put(t->accts, values, n_slots * 32);
    if (r == B200_OK) r = put(t->seg_offsets, seg_offsets, (n_segs + 1) * 8);
    if (r == B200_OK) r = trie_build_owned(t);
    if (r != B200_OK) {
        b200_trie_destroy(t);
        return r;
    }
    *out = t;
    return B200_OK;
}

// The accounts of one shard as a forest of 16 top-nibble bucket tries (seg_offsets: 17 bucket bounds on the host).
static int32_t bucket_forest_create_locked(b200_ctx *c, const void *keys, const void *accts, const void *sroots, uint64_t n,
                                           const uint64_t *bucket_offsets17, cudaMemcpyKind kind, b200_trie **out) {
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    b200_trie *t = new b200_trie();
    t->c = c;
    t->n = n;
    t->forest = t->forest_account = true;
    t->n_segs = 16;
    t->has_sroots = sroots != nullptr;
    int32_t r = B200_OK;
    auto put = [&](DevBuf &b, const void *src, size_t bytes, cudaMemcpyKind k) -> int32_t {
        TRY(trie_alloc(t, b, bytes ? bytes : 16));
        if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, k, c->stream));
        return B200_OK;
    };
    r = put(t->keys, keys, n * 32, kind);
    if (r == B200_OK) r = put(t->accts, accts, n * 72, kind);
    if (r == B200_OK && sroots) r = put(t->sroots, sroots, n * 32, kind);
    if (r == B200_OK) r = put(t->seg_offsets, bucket_offsets17, 17 * 8, cudaMemcpyHostToDevice);
    if (r == B200_OK) r = trie_build_owned(t);
    if (r != B200_OK) {
        b200_trie_destroy(t);
        return r;
    }
    *out = t;
    return B200_OK;
}

static int32_t trie_create_common(b200_ctx *c, const void *keys, const void *accts, const void *sroots, uint64_t n,
                                  cudaMemcpyKind kind, b200_trie **out, void *root_out) {
    if (!c || !out || (n && (!keys || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    return trie_create_locked(c, keys, accts, sroots, n, kind, out, root_out);
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
    static const uint64_t two_stage_min = [] {  // B200_WAVEFRONT_TWO_STAGE_MIN overrides the switch-over (tuning)
        const char *e = getenv("B200_WAVEFRONT_TWO_STAGE_MIN");
        return e ? strtoull(e, nullptr, 10) : (uint64_t)WARP_LEVEL_MAX;
    }();
    if (m > two_stage_min)
        for (int d = 0; d < 64; d++)
            if (std::min<uint64_t>(m, t->level_count[d]) > two_stage_min) {
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
    if (ps[0] != B200_DEVERR_NONE) return report_dev_error_now(c, (int)ps[0]);
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
