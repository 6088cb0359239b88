// eng_dtrie.inl — dynamic resident trie (b200_dtrie_*): the account trie as an arena of 16-slot branch nodes in HBM that
// takes a block's upserts and deletes in place and re-hashes only the touched paths (device side: tk_dtrie.cuh).
// Part of the single translation unit engine.cu (textually included, in this order).

struct IsKind {
    uint8_t k;
    __host__ __device__ bool operator()(uint8_t x) const { return x == k; }
};

struct b200_dtrie {
    b200_ctx *c = nullptr;
    bool has_sroots = false;
    uint64_t bytes = 0;
    uint32_t lcap = 0, ncap = 0;                         // capacities
    uint32_t leaf_alloc = 0, node_alloc = 0, n_leaves = 0;  // device counters as of the last apply
    DevBuf lkey, lacct, lsroot, lref, lmeta, lparent, lseed;
    DevBuf nchild, ndepth, nparent, nref, nmeta, nmasks, nkey, npending, nseed, ncur, nnext;
    DevBuf leaf_free, node_free, g, root;
    // per-apply scratch
    DevBuf in_keys, in_accts, in_sroots, in_present, kind, leaf_of, list_a, list_b, ins_idx, attach, seeds, built, removed,
        freed_now, flags, nh, sel, prefix, pick, out;
};

static int32_t dt_resize(b200_dtrie *t, DevBuf &b, size_t new_bytes, size_t keep_bytes, bool zero_fill) {
    b200_ctx *c = t->c;
    if (new_bytes <= b.cap) return B200_OK;
    void *p = nullptr;
    size_t want = new_bytes + 256;
    CU(cudaMalloc(&p, want));
    if (zero_fill) CU(cudaMemsetAsync(p, 0, want, c->stream));
    if (b.p && keep_bytes) CU(cudaMemcpyAsync(p, b.p, keep_bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (b.p) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(b.p));
        t->bytes -= b.cap;
    }
    b.p = p;
    b.cap = want;
    t->bytes += want;
    return B200_OK;
}
static int32_t dt_scratch(b200_dtrie *t, DevBuf &b, size_t bytes) { return dt_resize(t, b, bytes ? bytes : 16, 0, false); }

// capacity for `leaves` leaf slots and `nodes` node slots, keeping the contents of the slots allocated so far
static int32_t dt_reserve(b200_dtrie *t, uint64_t leaves, uint64_t nodes) {
    b200_ctx *c = t->c;
    if (leaves >= (1ull << 31) || nodes >= (1ull << 31)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^31-1 leaves");
    if (leaves > t->lcap) {
        uint64_t cap = std::max<uint64_t>(leaves, (uint64_t)t->lcap + t->lcap / 2) + 1024;
        size_t used = t->leaf_alloc;
        TRY(dt_resize(t, t->lkey, cap * 32, used * 32, false));
        TRY(dt_resize(t, t->lacct, cap * 72, used * 72, false));
        if (t->has_sroots) TRY(dt_resize(t, t->lsroot, cap * 32, used * 32, false));
        TRY(dt_resize(t, t->lref, cap * 32, used * 32, false));
        TRY(dt_resize(t, t->lmeta, cap, used, false));
        TRY(dt_resize(t, t->lparent, cap * 4, used * 4, false));
        TRY(dt_resize(t, t->lseed, cap, used, true));
        TRY(dt_resize(t, t->leaf_free, cap * 4, (size_t)t->lcap * 4, false));
        t->lcap = (uint32_t)cap;
    }
    if (nodes > t->ncap) {
        uint64_t cap = std::max<uint64_t>(nodes, (uint64_t)t->ncap + t->ncap / 2) + 1024;
        size_t used = t->node_alloc;
        TRY(dt_resize(t, t->nchild, cap * 64, used * 64, false));
        TRY(dt_resize(t, t->ndepth, cap, used, false));
        TRY(dt_resize(t, t->nparent, cap * 4, used * 4, false));
        TRY(dt_resize(t, t->nref, cap * 32, used * 32, false));
        TRY(dt_resize(t, t->nmeta, cap, used, false));
        TRY(dt_resize(t, t->nmasks, cap * 8, used * 8, false));
        TRY(dt_resize(t, t->nkey, cap * 32, used * 32, false));
        TRY(dt_resize(t, t->npending, cap * 4, used * 4, true));
        TRY(dt_resize(t, t->nseed, cap, used, true));
        TRY(dt_resize(t, t->ncur, cap, used, true));
        TRY(dt_resize(t, t->nnext, cap, used, true));
        TRY(dt_resize(t, t->node_free, cap * 4, (size_t)t->ncap * 4, false));
        t->ncap = (uint32_t)cap;
    }
    return B200_OK;
}

static DTrieDev dt_view(b200_dtrie *t) {
    b200_ctx *c = t->c;
    DTrieDev d{};
    d.lkey = static_cast<uint8_t *>(t->lkey.p);
    d.lacct = static_cast<uint8_t *>(t->lacct.p);
    d.lsroot = t->has_sroots ? static_cast<uint8_t *>(t->lsroot.p) : nullptr;
    d.lref = static_cast<uint8_t *>(t->lref.p);
    d.lmeta = static_cast<uint8_t *>(t->lmeta.p);
    d.lparent = static_cast<uint32_t *>(t->lparent.p);
    d.lseed = static_cast<uint8_t *>(t->lseed.p);
    d.nchild = static_cast<uint32_t *>(t->nchild.p);
    d.ndepth = static_cast<uint8_t *>(t->ndepth.p);
    d.nparent = static_cast<uint32_t *>(t->nparent.p);
    d.nref = static_cast<uint8_t *>(t->nref.p);
    d.nmeta = static_cast<uint8_t *>(t->nmeta.p);
    d.nmasks = static_cast<ushort4 *>(t->nmasks.p);
    d.nkey = static_cast<uint8_t *>(t->nkey.p);
    d.npending = static_cast<uint32_t *>(t->npending.p);
    d.nseed = static_cast<uint8_t *>(t->nseed.p);
    d.ncur = static_cast<uint8_t *>(t->ncur.p);
    d.nnext = static_cast<uint8_t *>(t->nnext.p);
    d.leaf_free = static_cast<uint32_t *>(t->leaf_free.p);
    d.node_free = static_cast<uint32_t *>(t->node_free.p);
    d.seeds = static_cast<uint32_t *>(t->seeds.p);
    d.built = static_cast<uint32_t *>(t->built.p);
    d.removed = static_cast<uint32_t *>(t->removed.p);
    d.freed_now = static_cast<uint32_t *>(t->freed_now.p);
    d.g = static_cast<uint32_t *>(t->g.p);
    d.err = reinterpret_cast<int *>(small_u32(c) + SM_ERR);
    d.counters = reinterpret_cast<unsigned long long *>(small_u32(c) + SM_COUNTERS);
    d.lcap = t->lcap;
    d.ncap = t->ncap;
    return d;
}

extern "C" B200_API void b200_dtrie_destroy(b200_dtrie *t) {
    if (!t) return;
    cudaSetDevice(t->c->device);
    cudaStreamSynchronize(t->c->stream);
    DevBuf *bufs[] = {&t->lkey, &t->lacct, &t->lsroot, &t->lref, &t->lmeta, &t->lparent, &t->lseed, &t->nchild, &t->ndepth,
                      &t->nparent, &t->nref, &t->nmeta, &t->nmasks, &t->nkey, &t->npending, &t->nseed, &t->ncur, &t->nnext,
                      &t->leaf_free, &t->node_free, &t->g, &t->root, &t->in_keys, &t->in_accts, &t->in_sroots, &t->in_present,
                      &t->kind, &t->leaf_of, &t->list_a, &t->list_b, &t->ins_idx, &t->attach, &t->seeds, &t->built, &t->removed,
                      &t->freed_now, &t->flags, &t->nh, &t->sel, &t->prefix, &t->pick, &t->out};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    delete t;
}
extern "C" B200_API uint64_t b200_dtrie_device_bytes(const b200_dtrie *t) { return t ? t->bytes : 0; }
extern "C" B200_API uint64_t b200_dtrie_leaves(const b200_dtrie *t) { return t ? t->n_leaves : 0; }
extern "C" B200_API uint64_t b200_dtrie_nodes(const b200_dtrie *t) { return t ? t->node_alloc : 0; }

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
    t->has_sroots = storage_roots32 != nullptr;
    auto body = [&]() -> int32_t {
        const uint32_t B = src->B;
        TRY(dt_reserve(t, n + n / 8 + 16, (uint64_t)B + B / 8 + 16));
        TRY(dt_resize(t, t->g, DG_WORDS * 4, 0, true));
        TRY(dt_resize(t, t->root, 64, 0, false));
        if (n) {
            CU(cudaMemcpyAsync(t->lkey.p, src->keys.p, n * 32, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(t->lacct.p, src->accts.p, n * 72, cudaMemcpyDeviceToDevice, st));
            if (t->has_sroots) CU(cudaMemcpyAsync(t->lsroot.p, src->sroots.p, n * 32, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(t->lref.p, src->leaf_ref.p, n * 32, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(t->lmeta.p, src->leaf_meta.p, n, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(t->lparent.p, src->leaf_parent.p, n * 4, cudaMemcpyDeviceToDevice, st));
        }
        uint32_t *ps = static_cast<uint32_t *>(c->pinned_small) + 256;  // 64 words of the readback page
        memset(ps, 0, DG_WORDS * 4);
        ps[DG_ROOT] = n == 1 ? (0u | DT_LEAF) : DT_NONE;  // n >= 2: the convert kernel writes the root node's id
        ps[DG_NLEAVES] = (uint32_t)n;
        ps[DG_LEAF_ALLOC] = (uint32_t)n;
        ps[DG_NODE_ALLOC] = B;
        CU(cudaMemcpyAsync(t->g.p, ps, DG_WORDS * 4, cudaMemcpyHostToDevice, st));
        t->leaf_alloc = t->n_leaves = (uint32_t)n;
        t->node_alloc = B;
        DTrieDev d = dt_view(t);
        CU(launch_dt_convert_nodes(src->f, B, static_cast<const uint32_t *>(src->node_parent.p), d, st));
        c->launches++;
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

// host copy of a device record set (same block layout as gather_and_copy)
static int32_t dt_collect_updates(b200_dtrie *t, const DTrieDev &d, uint32_t n_built, b200_updates *u) {
    b200_ctx *c = t->c;
    cudaStream_t st = c->stream;
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    uint32_t n_stored = 0, n_hashes = 0;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    const uint32_t *pick_ids = nullptr, *pick_prefix = nullptr;
    if (n_built) {
        TRY(dt_scratch(t, t->flags, n_built));
        TRY(dt_scratch(t, t->nh, (size_t)n_built * 4));
        TRY(dt_scratch(t, t->sel, (size_t)n_built * 4));
        TRY(dt_scratch(t, t->prefix, ((size_t)n_built + 1) * 4));
        TRY(dt_scratch(t, t->pick, (size_t)n_built * 8));
        uint8_t *flags = static_cast<uint8_t *>(t->flags.p);
        uint32_t *nh = static_cast<uint32_t *>(t->nh.p), *sel = static_cast<uint32_t *>(t->sel.p);
        uint32_t *prefix = static_cast<uint32_t *>(t->prefix.p);
        uint32_t *ids = static_cast<uint32_t *>(t->pick.p), *pref = ids + n_built;
        uint32_t *n_stored_p = small_u32(c) + SM_NSTORED;
        CU(launch_dt_stored_flags(d, n_built, flags, nh, st));
        size_t t_sel = 0, t_scan = 0;
        thrust::counting_iterator<uint32_t> counting(0);
        CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, flags, sel, n_stored_p, (int64_t)n_built, st));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, nh, prefix, (int64_t)n_built, st));
        ENSURE(cub_temp, std::max(t_sel, t_scan));
        CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, flags, sel, n_stored_p, (int64_t)n_built, st));
        CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t_scan, nh, prefix, (int64_t)n_built, st));
        c->launches += 3;
        CU(cudaMemcpyAsync(ps + 200, n_stored_p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 201, prefix + (n_built - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 202, nh + (n_built - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        n_stored = ps[200];
        n_hashes = ps[201] + ps[202];
        CU(launch_pick_subset(d.built, prefix, sel, n_stored, ids, pref, st));
        c->launches++;
        pick_ids = ids;
        pick_prefix = pref;
    }
    size_t o_tid = 0;
    size_t o_plen = align_up(o_tid + (size_t)n_stored * 4, 16);
    size_t o_path = align_up(o_plen + n_stored, 16);
    size_t o_sm = align_up(o_path + (size_t)n_stored * 32, 16);
    size_t o_tm = align_up(o_sm + (size_t)n_stored * 2, 16);
    size_t o_hm = align_up(o_tm + (size_t)n_stored * 2, 16);
    size_t o_ho32 = align_up(o_hm + (size_t)n_stored * 2, 16);
    size_t o_hash = align_up(o_ho32 + (size_t)n_stored * 4, 16);
    size_t o_ho64 = align_up(o_hash + (size_t)n_hashes * 32, 16);
    size_t dev_total = o_ho64, host_total = o_ho64 + ((size_t)n_stored + 1) * 8;
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
        TRY(dt_scratch(t, t->out, dev_total));
        uint8_t *dv = static_cast<uint8_t *>(t->out.p);
        UpdatesDev ud;
        ud.trie_id = reinterpret_cast<uint32_t *>(dv + o_tid);
        ud.path_len = dv + o_plen;
        ud.path_packed = dv + o_path;
        ud.state_mask = reinterpret_cast<uint16_t *>(dv + o_sm);
        ud.tree_mask = reinterpret_cast<uint16_t *>(dv + o_tm);
        ud.hash_mask = reinterpret_cast<uint16_t *>(dv + o_hm);
        ud.hash_offset = reinterpret_cast<uint32_t *>(dv + o_ho32);
        ud.hashes = dv + o_hash;
        CU(launch_dt_gather_updates(d, pick_ids, n_stored, pick_prefix, ud, st));
        c->launches++;
        CU(cudaMemcpyAsync(h, dv, dev_total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const uint32_t *ho32 = reinterpret_cast<const uint32_t *>(h + o_ho32);
        for (uint32_t i = 0; i < n_stored; i++) u->hash_offset[i] = ho32[i];
    }
    u->hash_offset[n_stored] = n_hashes;
    return B200_OK;
}

// removed_nodes as records without masks or hashes; paths that are also in `updated` are dropped (updated nodes take
// precedence over removed ones, crates/trie/common/src/updates.rs:160-167)
static int32_t dt_collect_removed(b200_dtrie *t, const DTrieDev &d, uint32_t n_removed, const b200_updates *updated,
                                  b200_updates *u) {
    b200_ctx *c = t->c;
    cudaStream_t st = c->stream;
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    size_t n = n_removed;
    size_t o_len = 0, o_path = align_up(n, 16), o_tid = align_up(o_path + n * 32, 16), o_masks = align_up(o_tid + n * 4, 16),
           o_ho = align_up(o_masks + n * 2, 16), total = o_ho + (n + 1) * 8;
    CU(cudaMallocHost(&owner->host, total));
    uint8_t *h = static_cast<uint8_t *>(owner->host);
    memset(h, 0, total);
    if (n) {
        TRY(dt_scratch(t, t->out, o_tid));
        uint8_t *dv = static_cast<uint8_t *>(t->out.p);
        CU(launch_dt_removed_paths(d, n_removed, dv + o_len, dv + o_path, st));
        c->launches++;
        CU(cudaMemcpyAsync(h, dv, o_tid, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    u->path_len = h + o_len;
    u->path_packed = h + o_path;
    u->trie_id = reinterpret_cast<uint32_t *>(h + o_tid);
    u->state_mask = u->tree_mask = u->hash_mask = reinterpret_cast<uint16_t *>(h + o_masks);  // all zero
    u->hash_offset = reinterpret_cast<uint64_t *>(h + o_ho);                                     // all zero
    u->hashes = h;
    // updated nodes take precedence; a path can be recorded once only, but sort + unique keeps this independent of that
    auto key_of = [](const uint8_t *packed, uint8_t len) { return std::string(reinterpret_cast<const char *>(packed), 32) + (char)len; };
    std::vector<std::string> upd, rem;
    if (updated)
        for (uint64_t i = 0; i < updated->n_nodes; i++) upd.push_back(key_of(updated->path_packed + 32 * i, updated->path_len[i]));
    std::sort(upd.begin(), upd.end());
    for (size_t i = 0; i < n; i++) {
        std::string k = key_of(u->path_packed + 32 * i, u->path_len[i]);
        if (!std::binary_search(upd.begin(), upd.end(), k)) rem.push_back(std::move(k));
    }
    std::sort(rem.begin(), rem.end());
    rem.erase(std::unique(rem.begin(), rem.end()), rem.end());
    size_t w = 0;
    for (const std::string &k : rem) {
        memcpy(u->path_packed + 32 * w, k.data(), 32);
        u->path_len[w] = (uint8_t)k[32];
        w++;
    }
    u->n_nodes = w;
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
    if (storage_roots32 && !t->has_sroots) return fail(c, B200_ERR_INVALID_ARG, "trie was created without storage roots");
    TRY(reset_build_state(c));
    uint32_t n_built = 0, n_removed = 0;
    if (m) {
        // every insert may take one leaf slot and one node slot from the bump region
        TRY(dt_reserve(t, (uint64_t)t->leaf_alloc + m, (uint64_t)t->node_alloc + m));
        const uint32_t max_list = (uint32_t)m + 16, max_seeds = (uint32_t)(6 * m + 64);
        const uint64_t max_built64 = std::min<uint64_t>((uint64_t)max_seeds * 64, (uint64_t)t->node_alloc + m) + 16;
        const uint32_t max_built = (uint32_t)max_built64;
        TRY(dt_scratch(t, t->in_keys, m * 32));
        TRY(dt_scratch(t, t->in_accts, m * 72));
        TRY(dt_scratch(t, t->kind, m));
        TRY(dt_scratch(t, t->leaf_of, m * 4));
        TRY(dt_scratch(t, t->list_a, (size_t)max_list * 4));
        TRY(dt_scratch(t, t->list_b, (size_t)max_list * 4));
        TRY(dt_scratch(t, t->ins_idx, m * 4));
        TRY(dt_scratch(t, t->attach, m * 8));
        TRY(dt_scratch(t, t->seeds, (size_t)max_seeds * 4));
        TRY(dt_scratch(t, t->built, (size_t)max_built * 4));
        TRY(dt_scratch(t, t->removed, ((size_t)max_built + max_list) * 4));
        TRY(dt_scratch(t, t->freed_now, (size_t)max_list * 4));
        TRY(dt_scratch(t, t->flags, max_list));  // per-entry defer flags of a collapse round (re-used for the output flags)
        CU(cudaMemcpyAsync(t->in_keys.p, keys32, m * 32, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(t->in_accts.p, accts, m * 72, cudaMemcpyHostToDevice, st));
        const uint8_t *d_present = nullptr, *d_sroots = nullptr;
        if (present) {
            TRY(dt_scratch(t, t->in_present, m));
            CU(cudaMemcpyAsync(t->in_present.p, present, m, cudaMemcpyHostToDevice, st));
            d_present = static_cast<const uint8_t *>(t->in_present.p);
        }
        if (storage_roots32) {
            TRY(dt_scratch(t, t->in_sroots, m * 32));
            CU(cudaMemcpyAsync(t->in_sroots.p, storage_roots32, m * 32, cudaMemcpyHostToDevice, st));
            d_sroots = static_cast<const uint8_t *>(t->in_sroots.p);
        }
        DTrieDev d = dt_view(t);
        const uint8_t *d_keys = static_cast<const uint8_t *>(t->in_keys.p), *d_accts = static_cast<const uint8_t *>(t->in_accts.p);
        uint8_t *kind = static_cast<uint8_t *>(t->kind.p);
        uint32_t *leaf_of = static_cast<uint32_t *>(t->leaf_of.p);
        CU(cudaMemsetAsync(d.g + DG_SEEDS, 0, (DG_WORDS - DG_SEEDS) * 4, st));  // the per-apply list lengths
        // ---- locate, value updates, detach deleted leaves
        CU(launch_dt_locate(d, d_keys, d_present, m, kind, leaf_of, st));
        uint32_t *list_cur = static_cast<uint32_t *>(t->list_a.p), *list_next = static_cast<uint32_t *>(t->list_b.p);
        uint32_t *cnt_cur = d.g + DG_LIST_A, *cnt_next = d.g + DG_LIST_B;
        CU(launch_dt_update_detach(d, d_accts, d_sroots, m, kind, leaf_of, list_cur, st));
        c->launches += 2;
        // ---- collapse rounds until no node is left that lost children
        uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
        for (int round = 0;; round++) {
            CU(cudaMemcpyAsync(ps + 200, cnt_cur, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(ps + 201, small_u32(c) + SM_ERR, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (ps[201] != B200_DEVERR_NONE) return map_dev_error(c, (int)ps[201]);
            if (ps[200] == 0) break;
            if (round > 200) return fail(c, B200_ERR_CUDA, "collapse rounds do not converge");
            CU(cudaMemsetAsync(cnt_next, 0, 4, st));
            CU(launch_dt_collapse_round(d, list_cur, cnt_cur, ps[200], static_cast<uint8_t *>(t->flags.p), list_next, cnt_next, st));
            c->launches += 4;
            std::swap(list_cur, list_next);
            std::swap(cnt_cur, cnt_next);
        }
        // ---- inserts: the dense list of insert keys, their attach points, one thread per run
        {
            uint32_t *ins_idx = static_cast<uint32_t *>(t->ins_idx.p);
            thrust::counting_iterator<uint32_t> counting(0);
            auto is_insert = thrust::make_transform_iterator(kind, IsKind{DK_INSERT});
            size_t t_sel = 0;
            CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, is_insert, ins_idx, d.g + DG_NINSERT, (int64_t)m, st));
            ENSURE(cub_temp, t_sel);
            CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, is_insert, ins_idx, d.g + DG_NINSERT, (int64_t)m, st));
            CU(launch_dt_insert(d, d_keys, d_accts, d_sroots, ins_idx, d.g + DG_NINSERT, m, static_cast<uint64_t *>(t->attach.p), st));
            c->launches += 3;
        }
        // ---- re-hash the dirty paths
        CU(launch_dt_rehash(d, max_seeds, static_cast<uint8_t *>(t->root.p), st));
        CU(launch_dt_finish(d, max_list, static_cast<uint8_t *>(t->root.p), st));
        c->launches += 5;
        c->stats.leaves_added += m;
        TRY(finish_build_state(c));
        // ---- counters back to the host
        CU(cudaMemcpyAsync(ps + 256, d.g, DG_WORDS * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, st));
        TRY(sync_and_status(c));
        t->n_leaves = ps[256 + DG_NLEAVES];
        t->leaf_alloc = ps[256 + DG_LEAF_ALLOC];
        t->node_alloc = ps[256 + DG_NODE_ALLOC];
        n_built = ps[256 + DG_BUILT];
        n_removed = ps[256 + DG_REMOVED];
        c->stats.branches_added = n_built;
        if (opt_updated || opt_removed) {
            b200_updates tmp{};
            b200_updates *upd = opt_updated ? opt_updated : &tmp;
            int32_t r = dt_collect_updates(t, d, n_built, upd);
            if (r == B200_OK && opt_removed) r = dt_collect_removed(t, d, n_removed, upd, opt_removed);
            if (!opt_updated) b200_updates_release(&tmp);
            if (r != B200_OK) {
                if (opt_updated) b200_updates_release(opt_updated);
                if (opt_removed) b200_updates_release(opt_removed);
                return r;
            }
        }
    } else {
        TRY(finish_build_state(c));
        CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, st));
        TRY(sync_and_status(c));
        if (opt_updated) TRY(dt_collect_updates(t, dt_view(t), 0, opt_updated));
        if (opt_removed) TRY(dt_collect_removed(t, dt_view(t), 0, opt_updated, opt_removed));
    }
    if (opt_stats) *opt_stats = c->stats;
    return B200_OK;
}
