// eng_dtrie.inl — dynamic resident tries (b200_dtrie_*: the account trie; b200_dstate_*: the account trie plus every
// storage trie): arenas of 16-slot branch nodes in HBM that take a block's upserts and deletes in place and re-hash only
// the touched paths (device side: tk_dtrie.cuh).
// Part of the single translation unit engine.cu (textually included, in this order).

struct IsKind {
    uint8_t k;
    __host__ __device__ bool operator()(uint8_t x) const { return x == k; }
};

// One arena: the account trie (account = true, a single trie) or the forest of all storage tries (account = false, trie id
// = id of the owning account leaf in the account arena).
struct DArena {
    b200_ctx *c = nullptr;
    uint64_t *bytes = nullptr;  // the owner's device-byte counter
    bool account = true, has_sroots = false, forest = false;
    uint32_t lcap = 0, ncap = 0, tcap = 0;                   // capacities: leaves, nodes, tries
    uint32_t leaf_alloc = 0, node_alloc = 0, n_leaves = 0;  // device counters as of the last apply
    uint8_t *top_out = nullptr;                              // where finished tries put their root hash
    uint32_t top_stride = 0;
    DevBuf lkey, lval, lsroot, lref, lmeta, lparent, ltrie, lseed;
    DevBuf nchild, ndepth, nparent, nref, nmeta, nmasks, nkey, npending, ntrie, nseed, ncur, nnext;
    DevBuf troot, leaf_free, node_free, g;
    // per-apply scratch
    DevBuf kind, leaf_of, list_a, list_b, ins_idx, attach, seeds, built, removed, freed_now, flags, nh, sel, prefix, pick, out;
    // outcome of the last apply
    uint32_t n_built = 0, n_removed = 0;
    uint32_t val_stride() const { return account ? 72u : 32u; }
};

static int32_t da_resize(DArena *a, DevBuf &b, size_t new_bytes, size_t keep_bytes, int fill /* -1 none, else byte */) {
    b200_ctx *c = a->c;
    if (new_bytes <= b.cap) return B200_OK;
    void *p = nullptr;
    size_t want = new_bytes + 256;
    CU(cudaMalloc(&p, want));
    if (fill >= 0) CU(cudaMemsetAsync(p, fill, want, c->stream));
    if (b.p && keep_bytes) CU(cudaMemcpyAsync(p, b.p, keep_bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (b.p) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(b.p));
        *a->bytes -= b.cap;
    }
    b.p = p;
    b.cap = want;
    *a->bytes += want;
    return B200_OK;
}
static int32_t da_scratch(DArena *a, DevBuf &b, size_t bytes) { return da_resize(a, b, bytes ? bytes : 16, 0, -1); }

// capacity for `leaves` leaf slots, `nodes` node slots and `tries` root words, keeping what is allocated so far
static int32_t da_reserve(DArena *a, uint64_t leaves, uint64_t nodes, uint64_t tries) {
    b200_ctx *c = a->c;
    if (leaves >= (1ull << 31) || nodes >= (1ull << 31)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^31-1 leaves per arena");
    if (leaves > a->lcap) {
        uint64_t cap = std::max<uint64_t>(leaves, (uint64_t)a->lcap + a->lcap / 2) + 1024;
        size_t used = a->leaf_alloc, vs = a->val_stride();
        TRY(da_resize(a, a->lkey, cap * 32, used * 32, -1));
        TRY(da_resize(a, a->lval, cap * vs, used * vs, -1));
        if (a->has_sroots) TRY(da_resize(a, a->lsroot, cap * 32, used * 32, -1));
        TRY(da_resize(a, a->lref, cap * 32, used * 32, -1));
        TRY(da_resize(a, a->lmeta, cap, used, -1));
        TRY(da_resize(a, a->lparent, cap * 4, used * 4, -1));
        if (a->forest) TRY(da_resize(a, a->ltrie, cap * 4, used * 4, -1));
        TRY(da_resize(a, a->lseed, cap, used, 0));
        TRY(da_resize(a, a->leaf_free, cap * 4, (size_t)a->lcap * 4, -1));
        a->lcap = (uint32_t)cap;
    }
    if (nodes > a->ncap) {
        uint64_t cap = std::max<uint64_t>(nodes, (uint64_t)a->ncap + a->ncap / 2) + 1024;
        size_t used = a->node_alloc;
        TRY(da_resize(a, a->nchild, cap * 64, used * 64, -1));
        TRY(da_resize(a, a->ndepth, cap, used, -1));
        TRY(da_resize(a, a->nparent, cap * 4, used * 4, -1));
        TRY(da_resize(a, a->nref, cap * 32, used * 32, -1));
        TRY(da_resize(a, a->nmeta, cap, used, -1));
        TRY(da_resize(a, a->nmasks, cap * 8, used * 8, -1));
        TRY(da_resize(a, a->nkey, cap * 32, used * 32, -1));
        TRY(da_resize(a, a->npending, cap * 4, used * 4, 0));
        if (a->forest) TRY(da_resize(a, a->ntrie, cap * 4, used * 4, -1));
        TRY(da_resize(a, a->nseed, cap, used, 0));
        TRY(da_resize(a, a->ncur, cap, used, 0));
        TRY(da_resize(a, a->nnext, cap, used, 0));
        TRY(da_resize(a, a->node_free, cap * 4, (size_t)a->ncap * 4, -1));
        a->ncap = (uint32_t)cap;
    }
    if (tries > a->tcap) {
        uint64_t cap = std::max<uint64_t>(tries, (uint64_t)a->tcap + a->tcap / 2) + 16;
        TRY(da_resize(a, a->troot, cap * 4, (size_t)a->tcap * 4, 0xFF));  // new tries are empty (DT_NONE)
        a->tcap = (uint32_t)cap;
    }
    return B200_OK;
}

static DTrieDev da_view(DArena *a) {
    b200_ctx *c = a->c;
    DTrieDev d{};
    d.lkey = static_cast<uint8_t *>(a->lkey.p);
    d.lval = static_cast<uint8_t *>(a->lval.p);
    d.lsroot = a->has_sroots ? static_cast<uint8_t *>(a->lsroot.p) : nullptr;
    d.lref = static_cast<uint8_t *>(a->lref.p);
    d.lmeta = static_cast<uint8_t *>(a->lmeta.p);
    d.lparent = static_cast<uint32_t *>(a->lparent.p);
    d.ltrie = a->forest ? static_cast<uint32_t *>(a->ltrie.p) : nullptr;
    d.lseed = static_cast<uint8_t *>(a->lseed.p);
    d.nchild = static_cast<uint32_t *>(a->nchild.p);
    d.ndepth = static_cast<uint8_t *>(a->ndepth.p);
    d.nparent = static_cast<uint32_t *>(a->nparent.p);
    d.nref = static_cast<uint8_t *>(a->nref.p);
    d.nmeta = static_cast<uint8_t *>(a->nmeta.p);
    d.nmasks = static_cast<ushort4 *>(a->nmasks.p);
    d.nkey = static_cast<uint8_t *>(a->nkey.p);
    d.npending = static_cast<uint32_t *>(a->npending.p);
    d.ntrie = a->forest ? static_cast<uint32_t *>(a->ntrie.p) : nullptr;
    d.nseed = static_cast<uint8_t *>(a->nseed.p);
    d.ncur = static_cast<uint8_t *>(a->ncur.p);
    d.nnext = static_cast<uint8_t *>(a->nnext.p);
    d.troot = static_cast<uint32_t *>(a->troot.p);
    d.top_out = a->top_out;
    d.top_stride = a->top_stride;
    d.val_stride = a->val_stride();
    d.account = a->account ? 1 : 0;
    d.leaf_free = static_cast<uint32_t *>(a->leaf_free.p);
    d.node_free = static_cast<uint32_t *>(a->node_free.p);
    d.seeds = static_cast<uint32_t *>(a->seeds.p);
    d.built = static_cast<uint32_t *>(a->built.p);
    d.removed = static_cast<uint32_t *>(a->removed.p);
    d.freed_now = static_cast<uint32_t *>(a->freed_now.p);
    d.g = static_cast<uint32_t *>(a->g.p);
    d.err = reinterpret_cast<int *>(small_u32(c) + SM_ERR);
    d.counters = reinterpret_cast<unsigned long long *>(small_u32(c) + SM_COUNTERS);
    d.lcap = a->lcap;
    d.ncap = a->ncap;
    return d;
}

static void da_free(DArena *a) {
    DevBuf *bufs[] = {&a->lkey, &a->lval, &a->lsroot, &a->lref, &a->lmeta, &a->lparent, &a->ltrie, &a->lseed, &a->nchild,
                      &a->ndepth, &a->nparent, &a->nref, &a->nmeta, &a->nmasks, &a->nkey, &a->npending, &a->ntrie, &a->nseed,
                      &a->ncur, &a->nnext, &a->troot, &a->leaf_free, &a->node_free, &a->g, &a->kind, &a->leaf_of, &a->list_a,
                      &a->list_b, &a->ins_idx, &a->attach, &a->seeds, &a->built, &a->removed, &a->freed_now, &a->flags, &a->nh,
                      &a->sel, &a->prefix, &a->pick, &a->out};
    for (DevBuf *b : bufs)
        if (b->p) {
            cudaFree(b->p);
            *b = DevBuf{};
        }
}

// Fills the arena from a finished resident build (`src`: the account trie, or a storage forest with its segment table).
// top_out / top_stride must be set.  Synchronises.
static int32_t da_from_build(DArena *a, b200_trie *src, uint64_t min_tries) {
    b200_ctx *c = a->c;
    cudaStream_t st = c->stream;
    const uint64_t n = src->n;
    const uint32_t B = src->B;
    const uint64_t n_tries = a->forest ? std::max<uint64_t>(src->n_segs, min_tries) : 1;
    TRY(da_reserve(a, n + n / 8 + 16, (uint64_t)B + B / 8 + 16, n_tries));
    TRY(da_resize(a, a->g, DG_WORDS * 4, 0, 0));
    if (n) {
        CU(cudaMemcpyAsync(a->lkey.p, src->keys.p, n * 32, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(a->lval.p, src->accts.p, n * a->val_stride(), cudaMemcpyDeviceToDevice, st));
        if (a->has_sroots) CU(cudaMemcpyAsync(a->lsroot.p, src->sroots.p, n * 32, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(a->lref.p, src->leaf_ref.p, n * 32, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(a->lmeta.p, src->leaf_meta.p, n, cudaMemcpyDeviceToDevice, st));
    }
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small) + 256;  // 64 words of the readback page
    memset(ps, 0, DG_WORDS * 4);
    ps[DG_NLEAVES] = (uint32_t)n;
    ps[DG_LEAF_ALLOC] = (uint32_t)n;
    ps[DG_NODE_ALLOC] = B;
    CU(cudaMemcpyAsync(a->g.p, ps, DG_WORDS * 4, cudaMemcpyHostToDevice, st));
    a->leaf_alloc = a->n_leaves = (uint32_t)n;
    a->node_alloc = B;
    const uint32_t *leaf_trie = nullptr;
    if (a->forest && n) {
        TRY(da_scratch(a, a->leaf_of, n * 4));
        CU(launch_dt_leaf_segments(static_cast<const uint64_t *>(src->seg_offsets.p), src->n_segs, n,
                                   static_cast<uint32_t *>(a->leaf_of.p), st));
        leaf_trie = static_cast<const uint32_t *>(a->leaf_of.p);
        c->launches++;
    }
    DTrieDev d = da_view(a);
    CU(launch_dt_convert(src->f, B, static_cast<const uint32_t *>(src->leaf_parent.p),
                         static_cast<const uint32_t *>(src->node_parent.p), leaf_trie, d, st));
    c->launches += 2;
    CU(cudaStreamSynchronize(st));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ one arena, one block
// Structural part of an apply: classify the m dirty entries, write value updates, detach deleted leaves, collapse, insert.
// Leaves the seeds in the arena; leaf_of[i] afterwards holds the leaf of every entry that exists (updated, touched or
// inserted), DT_NONE otherwise.  Every pointer is a device pointer.
static int32_t da_restructure(DArena *a, const uint32_t *d_trie_of_key, const uint8_t *d_keys, const uint8_t *d_vals,
                              const uint8_t *d_flags, const uint8_t *d_sroots, uint64_t m) {
    b200_ctx *c = a->c;
    cudaStream_t st = c->stream;
    // every insert may take one leaf slot and one node slot from the bump region
    TRY(da_reserve(a, (uint64_t)a->leaf_alloc + m, (uint64_t)a->node_alloc + m, a->tcap));
    const uint32_t max_list = (uint32_t)m + 16, max_seeds = (uint32_t)(6 * m + 64);
    const uint32_t max_built = (uint32_t)(std::min<uint64_t>((uint64_t)max_seeds * 64, (uint64_t)a->node_alloc + m) + 16);
    TRY(da_scratch(a, a->kind, m));
    TRY(da_scratch(a, a->leaf_of, m * 4));
    TRY(da_scratch(a, a->list_a, (size_t)max_list * 4));
    TRY(da_scratch(a, a->list_b, (size_t)max_list * 4));
    TRY(da_scratch(a, a->ins_idx, m * 4));
    TRY(da_scratch(a, a->attach, m * 8));
    TRY(da_scratch(a, a->seeds, (size_t)max_seeds * 4));
    TRY(da_scratch(a, a->built, (size_t)max_built * 4));
    TRY(da_scratch(a, a->removed, ((size_t)max_built + max_list) * 4));
    TRY(da_scratch(a, a->freed_now, (size_t)max_list * 4));
    TRY(da_scratch(a, a->flags, max_list));  // per-entry defer flags of a collapse round (re-used for the output flags)
    DTrieDev d = da_view(a);
    uint8_t *kind = static_cast<uint8_t *>(a->kind.p);
    uint32_t *leaf_of = static_cast<uint32_t *>(a->leaf_of.p);
    CU(cudaMemsetAsync(d.g + DG_SEEDS, 0, (DG_WORDS - DG_SEEDS) * 4, st));  // the per-apply list lengths
    // ---- locate, value updates, detach deleted leaves
    CU(launch_dt_locate(d, d_trie_of_key, d_keys, d_vals, d_flags, m, kind, leaf_of, st));
    uint32_t *list_cur = static_cast<uint32_t *>(a->list_a.p), *list_next = static_cast<uint32_t *>(a->list_b.p);
    uint32_t *cnt_cur = d.g + DG_LIST_A, *cnt_next = d.g + DG_LIST_B;
    CU(launch_dt_update_detach(d, d_vals, d_sroots, m, kind, leaf_of, list_cur, st));
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
        CU(launch_dt_collapse_round(d, list_cur, cnt_cur, ps[200], static_cast<uint8_t *>(a->flags.p), list_next, cnt_next, st));
        c->launches += 4;
        std::swap(list_cur, list_next);
        std::swap(cnt_cur, cnt_next);
    }
    // ---- inserts: the dense list of insert entries, their attach points, one thread per run
    uint32_t *ins_idx = static_cast<uint32_t *>(a->ins_idx.p);
    thrust::counting_iterator<uint32_t> counting(0);
    auto is_insert = thrust::make_transform_iterator(kind, IsKind{DK_INSERT});
    size_t t_sel = 0;
    CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, is_insert, ins_idx, d.g + DG_NINSERT, (int64_t)m, st));
    ENSURE(cub_temp, t_sel);
    CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, is_insert, ins_idx, d.g + DG_NINSERT, (int64_t)m, st));
    CU(launch_dt_insert(d, d_trie_of_key, d_keys, d_vals, d_sroots, ins_idx, d.g + DG_NINSERT, m,
                        static_cast<uint64_t *>(a->attach.p), leaf_of, st));
    c->launches += 3;
    return B200_OK;
}

// re-hash of the seeded paths, recycling of the freed nodes (asynchronous)
static int32_t da_rehash(DArena *a, uint64_t m) {
    b200_ctx *c = a->c;
    DTrieDev d = da_view(a);
    CU(launch_dt_rehash(d, (uint32_t)(6 * m + 64), c->stream));
    CU(launch_dt_finish(d, (uint32_t)m + 16, c->stream));
    c->launches += 5;
    return B200_OK;
}

// device counters -> host mirror (after a synchronisation point that covers the copy)
static int32_t da_pull_counters(DArena *a, uint32_t *pinned64) {
    b200_ctx *c = a->c;
    CU(cudaMemcpyAsync(pinned64, a->g.p, DG_WORDS * 4, cudaMemcpyDeviceToHost, c->stream));
    return B200_OK;
}
static void da_take_counters(DArena *a, const uint32_t *pinned64) {
    a->n_leaves = pinned64[DG_NLEAVES];
    a->leaf_alloc = pinned64[DG_LEAF_ALLOC];
    a->node_alloc = pinned64[DG_NODE_ALLOC];
    a->n_built = pinned64[DG_BUILT];
    a->n_removed = pinned64[DG_REMOVED];
}

// host copy of the re-hashed stored nodes (same block layout as gather_and_copy)
static int32_t da_collect_updates(DArena *a, b200_updates *u) {
    b200_ctx *c = a->c;
    cudaStream_t st = c->stream;
    const uint32_t n_built = a->n_built;
    DTrieDev d = da_view(a);
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    uint32_t n_stored = 0, n_hashes = 0;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    const uint32_t *pick_ids = nullptr, *pick_prefix = nullptr;
    if (n_built) {
        TRY(da_scratch(a, a->flags, n_built));
        TRY(da_scratch(a, a->nh, (size_t)n_built * 4));
        TRY(da_scratch(a, a->sel, (size_t)n_built * 4));
        TRY(da_scratch(a, a->prefix, ((size_t)n_built + 1) * 4));
        TRY(da_scratch(a, a->pick, (size_t)n_built * 8));
        uint8_t *flags = static_cast<uint8_t *>(a->flags.p);
        uint32_t *nh = static_cast<uint32_t *>(a->nh.p), *sel = static_cast<uint32_t *>(a->sel.p);
        uint32_t *prefix = static_cast<uint32_t *>(a->prefix.p);
        uint32_t *ids = static_cast<uint32_t *>(a->pick.p), *pref = ids + n_built;
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
        TRY(da_scratch(a, a->out, dev_total));
        uint8_t *dv = static_cast<uint8_t *>(a->out.p);
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

// removed_nodes as records without masks or hashes; paths that are also in `updated` (same trie) are dropped: updated
// nodes take precedence over removed ones (crates/trie/common/src/updates.rs:160-167)
static int32_t da_collect_removed(DArena *a, const b200_updates *updated, b200_updates *u) {
    b200_ctx *c = a->c;
    cudaStream_t st = c->stream;
    DTrieDev d = da_view(a);
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    const size_t n = a->n_removed;
    size_t o_len = 0, o_path = align_up(n, 16), o_tid = align_up(o_path + n * 32, 16), o_masks = align_up(o_tid + n * 4, 16),
           o_ho = align_up(o_masks + n * 2, 16), total = o_ho + (n + 1) * 8;
    CU(cudaMallocHost(&owner->host, total));
    uint8_t *h = static_cast<uint8_t *>(owner->host);
    memset(h, 0, total);
    if (n) {
        TRY(da_scratch(a, a->out, o_masks));
        uint8_t *dv = static_cast<uint8_t *>(a->out.p);
        CU(launch_dt_removed_paths(d, (uint32_t)n, dv + o_len, dv + o_path, reinterpret_cast<uint32_t *>(dv + o_tid), st));
        c->launches++;
        CU(cudaMemcpyAsync(h, dv, o_masks, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    u->path_len = h + o_len;
    u->path_packed = h + o_path;
    u->trie_id = reinterpret_cast<uint32_t *>(h + o_tid);
    u->state_mask = u->tree_mask = u->hash_mask = reinterpret_cast<uint16_t *>(h + o_masks);  // all zero
    u->hash_offset = reinterpret_cast<uint64_t *>(h + o_ho);                                     // all zero
    u->hashes = h;
    auto key_of = [](uint32_t trie, const uint8_t *packed, uint8_t len) {
        std::string k(reinterpret_cast<const char *>(&trie), 4);
        k.append(reinterpret_cast<const char *>(packed), 32);
        k.push_back((char)len);
        return k;
    };
    std::vector<std::string> upd, rem;
    if (updated)
        for (uint64_t i = 0; i < updated->n_nodes; i++)
            upd.push_back(key_of(updated->trie_id[i], updated->path_packed + 32 * i, updated->path_len[i]));
    std::sort(upd.begin(), upd.end());
    for (size_t i = 0; i < n; i++) {
        std::string k = key_of(u->trie_id[i], u->path_packed + 32 * i, u->path_len[i]);
        if (!std::binary_search(upd.begin(), upd.end(), k)) rem.push_back(std::move(k));
    }
    std::sort(rem.begin(), rem.end());
    rem.erase(std::unique(rem.begin(), rem.end()), rem.end());
    size_t w = 0;
    for (const std::string &k : rem) {
        memcpy(&u->trie_id[w], k.data(), 4);
        memcpy(u->path_packed + 32 * w, k.data() + 4, 32);
        u->path_len[w] = (uint8_t)k[36];
        w++;
    }
    u->n_nodes = w;
    return B200_OK;
}

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

static int32_t h2d_into(DArena *a, DevBuf &b, const void *src, size_t bytes) {
    b200_ctx *c = a->c;
    TRY(da_scratch(a, b, bytes));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
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

// ------------------------------------------------------------------------------------------------ b200_dstate: accounts + storage
// The whole hashed state resident: the account arena plus one forest arena holding every storage trie (trie id = id of
// the owning account leaf).  A block's HashedPostStateSorted is applied in place: account upserts / deletes, per-account
// slot upserts / deletes (zero value = delete) / wipes; storage roots flow into the account leaves on the device.
struct b200_dstate {
    b200_ctx *c = nullptr;
    uint64_t bytes = 0;
    DArena acc, sto;
    bool sharded = false;  // accounts as 16 top-nibble bucket tries: this state is one rank's shard
    DevBuf bucket_roots, frontier, acct_tries;
    DevBuf root, in_akeys, in_accts, in_aflags, in_skeys, in_svals, in_offs, trie_of_key, wipe_a, wipe_b, wipe_cnt;
};

extern "C" B200_API void b200_dstate_destroy(b200_dstate *t) {
    if (!t) return;
    cudaSetDevice(t->c->device);
    cudaStreamSynchronize(t->c->stream);
    da_free(&t->acc);
    da_free(&t->sto);
    DevBuf *bufs[] = {&t->root, &t->in_akeys, &t->in_accts, &t->in_aflags, &t->in_skeys, &t->in_svals, &t->in_offs,
                      &t->trie_of_key, &t->wipe_a, &t->wipe_b, &t->wipe_cnt, &t->bucket_roots, &t->frontier, &t->acct_tries};
    for (DevBuf *b : bufs) dbuf_free(*b);
    delete t;
}
extern "C" B200_API uint64_t b200_dstate_device_bytes(const b200_dstate *t) { return t ? t->bytes : 0; }
extern "C" B200_API uint64_t b200_dstate_accounts(const b200_dstate *t) { return t ? t->acc.n_leaves : 0; }
extern "C" B200_API uint64_t b200_dstate_slots(const b200_dstate *t) { return t ? t->sto.n_leaves : 0; }

extern "C" B200_API int32_t b200_dstate_root(b200_dstate *t, uint8_t root32[32]) {
    if (!t || !root32) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

// this shard's 16 frontier entries and the root they give on their own (enqueued; t->frontier / t->root)
static int32_t dstate_frontier_on_device(b200_dstate *t) {
    b200_ctx *c = t->c;
    DTrieDev d = da_view(&t->acc);
    CU(launch_dt_frontier(d, static_cast<const uint8_t *>(t->bucket_roots.p), static_cast<FrontierEntryDev *>(t->frontier.p),
                          c->stream));
    CU(launch_root_from_frontier(static_cast<const FrontierEntryDev *>(t->frontier.p), static_cast<uint8_t *>(t->root.p),
                                 c->stream));
    c->launches += 2;
    return B200_OK;
}

static int32_t dstate_create_impl(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts, uint64_t n_accounts,
                                  const uint8_t *slot_keys32, const uint8_t *values32_be, const uint64_t *seg_offsets,
                                  bool sharded, b200_dstate **out, uint8_t root32[32]) {
    if (!c || !out || !seg_offsets || (n_accounts && (!acct_keys32 || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    *out = nullptr;
    TRY(check_offsets_host(c, seg_offsets, n_accounts));
    const uint64_t n_slots = seg_offsets[n_accounts];
    if (n_slots && (!slot_keys32 || !values32_be)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    cudaStream_t st = c->stream;
    b200_trie *src_s = nullptr, *src_a = nullptr;
    TRY(forest_create_locked(c, slot_keys32, values32_be, seg_offsets, n_accounts, n_slots, cudaMemcpyHostToDevice, &src_s));
    // the account trie takes its storage roots straight from the forest build (device memory: cudaMemcpyDefault)
    int32_t r;
    if (sharded) {
        uint64_t bucket_offsets[17];
        for (uint32_t b = 0; b <= 16; b++) {  // first account whose top nibble >= b
            uint64_t lo = 0, hi = n_accounts;
            while (lo < hi) {
                uint64_t mid = (lo + hi) >> 1;
                if ((uint32_t)(acct_keys32[32 * mid] >> 4) < b) lo = mid + 1;
                else hi = mid;
            }
            bucket_offsets[b] = lo;
        }
        r = bucket_forest_create_locked(c, acct_keys32, accts, src_s->seg_roots.p, n_accounts, bucket_offsets, cudaMemcpyDefault, &src_a);
    } else {
        r = trie_create_locked(c, acct_keys32, accts, src_s->seg_roots.p, n_accounts, cudaMemcpyDefault, &src_a, nullptr);
    }
    if (r != B200_OK) {
        b200_trie_destroy(src_s);
        return r;
    }
    b200_dstate *t = new b200_dstate();
    t->c = c;
    t->acc.c = t->sto.c = c;
    t->acc.bytes = t->sto.bytes = &t->bytes;
    t->acc.account = true;
    t->acc.has_sroots = true;
    t->acc.forest = sharded;
    t->sharded = sharded;
    t->sto.account = false;
    t->sto.forest = true;
    auto body = [&]() -> int32_t {
        TRY(da_resize(&t->acc, t->root, 64, 0, -1));
        if (sharded) {
            TRY(da_resize(&t->acc, t->bucket_roots, 16 * 32, 0, 0));
            TRY(da_resize(&t->acc, t->frontier, 16 * sizeof(FrontierEntryDev), 0, 0));
            CU(cudaMemcpyAsync(t->bucket_roots.p, src_a->seg_roots.p, 16 * 32, cudaMemcpyDeviceToDevice, st));
            t->acc.top_out = static_cast<uint8_t *>(t->bucket_roots.p);
            t->acc.top_stride = 32;
        } else {
            t->acc.top_out = static_cast<uint8_t *>(t->root.p);
            t->acc.top_stride = 0;
        }
        TRY(da_from_build(&t->acc, src_a, sharded ? 16 : 1));
        t->sto.top_out = static_cast<uint8_t *>(t->acc.lsroot.p);
        t->sto.top_stride = 32;
        TRY(da_from_build(&t->sto, src_s, t->acc.lcap));
        if (sharded) TRY(dstate_frontier_on_device(t));
        else CU(cudaMemcpyAsync(t->root.p, src_a->root.p, 32, cudaMemcpyDeviceToDevice, st));
        if (root32) CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return B200_OK;
    };
    r = body();
    b200_trie_destroy(src_a);
    b200_trie_destroy(src_s);
    if (r != B200_OK) {
        b200_dstate_destroy(t);
        return r;
    }
    *out = t;
    return B200_OK;
}

extern "C" B200_API int32_t b200_dstate_create(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                               uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                               const uint64_t *seg_offsets, b200_dstate **out, uint8_t root32[32]) {
    return dstate_create_impl(c, acct_keys32, accts, n_accounts, slot_keys32, values32_be, seg_offsets, false, out, root32);
}
// One rank's shard of a state that is split by top key nibble (any subset of the 16 buckets).  root32 (nullable) receives
// the root this shard has on its own; the global root is b200_root_from_frontier over the gathered b200_dstate_frontier
// entries of all ranks.
extern "C" B200_API int32_t b200_dstate_create_sharded(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                                       uint64_t n_accounts, const uint8_t *slot_keys32,
                                                       const uint8_t *values32_be, const uint64_t *seg_offsets,
                                                       b200_dstate **out, uint8_t root32[32]) {
    return dstate_create_impl(c, acct_keys32, accts, n_accounts, slot_keys32, values32_be, seg_offsets, true, out, root32);
}
// The 16 top-nibble frontier entries of a sharded state as of its last apply (empty entries for buckets it does not hold).
extern "C" B200_API int32_t b200_dstate_frontier(b200_dstate *t, b200_frontier_entry out16[16]) {
    if (!t || !out16) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    if (!t->sharded) return fail(c, B200_ERR_INVALID_ARG, "not a sharded state (b200_dstate_create_sharded)");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(out16, t->frontier.p, 16 * sizeof(FrontierEntryDev), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

// One block.  Account entries: keys strictly ascending; acct_flags[i] bit 0 = the account exists after the block (0 =
// destroyed), bit 1 = its data is unchanged (only its storage changes: `accts[i]` is ignored), bit 2 = its storage is wiped
// before the block's slots are applied (NULL = every entry is a plain upsert).  Storage entries of account entry i are
// seg_offsets[i] .. seg_offsets[i+1]: slot keys ascending, zero value = delete.  Every account whose storage changes
// must have an entry.  storage_* records carry trie_id = index i of the account entry.
extern "C" B200_API int32_t b200_dstate_apply(b200_dstate *t, const uint8_t *acct_keys32, const b200_account *accts,
                                              const uint8_t *acct_flags, uint64_t m, const uint8_t *slot_keys32,
                                              const uint8_t *values32_be, const uint64_t *seg_offsets, uint8_t root32[32],
                                              b200_updates *opt_acct_updated, b200_updates *opt_acct_removed,
                                              b200_updates *opt_storage_updated, b200_updates *opt_storage_removed,
                                              uint8_t *opt_storage_deleted /* [m] */, b200_stats *opt_stats) {
    if (!t || !root32 || (m && (!acct_keys32 || !accts || !seg_offsets)))
        return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    b200_updates *outs[] = {opt_acct_updated, opt_acct_removed, opt_storage_updated, opt_storage_removed};
    for (b200_updates *u : outs)
        if (u) memset(u, 0, sizeof *u);
    if (m >= (1ull << 28)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^28-1 dirty accounts per apply");
    if (m) TRY(check_offsets_host(c, seg_offsets, m));
    const uint64_t n_entries = m ? seg_offsets[m] : 0;
    if (n_entries >= (1ull << 28)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^28-1 dirty slots per apply");
    if (n_entries && (!slot_keys32 || !values32_be)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> lock(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    DArena *A = &t->acc, *S = &t->sto;
    TRY(reset_build_state(c));
    A->n_built = A->n_removed = S->n_built = S->n_removed = 0;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    std::vector<uint8_t> h_kind;
    std::vector<uint32_t> h_leaf;
    if (m) {
        TRY(h2d_into(A, t->in_akeys, acct_keys32, m * 32));
        TRY(h2d_into(A, t->in_accts, accts, m * 72));
        if (acct_flags) TRY(h2d_into(A, t->in_aflags, acct_flags, m));
        TRY(h2d_into(A, t->in_offs, seg_offsets, (m + 1) * 8));
        const uint8_t *d_flags = acct_flags ? static_cast<const uint8_t *>(t->in_aflags.p) : nullptr;
        // ---- accounts: structure only (their leaves are re-hashed after the storage roots are known)
        const uint32_t *d_acct_tries = nullptr;
        if (t->sharded) {  // bucket trie of every account entry = its top key nibble
            TRY(da_scratch(A, t->acct_tries, m * 4));
            CU(launch_dt_nibble_tries(static_cast<const uint8_t *>(t->in_akeys.p), m, static_cast<uint32_t *>(t->acct_tries.p), st));
            c->launches++;
            d_acct_tries = static_cast<const uint32_t *>(t->acct_tries.p);
        }
        TRY(da_restructure(A, d_acct_tries, static_cast<const uint8_t *>(t->in_akeys.p), static_cast<const uint8_t *>(t->in_accts.p),
                           d_flags, nullptr, m));
        const uint8_t *a_kind = static_cast<const uint8_t *>(A->kind.p);
        const uint32_t *a_leaf = static_cast<const uint32_t *>(A->leaf_of.p);
        // ---- storage tries of destroyed / wiped accounts
        TRY(da_reserve(S, S->leaf_alloc, S->node_alloc, A->lcap));
        S->top_out = static_cast<uint8_t *>(A->lsroot.p);  // the account arena may have been re-allocated
        S->top_stride = 32;
        TRY(da_resize(A, t->wipe_cnt, 16, 0, 0));
        uint32_t *wc = static_cast<uint32_t *>(t->wipe_cnt.p);  // [0] tries to wipe, [1] / [2] BFS list lengths
        CU(cudaMemsetAsync(wc, 0, 16, st));
        TRY(da_scratch(A, t->wipe_a, std::max<size_t>((size_t)m, (size_t)S->node_alloc) * 4 + 16));
        TRY(da_scratch(A, t->wipe_b, std::max<size_t>((size_t)m, (size_t)S->node_alloc) * 4 + 16));
        TRY(da_scratch(A, t->trie_of_key, std::max<size_t>((size_t)m, (size_t)n_entries) * 4 + 16));
        uint32_t *wipe_tries = static_cast<uint32_t *>(t->trie_of_key.p);  // borrowed until the storage entries are expanded
        CU(launch_dt_wipe_list(a_kind, d_flags, a_leaf, m, wipe_tries, wc, st));
        c->launches++;
        {
            DTrieDev ds = da_view(S);
            uint32_t *cur = static_cast<uint32_t *>(t->wipe_a.p), *next = static_cast<uint32_t *>(t->wipe_b.p);
            uint32_t *cnt_cur = wc + 1, *cnt_next = wc + 2;
            CU(launch_dt_wipe_begin(ds, wipe_tries, wc, (uint32_t)m, cur, cnt_cur, st));
            c->launches++;
            for (int round = 0;; round++) {
                CU(cudaMemcpyAsync(ps + 200, cnt_cur, 4, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (ps[200] == 0) break;
                if (round > 70) return fail(c, B200_ERR_CUDA, "storage wipe does not terminate");
                CU(cudaMemsetAsync(cnt_next, 0, 4, st));
                CU(launch_dt_wipe_round(ds, cur, cnt_cur, ps[200], next, cnt_next, st));
                c->launches++;
                std::swap(cur, next);
                std::swap(cnt_cur, cnt_next);
            }
        }
        // ---- storage slots of the surviving accounts
        if (n_entries) {
            TRY(h2d_into(S, t->in_skeys, slot_keys32, n_entries * 32));
            TRY(h2d_into(S, t->in_svals, values32_be, n_entries * 32));
            uint32_t *trie_of_key = static_cast<uint32_t *>(t->trie_of_key.p);
            CU(launch_dt_expand_tries(static_cast<const uint64_t *>(t->in_offs.p), m, a_kind, a_leaf, n_entries, trie_of_key, st));
            c->launches++;
            TRY(da_restructure(S, trie_of_key, static_cast<const uint8_t *>(t->in_skeys.p),
                               static_cast<const uint8_t *>(t->in_svals.p), nullptr, nullptr, n_entries));
            TRY(da_rehash(S, n_entries));  // roots land in the account leaves' storage-root fields
            c->stats.leaves_added += n_entries;
        }
        // ---- accounts: re-hash
        TRY(da_rehash(A, m));
        if (t->sharded) TRY(dstate_frontier_on_device(t));
        c->stats.leaves_added += m;
        // what the host needs to label the storage records
        h_kind.resize(m);
        h_leaf.resize(m);
        CU(cudaMemcpyAsync(h_kind.data(), A->kind.p, m, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_leaf.data(), A->leaf_of.p, m * 4, cudaMemcpyDeviceToHost, st));
    }
    TRY(finish_build_state(c));
    if (m) {
        TRY(da_pull_counters(A, ps + 256));
        TRY(da_pull_counters(S, ps + 256 + DG_WORDS));
    }
    CU(cudaMemcpyAsync(root32, t->root.p, 32, cudaMemcpyDeviceToHost, st));
    TRY(sync_and_status(c));
    if (m) {
        da_take_counters(A, ps + 256);
        da_take_counters(S, ps + 256 + DG_WORDS);
        if (!n_entries) S->n_built = S->n_removed = 0;  // the storage arena's per-apply lists were not reset this block
    }
    c->stats.branches_added = A->n_built + S->n_built;
    if (opt_storage_deleted)
        for (uint64_t i = 0; i < m; i++)
            opt_storage_deleted[i] = (h_kind[i] == DK_DELETE || (acct_flags && (acct_flags[i] & 4) && (h_kind[i] == DK_UPDATE || h_kind[i] == DK_TOUCH))) ? 1 : 0;
    auto release_all = [&] {
        for (b200_updates *u : outs)
            if (u) b200_updates_release(u);
    };
    auto collect = [&](DArena *a, b200_updates *updated, b200_updates *removed) -> int32_t {
        if (!updated && !removed) return B200_OK;
        b200_updates tmp{};
        b200_updates *upd = updated ? updated : &tmp;
        int32_t r = da_collect_updates(a, upd);
        if (r == B200_OK && removed) r = da_collect_removed(a, upd, removed);
        if (!updated) b200_updates_release(&tmp);
        return r;
    };
    int32_t r = collect(A, opt_acct_updated, opt_acct_removed);
    if (r == B200_OK) r = collect(S, opt_storage_updated, opt_storage_removed);
    if (r != B200_OK) {
        release_all();
        return r;
    }
    // storage records: account leaf id -> index of the account entry
    if (opt_storage_updated || opt_storage_removed) {
        std::unordered_map<uint32_t, uint32_t> entry_of;
        for (uint64_t i = 0; i < m; i++)
            if (h_kind[i] == DK_UPDATE || h_kind[i] == DK_TOUCH || h_kind[i] == DK_INSERT) entry_of[h_leaf[i]] = (uint32_t)i;
        for (b200_updates *u : {opt_storage_updated, opt_storage_removed})
            if (u)
                for (uint64_t k = 0; k < u->n_nodes; k++) {
                    auto it = entry_of.find(u->trie_id[k]);
                    u->trie_id[k] = it == entry_of.end() ? 0xFFFFFFFFu : it->second;
                }
    }
    if (opt_stats) *opt_stats = c->stats;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ proofs
struct ProofsOwner {
    void *host = nullptr;
};
extern "C" B200_API void b200_proofs_release(b200_proofs *p) {
    if (!p) return;
    if (p->_owner) {
        ProofsOwner *o = static_cast<ProofsOwner *>(p->_owner);
        if (o->host) cudaFreeHost(o->host);
        delete o;
    }
    memset(p, 0, sizeof *p);
}

// proofs of n targets (device keys; optional device trie ids) out of one arena into a page-locked host block
static int32_t da_proofs(DArena *a, const uint32_t *d_trie_of_target, const uint8_t *d_keys, uint64_t n, b200_proofs *out) {
    b200_ctx *c = a->c;
    cudaStream_t st = c->stream;
    memset(out, 0, sizeof *out);
    ProofsOwner *owner = new ProofsOwner();
    out->_owner = owner;
    out->n_targets = n;
    DTrieDev d = da_view(a);
    uint64_t n_nodes = 0, n_bytes = 0;
    uint64_t *node_base = nullptr, *byte_base = nullptr;
    if (n) {
        // scratch: node_count u32[n] | byte_count u64[n] | node_base u64[n+1] | byte_base u64[n+1]
        TRY(da_scratch(a, a->nh, n * 4));
        TRY(da_scratch(a, a->prefix, n * 8));
        TRY(da_scratch(a, a->sel, (n + 1) * 8));
        TRY(da_scratch(a, a->pick, (n + 1) * 8));
        uint32_t *node_count = static_cast<uint32_t *>(a->nh.p);
        uint64_t *byte_count = static_cast<uint64_t *>(a->prefix.p);
        node_base = static_cast<uint64_t *>(a->sel.p);
        byte_base = static_cast<uint64_t *>(a->pick.p);
        CU(launch_dt_proof_sizes(d, d_trie_of_target, d_keys, n, node_count, byte_count, st));
        size_t t1 = 0, t2 = 0;
        CU(cub::DeviceScan::ExclusiveSum(nullptr, t1, node_count, node_base, (int64_t)n, st));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, t2, byte_count, byte_base, (int64_t)n, st));
        ENSURE(cub_temp, std::max(t1, t2));
        CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t1, node_count, node_base, (int64_t)n, st));
        CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t2, byte_count, byte_base, (int64_t)n, st));
        c->launches += 3;
        uint64_t *ps = reinterpret_cast<uint64_t *>(static_cast<uint32_t *>(c->pinned_small) + 200);
        uint32_t *ps32 = static_cast<uint32_t *>(c->pinned_small) + 210;
        CU(cudaMemcpyAsync(ps, node_base + (n - 1), 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 1, byte_base + (n - 1), 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 2, byte_count + (n - 1), 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps32, node_count + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        TRY(sync_and_status(c));
        n_nodes = ps[0] + ps32[0];
        n_bytes = ps[1] + ps[2];
    }
    // host block: node_offset u64[n+1] | rlp_offset u64[n_nodes+1] | rlp bytes
    size_t o_no = 0, o_ro = (n + 1) * 8, o_rlp = o_ro + (n_nodes + 1) * 8, total = o_rlp + n_bytes + 16;
    CU(cudaMallocHost(&owner->host, total));
    uint8_t *h = static_cast<uint8_t *>(owner->host);
    out->node_offset = reinterpret_cast<uint64_t *>(h + o_no);
    out->rlp_offset = reinterpret_cast<uint64_t *>(h + o_ro);
    out->rlp = h + o_rlp;
    out->n_nodes = n_nodes;
    if (n) {
        TRY(da_scratch(a, a->out, (n_nodes + 1) * 8 + n_bytes + 16));
        uint64_t *d_ro = static_cast<uint64_t *>(a->out.p);
        uint8_t *d_rlp = reinterpret_cast<uint8_t *>(d_ro + n_nodes + 1);
        CU(launch_dt_proof_write(d, d_trie_of_target, d_keys, n, node_base, byte_base, d_rlp, d_ro, st));
        c->launches++;
        CU(cudaMemcpyAsync(out->node_offset, node_base, n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->rlp_offset, d_ro, n_nodes * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->rlp, d_rlp, n_bytes, cudaMemcpyDeviceToHost, st));
        TRY(sync_and_status(c));
    }
    out->node_offset[n] = n_nodes;
    out->rlp_offset[n_nodes] = n_bytes;
    return B200_OK;
}

// Account proofs (eth_getProof / Proof::account_proof, crates/trie/trie/src/proof/mod.rs): for every target hashed address
// the nodes from the state root down to its leaf — or down to where the trie shows the account does not exist.
extern "C" B200_API int32_t b200_dstate_account_proofs(b200_dstate *t, const uint8_t *acct_keys32, uint64_t n, b200_proofs *out) {
    if (!t || !out || (n && !acct_keys32)) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    memset(out, 0, sizeof *out);
    if (t->sharded) return fail(c, B200_ERR_INVALID_ARG, "proofs of a sharded state start at the virtual root branch: not supported");
    if (n >= (1ull << 24)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^24-1 proof targets per call");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(reset_build_state(c));
    TRY(h2d_into(&t->acc, t->in_akeys, acct_keys32, n * 32));
    int32_t r = da_proofs(&t->acc, nullptr, static_cast<const uint8_t *>(t->in_akeys.p), n, out);
    if (r != B200_OK) b200_proofs_release(out);
    return r;
}

// Storage proofs of one account (Proof::storage_proof): slot targets are hashed slot keys.  storage_root32 receives the
// account's storage root (EMPTY_ROOT_HASH, and the one-node proof 0x80 per slot, if the account or its storage is absent,
// crates/trie/db/tests/proof.rs:105-132).
extern "C" B200_API int32_t b200_dstate_storage_proofs(b200_dstate *t, const uint8_t *acct_key32, const uint8_t *slot_keys32,
                                                       uint64_t n, uint8_t storage_root32[32], b200_proofs *out) {
    if (!t || !out || !acct_key32 || (n && !slot_keys32)) return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    memset(out, 0, sizeof *out);
    if (n >= (1ull << 24)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^24-1 proof targets per call");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    TRY(reset_build_state(c));
    TRY(h2d_into(&t->acc, t->in_akeys, acct_key32, 32));
    TRY(h2d_into(&t->sto, t->in_skeys, slot_keys32, n * 32));
    TRY(da_scratch(&t->acc, t->trie_of_key, (n + 1) * 4));
    uint32_t *d_tries = static_cast<uint32_t *>(t->trie_of_key.p);
    DTrieDev da = da_view(&t->acc);
    CU(launch_dt_find_leaf(da, static_cast<const uint8_t *>(t->in_akeys.p), d_tries, n + 1, st));  // [n] = the leaf itself
    c->launches++;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    CU(cudaMemcpyAsync(ps + 220, d_tries + n, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint32_t leaf = ps[220];
    if (storage_root32) {
        static const uint8_t EMPTY[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                          0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};
        if (leaf == 0xFFFFFFFFu) memcpy(storage_root32, EMPTY, 32);
        else CU(cudaMemcpyAsync(storage_root32, static_cast<uint8_t *>(t->acc.lsroot.p) + 32 * (size_t)leaf, 32, cudaMemcpyDeviceToHost, st));
    }
    t->sto.top_out = static_cast<uint8_t *>(t->acc.lsroot.p);
    t->sto.top_stride = 32;
    int32_t r = da_proofs(&t->sto, d_tries, static_cast<const uint8_t *>(t->in_skeys.p), n, out);
    if (r != B200_OK) b200_proofs_release(out);
    return r;
}
