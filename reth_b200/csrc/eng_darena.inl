// eng_darena.inl — the arena behind the dynamic resident tries (b200_dtrie_*: the account trie, eng_dtrie.inl; b200_dstate_*:
// the account trie plus every storage trie, eng_dstate.inl): 16-slot branch nodes in HBM that take a block's upserts and
// deletes in place and re-hash only the touched paths (device side: tk_dtrie.cuh).  Capacity, views, one block's
// restructure / re-hash, collection of TrieUpdates.
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
    DevBuf kind, leaf_of, list_a, list_b, ins_idx, attach, unlock, seeds, built, removed, freed_now, flags, nh, sel, prefix, pick, out;
    // outcome of the last apply
    uint32_t n_built = 0, n_removed = 0;
    bool marked = false;  // the last restructure also marked the dirty paths (fused small-block form)
    uint32_t val_stride() const { return account ? 72u : 32u; }
};

static int32_t da_resize(DArena *a, DevBuf &b, size_t new_bytes, size_t keep_bytes, int fill /* -1 none, else byte */) {
    b200_ctx *c = a->c;
    if (new_bytes <= b.cap) return B200_OK;
    void *p = nullptr;
    size_t want = new_bytes + 256;
    if (c->phase_timing) fprintf(stderr, "[b200 grow] %zu -> %zu bytes (keep %zu)\n", b.cap, want, keep_bytes);
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
// Per-block scratch: grows by half again, so that blocks of slowly rising size do not pay a cudaMalloc + cudaFree (each an
// implicit device synchronisation: 0.5 - 100 ms measured on a B200 with a few GB resident) every time one is the largest so far.
static int32_t da_scratch(DArena *a, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return B200_OK;
    return da_resize(a, b, bytes + bytes / 2 + 4096, 0, -1);
}

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
    d.unlock = static_cast<uint32_t *>(a->unlock.p);
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
                      &a->list_b, &a->ins_idx, &a->attach, &a->unlock, &a->seeds, &a->built, &a->removed, &a->freed_now, &a->flags, &a->nh,
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
// Every allocation one block of m entries can need (capacity for m inserts, all scratch): called for ALL arenas of a
// block before the first kernel mutates anything, so that running out of memory cannot leave a state half-applied.
static int32_t da_prepare(DArena *a, uint64_t m, uint64_t tries) {
    // every insert may take one leaf slot and one node slot from the bump region
    TRY(da_reserve(a, (uint64_t)a->leaf_alloc + m, (uint64_t)a->node_alloc + m, std::max<uint64_t>(tries, a->tcap)));
    const uint32_t max_list = (uint32_t)m + 16, max_seeds = (uint32_t)(6 * m + 64);
    const uint32_t max_built = (uint32_t)(std::min<uint64_t>((uint64_t)max_seeds * 64, (uint64_t)a->node_alloc + m) + 16);
    TRY(da_scratch(a, a->kind, m));
    TRY(da_scratch(a, a->leaf_of, m * 4));
    TRY(da_scratch(a, a->list_a, (size_t)max_list * 4));
    TRY(da_scratch(a, a->list_b, (size_t)max_list * 4));
    TRY(da_scratch(a, a->ins_idx, m * 4));
    TRY(da_scratch(a, a->attach, m * 8));
    TRY(da_scratch(a, a->unlock, m * 4));
    TRY(da_scratch(a, a->seeds, (size_t)max_seeds * 4));
    TRY(da_scratch(a, a->built, (size_t)max_built * 4));
    TRY(da_scratch(a, a->removed, ((size_t)max_built + max_list) * 4));
    TRY(da_scratch(a, a->freed_now, (size_t)max_list * 4));
    TRY(da_scratch(a, a->flags, max_list));  // per-entry defer flags of a collapse round (re-used for the output flags)
    TRY(da_scratch(a, a->nh, m + 16));       // per-entry "still to insert" flags of an insert round
    TRY(da_scratch(a, a->sel, ((size_t)max_seeds + 1) * 4));  // second insert list / hand-over list of the two-stage re-hash
    return B200_OK;
}

// Blocks of up to this many entries are restructured by one CTA in one launch (B200_DT_FUSED_MAX overrides; 0 = never).
static uint64_t dt_fused_max() {
    static const uint64_t v = [] {
        const char *e = getenv("B200_DT_FUSED_MAX");
        return e ? strtoull(e, nullptr, 10) : (uint64_t)8192;
    }();
    return v;
}

static int32_t da_restructure(DArena *a, const uint32_t *d_trie_of_key, const uint8_t *d_keys, const uint8_t *d_vals,
                              const uint8_t *d_flags, const uint8_t *d_sroots, uint64_t m) {
    b200_ctx *c = a->c;
    cudaStream_t st = c->stream;
    TRY(da_prepare(a, m, a->tcap));  // no-op when the caller prepared already
    DTrieDev d = da_view(a);
    uint8_t *kind = static_cast<uint8_t *>(a->kind.p);
    uint32_t *leaf_of = static_cast<uint32_t *>(a->leaf_of.p);
    CU(cudaMemsetAsync(d.g + DG_SEEDS, 0, (DG_WORDS - DG_SEEDS) * 4, st));  // the per-apply list lengths
    if (m <= dt_fused_max()) {  // small block: one CTA does every phase below without coming back to the host
        CU(launch_dt_restructure_fused(d, d_trie_of_key, d_keys, d_vals, d_flags, d_sroots, (uint32_t)m, kind, leaf_of,
                                       static_cast<uint32_t *>(a->list_a.p), static_cast<uint32_t *>(a->list_b.p),
                                       static_cast<uint8_t *>(a->flags.p), static_cast<uint32_t *>(a->ins_idx.p),
                                       static_cast<uint32_t *>(a->sel.p), static_cast<uint64_t *>(a->attach.p),
                                       static_cast<uint8_t *>(a->nh.p), 8, st));
        c->launches++;
        a->marked = true;
        return B200_OK;
    }
    a->marked = false;
    // ---- locate, value updates, detach deleted leaves
    CU(launch_dt_locate(d, d_trie_of_key, d_keys, d_vals, d_flags, m, kind, leaf_of, st));
    uint32_t *list_cur = static_cast<uint32_t *>(a->list_a.p), *list_next = static_cast<uint32_t *>(a->list_b.p);
    uint32_t *cnt_cur = d.g + DG_LIST_A, *cnt_next = d.g + DG_LIST_B;
    CU(launch_dt_update_detach(d, d_vals, d_sroots, m, kind, leaf_of, list_cur, st));
    c->launches += 2;
    phase_mark(c, "r:locate+detach");
    // ---- collapse rounds until no node is left that lost children
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    for (int round = 0;; round++) {
        CU(cudaMemcpyAsync(ps + 200, cnt_cur, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 201, small_u32(c) + SM_ERR, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (ps[201] != B200_DEVERR_NONE) return report_dev_error_now(c, (int)ps[201]);
        phase_mark(c, "r:collapse-sync");
        if (ps[200] == 0) break;
        if (round > 200) return fail(c, B200_ERR_CUDA, "collapse rounds do not converge");
        CU(cudaMemsetAsync(cnt_next, 0, 4, st));
        CU(launch_dt_collapse_round(d, list_cur, cnt_cur, ps[200], static_cast<uint8_t *>(a->flags.p), list_next, cnt_next, st));
        c->launches += 4;
        std::swap(list_cur, list_next);
        std::swap(cnt_cur, cnt_next);
    }
    // ---- inserts, in rounds: the dense list of insert entries; per round their attach points, then one thread per run
    // inserts up to DT_KEYS_PER_RUN of its keys; what is left is compacted (order kept) and goes round again
    constexpr uint32_t DT_KEYS_PER_RUN = 8;
    uint32_t *idx_cur = static_cast<uint32_t *>(a->ins_idx.p);
    thrust::counting_iterator<uint32_t> counting(0);
    auto is_insert = thrust::make_transform_iterator(kind, IsKind{DK_INSERT});
    size_t t_sel = 0;
    CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, is_insert, idx_cur, d.g + DG_NINSERT, (int64_t)m, st));
    ENSURE(cub_temp, t_sel);
    CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, is_insert, idx_cur, d.g + DG_NINSERT, (int64_t)m, st));
    c->launches++;
    TRY(da_scratch(a, a->nh, m + 16));  // per-entry "still to insert" flags of a round
    uint8_t *pending = static_cast<uint8_t *>(a->nh.p);
    uint32_t *leftover = d.g + DG_LIST_A;  // free again: the collapse rounds are over
    uint32_t *idx_next = nullptr;
    uint64_t bound = m;  // upper bound of the entries still to insert
    for (int round = 0;; round++) {
        CU(cudaMemsetAsync(leftover, 0, 4, st));
        CU(launch_dt_insert(d, d_trie_of_key, d_keys, d_vals, d_sroots, idx_cur, d.g + DG_NINSERT, bound,
                            static_cast<uint64_t *>(a->attach.p), leaf_of, DT_KEYS_PER_RUN, pending, leftover, st));
        c->launches += 3;
        CU(cudaMemcpyAsync(ps + 200, leftover, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 201, small_u32(c) + SM_ERR, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 202, d.g + DG_NINSERT, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (ps[201] != B200_DEVERR_NONE) return report_dev_error_now(c, (int)ps[201]);
        phase_mark(c, "r:insert-round");
        if (ps[200] == 0) break;
        if (round > 80) return fail(c, B200_ERR_CUDA, "insert rounds do not converge");
        if (!idx_next) {
            TRY(da_scratch(a, a->sel, m * 4));
            idx_next = static_cast<uint32_t *>(a->sel.p);
        }
        const int64_t live = ps[202];  // entries of this round: exactly those whose `pending` flag was just written
        size_t t2 = 0;
        CU(cub::DeviceSelect::Flagged(nullptr, t2, idx_cur, pending, idx_next, d.g + DG_NINSERT, live, st));
        ENSURE(cub_temp, t2);
        CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t2, idx_cur, pending, idx_next, d.g + DG_NINSERT, live, st));
        c->launches++;
        std::swap(idx_cur, idx_next);
        bound = ps[200];
    }
    return B200_OK;
}

// re-hash of the seeded paths, recycling of the freed nodes (asynchronous)
// Small dirty sets: one warp per seed (latency).  Large ones: one thread per seed through the populous levels, warps for
// the sparse top of the single account trie (split depth 3: <= 4096 nodes per level above it); a storage forest has no
// common sparse top, its threads climb all the way.  B200_DT_TWO_STAGE_MIN overrides the switch-over (tuning / tests).
static uint64_t dt_two_stage_min() {
    static const uint64_t v = [] {
        const char *e = getenv("B200_DT_TWO_STAGE_MIN");
        return e ? strtoull(e, nullptr, 10) : (uint64_t)WARP_LEVEL_MAX;
    }();
    return v;
}
static int32_t da_rehash(DArena *a, uint64_t m) {
    b200_ctx *c = a->c;
    const uint32_t max_seeds = (uint32_t)(6 * m + 64);
    uint32_t *handoff = nullptr, *handoff_count = nullptr;
    if (m > dt_two_stage_min()) {  // every seed hands over at most once
        TRY(da_scratch(a, a->sel, ((size_t)max_seeds + 1) * 4));
        handoff = static_cast<uint32_t *>(a->sel.p) + 1;
        handoff_count = static_cast<uint32_t *>(a->sel.p);
        CU(cudaMemsetAsync(handoff_count, 0, 4, c->stream));
    }
    DTrieDev d = da_view(a);
    CU(launch_dt_rehash(d, max_seeds, handoff, handoff_count, (a->forest && !a->account) ? 0 : 3, a->marked, c->stream));
    a->marked = false;
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
    const UpdatesLayout lay(n_stored, n_hashes);
    if (!(owner->host = pinned_block_alloc(lay.host_total ? lay.host_total : 16))) return fail(c, B200_ERR_OOM, "page-locked result block");
    uint8_t *h = static_cast<uint8_t *>(owner->host);
    lay.bind_host(u, h, n_stored);
    if (n_stored) {
        TRY(da_scratch(a, a->out, lay.dev_total));
        uint8_t *dv = static_cast<uint8_t *>(a->out.p);
        CU(launch_dt_gather_updates(d, pick_ids, n_stored, pick_prefix, lay.bind_dev(dv), st));
        c->launches++;
        CU(cudaMemcpyAsync(h, dv, lay.dev_total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    lay.finish_host(u, h, n_stored, n_hashes);
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
    if (!(owner->host = pinned_block_alloc(total))) return fail(c, B200_ERR_OOM, "page-locked result block");
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

