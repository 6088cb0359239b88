// eng_dstate.inl — b200_dstate_*: account arena + storage forest arena, optionally one rank's shard (eng_darena.inl).
// Part of the single translation unit engine.cu (textually included, in this order).

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
    DevBuf root, in_akeys, in_accts, in_aflags, in_skeys, in_svals, in_offs, trie_of_key, wipe_cnt;
};

extern "C" B200_API void b200_dstate_destroy(b200_dstate *t) {
    if (!t) return;
    cudaSetDevice(t->c->device);
    cudaStreamSynchronize(t->c->stream);
    da_free(&t->acc);
    da_free(&t->sto);
    DevBuf *bufs[] = {&t->root, &t->in_akeys, &t->in_accts, &t->in_aflags, &t->in_skeys, &t->in_svals, &t->in_offs,
                      &t->trie_of_key, &t->wipe_cnt, &t->bucket_roots, &t->frontier, &t->acct_tries};
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

// on_device: every input pointer is a device pointer and n_slots_dev gives the slot count (the segment table is then
// validated by the build itself: B200_ERR_INVALID_ARG through the sticky status)
static int32_t dstate_create_impl(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts, uint64_t n_accounts,
                                  const uint8_t *slot_keys32, const uint8_t *values32_be, const uint64_t *seg_offsets,
                                  bool sharded, bool on_device, uint64_t n_slots_dev, b200_dstate **out, uint8_t root32[32]) {
    if (!c || !out || !seg_offsets || (n_accounts && (!acct_keys32 || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    *out = nullptr;
    if (!on_device) TRY(check_offsets_host(c, seg_offsets, n_accounts));
    const uint64_t n_slots = on_device ? n_slots_dev : seg_offsets[n_accounts];
    if (n_slots && (!slot_keys32 || !values32_be)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const cudaMemcpyKind in_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    b200_trie *src_s = nullptr, *src_a = nullptr;
    TRY(forest_create_locked(c, slot_keys32, values32_be, seg_offsets, n_accounts, n_slots, in_kind, &src_s));
    // the account trie takes its storage roots straight from the forest build (device memory: cudaMemcpyDefault)
    int32_t r;
    if (sharded) {
        uint64_t bucket_offsets[17];
        if (on_device) {  // nibble_buckets_kernel: 17 binary searches on the device
            uint64_t *d_offs = reinterpret_cast<uint64_t *>(small_u32(c) + SM_HIST);
            cudaError_t e = launch_nibble_buckets(acct_keys32, n_accounts, d_offs, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(bucket_offsets, d_offs, sizeof bucket_offsets, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                b200_trie_destroy(src_s);
                return fail(c, B200_ERR_CUDA, "bucket offsets: %s", cudaGetErrorString(e));
            }
            c->launches++;
        } else {
            for (uint32_t b = 0; b <= 16; b++) {  // first account whose top nibble >= b
                uint64_t lo = 0, hi = n_accounts;
                while (lo < hi) {
                    uint64_t mid = (lo + hi) >> 1;
                    if ((uint32_t)(acct_keys32[32 * mid] >> 4) < b) lo = mid + 1;
                    else hi = mid;
                }
                bucket_offsets[b] = lo;
            }
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
        if (root32) CU(cudaMemcpyAsync(root32, t->root.p, 32, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
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
    return dstate_create_impl(c, acct_keys32, accts, n_accounts, slot_keys32, values32_be, seg_offsets, false, false, 0, out, root32);
}
// Device-resident seed (a state too large to stage through one host call is uploaded in pieces by the caller): every
// pointer is a device pointer, d_seg_offsets has n_accounts+1 entries, n_slots = d_seg_offsets[n_accounts]; `sharded`
// selects b200_dstate_create_sharded's layout.  d_root32 (nullable) is a device buffer.
extern "C" B200_API int32_t b200_dstate_create_dev(b200_ctx *c, const void *d_acct_keys32, const void *d_accts, uint64_t n_accounts,
                                                   const void *d_slot_keys32, const void *d_values32_be, const void *d_seg_offsets,
                                                   uint64_t n_slots, int32_t sharded, b200_dstate **out, void *d_root32) {
    return dstate_create_impl(c, static_cast<const uint8_t *>(d_acct_keys32), static_cast<const b200_account *>(d_accts), n_accounts,
                              static_cast<const uint8_t *>(d_slot_keys32), static_cast<const uint8_t *>(d_values32_be),
                              static_cast<const uint64_t *>(d_seg_offsets), sharded != 0, true, n_slots, out,
                              static_cast<uint8_t *>(d_root32));
}
// One rank's shard of a state that is split by top key nibble (any subset of the 16 buckets).  root32 (nullable) receives
// the root this shard has on its own; the global root is b200_root_from_frontier over the gathered b200_dstate_frontier
// entries of all ranks.
extern "C" B200_API int32_t b200_dstate_create_sharded(b200_ctx *c, const uint8_t *acct_keys32, const b200_account *accts,
                                                       uint64_t n_accounts, const uint8_t *slot_keys32,
                                                       const uint8_t *values32_be, const uint64_t *seg_offsets,
                                                       b200_dstate **out, uint8_t root32[32]) {
    return dstate_create_impl(c, acct_keys32, accts, n_accounts, slot_keys32, values32_be, seg_offsets, true, false, 0, out, root32);
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
// on_device: the block's arrays (and root32) are device memory, n_entries_dev = number of slot entries; the segment table
// cannot be checked on the host then (a malformed one mis-routes slots but stays in bounds).
static int32_t dstate_apply_impl(b200_dstate *t, const uint8_t *acct_keys32, const b200_account *accts,
                                 const uint8_t *acct_flags, uint64_t m, const uint8_t *slot_keys32,
                                 const uint8_t *values32_be, const uint64_t *seg_offsets, bool on_device, uint64_t n_entries_dev,
                                 uint8_t *root32, b200_updates *opt_acct_updated, b200_updates *opt_acct_removed,
                                 b200_updates *opt_storage_updated, b200_updates *opt_storage_removed,
                                 uint8_t *opt_storage_deleted /* [m] */, b200_stats *opt_stats) {
    if (!t || !root32 || (m && (!acct_keys32 || !accts || !seg_offsets)))
        return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    const cudaMemcpyKind in_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    b200_updates *outs[] = {opt_acct_updated, opt_acct_removed, opt_storage_updated, opt_storage_removed};
    for (b200_updates *u : outs)
        if (u) memset(u, 0, sizeof *u);
    if (m >= (1ull << 28)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^28-1 dirty accounts per apply");
    if (m && !on_device) TRY(check_offsets_host(c, seg_offsets, m));
    const uint64_t n_entries = m ? (on_device ? n_entries_dev : seg_offsets[m]) : 0;
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
        // all memory first: nothing below may fail for lack of it once the first arena has been touched
        TRY(da_prepare(A, m, 1));
        TRY(da_prepare(S, n_entries, A->lcap));
        TRY(da_resize(A, t->wipe_cnt, 16, 0, 0));
        TRY(da_scratch(A, t->trie_of_key, std::max<size_t>((size_t)m, (size_t)n_entries) * 4 + 16));
        if (t->sharded) TRY(da_scratch(A, t->acct_tries, m * 4));
        TRY(h2d_into(A, t->in_akeys, acct_keys32, m * 32, in_kind));
        TRY(h2d_into(A, t->in_accts, accts, m * 72, in_kind));
        if (acct_flags) TRY(h2d_into(A, t->in_aflags, acct_flags, m, in_kind));
        TRY(h2d_into(A, t->in_offs, seg_offsets, (m + 1) * 8, in_kind));
        if (n_entries) {
            TRY(h2d_into(S, t->in_skeys, slot_keys32, n_entries * 32, in_kind));
            TRY(h2d_into(S, t->in_svals, values32_be, n_entries * 32, in_kind));
        }
        const uint8_t *d_flags = acct_flags ? static_cast<const uint8_t *>(t->in_aflags.p) : nullptr;
        phase_mark(c, "block-in");
        // ---- accounts: structure only (their leaves are re-hashed after the storage roots are known)
        const uint32_t *d_acct_tries = nullptr;
        if (t->sharded) {  // bucket trie of every account entry = its top key nibble
            CU(launch_dt_nibble_tries(static_cast<const uint8_t *>(t->in_akeys.p), m, static_cast<uint32_t *>(t->acct_tries.p), st));
            c->launches++;
            d_acct_tries = static_cast<const uint32_t *>(t->acct_tries.p);
        }
        TRY(da_restructure(A, d_acct_tries, static_cast<const uint8_t *>(t->in_akeys.p), static_cast<const uint8_t *>(t->in_accts.p),
                           d_flags, nullptr, m));
        const uint8_t *a_kind = static_cast<const uint8_t *>(A->kind.p);
        const uint32_t *a_leaf = static_cast<const uint32_t *>(A->leaf_of.p);
        phase_mark(c, "acct-restructure");
        // ---- storage tries of destroyed / wiped accounts
        S->top_out = static_cast<uint8_t *>(A->lsroot.p);  // the account arena may have been re-allocated
        S->top_stride = 32;
        uint32_t *wc = static_cast<uint32_t *>(t->wipe_cnt.p);  // [0] tries to wipe
        CU(cudaMemsetAsync(wc, 0, 16, st));
        uint32_t *wipe_tries = static_cast<uint32_t *>(t->trie_of_key.p);  // borrowed until the storage entries are expanded
        CU(launch_dt_wipe_list(a_kind, d_flags, a_leaf, m, wipe_tries, wc, st));
        c->launches++;
        {
            DTrieDev ds = da_view(S);
            // the free stack is the BFS queue: every round visits what the round before pushed
            CU(cudaMemcpyAsync(ps + 200, ds.g + DG_NODE_FREE, 4, cudaMemcpyDeviceToHost, st));
            CU(launch_dt_wipe_begin(ds, wipe_tries, wc, (uint32_t)m, st));
            c->launches++;
            CU(cudaMemcpyAsync(ps + 201, ds.g + DG_NODE_FREE, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            uint32_t lo = ps[200], hi = ps[201];
            for (int round = 0; hi > lo; round++) {
                if (round > 70) return fail(c, B200_ERR_CUDA, "storage wipe does not terminate");
                CU(launch_dt_wipe_round(ds, lo, hi, st));
                c->launches++;
                CU(cudaMemcpyAsync(ps + 201, ds.g + DG_NODE_FREE, 4, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                lo = hi;
                hi = ps[201];
            }
        }
        phase_mark(c, "wipes");
        // ---- storage slots of the surviving accounts
        if (n_entries) {
            uint32_t *trie_of_key = static_cast<uint32_t *>(t->trie_of_key.p);
            CU(launch_dt_expand_tries(static_cast<const uint64_t *>(t->in_offs.p), m, a_kind, a_leaf, n_entries, trie_of_key, st));
            c->launches++;
            TRY(da_restructure(S, trie_of_key, static_cast<const uint8_t *>(t->in_skeys.p),
                               static_cast<const uint8_t *>(t->in_svals.p), nullptr, nullptr, n_entries));
            phase_mark(c, "storage-restructure");
            TRY(da_rehash(S, n_entries));  // roots land in the account leaves' storage-root fields
            phase_mark(c, "storage-rehash");
            c->stats.leaves_added += n_entries;
        }
        // ---- accounts: re-hash
        TRY(da_rehash(A, m));
        if (t->sharded) TRY(dstate_frontier_on_device(t));
        phase_mark(c, "acct-rehash");
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
    CU(cudaMemcpyAsync(root32, t->root.p, 32, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    TRY(sync_and_status(c));
    if (m) {
        da_take_counters(A, ps + 256);
        da_take_counters(S, ps + 256 + DG_WORDS);
        if (!n_entries) S->n_built = S->n_removed = 0;  // the storage arena's per-apply lists were not reset this block
    }
    c->stats.branches_added = A->n_built + S->n_built;
    if (opt_storage_deleted) {
        std::vector<uint8_t> h_flags;
        const uint8_t *fl = acct_flags;
        if (acct_flags && on_device && m) {  // the flags live on the device: bring them over
            h_flags.resize(m);
            CU(cudaMemcpy(h_flags.data(), acct_flags, m, cudaMemcpyDeviceToHost));
            fl = h_flags.data();
        }
        for (uint64_t i = 0; i < m; i++)
            opt_storage_deleted[i] = (h_kind[i] == DK_DELETE || (fl && (fl[i] & 4) && (h_kind[i] == DK_UPDATE || h_kind[i] == DK_TOUCH))) ? 1 : 0;
    }
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


extern "C" B200_API int32_t b200_dstate_apply(b200_dstate *t, const uint8_t *acct_keys32, const b200_account *accts,
                                              const uint8_t *acct_flags, uint64_t m, const uint8_t *slot_keys32,
                                              const uint8_t *values32_be, const uint64_t *seg_offsets, uint8_t root32[32],
                                              b200_updates *opt_acct_updated, b200_updates *opt_acct_removed,
                                              b200_updates *opt_storage_updated, b200_updates *opt_storage_removed,
                                              uint8_t *opt_storage_deleted /* [m] */, b200_stats *opt_stats) {
    return dstate_apply_impl(t, acct_keys32, accts, acct_flags, m, slot_keys32, values32_be, seg_offsets, false, 0, root32,
                             opt_acct_updated, opt_acct_removed, opt_storage_updated, opt_storage_removed, opt_storage_deleted,
                             opt_stats);
}
// The block already in device memory (hashed and sorted there, e.g. by b200_hash_sort_*): every input pointer and d_root32
// are device pointers, n_entries = d_seg_offsets[m].  The update records, if wanted, still arrive in host memory.
extern "C" B200_API int32_t b200_dstate_apply_dev(b200_dstate *t, const void *d_acct_keys32, const void *d_accts,
                                                  const void *d_acct_flags, uint64_t m, const void *d_slot_keys32,
                                                  const void *d_values32_be, const void *d_seg_offsets, uint64_t n_entries,
                                                  void *d_root32, b200_updates *opt_acct_updated, b200_updates *opt_acct_removed,
                                                  b200_updates *opt_storage_updated, b200_updates *opt_storage_removed,
                                                  uint8_t *opt_storage_deleted, b200_stats *opt_stats) {
    return dstate_apply_impl(t, static_cast<const uint8_t *>(d_acct_keys32), static_cast<const b200_account *>(d_accts),
                             static_cast<const uint8_t *>(d_acct_flags), m, static_cast<const uint8_t *>(d_slot_keys32),
                             static_cast<const uint8_t *>(d_values32_be), static_cast<const uint64_t *>(d_seg_offsets), true,
                             n_entries, static_cast<uint8_t *>(d_root32), opt_acct_updated, opt_acct_removed,
                             opt_storage_updated, opt_storage_removed, opt_storage_deleted, opt_stats);
}
