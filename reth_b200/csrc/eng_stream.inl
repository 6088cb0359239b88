// eng_stream.inl — streamed / resumable state root: the state arrives in ascending account-key ranges,
// each push is built and folded away, and only the open right edge of the account trie is carried to the next push.
// Part of the single translation unit engine.cu (textually included, in this order).
//
// What reth does with StateRoot::with_threshold / root_with_progress / with_intermediate_state
// (crates/trie/trie/src/trie.rs:73-85,156-330; progress.rs) and MerkleStage's chunked rebuild (merkle.rs:184-366): stop
// after `threshold` updates, hand back the HashBuilder stack + walker position, resume later — so that a state larger than
// memory (or than one database transaction) is committed range by range.  HashBuilder's open state is its stack: the
// unfinished nodes on the path to the last key.  The data-parallel equivalent of "everything left of the last key is
// finished" is: every top-nibble bucket of the hashed address space that a later key has closed is built completely
// (storage tries, account leaves, the bucket's subtrie) and survives only as one frontier entry (68 bytes, the same entry
// the multi-GPU path all-gathers); the accounts of the still open bucket — at most 1/16 of the state, 136 bytes each — stay
// in HBM.  b200_root_from_frontier folds the 16 entries at the end.  Stored nodes (TrieUpdates) leave with the push that
// closes them, like reth's per-chunk updates.

struct b200_root_stream {
    b200_ctx *c = nullptr;
    bool retain = false, finished = false;
    DevBuf ck, ca, cs;  // the open bucket: account keys, accounts, storage roots
    DevBuf wk, wa, ws;  // the buckets one push closes: carried + new accounts, contiguous
    uint64_t n_carry = 0;
    int carry_nibble = -1;
    bool have_last = false;
    uint8_t last_key[32] = {};
    b200_frontier_entry fr[16] = {};
    uint64_t accounts = 0, slots = 0;
    uint32_t closed_mask = 0;
};

static void stream_free_buffers(b200_root_stream *s) {
    for (DevBuf *b : {&s->ck, &s->ca, &s->cs, &s->wk, &s->wa, &s->ws})
        if (b->p) {
            cudaFree(b->p);
            s->c->dev_bytes -= b->cap;
            *b = DevBuf{};
        }
}

extern "C" B200_API int32_t b200_root_stream_begin(b200_ctx *c, int32_t retain_updates, b200_root_stream **out) {
    if (!c || !out) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    b200_root_stream *s = new b200_root_stream();
    s->c = c;
    s->retain = retain_updates != 0;
    *out = s;
    return B200_OK;
}

extern "C" B200_API void b200_root_stream_free(b200_root_stream *s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> g(s->c->mu);
        cudaSetDevice(s->c->device);
        cudaStreamSynchronize(s->c->stream);
        stream_free_buffers(s);
    }
    delete s;
}

// Builds the accounts w[0, n) (whole top-nibble buckets, ascending) as bucket tries, merges their frontier entries into
// the stream and, with `upd`, hands out their stored nodes.  Synchronises.
static int32_t stream_close_buckets(b200_root_stream *s, uint64_t n, b200_updates *upd) {
    b200_ctx *c = s->c;
    if (upd) memset(upd, 0, sizeof *upd);
    if (n == 0) return B200_OK;
    ENSURE(buckets, 17 * 8);
    ENSURE(out_a, 16 * sizeof(FrontierEntryDev));
    uint64_t *d_buckets = static_cast<uint64_t *>(c->buckets.p);
    const uint8_t *wk = static_cast<const uint8_t *>(s->wk.p), *wa = static_cast<const uint8_t *>(s->wa.p),
                  *ws = static_cast<const uint8_t *>(s->ws.p);
    Built ba;
    CU(launch_nibble_buckets(wk, n, d_buckets, c->stream));
    TRY(build_forest(c, wk, n, d_buckets, 16, true, wa, ws, s->retain && upd != nullptr, ba));
    CU(launch_frontier(ba.f, d_buckets, wa, ws, static_cast<FrontierEntryDev *>(c->out_a.p), c->stream));
    c->launches += 2;
    c->stats.leaves_added += n;
    c->stats.branches_added += ba.n_nodes;
    c->stats.levels += ba.levels;
    b200_frontier_entry got[16];
    CU(cudaMemcpyAsync(got, c->out_a.p, sizeof got, cudaMemcpyDeviceToHost, c->stream));
    TRY(finish_build_state(c));
    TRY(sync_and_status(c));
    for (int i = 0; i < 16; i++)
        if (got[i].as_child_len || got[i].as_root_len) {
            if (s->closed_mask & (1u << i)) return fail(c, B200_ERR_UNSORTED, "stream: bucket %d was closed by an earlier push", i);
            s->fr[i] = got[i];
            s->closed_mask |= 1u << i;
        }
    if (upd && s->retain) TRY(collect_updates(c, ba, d_buckets, 16, upd));
    return B200_OK;
}

extern "C" B200_API int32_t b200_root_stream_push(b200_root_stream *s, const uint8_t *acct_keys32, const b200_account *accts,
                                                  uint64_t n_accounts, const uint8_t *slot_keys32, const uint8_t *values32_be,
                                                  const uint64_t *seg_offsets, b200_updates *opt_account_updates,
                                                  b200_updates *opt_storage_updates, b200_stream_progress *opt_progress) {
    if (!s) return B200_ERR_INVALID_ARG;
    b200_ctx *c = s->c;
    if (opt_account_updates) memset(opt_account_updates, 0, sizeof *opt_account_updates);
    if (opt_storage_updates) memset(opt_storage_updates, 0, sizeof *opt_storage_updates);
    std::lock_guard<std::mutex> g(c->mu);
    if (s->finished) return fail(c, B200_ERR_INVALID_ARG, "stream: already finished");
    if (!seg_offsets || (n_accounts && (!acct_keys32 || !accts))) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    TRY(check_offsets_host(c, seg_offsets, n_accounts));
    const uint64_t n_slots = seg_offsets[n_accounts];
    if (n_slots && (!slot_keys32 || !values32_be)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n_accounts == 0) {
        if (opt_progress) *opt_progress = b200_stream_progress{s->accounts, s->slots, s->n_carry, s->closed_mask};
        return B200_OK;
    }
    if (s->have_last && memcmp(acct_keys32, s->last_key, 32) <= 0)
        return fail(c, B200_ERR_UNSORTED, "stream: a push must start after the last key of the previous push");
    CU(cudaSetDevice(c->device));
    const uint64_t n = n_accounts;
    TRY(h2d(c, c->in_a, slot_keys32, n_slots * 32));
    TRY(h2d(c, c->in_b, values32_be, n_slots * 32));
    TRY(h2d(c, c->in_c, seg_offsets, (n + 1) * 8));
    TRY(h2d(c, c->in_d, acct_keys32, n * 32));
    TRY(h2d(c, c->in_e, accts, n * sizeof(b200_account)));
    ENSURE(sroots, n * 32);
    // ---- storage tries of this range: complete on their own
    TRY(reset_build_state(c));
    Built bs;
    TRY(storage_roots_on_device(c, static_cast<const uint8_t *>(c->in_a.p), static_cast<const uint8_t *>(c->in_b.p),
                                static_cast<const uint64_t *>(c->in_c.p), n, n_slots, static_cast<uint8_t *>(c->sroots.p),
                                s->retain && opt_storage_updates, bs));
    if (s->retain && opt_storage_updates) {  // the forest's scratch is reused by the account build: gather first
        TRY(finish_build_state(c));
        TRY(sync_and_status(c));
        TRY(collect_updates(c, bs, static_cast<const uint64_t *>(c->in_c.p), n, opt_storage_updates));
        TRY(reset_build_state(c));
    }
    // ---- accounts: everything below the top nibble of the last key closes now, the rest stays open
    const int last_nib = acct_keys32[32 * (n - 1)] >> 4;
    uint64_t split = 0;  // first account of this push inside the bucket that stays open
    {
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if ((acct_keys32[32 * mid] >> 4) < last_nib) lo = mid + 1;
            else hi = mid;
        }
        split = lo;
    }
    cudaStream_t st = c->stream;
    const uint8_t *nk = static_cast<const uint8_t *>(c->in_d.p), *na = static_cast<const uint8_t *>(c->in_e.p),
                  *ns = static_cast<const uint8_t *>(c->sroots.p);
    const size_t AS = sizeof(b200_account);
    int32_t r = B200_OK;
    if (s->n_carry && s->carry_nibble == last_nib) {  // the whole push stays inside the open bucket
        const uint64_t tot = s->n_carry + n;
        // grow the carry (ensure() would drop the contents): allocate, copy, swap
        for (int which = 0; which < 3; which++) {
            DevBuf &b = which == 0 ? s->ck : (which == 1 ? s->ca : s->cs);
            const size_t unit = which == 1 ? AS : 32;
            if (tot * unit > b.cap) {
                DevBuf nb;
                TRY(ensure(c, nb, tot * unit + tot * unit / 2));
                CU(cudaMemcpyAsync(nb.p, b.p, s->n_carry * unit, cudaMemcpyDeviceToDevice, st));
                CU(cudaStreamSynchronize(st));
                cudaFree(b.p);
                c->dev_bytes -= b.cap;
                b = nb;
            }
            const uint8_t *src = which == 0 ? nk : (which == 1 ? na : ns);
            CU(cudaMemcpyAsync(static_cast<uint8_t *>(b.p) + s->n_carry * unit, src, n * unit, cudaMemcpyDeviceToDevice, st));
        }
        s->n_carry = tot;
        TRY(finish_build_state(c));
        r = sync_and_status(c);
    } else {
        const uint64_t n_close = s->n_carry + split, n_open = n - split;
        TRY(ensure(c, s->wk, (n_close ? n_close : 1) * 32));
        TRY(ensure(c, s->wa, (n_close ? n_close : 1) * AS));
        TRY(ensure(c, s->ws, (n_close ? n_close : 1) * 32));
        if (s->n_carry) {
            CU(cudaMemcpyAsync(s->wk.p, s->ck.p, s->n_carry * 32, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(s->wa.p, s->ca.p, s->n_carry * AS, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(s->ws.p, s->cs.p, s->n_carry * 32, cudaMemcpyDeviceToDevice, st));
        }
        if (split) {
            CU(cudaMemcpyAsync(static_cast<uint8_t *>(s->wk.p) + s->n_carry * 32, nk, split * 32, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(static_cast<uint8_t *>(s->wa.p) + s->n_carry * AS, na, split * AS, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(static_cast<uint8_t *>(s->ws.p) + s->n_carry * 32, ns, split * 32, cudaMemcpyDeviceToDevice, st));
        }
        TRY(ensure(c, s->ck, n_open * 32));  // (the old carry has been copied out on the same stream; ensure() syncs before freeing)
        TRY(ensure(c, s->ca, n_open * AS));
        TRY(ensure(c, s->cs, n_open * 32));
        CU(cudaMemcpyAsync(s->ck.p, nk + split * 32, n_open * 32, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(s->ca.p, na + split * AS, n_open * AS, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(s->cs.p, ns + split * 32, n_open * 32, cudaMemcpyDeviceToDevice, st));
        s->n_carry = n_open;
        s->carry_nibble = last_nib;
        if (n_close) {
            r = stream_close_buckets(s, n_close, opt_account_updates);
        } else {
            TRY(finish_build_state(c));
            r = sync_and_status(c);
        }
    }
    if (r != B200_OK) {
        if (opt_account_updates) b200_updates_release(opt_account_updates);
        if (opt_storage_updates) b200_updates_release(opt_storage_updates);
        return r;
    }
    memcpy(s->last_key, acct_keys32 + 32 * (n - 1), 32);
    s->have_last = true;
    s->accounts += n;
    s->slots += n_slots;
    if (opt_progress) *opt_progress = b200_stream_progress{s->accounts, s->slots, s->n_carry, s->closed_mask};
    return B200_OK;
}

extern "C" B200_API int32_t b200_root_stream_finish(b200_root_stream *s, uint8_t root32[32], b200_updates *opt_account_updates) {
    if (!s || !root32) return B200_ERR_INVALID_ARG;
    b200_ctx *c = s->c;
    if (opt_account_updates) memset(opt_account_updates, 0, sizeof *opt_account_updates);
    {
        std::lock_guard<std::mutex> g(c->mu);
        if (s->finished) return fail(c, B200_ERR_INVALID_ARG, "stream: already finished");
        CU(cudaSetDevice(c->device));
        if (s->n_carry) {  // the open bucket closes: it becomes the work set
            std::swap(s->wk, s->ck);
            std::swap(s->wa, s->ca);
            std::swap(s->ws, s->cs);
            const uint64_t n_close = s->n_carry;
            s->n_carry = 0;
            TRY(reset_build_state(c));
            int32_t r = stream_close_buckets(s, n_close, opt_account_updates);
            if (r != B200_OK) {
                if (opt_account_updates) b200_updates_release(opt_account_updates);
                return r;
            }
        }
        s->finished = true;
        stream_free_buffers(s);
    }
    return b200_root_from_frontier(c, s->fr, root32);  // takes the lock itself; 16 empty entries give EMPTY_ROOT_HASH
}

// The resumable part of a stream, to be kept across a restart (the role of MerkleCheckpoint, crates/stages/types/src/
// checkpoints.rs, written by MerkleStage::save_execution_checkpoint, merkle.rs:118-148): the closed buckets' frontier entries
// and the nibble of the open bucket.  The open bucket's accounts are NOT part of it — after b200_root_stream_resume the
// caller pushes again from the first key of that bucket (top nibble == *resume_nibble; 16 = nothing left).
extern "C" B200_API int32_t b200_root_stream_checkpoint(const b200_root_stream *s, b200_stream_checkpoint *out) {
    if (!s || !out) return B200_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    memcpy(out->frontier, s->fr, sizeof s->fr);
    out->closed_mask = s->closed_mask;
    // the bucket of the last key is always open while the stream runs: it is re-pushed after a resume
    out->resume_nibble = s->finished ? 16u : (s->n_carry ? (uint32_t)s->carry_nibble : (s->have_last ? (uint32_t)(s->last_key[0] >> 4) + 1u : 0u));
    out->retain_updates = s->retain ? 1 : 0;
    return B200_OK;
}
extern "C" B200_API int32_t b200_root_stream_resume(b200_ctx *c, const b200_stream_checkpoint *cp, b200_root_stream **out) {
    if (!c || !cp || !out || cp->resume_nibble > 16) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    for (int i = 0; i < 16; i++) {
        const bool closed = (cp->closed_mask >> i) & 1;
        const bool has = cp->frontier[i].as_child_len || cp->frontier[i].as_root_len;
        if ((has && !closed) || (closed && (uint32_t)i >= cp->resume_nibble) || cp->frontier[i].as_child_len > 33)
            return fail(c, B200_ERR_INVALID_ARG, "malformed stream checkpoint (bucket %d)", i);
    }
    b200_root_stream *s = new b200_root_stream();
    s->c = c;
    s->retain = cp->retain_updates != 0;
    memcpy(s->fr, cp->frontier, sizeof s->fr);
    s->closed_mask = cp->closed_mask;
    if (cp->resume_nibble > 0) {  // keys below the resume bucket are history: the next push must start at or after it
        memset(s->last_key, 0xFF, 32);
        s->last_key[0] = (uint8_t)(((cp->resume_nibble - 1) << 4) | 0x0F);
        s->have_last = true;
    }
    *out = s;
    return B200_OK;
}
