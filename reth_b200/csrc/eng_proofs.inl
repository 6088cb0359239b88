// eng_proofs.inl — Merkle proofs out of the dynamic state (device side: tk_proofs.cuh).
// Part of the single translation unit engine.cu (textually included, in this order).

// ------------------------------------------------------------------------------------------------ proofs
struct ProofsOwner {
    void *host = nullptr;
};
extern "C" B200_API void b200_proofs_release(b200_proofs *p) {
    if (!p) return;
    if (p->_owner) {
        ProofsOwner *o = static_cast<ProofsOwner *>(p->_owner);
        pinned_block_free(o->host);
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
    // host block: node_offset u64[n+1] | rlp_offset u64[n_nodes+1] | node_masks u32[n_nodes] | node_depth u8[n_nodes] | rlp bytes
    size_t o_no = 0, o_ro = (n + 1) * 8, o_nm = o_ro + (n_nodes + 1) * 8, o_nd = o_nm + n_nodes * 4,
           o_rlp = align_up(o_nd + n_nodes, 16), total = o_rlp + n_bytes + 16;
    if (!(owner->host = pinned_block_alloc(total))) return fail(c, B200_ERR_OOM, "page-locked result block");
    uint8_t *h = static_cast<uint8_t *>(owner->host);
    out->node_offset = reinterpret_cast<uint64_t *>(h + o_no);
    out->rlp_offset = reinterpret_cast<uint64_t *>(h + o_ro);
    out->rlp = h + o_rlp;
    out->node_depth = h + o_nd;
    out->node_masks = reinterpret_cast<uint32_t *>(h + o_nm);
    out->n_nodes = n_nodes;
    if (n) {
        TRY(da_scratch(a, a->out, (n_nodes + 1) * 8 + n_nodes * 5 + n_bytes + 32));
        uint64_t *d_ro = static_cast<uint64_t *>(a->out.p);
        uint32_t *d_nm = reinterpret_cast<uint32_t *>(d_ro + n_nodes + 1);
        uint8_t *d_nd = reinterpret_cast<uint8_t *>(d_nm + n_nodes);
        uint8_t *d_rlp = d_nd + n_nodes;
        CU(launch_dt_proof_write(d, d_trie_of_target, d_keys, n, node_base, byte_base, d_rlp, d_ro, d_nd, d_nm, st));
        c->launches++;
        CU(cudaMemcpyAsync(out->node_offset, node_base, n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->rlp_offset, d_ro, n_nodes * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->rlp, d_rlp, n_bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->node_depth, d_nd, n_nodes, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->node_masks, d_nm, n_nodes * 4, cudaMemcpyDeviceToHost, st));
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


// Multiproof batch (Proof::multiproof over MultiProofTargets, crates/trie/trie/src/proof/mod.rs:143-193; the unit of work of
// the proof workers, crates/trie/parallel/src/proof_task.rs): n target accounts, account i with the slot targets
// slot_seg_offsets[i] .. [i+1].  One call: the account proofs, every account's storage root, and the proofs of all slot
// targets of all accounts (one pass over the storage forest).  Every proof node carries its depth, so
// MultiProof::account_subtree is { target[..node_depth] -> rlp } over account_proofs and StorageMultiProof::subtree the same
// over the slot targets of one account.
extern "C" B200_API int32_t b200_dstate_multiproof(b200_dstate *t, const uint8_t *acct_keys32, uint64_t n_accounts,
                                                   const uint64_t *slot_seg_offsets, const uint8_t *slot_keys32,
                                                   b200_proofs *account_proofs, uint8_t *storage_roots32, b200_proofs *storage_proofs) {
    if (!t || !account_proofs || !storage_proofs || !slot_seg_offsets || (n_accounts && (!acct_keys32 || !storage_roots32)))
        return fail(t ? t->c : nullptr, B200_ERR_INVALID_ARG, "bad argument");
    b200_ctx *c = t->c;
    memset(account_proofs, 0, sizeof *account_proofs);
    memset(storage_proofs, 0, sizeof *storage_proofs);
    if (t->sharded) return fail(c, B200_ERR_INVALID_ARG, "proofs of a sharded state start at the virtual root branch: not supported");
    TRY(check_offsets_host(c, slot_seg_offsets, n_accounts));
    const uint64_t m = slot_seg_offsets[n_accounts];
    if (m && !slot_keys32) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    if (n_accounts >= (1ull << 24) || m >= (1ull << 24)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^24-1 proof targets per call");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    TRY(reset_build_state(c));
    TRY(h2d_into(&t->acc, t->in_akeys, acct_keys32, n_accounts * 32));
    TRY(h2d_into(&t->sto, t->in_skeys, slot_keys32, m * 32));
    TRY(h2d_into(&t->acc, t->in_offs, slot_seg_offsets, (n_accounts + 1) * 8));
    int32_t r = da_proofs(&t->acc, nullptr, static_cast<const uint8_t *>(t->in_akeys.p), n_accounts, account_proofs);
    if (r == B200_OK) {
        // leaf (= storage trie id) and storage root of every target account, then the trie of every slot target
        TRY(da_scratch(&t->acc, t->trie_of_key, (n_accounts + m + 1) * 4));
        TRY(da_scratch(&t->acc, t->in_svals, (n_accounts ? n_accounts : 1) * 32));
        uint32_t *d_leaf = static_cast<uint32_t *>(t->trie_of_key.p), *d_tries = d_leaf + n_accounts;
        DTrieDev da = da_view(&t->acc);
        CU(launch_dt_find_leaves(da, static_cast<const uint8_t *>(t->in_akeys.p), n_accounts, d_leaf, static_cast<uint8_t *>(t->in_svals.p), st));
        CU(launch_dt_target_tries(static_cast<const uint64_t *>(t->in_offs.p), n_accounts, d_leaf, m, d_tries, st));
        c->launches += 2;
        if (n_accounts) CU(cudaMemcpyAsync(storage_roots32, t->in_svals.p, n_accounts * 32, cudaMemcpyDeviceToHost, st));
        t->sto.top_out = static_cast<uint8_t *>(t->acc.lsroot.p);
        t->sto.top_stride = 32;
        r = da_proofs(&t->sto, d_tries, static_cast<const uint8_t *>(t->in_skeys.p), m, storage_proofs);
    }
    if (r != B200_OK) {
        b200_proofs_release(account_proofs);
        b200_proofs_release(storage_proofs);
    }
    return r;
}
