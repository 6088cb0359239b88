// eng_updates.inl — TrieUpdates collection (stored BranchNodeCompact records) into page-locked host memory.
// Part of the single translation unit engine.cu (textually included, in this order).

// ------------------------------------------------------------------------------------------------ updates
struct UpdatesOwner {
    void *host = nullptr;  // one page-locked block holding every array
};

extern "C" B200_API void b200_updates_release(b200_updates *u) {
    if (!u) return;
    if (u->_owner) {
        UpdatesOwner *o = static_cast<UpdatesOwner *>(u->_owner);
        pinned_block_free(o->host);
        delete o;
    }
    memset(u, 0, sizeof *u);
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// One block holding every array of a b200_updates; the same layout on the device (gather target) and in page-locked host
// memory (what the caller receives), so a single copy brings the records over.
struct UpdatesLayout {
    size_t o_tid, o_plen, o_path, o_sm, o_tm, o_hm, o_ho32, o_hash, o_ho64, dev_total, host_total;
    UpdatesLayout(uint32_t n_stored, uint32_t n_hashes) {
        o_tid = 0;
        o_plen = align_up(o_tid + (size_t)n_stored * 4, 16);
        o_path = align_up(o_plen + n_stored, 16);
        o_sm = align_up(o_path + (size_t)n_stored * 32, 16);
        o_tm = align_up(o_sm + (size_t)n_stored * 2, 16);
        o_hm = align_up(o_tm + (size_t)n_stored * 2, 16);
        o_ho32 = align_up(o_hm + (size_t)n_stored * 2, 16);
        o_hash = align_up(o_ho32 + (size_t)n_stored * 4, 16);
        o_ho64 = align_up(o_hash + (size_t)n_hashes * 32, 16);
        dev_total = o_ho64;
        host_total = o_ho64 + ((size_t)n_stored + 1) * 8;
    }
    void bind_host(b200_updates *u, uint8_t *h, uint32_t n_stored) const {
        u->n_nodes = n_stored;
        u->trie_id = reinterpret_cast<uint32_t *>(h + o_tid);
        u->path_len = h + o_plen;
        u->path_packed = h + o_path;
        u->state_mask = reinterpret_cast<uint16_t *>(h + o_sm);
        u->tree_mask = reinterpret_cast<uint16_t *>(h + o_tm);
        u->hash_mask = reinterpret_cast<uint16_t *>(h + o_hm);
        u->hashes = h + o_hash;
        u->hash_offset = reinterpret_cast<uint64_t *>(h + o_ho64);
    }
    UpdatesDev bind_dev(uint8_t *d) const {
        UpdatesDev ud;
        ud.trie_id = reinterpret_cast<uint32_t *>(d + o_tid);
        ud.path_len = d + o_plen;
        ud.path_packed = d + o_path;
        ud.state_mask = reinterpret_cast<uint16_t *>(d + o_sm);
        ud.tree_mask = reinterpret_cast<uint16_t *>(d + o_tm);
        ud.hash_mask = reinterpret_cast<uint16_t *>(d + o_hm);
        ud.hash_offset = reinterpret_cast<uint32_t *>(d + o_ho32);
        ud.hashes = d + o_hash;
        return ud;
    }
    // after the device block has been copied to h: widen the 32-bit hash offsets and append the total
    void finish_host(b200_updates *u, const uint8_t *h, uint32_t n_stored, uint32_t n_hashes) const {
        const uint32_t *ho32 = reinterpret_cast<const uint32_t *>(h + o_ho32);
        for (uint32_t i = 0; i < n_stored; i++) u->hash_offset[i] = ho32[i];
        u->hash_offset[n_stored] = n_hashes;
    }
};

// Gathers the records of `n_stored` stored nodes (ids on the device) into `u` (host, page-locked).
static int32_t gather_and_copy(b200_ctx *c, const ForestDev &f, const uint32_t *d_stored_ids, uint32_t n_stored,
                               uint32_t n_hashes, const uint32_t *d_prefix_by_node, const uint32_t *d_prefix_by_record,
                               const uint64_t *d_seg_offsets, uint64_t n_segs, b200_updates *u, UpdatesOwner *owner) {
    cudaStream_t st = c->stream;
    const UpdatesLayout lay(n_stored, n_hashes);
    if (!(owner->host = pinned_block_alloc(lay.host_total ? lay.host_total : 16))) return fail(c, B200_ERR_OOM, "page-locked result block");
    uint8_t *h = static_cast<uint8_t *>(owner->host);
    lay.bind_host(u, h, n_stored);
    if (n_stored) {
        ENSURE(out_a, lay.dev_total);
        uint8_t *d = static_cast<uint8_t *>(c->out_a.p);
        CU(launch_gather_updates(f, d_stored_ids, n_stored, d_prefix_by_node, d_prefix_by_record, d_seg_offsets, n_segs,
                                 lay.bind_dev(d), st));
        c->launches++;
        CU(cudaMemcpyAsync(h, d, lay.dev_total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    lay.finish_host(u, h, n_stored, n_hashes);
    return B200_OK;
}

// The stored nodes of a finished build in table order (ascending trie, then path): selected in (depth, position) order,
// then one radix sort by (leftmost leaf, depth) = pre-order.  *ids_out points into ctx scratch (upd_ids2).
static int32_t stored_ids_in_table_order(b200_ctx *c, const Built &b, uint32_t *n_stored_out, uint32_t *n_hashes_out,
                                         const uint32_t **ids_out) {
    cudaStream_t st = c->stream;
    const uint32_t B = b.n_nodes;
    *n_stored_out = *n_hashes_out = 0;
    *ids_out = nullptr;
    if (!B) return B200_OK;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    ENSURE(upd_flags, B);
    ENSURE(upd_nh, (size_t)B * 4);
    ENSURE(upd_ids, (size_t)B * 4);
    ENSURE(upd_prefix, (size_t)(B + 1) * 4);
    uint8_t *flags = static_cast<uint8_t *>(c->upd_flags.p);
    uint32_t *nh = static_cast<uint32_t *>(c->upd_nh.p);
    uint32_t *ids = static_cast<uint32_t *>(c->upd_ids.p);
    uint32_t *prefix = static_cast<uint32_t *>(c->upd_prefix.p);
    uint32_t *n_stored_p = small_u32(c) + SM_NSTORED;
    CU(launch_stored_flags(b.f, B, flags, nh, st));
    size_t t_sel = 0, t_scan = 0, t_sort = 0;
    thrust::counting_iterator<uint32_t> counting(0);
    CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, flags, ids, n_stored_p, (int64_t)B, st));
    CU(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, nh, prefix, (int64_t)B, st));
    ENSURE(cub_temp, std::max(t_sel, t_scan));
    CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, flags, ids, n_stored_p, (int64_t)B, st));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t_scan, nh, prefix, (int64_t)B, st));
    c->launches += 3;
    CU(cudaMemcpyAsync(ps + 200, n_stored_p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ps + 201, prefix + (B - 1), 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ps + 202, nh + (B - 1), 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint32_t n_stored = ps[200];
    *n_stored_out = n_stored;
    *n_hashes_out = ps[201] + ps[202];
    if (!n_stored) return B200_OK;
    ENSURE(upd_key, (size_t)n_stored * 8);
    ENSURE(upd_key2, (size_t)n_stored * 8);
    ENSURE(upd_ids2, (size_t)n_stored * 4);
    uint32_t *ids2 = static_cast<uint32_t *>(c->upd_ids2.p);
    uint64_t *key = static_cast<uint64_t *>(c->upd_key.p), *key2 = static_cast<uint64_t *>(c->upd_key2.p);
    CU(launch_table_order_keys(b.f, ids, n_stored, key, st));
    CU(cub::DeviceRadixSort::SortPairs(nullptr, t_sort, key, key2, ids, ids2, (int64_t)n_stored, 0, 40, st));
    ENSURE(cub_temp, t_sort);
    CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, t_sort, key, key2, ids, ids2, (int64_t)n_stored, 0, 40, st));
    c->launches += 2;
    *ids_out = ids2;
    return B200_OK;
}

// Collects the stored BranchNodeCompact records of a finished build into `u` (host, page-locked), in table order; the
// hash offsets are scanned in that order so that hashes follow their records.
static int32_t collect_updates(b200_ctx *c, const Built &b, const uint64_t *d_seg_offsets, uint64_t n_segs,
                               b200_updates *u) {
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    cudaStream_t st = c->stream;
    uint32_t n_stored = 0, n_hashes = 0;
    const uint32_t *ids = nullptr;
    TRY(stored_ids_in_table_order(c, b, &n_stored, &n_hashes, &ids));
    if (n_stored == 0) return gather_and_copy(c, b.f, nullptr, 0, 0, nullptr, nullptr, d_seg_offsets, n_segs, u, owner);
    uint8_t *flags = static_cast<uint8_t *>(c->upd_flags.p);
    uint32_t *nh = static_cast<uint32_t *>(c->upd_nh.p), *prefix = static_cast<uint32_t *>(c->upd_prefix.p);
    size_t t_scan = 0;
    CU(launch_stored_flags_subset(b.f, ids, n_stored, flags, nh, st));  // (all flagged; nh in record order)
    CU(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, nh, prefix, (int64_t)n_stored, st));
    ENSURE(cub_temp, t_scan);
    CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t_scan, nh, prefix, (int64_t)n_stored, st));
    c->launches += 2;
    return gather_and_copy(c, b.f, ids, n_stored, n_hashes, nullptr, prefix, d_seg_offsets, n_segs, u, owner);
}

// The same stored nodes as finished AccountsTrie / StoragesTrie rows (b200_rows, page-locked host memory): sizes, one
// scan, one warp per row writing table bytes, one D2H (tk_outputs.cuh; byte-identical to csrc/table_rows.cu's layout).
struct RowsOwner {  // shared with table_rows.cu (b200_rows_release)
    void *block;
    int pinned;
};
static int32_t collect_rows(b200_ctx *c, const Built &b, const uint64_t *d_seg_offsets, uint64_t n_segs,
                            const uint8_t *d_acct_keys, int32_t fmt, bool storage, b200_rows *rows) {
    memset(rows, 0, sizeof *rows);
    cudaStream_t st = c->stream;
    uint32_t n_stored = 0, n_hashes = 0;
    const uint32_t *ids = nullptr;
    TRY(stored_ids_in_table_order(c, b, &n_stored, &n_hashes, &ids));
    const int packed = fmt == B200_KEYS_PACKED ? 1 : 0;
    const size_t n = n_stored;
    // device: [size (n+1) u64][row_off (n+1) u64][key_len n u32]
    ENSURE(upd_key, (n + 1) * 8);
    ENSURE(upd_key2, (n + 1) * 8);
    ENSURE(upd_nh, (n + 1) * 4);
    uint64_t *size = static_cast<uint64_t *>(c->upd_key.p), *row_off = static_cast<uint64_t *>(c->upd_key2.p);
    uint32_t *key_len = static_cast<uint32_t *>(c->upd_nh.p);
    size_t t_scan = 0;
    CU(launch_row_sizes(b.f, ids, n_stored, packed, storage ? 1 : 0, size, key_len, st));
    CU(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, size, row_off, (int64_t)(n + 1), st));
    ENSURE(cub_temp, t_scan);
    CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t_scan, size, row_off, (int64_t)(n + 1), st));
    c->launches += 2;
    uint64_t *h_total = reinterpret_cast<uint64_t *>(static_cast<uint32_t *>(c->pinned_small) + 204);
    CU(cudaMemcpyAsync(h_total, row_off + n, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint64_t total = *h_total;
    const size_t off_bytes = (n + 1) * sizeof(uint64_t), kl_bytes = ((n * sizeof(uint32_t)) + 7) & ~size_t(7);
    RowsOwner *owner = new RowsOwner{nullptr, 1};
    rows->_owner = owner;
    if (!(owner->block = pinned_block_alloc(off_bytes + kl_bytes + (total ? total : 1)))) return fail(c, B200_ERR_OOM, "page-locked result block");
    uint8_t *h = static_cast<uint8_t *>(owner->block);
    rows->row_offset = reinterpret_cast<uint64_t *>(h);
    rows->key_len = reinterpret_cast<uint32_t *>(h + off_bytes);
    rows->bytes = h + off_bytes + kl_bytes;
    rows->n_rows = n;
    if (n) {
        ENSURE(out_a, total);
        uint8_t *d_bytes = static_cast<uint8_t *>(c->out_a.p);
        CU(launch_encode_rows(b.f, ids, n_stored, packed, storage ? 1 : 0, d_seg_offsets, n_segs, d_acct_keys, row_off, d_bytes,
                              st));
        c->launches++;
        CU(cudaMemcpyAsync(rows->key_len, key_len, n * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(rows->bytes, d_bytes, total, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(rows->row_offset, row_off, off_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return B200_OK;
}

// Same for a subset of nodes given by id (the dirty nodes of an incremental update), in list order.
static int32_t collect_updates_subset(b200_ctx *c, const ForestDev &f, const uint32_t *d_ids, uint32_t count,
                                      b200_updates *u) {
    memset(u, 0, sizeof *u);
    UpdatesOwner *owner = new UpdatesOwner();
    u->_owner = owner;
    cudaStream_t st = c->stream;
    uint32_t n_stored = 0, n_hashes = 0;
    uint32_t *ps = static_cast<uint32_t *>(c->pinned_small);
    if (count) {
        ENSURE(upd_flags, count);
        ENSURE(upd_nh, (size_t)count * 4);
        ENSURE(upd_ids, (size_t)count * 4 * 3);  // selected positions | picked ids | picked prefixes
        ENSURE(upd_prefix, (size_t)(count + 1) * 4);
        uint8_t *flags = static_cast<uint8_t *>(c->upd_flags.p);
        uint32_t *nh = static_cast<uint32_t *>(c->upd_nh.p);
        uint32_t *sel = static_cast<uint32_t *>(c->upd_ids.p), *pick_ids = sel + count, *pick_prefix = sel + 2 * (size_t)count;
        uint32_t *prefix = static_cast<uint32_t *>(c->upd_prefix.p);
        uint32_t *n_stored_p = small_u32(c) + SM_NSTORED;
        CU(launch_stored_flags_subset(f, d_ids, count, flags, nh, st));
        size_t t_sel = 0, t_scan = 0;
        thrust::counting_iterator<uint32_t> counting(0);
        CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, flags, sel, n_stored_p, (int64_t)count, st));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, nh, prefix, (int64_t)count, st));
        ENSURE(cub_temp, std::max(t_sel, t_scan));
        CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, flags, sel, n_stored_p, (int64_t)count, st));
        CU(cub::DeviceScan::ExclusiveSum(c->cub_temp.p, t_scan, nh, prefix, (int64_t)count, st));
        c->launches += 3;
        CU(cudaMemcpyAsync(ps + 200, n_stored_p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 201, prefix + (count - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ps + 202, nh + (count - 1), 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        n_stored = ps[200];
        n_hashes = ps[201] + ps[202];
        CU(launch_pick_subset(d_ids, prefix, sel, n_stored, pick_ids, pick_prefix, st));
        c->launches++;
        return gather_and_copy(c, f, pick_ids, n_stored, n_hashes, nullptr, pick_prefix, nullptr, 0, u, owner);
    }
    return gather_and_copy(c, f, nullptr, 0, 0, nullptr, nullptr, nullptr, 0, u, owner);
}
