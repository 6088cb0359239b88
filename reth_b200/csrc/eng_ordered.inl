// eng_ordered.inl — b200_ordered_roots: transactions / receipts / withdrawals roots of a batch of lists.
// Part of the single translation unit engine.cu (textually included, in this order).

// ------------------------------------------------------------------------------------------------ ordered roots
// One forest over every list: keys are synthesized on the device (rlp(index) in adjust_index_for_rlp order, zero padded:
// tk_ordered.cuh), the structure and branch passes are the ones every other build uses, the leaf pass streams the items.
static int32_t ordered_roots_on_device(b200_ctx *c, const uint8_t *d_values, uint64_t blob_len, const uint64_t *d_val_off,
                                       const uint64_t *d_seg_offsets, uint64_t n_lists, uint64_t n_items, uint8_t *d_roots) {
    if (n_items >= (1ull << 31)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^31-1 items per call");
    cudaStream_t st = c->stream;
    Built b;
    OrderedLeavesDev o{};
    if (n_items) {
        ENSURE(ord_keys, n_items * 32);
        ENSURE(ord_knib, n_items);
        ENSURE(ord_item, n_items * 4);
        ENSURE(ord_sched, n_items * 2);
        ENSURE(ord_sched2, n_items * 2);
        ENSURE(ord_pos, n_items * 4);
        ENSURE(ord_order, n_items * 4);
        int *err = reinterpret_cast<int *>(small_u32(c) + SM_ERR);
        uint16_t *sched = static_cast<uint16_t *>(c->ord_sched.p), *sched2 = static_cast<uint16_t *>(c->ord_sched2.p);
        uint32_t *pos = static_cast<uint32_t *>(c->ord_pos.p), *order = static_cast<uint32_t *>(c->ord_order.p);
        CU(launch_ordered_keys(d_seg_offsets, n_lists, n_items, d_val_off, static_cast<uint8_t *>(c->ord_keys.p),
                               static_cast<uint8_t *>(c->ord_knib.p), static_cast<uint32_t *>(c->ord_item.p), sched, pos, err,
                               st));
        // leaf visiting order: items of about equal length share a warp (stable: neighbours stay neighbours)
        size_t t_sort = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, t_sort, sched, sched2, pos, order, (int64_t)n_items, 0, 16, st));
        ENSURE(cub_temp, t_sort);
        CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, t_sort, sched, sched2, pos, order, (int64_t)n_items, 0, 16, st));
        c->launches += 2;
        o.key_nibs = static_cast<const uint8_t *>(c->ord_knib.p);
        o.item = static_cast<const uint32_t *>(c->ord_item.p);
        o.order = order;
        o.sched_sorted = sched2;
        o.n_long = small_u32(c) + SM_ORD_NLONG;
        o.values = d_values;
        o.val_off = d_val_off;
        o.blob_len = blob_len;
    }
    TRY(build_forest(c, static_cast<const uint8_t *>(c->ord_keys.p), n_items, d_seg_offsets, n_lists, false, nullptr, nullptr,
                     false, b, &o));
    CU(launch_segment_roots(b.f, d_seg_offsets, n_lists, d_roots, st));
    c->launches++;
    c->stats.leaves_added += n_items;
    c->stats.branches_added += b.n_nodes;
    c->stats.levels += b.levels;
    return B200_OK;
}

extern "C" B200_API int32_t b200_ordered_roots_dev(b200_ctx *c, const void *d_values, uint64_t values_len,
                                                   const void *d_value_offsets, const void *d_seg_offsets, uint64_t n_lists,
                                                   uint64_t n_items, void *d_roots32) {
    if (!c || !d_seg_offsets || (n_lists && !d_roots32) || (n_items && !d_value_offsets) || (values_len && !d_values))
        return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    const bool offsets_aligned =
        !(reinterpret_cast<uintptr_t>(d_value_offsets) & 7) && !(reinterpret_cast<uintptr_t>(d_seg_offsets) & 7);
    if (!aligned16(d_roots32) || !offsets_aligned)
        return fail(c, B200_ERR_INVALID_ARG, "device buffers must be aligned (roots 16, offsets 8)");
    if (n_items && !n_lists) return fail(c, B200_ERR_INVALID_ARG, "items without a list");  // (the key pass reads seg_offsets[1])
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(reset_build_state(c));
    TRY(ordered_roots_on_device(c, static_cast<const uint8_t *>(d_values), values_len,
                                static_cast<const uint64_t *>(d_value_offsets), static_cast<const uint64_t *>(d_seg_offsets),
                                n_lists, n_items, static_cast<uint8_t *>(d_roots32)));
    return finish_build_state(c);
}

extern "C" B200_API int32_t b200_ordered_roots(b200_ctx *c, const uint8_t *values, const uint64_t *value_offsets,
                                               const uint64_t *seg_offsets, uint64_t n_lists, uint8_t *roots32,
                                               b200_stats *opt_stats) {
    if (!c || !seg_offsets || (n_lists && !roots32)) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    TRY(check_offsets_host(c, seg_offsets, n_lists));
    const uint64_t n_items = seg_offsets[n_lists];
    if (n_items && !value_offsets) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
    uint64_t blob_len = 0;
    if (n_items) {
        TRY(check_offsets_host(c, value_offsets, n_items));
        blob_len = value_offsets[n_items];
        if (blob_len && !values) return fail(c, B200_ERR_INVALID_ARG, "bad argument");
        for (uint64_t i = 0; i < n_items; i++)
            if (value_offsets[i + 1] - value_offsets[i] >= (1ull << 31))
                return fail(c, B200_ERR_INVALID_ARG, "item %llu is 2 GiB or larger", (unsigned long long)i);
    }
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    TRY(h2d(c, c->in_a, values, blob_len));
    TRY(h2d(c, c->in_b, value_offsets, n_items ? (n_items + 1) * 8 : 0));
    TRY(h2d(c, c->in_c, seg_offsets, (n_lists + 1) * 8));
    ENSURE(sroots, (n_lists ? n_lists : 1) * 32);
    TRY(reset_build_state(c));
    TRY(ordered_roots_on_device(c, static_cast<const uint8_t *>(c->in_a.p), blob_len, static_cast<const uint64_t *>(c->in_b.p),
                                static_cast<const uint64_t *>(c->in_c.p), n_lists, n_items,
                                static_cast<uint8_t *>(c->sroots.p)));
    TRY(finish_build_state(c));
    if (n_lists) CU(cudaMemcpyAsync(roots32, c->sroots.p, n_lists * 32, cudaMemcpyDeviceToHost, c->stream));
    int32_t r = sync_and_status(c);
    if (opt_stats) *opt_stats = c->stats;
    return r;
}
