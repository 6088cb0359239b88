// eng_build.inl — one forest build: structure pass, leaf pass, level-by-level branch passes.
// Part of the single translation unit engine.cu (textually included, in this order).

// ------------------------------------------------------------------------------------------------ forest build
struct IsHead {  // sorted gap key (depth | flags << 8, tk_structure.cuh): the gap starts a branch node
    __host__ __device__ uint8_t operator()(uint16_t k) const { return (uint8_t)((k >> 8) & 1u); }
};

struct Built {
    ForestDev f{};
    uint32_t n_nodes = 0;
    uint32_t levels = 0;
    uint32_t level_count[64] = {};  // branch nodes per depth
};

// Builds every trie of a forest over d_keys (n leaves).  d_seg_offsets == nullptr: one trie.
// account: leaves are accounts (d_values = b200_account[n], d_sroots = storage roots or null); else storage
// slots (d_values = U256 BE [n][32]).  ordered != nullptr: leaves of index-keyed tries (eng_ordered.inl); items != nullptr:
// leaves mixed with stored hashes of unchanged subtrees (eng_items.inl).
static int32_t build_forest(b200_ctx *c, const uint8_t *d_keys, uint64_t n, const uint64_t *d_seg_offsets,
                            uint64_t n_segs, bool account, const uint8_t *d_values, const uint8_t *d_sroots,
                            bool retain_updates, Built &out, const OrderedLeavesDev *ordered = nullptr,
                            const ItemLeavesDev *items = nullptr) {
    if (n >= (1ull << 31)) return fail(c, B200_ERR_INVALID_ARG, "at most 2^31-1 leaves per build");
    cudaStream_t st = c->stream;
    ForestDev &f = out.f;
    f.n = n;
    f.keys = d_keys;
    f.err = reinterpret_cast<int *>(small_u32(c) + SM_ERR);
    f.counters = reinterpret_cast<unsigned long long *>(small_u32(c) + SM_COUNTERS);
    f.retain_updates = retain_updates ? 1 : 0;
    out.n_nodes = 0;
    out.levels = 0;
    if (n == 0) return B200_OK;
    c->extra_blocks_valid = !ordered;  // (items of an ordered trie span several rate blocks themselves: not counted)

    ENSURE(Lp, n + 1);
    ENSURE(nibs, n + 1);
    ENSURE(leaf_ref, n * 32);
    ENSURE(leaf_meta, n);
    ENSURE(S, n * 4);
    ENSURE(E, n * 4);
    f.Lp = static_cast<uint8_t *>(c->Lp.p);
    f.nibs = static_cast<uint8_t *>(c->nibs.p);
    f.leaf_ref = static_cast<uint8_t *>(c->leaf_ref.p);
    f.leaf_meta = static_cast<uint8_t *>(c->leaf_meta.p);
    f.S = static_cast<uint32_t *>(c->S.p);
    f.E = static_cast<uint32_t *>(c->E.p);

    CU(cudaMemsetAsync(f.Lp, 0, n + 1, st));
    if (d_seg_offsets) {
        CU(launch_mark_boundaries(d_seg_offsets, n_segs, n, f.Lp, f.err, st));
        c->launches++;
    }
    CU(launch_lcp(d_keys, n, f.Lp, f.nibs, f.err, st));
    c->launches++;
    // The leaf pass (ALU-bound, the longest kernel of a build) and the structure pass (sorts / scans / flags: memory- and
    // latency-bound, one host read-back at its end) both need only Lp: they run side by side, the structure pass on the
    // high-priority aux stream; the branch levels join them again.  A single leaf has no structure pass.
    static const bool overlap = [] {  // B200_OVERLAP_STRUCTURE=0: one stream (measurement aid)
        const char *e = getenv("B200_OVERLAP_STRUCTURE");
        return !e || atoi(e) != 0;
    }();
    cudaStream_t sa = (n >= 2 && overlap) ? c->aux_stream : st;
    if (sa != st) {
        CU(cudaEventRecord(c->ev_fork, st));
        CU(cudaStreamWaitEvent(sa, c->ev_fork, 0));
    }
    if (ordered) CU(launch_ordered_leaves(f, *ordered, st));  // index-keyed tries: variable-length keys and values
    else if (items) CU(launch_item_leaves(f, *items, d_values, d_sroots, st));  // leaves + hashes of unchanged subtrees
    else CU(launch_leaves(f, account, d_values, d_sroots, st));
    c->launches++;
    phase_mark(c, account ? "lcp+leaves(acct)" : "lcp+leaves");
    if (n < 2) return B200_OK;

    // ---- gaps sorted by depth (stable: position order inside a depth) -> branch nodes in CSR form
    const uint64_t G = n - 1;
    ENSURE(iota, G * 4);
    ENSURE(head, G * 2);          // sort keys in position order (depth | flags << 8)
    ENSURE(depth_sorted, G * 2);  // the same, sorted
    ENSURE(gap_sorted, G * 4);
    ENSURE(node_start, (G + 1) * 4);
    uint32_t *bucket_off = small_u32(c) + SM_BUCKET_OFF;
    uint32_t *level_lo = small_u32(c) + SM_LEVEL_LO;
    uint32_t *n_nodes_p = small_u32(c) + SM_NNODES;
    uint32_t *unresolved = small_u32(c) + SM_UNRESOLVED;
    uint16_t *gap_key = static_cast<uint16_t *>(c->head.p);
    uint16_t *key_sorted = static_cast<uint16_t *>(c->depth_sorted.p);
    uint32_t *gap_sorted = static_cast<uint32_t *>(c->gap_sorted.p);
    uint32_t *node_start = static_cast<uint32_t *>(c->node_start.p);

    auto mark = [&](const char *name) {  // phase marks of the structure pass (serial mode only: they record on c->stream)
        if (c->phase_timing && sa == st) phase_mark(c, name);
    };
    // sort input in position order; most gaps learn here whether they start a node (tk_structure.cuh gap_keys_kernel)
    CU(cudaMemsetAsync(unresolved, 0, 4, sa));
    CU(launch_gap_keys(f.Lp, G, gap_key, static_cast<uint32_t *>(c->iota.p), unresolved, sa));
    c->launches++;
    mark("s:gap-keys");
    size_t t_sort = 0, t_sel = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, t_sort, gap_key, key_sorted, static_cast<uint32_t *>(c->iota.p), gap_sorted,
                                       (int64_t)G, 0, 8, sa));
    thrust::counting_iterator<uint32_t> counting(0);
    auto head = thrust::make_transform_iterator(static_cast<const uint16_t *>(key_sorted), IsHead());
    CU(cub::DeviceSelect::Flagged(nullptr, t_sel, counting, head, node_start, n_nodes_p, (int64_t)G, sa));
    size_t t_max = std::max(t_sort, t_sel);
    ENSURE(cub_temp, t_max);
    CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, t_sort, gap_key, key_sorted, static_cast<uint32_t *>(c->iota.p),
                                       gap_sorted, (int64_t)G, 0, 8, sa));
    mark("s:gap-sort");
    CU(launch_bucket_offsets(key_sorted, G, bucket_off, sa));
    CU(launch_head_fix(d_keys, key_sorted, gap_sorted, d_seg_offsets, n_segs, unresolved, G, sa));
    mark("s:offsets+head-fix");
    CU(cub::DeviceSelect::Flagged(c->cub_temp.p, t_sel, counting, head, node_start, n_nodes_p, (int64_t)G, sa));
    CU(launch_level_ranges(node_start, n_nodes_p, bucket_off, level_lo, sa));
    mark("s:select+ranges");
    // (depth, child-count class) of every node + histogram, still without knowing the node count on the host
    ENSURE(node_key, G);
    ENSURE(node_ids, G * 4);
    uint8_t *nk = static_cast<uint8_t *>(c->node_key.p);
    uint32_t *nids = static_cast<uint32_t *>(c->node_ids.p);
    uint32_t *hist = small_u32(c) + SM_HIST;
    CU(cudaMemsetAsync(hist, 0, 256 * 4, sa));
    CU(launch_node_class_keys(node_start, key_sorted, n_nodes_p, G, nk, nids, hist, sa));
    c->launches += 7;
    mark("s:class-keys");
    uint32_t *h_level = static_cast<uint32_t *>(c->pinned_small) + 64;
    uint32_t *h_hist = static_cast<uint32_t *>(c->pinned_small) + 256;
    CU(cudaMemcpyAsync(h_level, level_lo, 66 * 4, cudaMemcpyDeviceToHost, sa));
    CU(cudaMemcpyAsync(h_hist, hist, 256 * 4, cudaMemcpyDeviceToHost, sa));
    CU(cudaStreamSynchronize(sa));  // the only host round trip of a build: 322 integers (the leaf pass keeps running)
    mark("s:readback");
    const uint32_t B = h_level[65];
    out.n_nodes = B;
    f.gap_sorted = gap_sorted;
    f.node_start = node_start;
    if (B == 0) {  // every trie has at most one leaf
        if (sa != st) {
            CU(cudaEventRecord(c->ev_join, sa));
            CU(cudaStreamWaitEvent(st, c->ev_join, 0));
        }
        return B200_OK;
    }

    ENSURE(node_ref, (size_t)B * 32);
    ENSURE(node_meta, B);
    ENSURE(node_l, (size_t)B * 4);
    ENSURE(node_r, (size_t)B * 4);
    ENSURE(node_masks, (size_t)B * 8);
    f.node_ref = static_cast<uint8_t *>(c->node_ref.p);
    f.node_meta = static_cast<uint8_t *>(c->node_meta.p);
    f.node_l = static_cast<uint32_t *>(c->node_l.p);
    f.node_r = static_cast<uint32_t *>(c->node_r.p);
    f.node_masks = static_cast<ushort4 *>(c->node_masks.p);

    // ---- node visiting order: (depth descending, child-count class); ids stay what they are
    ENSURE(node_key2, B);
    ENSURE(node_order, (size_t)B * 4);
    uint8_t *nk2 = static_cast<uint8_t *>(c->node_key2.p);
    uint32_t *norder = static_cast<uint32_t *>(c->node_order.p);
    size_t t_ns = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, t_ns, nk, nk2, nids, norder, (int64_t)B, 0, 8, sa));
    ENSURE(cub_temp, t_ns);
    CU(cub::DeviceRadixSort::SortPairs(c->cub_temp.p, t_ns, nk, nk2, nids, norder, (int64_t)B, 0, 8, sa));
    c->launches += 1;
    if (sa != st) {
        CU(cudaEventRecord(c->ev_join, sa));
        CU(cudaStreamWaitEvent(st, c->ev_join, 0));  // the branch levels need the leaves (st) and the structure (sa)
    }
    phase_mark(c, "leaves||structure");

    // ---- deepest level first; the per-level frontier stays in HBM.  Big levels get one launch per child-count
    // class (strip size and unrolling fit the class), small ones a single launch.
    uint32_t pos = 0;
    for (int d = 63; d >= 0; d--) {
        const uint32_t *hc = h_hist + 4 * (63 - d);
        uint32_t cnt = hc[0] + hc[1] + hc[2] + hc[3];
        if (!cnt) continue;
        out.levels++;
        out.level_count[d] = cnt;
        c->extra_blocks += (uint64_t)hc[1] + 2ull * hc[2] + 3ull * hc[3];
        if (cnt <= WARP_LEVEL_MAX) {  // about one wave of warps: latency-bound, one warp per node
            CU(launch_branch_level(f, norder, pos, pos + cnt, d, -1, st));
            c->launches++;
            pos += cnt;
            phase_mark(c, "small-level");
        } else {
            for (int cls = 0; cls < 4; cls++) {
                if (!hc[cls]) continue;
                // a sparsely populated class of a big level is latency-bound too: one warp per node
                CU(launch_branch_level(f, norder, pos, pos + hc[cls], d, hc[cls] <= WARP_LEVEL_MAX / 4 ? -1 : cls, st));
                c->launches++;
                pos += hc[cls];
                phase_mark(c, hc[cls] <= WARP_LEVEL_MAX / 4 ? "small-class" : (cls == 0 ? "big<=3" : cls == 1 ? "big<=7" : cls == 2 ? "big<=12" : "big<=16"));
            }
        }
    }
    if (pos != B) return fail(c, B200_ERR_CUDA, "internal: level histogram (%u) != node count (%u)", pos, B);
    return B200_OK;
}
