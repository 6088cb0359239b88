// tk_outputs.cuh — trie roots and TrieUpdates gathering.
// Part of the single translation unit trie_kernels.cu (included inside namespace b200, in this order: the later
// files use the device functions of the earlier ones).

// ------------------------------------------------------------------------------------------------ roots
// Root of every trie of the forest: the frontier item that starts at the segment's first leaf.
__global__ void segment_roots_kernel(ForestDev f, const uint64_t *__restrict__ seg_offsets, uint64_t n_segs,
                                     uint8_t *__restrict__ roots) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    if (*(volatile int *)f.err != B200_DEVERR_NONE) return;
    uint32_t ref[8];
    uint64_t lo = seg_offsets ? seg_offsets[s] : 0, hi = seg_offsets ? seg_offsets[s + 1] : f.n;
    if (lo == hi) {  // StorageRoot::calculate short circuit, trie.rs:622-629
        ref[0] = 0x171fe856u; ref[1] = 0xa655cc1bu; ref[2] = 0xe64583ffu; ref[3] = 0x6ef8c092u;
        ref[4] = 0x1be0485bu; ref[5] = 0xc0ad6c99u; ref[6] = 0xb52f6201u; ref[7] = 0x21b463e3u;
    } else {
        uint32_t item = f.S[lo];
        const uint8_t *rp =
            item < f.n ? f.leaf_ref + 32 * (uint64_t)item : f.node_ref + 32 * (uint64_t)(item - (uint32_t)f.n);
        load32_nc(rp, ref);
    }
    store32(roots + 32 * s, ref);
}

// ------------------------------------------------------------------------------------------------ updates
// stored[v] flags for DeviceSelect; depth 0 (empty path) is excluded like TrieUpdates::finalize does
// (crates/trie/common/src/updates.rs:147, exclude_empty_from_pair :822-832).
__global__ void stored_flags_kernel(ForestDev f, uint32_t n_nodes, uint8_t *__restrict__ flags,
                                    uint32_t *__restrict__ n_hashes) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    bool st = (f.node_meta[v] & META_STORED) && f.node_masks[v].w != 0;
    flags[v] = st ? 1 : 0;
    n_hashes[v] = st ? __popc((uint32_t)f.node_masks[v].z) : 0;
}

// Same over a list of node ids (the dirty nodes of an incremental update).
__global__ void stored_flags_subset_kernel(ForestDev f, const uint32_t *__restrict__ ids, uint32_t count,
                                           uint8_t *__restrict__ flags, uint32_t *__restrict__ n_hashes) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint32_t v = ids[t];
    bool st = (f.node_meta[v] & META_STORED) && f.node_masks[v].w != 0;
    flags[t] = st ? 1 : 0;
    n_hashes[t] = st ? __popc((uint32_t)f.node_masks[v].z) : 0;
}
// compacts (node id, hash prefix) of the selected positions
__global__ void pick_subset_kernel(const uint32_t *__restrict__ ids, const uint32_t *__restrict__ prefix,
                                   const uint32_t *__restrict__ sel_pos, uint32_t n_sel, uint32_t *__restrict__ out_ids,
                                   uint32_t *__restrict__ out_prefix) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_sel) return;
    uint32_t p = sel_pos[t];
    out_ids[t] = ids[p];
    out_prefix[t] = prefix[p];
}

// Table-order key of a stored node: (leftmost leaf, depth) ascending is pre-order over the forest — ascending trie id,
// then ascending path with a prefix before its extensions — the key order of AccountsTrie / StoragesTrie
// (crates/trie/common/src/nibbles.rs StoredNibbles / StoredNibblesSubKey; MDBX memcmp order).
__global__ void table_order_keys_kernel(ForestDev f, const uint32_t *__restrict__ ids, uint32_t count,
                                        uint64_t *__restrict__ keys) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint32_t v = ids[t];
    keys[t] = ((uint64_t)f.node_l[v] << 8) | (uint64_t)(f.node_masks[v].w & 0xff);
}

// One thread per stored node: path, masks and the child hashes under hash_mask, ascending nibble.
__global__ void gather_updates_kernel(ForestDev f, const uint32_t *__restrict__ stored_ids, uint32_t n_stored,
                                      const uint32_t *__restrict__ hash_prefix /* exclusive, over all nodes */,
                                      const uint32_t *__restrict__ prefix_by_record /* or null */,
                                      const uint64_t *__restrict__ seg_offsets, uint64_t n_segs,
                                      UpdatesDev out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_stored) return;
    uint32_t v = stored_ids[t];
    ushort4 mk = f.node_masks[v];
    uint32_t d = mk.w, l = f.node_l[v];
    // trie id = segment containing leaf l
    uint32_t tid = 0;
    if (seg_offsets) {
        uint64_t lo = 0, hi = n_segs;  // last s with seg_offsets[s] <= l
        while (lo + 1 < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (seg_offsets[mid] <= l) lo = mid;
            else hi = mid;
        }
        tid = (uint32_t)lo;
    }
    out.trie_id[t] = tid;
    out.path_len[t] = (uint8_t)d;
    const uint8_t *key = f.keys + 32 * (uint64_t)l;
    uint8_t *pp = out.path_packed + 32 * (uint64_t)t;
    for (uint32_t b = 0; b < 32; b++) {
        uint32_t x = 0;
        if (2 * b < d) x = key[b] & 0xF0;
        if (2 * b + 1 < d) x |= key[b] & 0x0F;
        pp[b] = (uint8_t)x;
    }
    out.state_mask[t] = mk.x;
    out.tree_mask[t] = mk.y;
    out.hash_mask[t] = mk.z;
    uint32_t ho = prefix_by_record ? prefix_by_record[t] : hash_prefix[v];
    out.hash_offset[t] = ho;
    uint32_t j0 = f.node_start[v], k = f.node_start[v + 1] - j0;
    for (uint32_t c = 0; c <= k; c++) {
        ChildInfo ci = fetch_child(f, j0, c);
        if (mk.z & (1u << ci.nib)) {
            uint32_t ref[8];
            load32_nc(f.node_ref + 32 * (uint64_t)(ci.id - (uint32_t)f.n), ref);
            store32(out.hashes + 32 * (uint64_t)ho, ref);
            ho++;
        }
    }
}
