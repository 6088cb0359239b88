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
            // (a hash child is a node of this build or, in an items build, the stored hash of an unchanged subtree)
            load32_nc(ci.id < f.n ? f.leaf_ref + 32 * (uint64_t)ci.id : f.node_ref + 32 * (uint64_t)(ci.id - (uint32_t)f.n), ref);
            store32(out.hashes + 32 * (uint64_t)ho, ref);
            ho++;
        }
    }
}

// ------------------------------------------------------------------------------------------------ table rows on the device
// AccountsTrie / StoragesTrie rows of the stored nodes, byte for byte what csrc/table_rows.cu lays out on the host
// (StoredNibbles / StoredNibblesSubKey / packed keys ‖ BranchNodeCompact: 3 × u16 BE masks ‖ hashes), encoded where the
// nodes are: one D2H of finished table bytes instead of records + a host pass.
__device__ __forceinline__ uint32_t row_key_bytes(int packed, int storage, uint32_t d) {
    return packed ? 33u : (storage ? 65u : d);
}

__global__ void row_sizes_kernel(ForestDev f, const uint32_t *__restrict__ ids, uint32_t count, int packed, int storage,
                                 uint64_t *__restrict__ size, uint32_t *__restrict__ key_len) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > count) return;
    if (t == count) {  // the scan's last output is the total
        size[t] = 0;
        return;
    }
    ushort4 mk = f.node_masks[ids[t]];
    uint32_t kb = row_key_bytes(packed, storage, mk.w);
    size[t] = (storage ? 32u : 0u) + kb + 6u + 32u * (uint32_t)__popc((uint32_t)mk.z);
    key_len[t] = storage ? 32u : kb;
}

// One warp per row; lanes stride over the row's bytes.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) encode_rows_kernel(ForestDev f, const uint32_t *__restrict__ ids, uint32_t count,
                                                                 int packed, int storage,
                                                                 const uint64_t *__restrict__ seg_offsets, uint64_t n_segs,
                                                                 const uint8_t *__restrict__ acct_keys,
                                                                 const uint64_t *__restrict__ row_off,
                                                                 uint8_t *__restrict__ out) {
    __shared__ uint32_t hashed_child[WARPS][16];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * WARPS + w;
    if (t >= count) return;
    const uint32_t v = ids[t];
    const ushort4 mk = f.node_masks[v];
    const uint32_t d = mk.w, l = f.node_l[v];
    uint32_t tid = 0;
    if (storage) {  // trie id = segment containing leaf l
        uint64_t lo = 0, hi = n_segs;
        while (lo + 1 < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (seg_offsets[mid] <= l) lo = mid;
            else hi = mid;
        }
        tid = (uint32_t)lo;
    }
    // children under hash_mask, by rank: lane c looks at child c
    const uint32_t j0 = f.node_start[v], k = f.node_start[v + 1] - j0;
    if (lane <= k) {
        ChildInfo ci = fetch_child(f, j0, lane);
        if (mk.z & (1u << ci.nib)) hashed_child[w][__popc((uint32_t)mk.z & ((1u << ci.nib) - 1u))] = ci.id - (uint32_t)f.n;
    }
    __syncwarp();
    const uint8_t *key = f.keys + 32 * (uint64_t)l;
    const uint32_t a_len = storage ? 32u : 0u, kb = row_key_bytes(packed, storage, d);
    const uint32_t m0 = a_len + kb, h0 = m0 + 6u, total = h0 + 32u * (uint32_t)__popc((uint32_t)mk.z);
    uint8_t *row = out + row_off[t];
    for (uint32_t j = lane; j < total; j += 32) {
        uint32_t x;
        if (j < a_len) {
            x = acct_keys[32 * (uint64_t)tid + j];
        } else if (j < m0) {
            uint32_t jj = j - a_len;
            if (jj == kb - 1 && (packed || storage)) {
                x = d;  // trailing length byte of the packed key / of the subkey
            } else if (packed) {
                x = 0;
                if (2 * jj < d) x = key[jj] & 0xF0;
                if (2 * jj + 1 < d) x |= key[jj] & 0x0F;
            } else {
                x = jj < d ? ((jj & 1) ? (key[jj >> 1] & 15u) : (key[jj >> 1] >> 4)) : 0u;
            }
        } else if (j < h0) {
            uint32_t q = j - m0;
            uint32_t m = q < 2 ? mk.x : (q < 4 ? mk.y : mk.z);
            x = (q & 1) ? (m & 0xff) : (m >> 8);
        } else {
            uint32_t q = j - h0;
            x = f.node_ref[32 * (uint64_t)hashed_child[w][q >> 5] + (q & 31)];
        }
        row[j] = (uint8_t)x;
    }
}
