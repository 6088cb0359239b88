// tk_proofs.cuh — Merkle proofs from the dynamic arenas (tk_dtrie.cuh).
// Part of the single translation unit trie_kernels.cu (included inside namespace b200, in this order: the later
// files use the device functions of the earlier ones).

// ------------------------------------------------------------------------------------------------ proofs
// Merkle proofs from the resident arenas: for a target key, the RLP of every node whose position is a prefix
// of the key, root first — what alloy-trie's ProofRetainer keeps while reth's Proof::account_proof / storage_proof walk
// the trie (crates/trie/trie/src/proof/mod.rs).  An extension node and the branch below it are two proof nodes; the walk
// stops at a leaf (inclusion, or exclusion by a different key), at an empty slot, or inside an extension whose nibbles
// differ from the key.  One thread per target; two passes (sizes, then bytes) around an exclusive scan.
struct CountBuf {  // sizing pass: same interface as LinBuf, nothing is written
    uint32_t n;
    __device__ __forceinline__ void byte(uint32_t) { n++; }
    __device__ __forceinline__ void tail32(const uint32_t (&)[8], uint32_t b0) { n += 32 - b0; }
    __device__ __forceinline__ void words8(const uint32_t (&)[8]) { n += 32; }
};

// keccak256 of `len` bytes at an arbitrarily aligned global address (thread-serial; proofs are not a throughput path)
static __device__ void dt_keccak_global(const uint8_t *p, uint32_t len, uint32_t (&dig)[8]) {
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    uint32_t off = 0;
    for (;;) {
        uint32_t take = len - off < 136 ? len - off : 136;
        for (uint32_t lane = 0; lane < 17; lane++) {
            uint64_t w = 0;
            for (uint32_t b = 0; b < 8; b++) {
                uint32_t i = 8 * lane + b;
                uint32_t x = i < take ? p[off + i] : 0;
                if (take < 136 && i == take) x ^= 0x01;
                if (take < 136 && i == 135) x ^= 0x80;
                w |= (uint64_t)x << (8 * b);
            }
#pragma unroll
            for (int q = 0; q < 17; q++)
                if ((uint32_t)q == lane) a[q] ^= w;
        }
        keccak_f1600(a);
        off += take;
        if (take < 136) break;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        dig[2 * i] = (uint32_t)a[i];
        dig[2 * i + 1] = (uint32_t)(a[i] >> 32);
    }
}

static __device__ __forceinline__ uint32_t dt_branch_rlp_len(const DTrieDev &t, uint32_t v, uint32_t &payload) {
    payload = 1;
    for (int s = 0; s < 16; s++) {
        uint32_t cw = t.nchild[16 * (uint64_t)v + s];
        if (cw == DT_NONE) {
            payload += 1;
        } else {
            uint32_t m = (cw & DT_LEAF) ? t.lmeta[cw & ~DT_LEAF] : t.nmeta[cw];
            payload += (m & META_LEN) ? (m & META_LEN) : 33u;
        }
    }
    return list_header_len(payload) + payload;
}
static __device__ void dt_write_branch_rlp(const DTrieDev &t, uint32_t v, uint32_t payload, uint8_t *dst) {
    LinBuf lb{dst, 0};
    put_list_header(lb, payload);
    for (int s = 0; s < 16; s++) {
        uint32_t cw = t.nchild[16 * (uint64_t)v + s];
        if (cw == DT_NONE) {
            lb.byte(0x80);
            continue;
        }
        bool leaf = (cw & DT_LEAF) != 0;
        uint32_t id = cw & ~DT_LEAF, m = leaf ? t.lmeta[id] : t.nmeta[id];
        uint32_t ref[8];
        load32_nc((leaf ? t.lref : t.nref) + 32 * (uint64_t)id, ref);
        uint32_t il = m & META_LEN;
        if (il == 0) {
            lb.byte(0xa0);
            lb.words8(ref);
        } else {
            for (uint32_t b = 0; b < il; b++) lb.byte(byte_at(ref, b));
        }
    }
    lb.byte(0x80);
}

// Walks target `key` in trie `trie`.  WRITE = false: returns node / byte counts.  WRITE = true: writes the nodes at
// rlp + byte_base and their start offsets at rlp_offset[node_base ..].
template <bool WRITE>
static __device__ void dt_proof_walk(const DTrieDev &t, uint32_t trie, const uint8_t *key, uint32_t &n_nodes, uint64_t &n_bytes,
                                     uint8_t *rlp, uint64_t byte_base, uint64_t *rlp_offset, uint8_t *node_depth, uint32_t *node_masks,
                                     uint64_t node_base) {
    n_nodes = 0;
    n_bytes = 0;
    uint32_t cur = t.troot[trie];
    int pd = -1;
    // depth = number of key nibbles that lead to the node: its path in a ProofNodes / MultiProof map is key[..depth]
    // masks: hash_mask << 16 | tree_mask of a branch node reth would store (BranchNodeMasks), 0 for everything else
    auto begin_node = [&](uint32_t len, int depth, uint32_t masks = 0) {
        if (WRITE) {
            rlp_offset[node_base + n_nodes] = byte_base + n_bytes;
            node_depth[node_base + n_nodes] = (uint8_t)depth;
            node_masks[node_base + n_nodes] = masks;
        }
        n_nodes++;
        n_bytes += len;
    };
    if (cur == DT_NONE) {  // empty trie: the proof is the empty string (EMPTY_STRING_CODE), proof.rs:121-126
        if (WRITE) rlp[byte_base] = 0x80;
        begin_node(1, 0);
        return;
    }
    for (int hops = 0; hops <= DT_MAX_HOPS; hops++) {
        if (cur & DT_LEAF) {
            const uint32_t x = cur & ~DT_LEAF;
            uint32_t k[8];
            load32_nc(t.lkey + 32 * (uint64_t)x, k);
            const uint8_t *val = t.lval + (uint64_t)t.val_stride * x;
            const uint8_t *sr = t.lsroot ? t.lsroot + 32 * (uint64_t)x : nullptr;
            CountBuf cb{0};
            uint32_t len = t.account ? encode_leaf<CountBuf, true>(cb, k, pd, val, sr, t.err) : encode_leaf<CountBuf, false>(cb, k, pd, val, nullptr, t.err);
            if (WRITE) {
                LinBuf lb{rlp + byte_base + n_bytes, 0};
                if (t.account) encode_leaf<LinBuf, true>(lb, k, pd, val, sr, t.err);
                else encode_leaf<LinBuf, false>(lb, k, pd, val, nullptr, t.err);
            }
            begin_node(len, pd + 1);
            return;
        }
        const uint32_t v = cur;
        const int d = t.ndepth[v];
        const uint8_t *nk = t.nkey + 32 * (uint64_t)v;
        uint32_t payload;
        const uint32_t blen = dt_branch_rlp_len(t, v, payload);
        const ushort4 mk = t.nmasks[v];
        const uint32_t masks = ((uint32_t)mk.z << 16) | mk.y;
        const bool ext = pd + 1 < d;
        const bool matches = dt_lcp(key, nk, (uint32_t)(pd + 1), (uint32_t)d) == (uint32_t)d;
        if (ext) {  // the extension node sits at a prefix of the key (we got here); the branch only if its nibbles match
            uint32_t m = (uint32_t)(d - (pd + 1)), hp_len = 1 + (m >> 1), path_str = hp_len == 1 ? 1 : 1 + hp_len;
            uint32_t clen = blen >= 32 ? 33 : blen;
            uint32_t epayload = path_str + clen, elen = list_header_len(epayload) + epayload;
            if (WRITE) {
                // the branch's RLP is needed first (its hash, or itself when shorter than 32 bytes, is the extension's
                // child): written to its final place right after the extension when it belongs to the proof, to a
                // thread-local buffer otherwise
                uint8_t *ext_at = rlp + byte_base + n_bytes;
                uint8_t tmp[544];
                uint8_t *br_at = matches ? ext_at + elen : tmp;
                dt_write_branch_rlp(t, v, payload, br_at);
                uint32_t child[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (blen >= 32) dt_keccak_global(br_at, blen, child);
                else
                    for (uint32_t b = 0; b < blen; b++) child[b >> 2] |= (uint32_t)br_at[b] << (8 * (b & 3));
                LinBuf lb{ext_at, 0};
                encode_extension(lb, nk, (uint32_t)(pd + 1), (uint32_t)d, child, blen >= 32 ? 0u : blen);
            }
            begin_node(elen, pd + 1);
            if (!matches) return;
            begin_node(blen, d, masks);
        } else {
            if (WRITE) dt_write_branch_rlp(t, v, payload, rlp + byte_base + n_bytes);
            begin_node(blen, d, masks);
        }
        pd = d;
        cur = t.nchild[16 * (uint64_t)v + dt_nib(key, (uint32_t)d)];
        if (cur == DT_NONE) return;  // exclusion: the branch has no child for the key's next nibble
    }
    atomicExch(t.err, B200_DEVERR_CORRUPT);
}

// trie_of_target: nullptr = trie 0; DT_NONE entries (storage of an absent account) prove against the empty trie
__global__ void dt_proof_size_kernel(DTrieDev t, const uint32_t *__restrict__ trie_of_target, const uint8_t *__restrict__ keys,
                                     uint64_t n, uint32_t *__restrict__ node_count, uint64_t *__restrict__ byte_count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t trie = trie_of_target ? trie_of_target[i] : 0;
    uint32_t nn;
    uint64_t nb;
    if (trie == DT_NONE) {
        nn = 1;
        nb = 1;
    } else {
        dt_proof_walk<false>(t, trie, keys + 32 * i, nn, nb, nullptr, 0, nullptr, nullptr, nullptr, 0);
    }
    node_count[i] = nn;
    byte_count[i] = nb;
}
__global__ void dt_proof_write_kernel(DTrieDev t, const uint32_t *__restrict__ trie_of_target, const uint8_t *__restrict__ keys,
                                      uint64_t n, const uint64_t *__restrict__ node_base, const uint64_t *__restrict__ byte_base,
                                      uint8_t *__restrict__ rlp, uint64_t *__restrict__ rlp_offset, uint8_t *__restrict__ node_depth,
                                      uint32_t *__restrict__ node_masks) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t trie = trie_of_target ? trie_of_target[i] : 0;
    uint32_t nn;
    uint64_t nb;
    if (trie == DT_NONE) {
        rlp[byte_base[i]] = 0x80;
        rlp_offset[node_base[i]] = byte_base[i];
        node_depth[node_base[i]] = 0;
        node_masks[node_base[i]] = 0;
    } else {
        dt_proof_walk<true>(t, trie, keys + 32 * i, nn, nb, rlp, byte_base[i], rlp_offset, node_depth, node_masks, node_base[i]);
    }
}
// the account leaf (= storage trie id) of one account key, DT_NONE when the account does not exist
__global__ void dt_find_leaf_kernel(DTrieDev t, const uint8_t *__restrict__ key, uint32_t *__restrict__ out, uint64_t n_copies) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_copies) return;
    DtLoc loc = dt_descend(t, t.ltrie ? (uint32_t)(key[0] >> 4) : 0u, key);
    out[i] = loc.found ? (loc.child & ~DT_LEAF) : DT_NONE;
}


// ---- multiproof batch (MultiProofTargets: accounts with their slot targets)
// leaf_out[i] = the account leaf (= storage trie id) of account key i, DT_NONE when the account does not exist; its storage
// root goes to sroot_out (EMPTY_ROOT_HASH for a missing account)
__global__ void dt_find_leaves_kernel(DTrieDev t, const uint8_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ leaf_out,
                                      uint8_t *__restrict__ sroot_out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *key = keys + 32 * i;
    DtLoc loc = dt_descend(t, t.ltrie ? (uint32_t)(key[0] >> 4) : 0u, key);
    uint32_t leaf = loc.found ? (loc.child & ~DT_LEAF) : DT_NONE;
    leaf_out[i] = leaf;
    if (leaf != DT_NONE && t.lsroot) dt_copy32(sroot_out + 32 * i, t.lsroot + 32 * (uint64_t)leaf);
    else dt_put_empty_root(sroot_out + 32 * i);
}
// trie_of_target[j] = leaf of the account whose slot-target segment holds j
__global__ void dt_target_tries_kernel(const uint64_t *__restrict__ seg_offsets, uint64_t n_accounts, const uint32_t *__restrict__ leaf_of,
                                       uint64_t n_targets, uint32_t *__restrict__ trie_of_target) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_targets) return;
    uint64_t lo = 0, hi = n_accounts;  // last account with offset <= j
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (seg_offsets[mid] <= j) lo = mid;
        else hi = mid;
    }
    trie_of_target[j] = leaf_of[lo];
}
