// tk_dtrie.cuh — dynamic resident trie: an arena of branch nodes with 16 child slots each, kept in HBM, that takes
// inserts and deletes in place (SURVEY.md §8 f1 "sparse-trie update path" / row a10).
// Part of the single translation unit trie_kernels.cu (included inside namespace b200, in this order: the later
// files use the device functions of the earlier ones).
//
// What reth does with ParallelSparseTrie (crates/trie/sparse/src/parallel.rs: update_leaf / remove_leaf, then
// update_subtrie_hashes / root) and, on the database path, with TrieWalker + prefix sets (crates/trie/trie/src/walker.rs
// :161-202): touch only the paths of the changed keys.  Here:
//   * a child word is NONE, a leaf id (DT_LEAF set) or a node id; extension nodes stay implicit (a node whose parent
//     is more than one nibble shallower), exactly as in the level-synchronous builder;
//   * every phase is its own launch and is written so that concurrent threads never touch the same word:
//       locate        read-only descent per dirty key
//       update/detach value writes to found leaves; deleted leaves leave their parent's slot (distinct slots)
//       collapse      rounds over the nodes that lost children: a node acts only if its parent is not in the same
//                     round (so the acting nodes are never adjacent); 0 children -> the node leaves its parent,
//                     1 child -> the child moves up (path compression), >= 2 -> the node is just dirty
//       insert        keys that attach at the same (parent, slot) form a run of the sorted dirty list; one thread
//                     inserts a run serially, so different threads work in disjoint subtrees
//       mark / wavefront   dirty leaves and nodes are seeds; pending[] counts dirty children; the last dirty child
//                     to arrive re-hashes the parent (same scheme as tk_wavefront.cuh)
//   * stored-node changes are reported as reth's TrieUpdates: re-hashed nodes with tree|hash mask != 0 are "updated",
//     freed nodes and nodes whose masks became empty are "removed" (crates/trie/common/src/updates.rs:17-26).

// ------------------------------------------------------------------------------------------------ helpers
static __device__ __forceinline__ uint32_t dt_nib(const uint8_t *key, uint32_t i) { return key_nibble_mem(key, i); }

// common prefix of two keys in nibbles, at most `limit`, knowing the first `from` nibbles agree
static __device__ __forceinline__ uint32_t dt_lcp(const uint8_t *a, const uint8_t *b, uint32_t from, uint32_t limit) {
    uint32_t i = from & ~1u;
    for (; i < limit; i += 2) {
        uint32_t x = a[i >> 1] ^ b[i >> 1];
        if (x) {
            uint32_t l = (x & 0xF0) ? i : i + 1;
            return l < limit ? l : limit;
        }
    }
    return limit;
}

static __device__ __forceinline__ void dt_copy32(uint8_t *dst, const uint8_t *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    d[0] = s[0];
    d[1] = s[1];
}
static __device__ __forceinline__ void dt_copy72(uint8_t *dst, const uint8_t *src) {
    const uint64_t *s = reinterpret_cast<const uint64_t *>(src);
    uint64_t *d = reinterpret_cast<uint64_t *>(dst);
#pragma unroll
    for (int w = 0; w < 9; w++) d[w] = s[w];
}

static __device__ __forceinline__ uint32_t dt_trie_of(const DTrieDev &t, uint32_t word) {
    if (!t.ltrie) return 0;
    return (word & DT_LEAF) ? t.ltrie[word & ~DT_LEAF] : t.ntrie[word];
}
static __device__ __forceinline__ void dt_put_empty_root(uint8_t *dst) {  // EMPTY_ROOT_HASH
    uint32_t *w = reinterpret_cast<uint32_t *>(dst);
    w[0] = 0x171fe856u; w[1] = 0xa655cc1bu; w[2] = 0xe64583ffu; w[3] = 0x6ef8c092u;
    w[4] = 0x1be0485bu; w[5] = 0xc0ad6c99u; w[6] = 0xb52f6201u; w[7] = 0x21b463e3u;
}
// parent == NONE addresses the root word of `trie`; a trie that becomes empty gets EMPTY_ROOT_HASH as its root right
// away (no wavefront will visit it)
static __device__ __forceinline__ void dt_set_child(const DTrieDev &t, uint32_t trie, uint32_t parent, uint32_t slot, uint32_t word) {
    if (parent == DT_NONE) {
        t.troot[trie] = word;
        if (word == DT_NONE) dt_put_empty_root(t.top_out + (uint64_t)t.top_stride * trie);
    } else {
        t.nchild[16 * (uint64_t)parent + slot] = word;
    }
}
static __device__ __forceinline__ void dt_copy_val(const DTrieDev &t, uint32_t x, const uint8_t *src) {
    if (t.account) dt_copy72(t.lval + 72 * (uint64_t)x, src);
    else dt_copy32(t.lval + 32 * (uint64_t)x, src);
}
static __device__ __forceinline__ void dt_set_parent(const DTrieDev &t, uint32_t word, uint32_t parent) {
    if (word & DT_LEAF) t.lparent[word & ~DT_LEAF] = parent;
    else t.nparent[word] = parent;
}
// exact test-and-set on a byte-flag array (the flag arrays are 4-byte aligned)
static __device__ __forceinline__ bool dt_test_and_set(uint8_t *base, uint32_t idx) {
    unsigned int *w = reinterpret_cast<unsigned int *>(base + (idx & ~3u));
    unsigned int bit = 1u << (8 * (idx & 3u));
    return (atomicOr(w, bit) & bit) != 0;
}
// a dirty item: must be re-hashed even if nothing below it changed
static __device__ __forceinline__ void dt_seed(const DTrieDev &t, uint32_t word) {
    bool was = (word & DT_LEAF) ? dt_test_and_set(t.lseed, word & ~DT_LEAF) : dt_test_and_set(t.nseed, word);
    if (!was) t.seeds[atomicAdd(&t.g[DG_SEEDS], 1u)] = word;
}

static __device__ __forceinline__ uint32_t dt_pop(uint32_t *count, const uint32_t *stack, uint32_t *bump) {
    for (int tries = 0; tries < 64; tries++) {  // under heavy contention give up on recycling and take a fresh slot
        uint32_t c = *(volatile uint32_t *)count;
        if (c == 0) break;
        if (atomicCAS(count, c, c - 1) == c) return stack[c - 1];
    }
    return atomicAdd(bump, 1u);
}
static __device__ __forceinline__ uint32_t dt_alloc_leaf(const DTrieDev &t) {
    uint32_t id = dt_pop(&t.g[DG_LEAF_FREE], t.leaf_free, &t.g[DG_LEAF_ALLOC]);
    t.lseed[id] = 0;
    return id;
}
static __device__ __forceinline__ uint32_t dt_alloc_node(const DTrieDev &t, uint32_t trie, uint32_t depth, const uint8_t *key,
                                                        uint32_t parent) {
    uint32_t v = dt_pop(&t.g[DG_NODE_FREE], t.node_free, &t.g[DG_NODE_ALLOC]);
    if (t.ntrie) t.ntrie[v] = trie;
    uint4 none = make_uint4(DT_NONE, DT_NONE, DT_NONE, DT_NONE);
    uint4 *ch = reinterpret_cast<uint4 *>(t.nchild + 16 * (uint64_t)v);
    ch[0] = none; ch[1] = none; ch[2] = none; ch[3] = none;
    t.ndepth[v] = (uint8_t)depth;
    t.nparent[v] = parent;
    t.nmeta[v] = 0;
    t.nmasks[v] = make_ushort4(0, 0, 0, (unsigned short)depth);
    dt_copy32(t.nkey + 32 * (uint64_t)v, key);
    t.npending[v] = 0;
    t.nseed[v] = 0;
    t.ncur[v] = 0;
    t.nnext[v] = 0;
    return v;
}
// A stored node that ceases to exist (or to be stored) is one of reth's removed_nodes; its path is read from nkey /
// nmasks.w when the apply gathers its output, so a freed slot is only recycled after that (freed_now list).
static __device__ __forceinline__ void dt_record_removed(const DTrieDev &t, uint32_t v) {
    if (t.nmasks[v].w == 0) return;  // the empty path is never stored (updates.rs:140-158)
    t.removed[atomicAdd(&t.g[DG_REMOVED], 1u)] = v;
}
static __device__ __forceinline__ void dt_free_node(const DTrieDev &t, uint32_t v) {
    if (t.nmeta[v] & META_STORED) dt_record_removed(t, v);
    t.ndepth[v] = DT_DEAD;
    t.nmeta[v] = 0;
    t.freed_now[atomicAdd(&t.g[DG_FREED_NOW], 1u)] = v;
}
__global__ void dt_recycle_kernel(DTrieDev t) {  // end of an apply: this apply's freed nodes become allocatable
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.g[DG_FREED_NOW]) return;
    t.node_free[t.g[DG_NODE_FREE] + i] = t.freed_now[i];
}
__global__ void dt_removed_paths_kernel(DTrieDev t, uint32_t n_removed, uint8_t *__restrict__ path_len, uint8_t *__restrict__ path_packed,
                                        uint32_t *__restrict__ trie_id) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_removed) return;
    uint32_t v = t.removed[i];
    uint32_t d = t.nmasks[v].w;
    const uint8_t *key = t.nkey + 32 * (uint64_t)v;
    uint8_t *pp = path_packed + 32 * (uint64_t)i;
    for (uint32_t b = 0; b < 32; b++) pp[b] = (uint8_t)(2 * b + 1 < d ? key[b] : (2 * b < d ? (key[b] & 0xF0) : 0));
    path_len[i] = (uint8_t)d;
    trie_id[i] = t.ntrie ? t.ntrie[v] : 0;
}

// ------------------------------------------------------------------------------------------------ create
// Conversion of a finished level-synchronous build (ForestDev, one trie) into the arena: node ids and leaf ids carry
// over unchanged.
__global__ void dt_convert_nodes_kernel(ForestDev f, uint32_t n_nodes, const uint32_t *__restrict__ node_parent,
                                        const uint32_t *__restrict__ leaf_trie, DTrieDev t) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    uint32_t j0 = f.node_start[v], k = f.node_start[v + 1] - j0;
    uint32_t *ch = t.nchild + 16 * (uint64_t)v;
    for (int s = 0; s < 16; s++) ch[s] = DT_NONE;
    for (uint32_t c = 0; c <= k; c++) {
        ChildInfo ci = fetch_child(f, j0, c);
        ch[ci.nib] = ci.id < f.n ? (ci.id | DT_LEAF) : ci.id - (uint32_t)f.n;
    }
    ushort4 m = f.node_masks[v];
    t.ndepth[v] = (uint8_t)m.w;
    t.nmasks[v] = m;
    t.nparent[v] = node_parent[v];
    t.nmeta[v] = f.node_meta[v];
    dt_copy32(t.nref + 32 * (uint64_t)v, f.node_ref + 32 * (uint64_t)v);
    dt_copy32(t.nkey + 32 * (uint64_t)v, f.keys + 32 * (uint64_t)f.node_l[v]);
    t.npending[v] = 0;
    t.nseed[v] = 0;
    t.ncur[v] = 0;
    t.nnext[v] = 0;
    uint32_t trie = leaf_trie ? leaf_trie[f.node_l[v]] : 0;
    if (t.ntrie) t.ntrie[v] = trie;
    if (node_parent[v] == DT_NONE) t.troot[trie] = v;
}
// single-leaf tries: the root word is the leaf itself (leaf_parent == NONE)
__global__ void dt_convert_leaves_kernel(uint64_t n, const uint32_t *__restrict__ leaf_parent, const uint32_t *__restrict__ leaf_trie,
                                         DTrieDev t) {
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    uint32_t trie = leaf_trie ? leaf_trie[x] : 0;
    if (t.ltrie) t.ltrie[x] = trie;
    t.lparent[x] = leaf_parent[x];
    if (leaf_parent[x] == DT_NONE) t.troot[trie] = (uint32_t)x | DT_LEAF;
}
// leaf_trie[x] = segment (trie) of leaf x of a forest build
__global__ void dt_leaf_segments_kernel(const uint64_t *__restrict__ seg_offsets, uint64_t n_segs, uint64_t n, uint32_t *__restrict__ leaf_trie) {
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    uint64_t lo = 0, hi = n_segs;  // last segment with offset <= x
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (seg_offsets[mid] <= x) lo = mid;
        else hi = mid;
    }
    leaf_trie[x] = (uint32_t)lo;
}

// ------------------------------------------------------------------------------------------------ locate

struct DtLoc {
    uint32_t parent, slot, child;  // attach point and what hangs there now
    bool found;                    // child is the leaf holding exactly this key
};
static __device__ __forceinline__ DtLoc dt_descend(const DTrieDev &t, uint32_t trie, const uint8_t *key) {
    DtLoc r;
    r.parent = DT_NONE;
    r.slot = 0;
    r.found = false;
    uint32_t cur = t.troot[trie];
    uint32_t matched = 0;  // nibbles known to agree with everything below `cur`
    for (int hops = 0;; hops++) {
        if (hops > DT_MAX_HOPS) {  // depth grows with every hop: more than 64 means a damaged structure, never a long path
            atomicExch(t.err, B200_DEVERR_CORRUPT);
            r.child = DT_NONE;
            return r;
        }
        r.child = cur;
        if (cur == DT_NONE) return r;
        if (cur & DT_LEAF) {
            r.found = dt_lcp(key, t.lkey + 32 * (uint64_t)(cur & ~DT_LEAF), matched, 64) == 64;
            return r;
        }
        uint32_t d = t.ndepth[cur];
        if (dt_lcp(key, t.nkey + 32 * (uint64_t)cur, matched, d) < d) return r;  // diverges inside the edge above `cur`
        r.parent = cur;
        r.slot = dt_nib(key, d);
        matched = d + 1;
        cur = t.nchild[16 * (uint64_t)cur + r.slot];
    }
}

// flags[i] (accounts): bit 0 = present (0 deletes), bit 1 = touch only (the leaf's data is unchanged but it must be
// re-hashed: its storage root changes); nullptr = all present.  Storage slots: a zero value deletes.
// trie_of_key (forest arenas): the trie each key belongs to, DT_NONE = skip the entry.
__global__ void dt_locate_kernel(DTrieDev t, const uint32_t *__restrict__ trie_of_key, const uint8_t *__restrict__ keys,
                                 const uint8_t *__restrict__ vals, const uint8_t *__restrict__ flags, uint64_t m,
                                 uint8_t *__restrict__ kind, uint32_t *__restrict__ leaf_of) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t trie = trie_of_key ? trie_of_key[i] : 0;
    if (trie == DT_NONE) {
        kind[i] = DK_NOOP;
        leaf_of[i] = DT_NONE;
        return;
    }
    if (i && (!trie_of_key || trie_of_key[i - 1] == trie)) {  // strictly ascending inside a trie: the insert runs rely on it
        const uint8_t *a = keys + 32 * (i - 1), *b = keys + 32 * i;
        uint32_t l = dt_lcp(a, b, 0, 64);
        if (l == 64 || dt_nib(a, l) > dt_nib(b, l)) atomicExch(t.err, B200_DEVERR_UNSORTED);
    }
    DtLoc loc = dt_descend(t, trie, keys + 32 * i);
    bool want, touch = false;
    if (t.account) {
        want = flags == nullptr || (flags[i] & 1);
        touch = flags != nullptr && (flags[i] & 2);
    } else {
        const uint64_t *v = reinterpret_cast<const uint64_t *>(vals + 32 * i);
        want = (v[0] | v[1] | v[2] | v[3]) != 0;
    }
    uint8_t k;
    if (touch) k = (want && loc.found) ? DK_TOUCH : DK_NOOP;
    else k = loc.found ? (want ? DK_UPDATE : DK_DELETE) : (want ? DK_INSERT : DK_NOOP);
    kind[i] = k;
    leaf_of[i] = loc.found ? (loc.child & ~DT_LEAF) : DT_NONE;
}

// value changes of existing leaves; deleted leaves leave their parent's slot
__global__ void dt_update_detach_kernel(DTrieDev t, const uint8_t *__restrict__ vals, const uint8_t *__restrict__ sroots,
                                        uint64_t m, const uint8_t *__restrict__ kind, const uint32_t *__restrict__ leaf_of,
                                        uint32_t *__restrict__ touched) {
    if (*(volatile int *)t.err != B200_DEVERR_NONE) return;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t x = leaf_of[i];
    if (kind[i] == DK_UPDATE) {
        dt_copy_val(t, x, vals + (uint64_t)t.val_stride * i);
        if (sroots && t.lsroot) dt_copy32(t.lsroot + 32 * (uint64_t)x, sroots + 32 * i);
        dt_seed(t, x | DT_LEAF);
    } else if (kind[i] == DK_TOUCH) {
        dt_seed(t, x | DT_LEAF);
    } else if (kind[i] == DK_DELETE) {
        uint32_t p = t.lparent[x];
        if (p == DT_NONE) {
            dt_set_child(t, dt_trie_of(t, x | DT_LEAF), DT_NONE, 0, DT_NONE);
        } else {
            t.nchild[16 * (uint64_t)p + dt_nib(t.lkey + 32 * (uint64_t)x, t.ndepth[p])] = DT_NONE;
            if (!dt_test_and_set(t.nnext, p)) touched[atomicAdd(&t.g[DG_LIST_A], 1u)] = p;
        }
        t.lmeta[x] = DT_DEAD;
        t.lseed[x] = 0;
        t.leaf_free[atomicAdd(&t.g[DG_LEAF_FREE], 1u)] = x;
        atomicSub(&t.g[DG_NLEAVES], 1u);
    }
}

// ------------------------------------------------------------------------------------------------ collapse rounds
// ncur = member of the round being processed, nnext = already listed for the following round
__global__ void dt_round_begin_kernel(DTrieDev t, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count_p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t v = list[i];
    t.ncur[v] = 1;
    t.nnext[v] = 0;
}
// snapshot of "my parent is a member of this round", taken before anybody acts: the decision must not depend on
// parent links that an acting node re-writes during the round
__global__ void dt_round_defer_kernel(DTrieDev t, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count_p,
                                      uint8_t *__restrict__ defer) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t v = list[i];
    uint32_t gp = t.ndepth[v] == DT_DEAD ? DT_NONE : t.nparent[v];
    defer[i] = (gp != DT_NONE && t.ncur[gp]) ? 1 : 0;
}
__global__ void dt_round_end_kernel(DTrieDev t, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count_p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    t.ncur[list[i]] = 0;
}

// One round over the nodes that lost children.  ncur marks the members of this round; a node whose parent was a
// member when the round began waits for the next round, so two acting nodes are never parent and child and every
// word has one writer.
__global__ void dt_collapse_round_kernel(DTrieDev t, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count_p,
                                         const uint8_t *__restrict__ defer, uint32_t *__restrict__ next, uint32_t *next_count) {
    if (*(volatile int *)t.err != B200_DEVERR_NONE) return;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t v = list[i];
    if (t.ndepth[v] == DT_DEAD) return;
    uint32_t gp = t.nparent[v];
    auto push_next = [&](uint32_t node) {
        if (!dt_test_and_set(t.nnext, node)) next[atomicAdd(next_count, 1u)] = node;
    };
    if (defer[i]) {
        push_next(v);
        return;
    }
    const uint32_t *ch = t.nchild + 16 * (uint64_t)v;
    uint32_t cnt = 0, only = DT_NONE;
    for (int s = 0; s < 16; s++)
        if (ch[s] != DT_NONE) {
            cnt++;
            only = ch[s];
        }
    if (cnt >= 2) {
        dt_seed(t, v);
        return;
    }
    uint32_t slot = gp == DT_NONE ? 0 : dt_nib(t.nkey + 32 * (uint64_t)v, t.ndepth[gp]);
    const uint32_t trie = dt_trie_of(t, v);
    if (cnt == 1) {  // path compression: the only child takes this node's place
        dt_set_child(t, trie, gp, slot, only);
        dt_set_parent(t, only, gp);
        dt_seed(t, only);  // its parent depth changed: the leaf path / extension above it is different now
    } else {
        dt_set_child(t, trie, gp, slot, DT_NONE);
        if (gp != DT_NONE) push_next(gp);
    }
    dt_free_node(t, v);
}

// ------------------------------------------------------------------------------------------------ insert
// attach[j] identifies where insert key j hangs in the structure left by the deletes: (parent << 4 | slot), or, at a
// trie's root word, (1 << 63 | trie).  Keys with equal attach words are consecutive (same trie, same leading prefix).
__global__ void dt_insert_locate_kernel(DTrieDev t, const uint32_t *__restrict__ trie_of_key, const uint8_t *__restrict__ keys,
                                        const uint32_t *__restrict__ ins_idx, const uint32_t *__restrict__ n_ins_p,
                                        uint64_t *__restrict__ attach) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *n_ins_p) return;
    const uint32_t i = ins_idx[j];
    const uint32_t trie = trie_of_key ? trie_of_key[i] : 0;
    DtLoc loc = dt_descend(t, trie, keys + 32 * (uint64_t)i);
    attach[j] = loc.parent == DT_NONE ? ((1ull << 63) | trie) : (((uint64_t)loc.parent << 4) | loc.slot);
}

// returns the id of the new leaf
static __device__ __forceinline__ uint32_t dt_insert_one(const DTrieDev &t, uint32_t trie, const uint8_t *key, const uint8_t *val,
                                                         const uint8_t *sroot, uint32_t parent, uint32_t slot) {
    uint32_t matched = parent == DT_NONE ? 0 : (uint32_t)t.ndepth[parent] + 1;
    uint32_t cur = parent == DT_NONE ? t.troot[trie] : t.nchild[16 * (uint64_t)parent + slot];
    // the new leaf
    uint32_t x = dt_alloc_leaf(t);
    dt_copy32(t.lkey + 32 * (uint64_t)x, key);
    dt_copy_val(t, x, val);
    if (t.ltrie) t.ltrie[x] = trie;
    if (t.lsroot) {
        if (sroot) dt_copy32(t.lsroot + 32 * (uint64_t)x, sroot);
        else dt_put_empty_root(t.lsroot + 32 * (uint64_t)x);
    }
    t.lmeta[x] = 0;
    atomicAdd(&t.g[DG_NLEAVES], 1u);
    for (int hops = 0;; hops++) {
        if (hops > DT_MAX_HOPS) {
            atomicExch(t.err, B200_DEVERR_CORRUPT);
            break;
        }
        if (cur == DT_NONE) {
            dt_set_child(t, trie, parent, slot, x | DT_LEAF);
            t.lparent[x] = parent;
            break;
        }
        const uint8_t *other = (cur & DT_LEAF) ? t.lkey + 32 * (uint64_t)(cur & ~DT_LEAF) : t.nkey + 32 * (uint64_t)cur;
        uint32_t limit = (cur & DT_LEAF) ? 64u : (uint32_t)t.ndepth[cur];
        uint32_t l = dt_lcp(key, other, matched, limit);
        if (l < limit) {  // diverges above `cur`: a new branch at depth l holds both
            uint32_t b = dt_alloc_node(t, trie, l, key, parent);
            t.nchild[16 * (uint64_t)b + dt_nib(key, l)] = x | DT_LEAF;
            t.nchild[16 * (uint64_t)b + dt_nib(other, l)] = cur;
            t.lparent[x] = b;
            dt_set_parent(t, cur, b);
            dt_set_child(t, trie, parent, slot, b);
            dt_seed(t, b);
            dt_seed(t, cur);  // parent depth changed
            break;
        }
        // (a leaf can never match all 64 nibbles here: the key is new)
        parent = cur;
        slot = dt_nib(key, limit);
        matched = limit + 1;
        cur = t.nchild[16 * (uint64_t)cur + slot];
    }
    dt_seed(t, x | DT_LEAF);
    return x;
}

__global__ void dt_insert_runs_kernel(DTrieDev t, const uint32_t *__restrict__ trie_of_key, const uint8_t *__restrict__ keys,
                                      const uint8_t *__restrict__ vals, const uint8_t *__restrict__ sroots,
                                      const uint32_t *__restrict__ ins_idx, const uint32_t *__restrict__ n_ins_p,
                                      const uint64_t *__restrict__ attach, uint32_t *__restrict__ leaf_of) {
    if (*(volatile int *)t.err != B200_DEVERR_NONE) return;
    const uint32_t n_ins = *n_ins_p;
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_ins) return;
    uint64_t a = attach[j];
    if (j && attach[j - 1] == a) return;  // not the head of its run
    const bool at_root = (a >> 63) != 0;
    uint32_t parent = at_root ? DT_NONE : (uint32_t)(a >> 4), slot = at_root ? 0u : (uint32_t)(a & 15);
    for (uint32_t q = j; q < n_ins && attach[q] == a; q++) {
        uint64_t i = ins_idx[q];
        uint32_t trie = trie_of_key ? trie_of_key[i] : 0;
        leaf_of[i] = dt_insert_one(t, trie, keys + 32 * i, vals + (uint64_t)t.val_stride * i, sroots ? sroots + 32 * i : nullptr,
                                   parent, slot);
    }
}

// ------------------------------------------------------------------------------------------------ mark + wavefront
static __device__ __forceinline__ bool dt_alive(const DTrieDev &t, uint32_t word) {
    return (word & DT_LEAF) ? t.lmeta[word & ~DT_LEAF] != DT_DEAD : t.ndepth[word] != DT_DEAD;
}
static __device__ __forceinline__ uint32_t dt_parent_of(const DTrieDev &t, uint32_t word) {
    return (word & DT_LEAF) ? t.lparent[word & ~DT_LEAF] : t.nparent[word];
}

// pending[p] = number of dirty children of p.  A walk stops at the first ancestor somebody already reached, and at a
// seed (which walks on its own behalf).
__global__ void dt_mark_kernel(DTrieDev t, const uint32_t *__restrict__ count_p) {
    if (*(volatile int *)t.err != B200_DEVERR_NONE) return;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t s = t.seeds[i];
    if (!dt_alive(t, s)) return;
    uint32_t p = dt_parent_of(t, s);
    for (int hops = 0; p != DT_NONE; hops++) {
        if (hops > DT_MAX_HOPS) {
            atomicExch(t.err, B200_DEVERR_CORRUPT);
            break;
        }
        if (atomicAdd(&t.npending[p], 1u) != 0u) break;
        if (t.nseed[p]) break;
        p = t.nparent[p];
    }
}

// After marking: clears the seed flags and keeps only the wavefront's starting points in the list — live leaves, and
// live nodes without a dirty child (a seed node that has dirty children is re-hashed by the last of them to arrive).
__global__ void dt_starts_kernel(DTrieDev t, const uint32_t *__restrict__ count_p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t s = t.seeds[i];
    if (s & DT_LEAF) {
        t.lseed[s & ~DT_LEAF] = 0;
        if (t.lmeta[s & ~DT_LEAF] == DT_DEAD) t.seeds[i] = DT_NONE;
    } else {
        t.nseed[s] = 0;
        if (t.ndepth[s] == DT_DEAD || t.npending[s] != 0u) t.seeds[i] = DT_NONE;
    }
}

// One warp builds node v from its 16 child slots (lane = nibble).  All 32 lanes must call.
// PD_OVERRIDE >= -1: encode as if the parent were at that depth and leave the arena untouched (multi-GPU frontier:
// a bucket's top node as child of the depth-0 root branch); returns the RlpNode meta (inline length | META_EXT ...).
template <int PD_OVERRIDE = -2>
__device__ __forceinline__ uint32_t dt_warp_build_node(const DTrieDev &t, uint32_t v, uint8_t *buf, const WarpKeccak &kw, int lane,
                                                       uint32_t &hashed, uint32_t &exts, uint32_t (&out)[8]) {
    uint32_t *bufw = reinterpret_cast<uint32_t *>(buf);
    const int d = t.ndepth[v];
    const uint32_t cw = lane < 16 ? t.nchild[16 * (uint64_t)v + lane] : DT_NONE;
    const bool has = cw != DT_NONE;
    const bool is_leaf = has && (cw & DT_LEAF);
    const uint32_t id = cw & ~DT_LEAF;
    uint32_t ref[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t cmeta = 0;
    if (has) {
        cmeta = __ldcg(is_leaf ? t.lmeta + id : t.nmeta + id);
        const uint4 *q = reinterpret_cast<const uint4 *>((is_leaf ? t.lref : t.nref) + 32 * (uint64_t)id);
        uint4 x = __ldcg(q), y = __ldcg(q + 1);
        ref[0] = x.x; ref[1] = x.y; ref[2] = x.z; ref[3] = x.w;
        ref[4] = y.x; ref[5] = y.y; ref[6] = y.z; ref[7] = y.w;
    }
    const uint32_t clen = lane < 16 ? (has ? ((cmeta & META_LEN) ? (cmeta & META_LEN) : 33u) : 1u) : 0u;
    const uint32_t bit = has ? (1u << lane) : 0u;
    const bool is_branch = has && !is_leaf;
    const uint32_t hbit = (is_branch && !(cmeta & META_EXT)) ? bit : 0u;
    const uint32_t tbit = (is_branch && (cmeta & META_STORED)) ? bit : 0u;
    if (hbit && (cmeta & META_LEN)) atomicExch(t.err, B200_DEVERR_INLINE_HASH_CHILD);
    const uint32_t state_mask = __reduce_or_sync(0xffffffffu, bit);
    const uint32_t hash_mask = __reduce_or_sync(0xffffffffu, hbit);
    const uint32_t tree_mask = __reduce_or_sync(0xffffffffu, tbit);
    uint32_t incl = clen;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const uint32_t payload = __shfl_sync(0xffffffffu, incl, 15) + 1;  // + the empty value slot
    const uint32_t hdr = list_header_len(payload), total = hdr + payload;
    const uint32_t blocks = total / 136 + 1;
    for (uint32_t w = lane; w < blocks * 34; w += 32) bufw[w] = 0;
    __syncwarp();
    if (lane == 0) {
        LinBuf lb{buf, 0};
        put_list_header(lb, payload);
    }
    if (lane < 16) {
        uint32_t off = hdr + incl - clen;
        if (!has) {
            buf[off] = 0x80;
        } else if ((cmeta & META_LEN) == 0) {
            buf[off++] = 0xa0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                buf[off++] = (uint8_t)ref[i];
                buf[off++] = (uint8_t)(ref[i] >> 8);
                buf[off++] = (uint8_t)(ref[i] >> 16);
                buf[off++] = (uint8_t)(ref[i] >> 24);
            }
        } else {
            for (uint32_t b = 0; b < clen; b++) buf[off++] = (uint8_t)byte_at(ref, b);
        }
    }
    if (lane == 16) {
        buf[total - 1] = 0x80;
        buf[total] |= 0x01;
        buf[blocks * 136 - 1] |= 0x80;
    }
    __syncwarp();
    const uint32_t par = t.nparent[v];
    const int pd = PD_OVERRIDE >= -1 ? PD_OVERRIDE : (par == DT_NONE ? -1 : (int)t.ndepth[par]);
    uint32_t meta = warp_finish_node(buf, total, blocks, d, pd, t.nkey + 32 * (uint64_t)v, kw, lane, hashed, exts, out);
    if (PD_OVERRIDE >= -1) {
        __syncwarp();
        return meta;
    }
    if (lane == 0) {
        if ((tree_mask | hash_mask) != 0) meta |= META_STORED;
        if ((t.nmeta[v] & META_STORED) && !(meta & META_STORED)) dt_record_removed(t, v);
        store32(t.nref + 32 * (uint64_t)v, out);
        t.nmeta[v] = (uint8_t)meta;
        t.nmasks[v] = make_ushort4((unsigned short)state_mask, (unsigned short)tree_mask, (unsigned short)hash_mask,
                                   (unsigned short)d);
        t.built[atomicAdd(&t.g[DG_BUILT], 1u)] = v;
    }
    __syncwarp();
    return meta;
}

// One warp per seed: re-hash the item if nothing below it is dirty, then climb; the last dirty child to arrive at a
// node re-hashes it.  The warp that runs out of parents holds the new root reference.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) dt_wavefront_kernel(DTrieDev t, const uint32_t *__restrict__ count_p) {
    __shared__ __align__(16) uint8_t sbuf[WARPS][WARP_BUF];
    if (*(volatile int *)t.err != B200_DEVERR_NONE) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *buf = sbuf[warp];
    uint32_t *bufw = reinterpret_cast<uint32_t *>(buf);
    WarpKeccak kw;
    kw.init(lane);
    uint32_t hashed = 0, exts = 0;
    const uint32_t count = *count_p;
    for (uint32_t e = blockIdx.x * WARPS + warp; e < count; e += gridDim.x * WARPS) {
        const uint32_t s = t.seeds[e];
        if (s == DT_NONE) continue;  // not a starting point (dt_starts_kernel)
        uint32_t out[8];
        uint32_t p;
        if (s & DT_LEAF) {
            const uint32_t x = s & ~DT_LEAF;
            p = t.lparent[x];
            const int pd = p == DT_NONE ? -1 : (int)t.ndepth[p];
            for (uint32_t w = lane; w < 68; w += 32) bufw[w] = 0;
            __syncwarp();
            uint32_t len = 0;
            if (lane == 0) {
                uint32_t k[8];
                load32_nc(t.lkey + 32 * (uint64_t)x, k);
                LinBuf lb{buf, 0};
                if (t.account)
                    len = encode_leaf<LinBuf, true>(lb, k, pd, t.lval + 72 * (uint64_t)x,
                                                    t.lsroot ? t.lsroot + 32 * (uint64_t)x : nullptr, t.err);
                else
                    len = encode_leaf<LinBuf, false>(lb, k, pd, t.lval + 32 * (uint64_t)x, nullptr, t.err);
            }
            len = __shfl_sync(0xffffffffu, len, 0);
            uint32_t lmeta;
            if (len >= 32 || pd < 0) {  // account leaves are >= 70 bytes; a short storage leaf is hashed only as a whole trie
                if (lane == 0) {
                    buf[len] |= 0x01;
                    buf[(len / 136 + 1) * 136 - 1] |= 0x80;
                }
                __syncwarp();
                uint64_t a = kw.hash(buf, len / 136 + 1, lane);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint64_t w = shfl64(a, q);
                    out[2 * q] = (uint32_t)w;
                    out[2 * q + 1] = (uint32_t)(w >> 32);
                }
                hashed += lane == 0;
                lmeta = 0;
            } else {
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; q++) out[q] = bufw[q];
                lmeta = len;
            }
            if (lane == 0) {
                store32(t.lref + 32 * (uint64_t)x, out);
                t.lmeta[x] = (uint8_t)lmeta;
            }
            __syncwarp();
        } else {
            dt_warp_build_node(t, s, buf, kw, lane, hashed, exts, out);
            p = t.nparent[s];
        }
        bool top = true;
        for (int hops = 0; p != DT_NONE; hops++) {
            if (hops > DT_MAX_HOPS) {  // uniform across the warp
                if (lane == 0) atomicExch(t.err, B200_DEVERR_CORRUPT);
                top = false;
                break;
            }
            uint32_t last = 0;
            if (lane == 0) {
                __threadfence();
                last = atomicSub(&t.npending[p], 1u) == 1u;
                __threadfence();
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (!last) {
                top = false;
                break;
            }
            dt_warp_build_node(t, p, buf, kw, lane, hashed, exts, out);
            p = t.nparent[p];
        }
        if (top && lane == 0) store32(t.top_out + (uint64_t)t.top_stride * dt_trie_of(t, s), out);  // the trie's new root
    }
    if (lane == 0) {
        if (hashed) atomicAdd(&t.counters[CNT_HASHED], (unsigned long long)hashed);
        if (exts) atomicAdd(&t.counters[CNT_EXT], (unsigned long long)exts);
    }
}

// ------------------------------------------------------------------------------------------------ TrieUpdates
// flags[i] = 1 iff re-hashed node built[i] is stored (tree|hash mask != 0, path not empty); n_hashes[i] its hash count
__global__ void dt_stored_flags_kernel(DTrieDev t, const uint32_t *__restrict__ count_p, uint8_t *__restrict__ flags,
                                       uint32_t *__restrict__ n_hashes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t v = t.built[i];
    bool stored = t.ndepth[v] != DT_DEAD && t.ndepth[v] != 0 && (t.nmeta[v] & META_STORED);
    flags[i] = stored ? 1 : 0;
    n_hashes[i] = stored ? (uint32_t)__popc(t.nmasks[v].z) : 0u;
}
__global__ void dt_gather_updates_kernel(DTrieDev t, const uint32_t *__restrict__ stored_ids, uint32_t n_stored,
                                         const uint32_t *__restrict__ hash_prefix_by_record, UpdatesDev out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_stored) return;
    uint32_t v = stored_ids[i];
    ushort4 m = t.nmasks[v];
    uint32_t d = m.w;
    out.trie_id[i] = t.ntrie ? t.ntrie[v] : 0;
    out.path_len[i] = (uint8_t)d;
    const uint8_t *key = t.nkey + 32 * (uint64_t)v;
    uint8_t *pp = out.path_packed + 32 * (uint64_t)i;
    for (uint32_t b = 0; b < 32; b++) pp[b] = (uint8_t)(2 * b + 1 < d ? key[b] : (2 * b < d ? (key[b] & 0xF0) : 0));
    out.state_mask[i] = m.x;
    out.tree_mask[i] = m.y;
    out.hash_mask[i] = m.z;
    uint32_t h = hash_prefix_by_record[i];
    out.hash_offset[i] = h;
    const uint32_t *ch = t.nchild + 16 * (uint64_t)v;
    for (int s = 0; s < 16; s++)
        if ((m.z >> s) & 1) {
            dt_copy32(out.hashes + 32 * (uint64_t)h, t.nref + 32 * (uint64_t)ch[s]);
            h++;
        }
}

// ------------------------------------------------------------------------------------------------ sharded accounts
// Multi-GPU layout of §6 for the dynamic state: the account arena holds one trie per top-nibble bucket (trie id = nibble,
// every bucket a trie of its own, so its root hash is the frontier's as_root); as_child re-encodes the bucket's top item
// as a child of the depth-0 root branch.  One warp per bucket.
__global__ void dt_nibble_tries_kernel(const uint8_t *__restrict__ keys, uint64_t m, uint32_t *__restrict__ trie_of_key) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) trie_of_key[i] = keys[32 * i] >> 4;
}
__global__ void __launch_bounds__(512) dt_frontier_kernel(DTrieDev t, const uint8_t *__restrict__ bucket_roots,
                                                          FrontierEntryDev *__restrict__ out) {
    __shared__ __align__(16) uint8_t sbuf[16][WARP_BUF];
    const int lane = threadIdx.x & 31, b = threadIdx.x >> 5;
    uint8_t *buf = sbuf[b];
    uint32_t *bufw = reinterpret_cast<uint32_t *>(buf);
    WarpKeccak kw;
    kw.init(lane);
    FrontierEntryDev &e = out[b];
    for (int i = lane; i < (int)sizeof(FrontierEntryDev); i += 32) reinterpret_cast<uint8_t *>(&e)[i] = 0;
    __syncwarp();
    const uint32_t w = t.troot[b];
    if (w == DT_NONE || *(volatile int *)t.err != B200_DEVERR_NONE) return;
    uint32_t out8[8], hashed = 0, exts = 0, meta;
    if (w & DT_LEAF) {
        const uint32_t x = w & ~DT_LEAF;
        for (uint32_t q = lane; q < 68; q += 32) bufw[q] = 0;
        __syncwarp();
        uint32_t len = 0;
        if (lane == 0) {
            uint32_t k[8];
            load32_nc(t.lkey + 32 * (uint64_t)x, k);
            LinBuf lb{buf, 0};
            len = encode_leaf<LinBuf, true>(lb, k, 0, t.lval + 72 * (uint64_t)x, t.lsroot ? t.lsroot + 32 * (uint64_t)x : nullptr, t.err);
            buf[len] |= 0x01;
            buf[(len / 136 + 1) * 136 - 1] |= 0x80;
        }
        len = __shfl_sync(0xffffffffu, len, 0);
        __syncwarp();
        uint64_t a = kw.hash(buf, len / 136 + 1, lane);  // account leaves are >= 70 bytes: always a hash reference
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint64_t v = shfl64(a, q);
            out8[2 * q] = (uint32_t)v;
            out8[2 * q + 1] = (uint32_t)(v >> 32);
        }
        meta = 0;
    } else {
        meta = dt_warp_build_node<0>(t, w, buf, kw, lane, hashed, exts, out8);
    }
    if (lane == 0) {
        e.as_root_len = 32;
        for (int i = 0; i < 32; i++) e.as_root[i] = bucket_roots[32 * b + i];
        uint32_t il = meta & META_LEN;
        if (il == 0) {
            e.as_child_len = 33;
            e.as_child[0] = 0xa0;
            for (int i = 0; i < 32; i++) e.as_child[1 + i] = (uint8_t)(out8[i >> 2] >> (8 * (i & 3)));
        } else {
            e.as_child_len = (uint8_t)il;
            for (uint32_t i = 0; i < il; i++) e.as_child[i] = (uint8_t)(out8[i >> 2] >> (8 * (i & 3)));
        }
    }
}

// ------------------------------------------------------------------------------------------------ dynamic state glue
// Which storage tries a block wipes: the tries of destroyed accounts and of accounts flagged "storage wiped"
// (HashedStorage::wiped, crates/trie/common/src/hashed_state.rs:423-428).  Trie id = id of the account's leaf.
__global__ void dt_wipe_list_kernel(const uint8_t *__restrict__ kind, const uint8_t *__restrict__ flags,
                                    const uint32_t *__restrict__ leaf_of, uint64_t m, uint32_t *__restrict__ tries,
                                    uint32_t *__restrict__ count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint8_t k = kind[i];
    bool wiped = flags != nullptr && (flags[i] & 4);
    if (k == DK_DELETE || (wiped && (k == DK_UPDATE || k == DK_TOUCH))) tries[atomicAdd(count, 1u)] = leaf_of[i];
}
static __device__ __forceinline__ void dt_wipe_leaf(const DTrieDev &t, uint32_t x) {
    t.lmeta[x] = DT_DEAD;
    t.lseed[x] = 0;
    t.leaf_free[atomicAdd(&t.g[DG_LEAF_FREE], 1u)] = x;
    atomicSub(&t.g[DG_NLEAVES], 1u);
}
// breadth-first release of whole tries: no removed-node records (reth reports a wiped storage trie as is_deleted)
__global__ void dt_wipe_begin_kernel(DTrieDev t, const uint32_t *__restrict__ tries, const uint32_t *__restrict__ count_p,
                                     uint32_t *__restrict__ next, uint32_t *next_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t r = tries[i], w = t.troot[r];
    if (w == DT_NONE) return;
    dt_set_child(t, r, DT_NONE, 0, DT_NONE);
    if (w & DT_LEAF) dt_wipe_leaf(t, w & ~DT_LEAF);
    else next[atomicAdd(next_count, 1u)] = w;
}
__global__ void dt_wipe_round_kernel(DTrieDev t, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count_p,
                                     uint32_t *__restrict__ next, uint32_t *next_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_p) return;
    uint32_t v = list[i];
    const uint32_t *ch = t.nchild + 16 * (uint64_t)v;
    for (int s = 0; s < 16; s++) {
        uint32_t w = ch[s];
        if (w == DT_NONE) continue;
        if (w & DT_LEAF) dt_wipe_leaf(t, w & ~DT_LEAF);
        else next[atomicAdd(next_count, 1u)] = w;
    }
    t.ndepth[v] = DT_DEAD;
    t.nmeta[v] = 0;
    t.npending[v] = 0;
    t.node_free[atomicAdd(&t.g[DG_NODE_FREE], 1u)] = v;
}
// trie_of_key[j] for storage entry j of account entry i (seg_offsets[i] <= j < seg_offsets[i+1]): the account's leaf if
// the account exists after the block, DT_NONE (entry ignored) otherwise
__global__ void dt_expand_tries_kernel(const uint64_t *__restrict__ seg_offsets, uint64_t m, const uint8_t *__restrict__ kind,
                                       const uint32_t *__restrict__ leaf_of, uint64_t n_entries, uint32_t *__restrict__ trie_of_key) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_entries) return;
    uint64_t lo = 0, hi = m;  // last account with offset <= j
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (seg_offsets[mid] <= j) lo = mid;
        else hi = mid;
    }
    uint8_t k = kind[lo];
    trie_of_key[j] = (k == DK_UPDATE || k == DK_TOUCH || k == DK_INSERT) ? leaf_of[lo] : DT_NONE;
}

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
                                     uint8_t *rlp, uint64_t byte_base, uint64_t *rlp_offset, uint64_t node_base) {
    n_nodes = 0;
    n_bytes = 0;
    uint32_t cur = t.troot[trie];
    int pd = -1;
    auto begin_node = [&](uint32_t len) {
        if (WRITE) rlp_offset[node_base + n_nodes] = byte_base + n_bytes;
        n_nodes++;
        n_bytes += len;
    };
    if (cur == DT_NONE) {  // empty trie: the proof is the empty string (EMPTY_STRING_CODE), proof.rs:121-126
        if (WRITE) rlp[byte_base] = 0x80;
        begin_node(1);
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
            begin_node(len);
            return;
        }
        const uint32_t v = cur;
        const int d = t.ndepth[v];
        const uint8_t *nk = t.nkey + 32 * (uint64_t)v;
        uint32_t payload;
        const uint32_t blen = dt_branch_rlp_len(t, v, payload);
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
            begin_node(elen);
            if (!matches) return;
            begin_node(blen);
        } else {
            if (WRITE) dt_write_branch_rlp(t, v, payload, rlp + byte_base + n_bytes);
            begin_node(blen);
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
        dt_proof_walk<false>(t, trie, keys + 32 * i, nn, nb, nullptr, 0, nullptr, 0);
    }
    node_count[i] = nn;
    byte_count[i] = nb;
}
__global__ void dt_proof_write_kernel(DTrieDev t, const uint32_t *__restrict__ trie_of_target, const uint8_t *__restrict__ keys,
                                      uint64_t n, const uint64_t *__restrict__ node_base, const uint64_t *__restrict__ byte_base,
                                      uint8_t *__restrict__ rlp, uint64_t *__restrict__ rlp_offset) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t trie = trie_of_target ? trie_of_target[i] : 0;
    uint32_t nn;
    uint64_t nb;
    if (trie == DT_NONE) {
        rlp[byte_base[i]] = 0x80;
        rlp_offset[node_base[i]] = byte_base[i];
    } else {
        dt_proof_walk<true>(t, trie, keys + 32 * i, nn, nb, rlp, byte_base[i], rlp_offset, node_base[i]);
    }
}
// the account leaf (= storage trie id) of one account key, DT_NONE when the account does not exist
__global__ void dt_find_leaf_kernel(DTrieDev t, const uint8_t *__restrict__ key, uint32_t *__restrict__ out, uint64_t n_copies) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_copies) return;
    DtLoc loc = dt_descend(t, t.ltrie ? (uint32_t)(key[0] >> 4) : 0u, key);
    out[i] = loc.found ? (loc.child & ~DT_LEAF) : DT_NONE;
}

// ------------------------------------------------------------------------------------------------ launchers
// leaf_trie: trie (segment) of every leaf of a forest build, nullptr for a single trie
cudaError_t launch_dt_convert(const ForestDev &f, uint32_t n_nodes, const uint32_t *leaf_parent, const uint32_t *node_parent,
                              const uint32_t *leaf_trie, const DTrieDev &t, cudaStream_t st) {
    if (f.n) dt_convert_leaves_kernel<<<blocks_for(f.n, 256), 256, 0, st>>>(f.n, leaf_parent, leaf_trie, t);
    if (n_nodes) dt_convert_nodes_kernel<<<blocks_for(n_nodes, 128), 128, 0, st>>>(f, n_nodes, node_parent, leaf_trie, t);
    return cudaGetLastError();
}
cudaError_t launch_dt_leaf_segments(const uint64_t *seg_offsets, uint64_t n_segs, uint64_t n, uint32_t *leaf_trie, cudaStream_t st) {
    if (n) dt_leaf_segments_kernel<<<blocks_for(n, 256), 256, 0, st>>>(seg_offsets, n_segs, n, leaf_trie);
    return cudaGetLastError();
}
cudaError_t launch_dt_locate(const DTrieDev &t, const uint32_t *trie_of_key, const uint8_t *keys, const uint8_t *vals,
                             const uint8_t *flags, uint64_t m, uint8_t *kind, uint32_t *leaf_of, cudaStream_t st) {
    dt_locate_kernel<<<blocks_for(m, 128), 128, 0, st>>>(t, trie_of_key, keys, vals, flags, m, kind, leaf_of);
    return cudaGetLastError();
}
cudaError_t launch_dt_update_detach(const DTrieDev &t, const uint8_t *accts, const uint8_t *sroots, uint64_t m,
                                    const uint8_t *kind, const uint32_t *leaf_of, uint32_t *touched, cudaStream_t st) {
    dt_update_detach_kernel<<<blocks_for(m, 128), 128, 0, st>>>(t, accts, sroots, m, kind, leaf_of, touched);
    return cudaGetLastError();
}
// one collapse round over `list` (count on the device, at most max_count): begin / act / end
cudaError_t launch_dt_collapse_round(const DTrieDev &t, const uint32_t *list, const uint32_t *count_p, uint32_t max_count,
                                     uint8_t *defer, uint32_t *next, uint32_t *next_count, cudaStream_t st) {
    unsigned blocks = blocks_for(max_count, 128);
    dt_round_begin_kernel<<<blocks, 128, 0, st>>>(t, list, count_p);
    dt_round_defer_kernel<<<blocks, 128, 0, st>>>(t, list, count_p, defer);
    dt_collapse_round_kernel<<<blocks, 128, 0, st>>>(t, list, count_p, defer, next, next_count);
    dt_round_end_kernel<<<blocks, 128, 0, st>>>(t, list, count_p);
    return cudaGetLastError();
}
cudaError_t launch_dt_insert(const DTrieDev &t, const uint32_t *trie_of_key, const uint8_t *keys, const uint8_t *vals,
                             const uint8_t *sroots, const uint32_t *ins_idx, const uint32_t *n_ins_p, uint64_t max_ins,
                             uint64_t *attach, uint32_t *leaf_of, cudaStream_t st) {
    unsigned blocks = blocks_for(max_ins, 128);
    dt_insert_locate_kernel<<<blocks, 128, 0, st>>>(t, trie_of_key, keys, ins_idx, n_ins_p, attach);
    dt_insert_runs_kernel<<<blocks, 128, 0, st>>>(t, trie_of_key, keys, vals, sroots, ins_idx, n_ins_p, attach, leaf_of);
    return cudaGetLastError();
}
// mark -> starts -> wavefront -> finish (empty-trie root, recycling of this apply's freed nodes)
__global__ void dt_finish_kernel(DTrieDev t) {
    t.g[DG_NODE_FREE] += t.g[DG_FREED_NOW];
    t.g[DG_FREED_NOW] = 0;
}
cudaError_t launch_dt_rehash(const DTrieDev &t, uint32_t max_seeds, cudaStream_t st) {
    constexpr int WARPS = 4;
    const uint32_t *count_p = t.g + DG_SEEDS;
    unsigned blocks = blocks_for(max_seeds, 128);
    dt_mark_kernel<<<blocks, 128, 0, st>>>(t, count_p);
    dt_starts_kernel<<<blocks, 128, 0, st>>>(t, count_p);
    unsigned wblocks = blocks_for(max_seeds, WARPS), cap = (unsigned)sms() * 16;
    dt_wavefront_kernel<WARPS><<<wblocks < cap ? wblocks : cap, WARPS * 32, 0, st>>>(t, count_p);
    return cudaGetLastError();
}
cudaError_t launch_dt_finish(const DTrieDev &t, uint32_t max_freed, cudaStream_t st) {
    if (max_freed) dt_recycle_kernel<<<blocks_for(max_freed, 128), 128, 0, st>>>(t);
    dt_finish_kernel<<<1, 1, 0, st>>>(t);
    return cudaGetLastError();
}
cudaError_t launch_dt_stored_flags(const DTrieDev &t, uint32_t max_built, uint8_t *flags, uint32_t *n_hashes, cudaStream_t st) {
    if (max_built) dt_stored_flags_kernel<<<blocks_for(max_built, 256), 256, 0, st>>>(t, t.g + DG_BUILT, flags, n_hashes);
    return cudaGetLastError();
}
cudaError_t launch_dt_gather_updates(const DTrieDev &t, const uint32_t *stored_ids, uint32_t n_stored,
                                     const uint32_t *hash_prefix_by_record, const UpdatesDev &out, cudaStream_t st) {
    if (n_stored) dt_gather_updates_kernel<<<blocks_for(n_stored, 128), 128, 0, st>>>(t, stored_ids, n_stored, hash_prefix_by_record, out);
    return cudaGetLastError();
}
cudaError_t launch_dt_removed_paths(const DTrieDev &t, uint32_t n_removed, uint8_t *path_len, uint8_t *path_packed,
                                    uint32_t *trie_id, cudaStream_t st) {
    if (n_removed) dt_removed_paths_kernel<<<blocks_for(n_removed, 128), 128, 0, st>>>(t, n_removed, path_len, path_packed, trie_id);
    return cudaGetLastError();
}
cudaError_t launch_dt_wipe_list(const uint8_t *kind, const uint8_t *flags, const uint32_t *leaf_of, uint64_t m, uint32_t *tries,
                                uint32_t *count, cudaStream_t st) {
    if (m) dt_wipe_list_kernel<<<blocks_for(m, 256), 256, 0, st>>>(kind, flags, leaf_of, m, tries, count);
    return cudaGetLastError();
}
cudaError_t launch_dt_wipe_begin(const DTrieDev &t, const uint32_t *tries, const uint32_t *count_p, uint32_t max_count,
                                 uint32_t *next, uint32_t *next_count, cudaStream_t st) {
    if (max_count) dt_wipe_begin_kernel<<<blocks_for(max_count, 128), 128, 0, st>>>(t, tries, count_p, next, next_count);
    return cudaGetLastError();
}
cudaError_t launch_dt_wipe_round(const DTrieDev &t, const uint32_t *list, const uint32_t *count_p, uint32_t max_count,
                                 uint32_t *next, uint32_t *next_count, cudaStream_t st) {
    if (max_count) dt_wipe_round_kernel<<<blocks_for(max_count, 128), 128, 0, st>>>(t, list, count_p, next, next_count);
    return cudaGetLastError();
}
cudaError_t launch_dt_expand_tries(const uint64_t *seg_offsets, uint64_t m, const uint8_t *kind, const uint32_t *leaf_of,
                                   uint64_t n_entries, uint32_t *trie_of_key, cudaStream_t st) {
    if (n_entries) dt_expand_tries_kernel<<<blocks_for(n_entries, 256), 256, 0, st>>>(seg_offsets, m, kind, leaf_of, n_entries, trie_of_key);
    return cudaGetLastError();
}
cudaError_t launch_dt_nibble_tries(const uint8_t *keys, uint64_t m, uint32_t *trie_of_key, cudaStream_t st) {
    if (m) dt_nibble_tries_kernel<<<blocks_for(m, 256), 256, 0, st>>>(keys, m, trie_of_key);
    return cudaGetLastError();
}
cudaError_t launch_dt_frontier(const DTrieDev &t, const uint8_t *bucket_roots, FrontierEntryDev *out, cudaStream_t st) {
    dt_frontier_kernel<<<1, 512, 0, st>>>(t, bucket_roots, out);
    return cudaGetLastError();
}
cudaError_t launch_dt_proof_sizes(const DTrieDev &t, const uint32_t *trie_of_target, const uint8_t *keys, uint64_t n,
                                  uint32_t *node_count, uint64_t *byte_count, cudaStream_t st) {
    if (n) dt_proof_size_kernel<<<blocks_for(n, 64), 64, 0, st>>>(t, trie_of_target, keys, n, node_count, byte_count);
    return cudaGetLastError();
}
cudaError_t launch_dt_proof_write(const DTrieDev &t, const uint32_t *trie_of_target, const uint8_t *keys, uint64_t n,
                                  const uint64_t *node_base, const uint64_t *byte_base, uint8_t *rlp, uint64_t *rlp_offset,
                                  cudaStream_t st) {
    if (n) dt_proof_write_kernel<<<blocks_for(n, 64), 64, 0, st>>>(t, trie_of_target, keys, n, node_base, byte_base, rlp, rlp_offset);
    return cudaGetLastError();
}
cudaError_t launch_dt_find_leaf(const DTrieDev &t, const uint8_t *key, uint32_t *out, uint64_t n_copies, cudaStream_t st) {
    if (n_copies) dt_find_leaf_kernel<<<blocks_for(n_copies, 128), 128, 0, st>>>(t, key, out, n_copies);
    return cudaGetLastError();
}
