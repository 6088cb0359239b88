// trie_kernels.h — device-side data layout of one forest build and the kernel launch interface
// (internal to libb200trie.so).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

// sticky device error codes (mapped to b200_status by the engine)
enum : int {
    B200_DEVERR_NONE = 0,
    B200_DEVERR_UNSORTED = 1,
    B200_DEVERR_ZERO_VALUE = 2,
    B200_DEVERR_INLINE_HASH_CHILD = 3,
    B200_DEVERR_BAD_OFFSETS = 4,
    B200_DEVERR_NOT_FOUND = 5,
    B200_DEVERR_CORRUPT = 6,  // dynamic trie: a walk did not terminate within 64 hops
};

// node / leaf meta byte
// META_ISNODE (leaf_meta only): the position holds the hash of a whole unchanged subtree (HashBuilder::add_branch,
// tk_items.cuh), which its parent treats like a branch child (hash / tree mask bits)
enum : uint32_t { META_LEN = 31u, META_EXT = 32u, META_STORED = 64u, META_ISNODE = 128u };

enum : int { CNT_HASHED = 0, CNT_EXT = 1, CNT_COUNT = 4 };

struct b200_account_dev {  // == b200_account (include/b200trie.h), 72 bytes
    uint64_t nonce;
    uint8_t balance_be[32];
    uint8_t code_hash[32];
};
static_assert(sizeof(b200_account_dev) == 72, "account layout");

struct FrontierEntryDev {  // == b200_frontier_entry
    uint8_t as_child_len;
    uint8_t as_child[33];
    uint8_t as_root_len;
    uint8_t as_root[33];
};
static_assert(sizeof(FrontierEntryDev) == 68, "frontier layout");

// All arrays live in HBM for the duration of one build (and stay resident afterwards: the node-hash
// frontier of every level is exactly node_ref/leaf_ref).
//
//   per leaf  (n)    : keys 32 B (input) | Lp 1 | nibs 1 | leaf_ref 32 | leaf_meta 1 | S 4 | E 4
//   per gap   (n-1)  : key_sorted 2 (depth | head flag << 8) | gap_sorted 4   (gaps ordered by (depth, position))
//   per branch (B)   : node_start 4 (CSR into gap_sorted) | node_ref 32 | node_meta 1 | node_l 4 | node_r 4 |
//                      node_masks 8 (state, tree, hash, depth)
struct ForestDev {
    uint64_t n;             // leaves
    const uint8_t *keys;    // [n][32] sorted inside each trie
    uint8_t *Lp;            // [n+1] gap depth; Lp[0] = Lp[n] = 0xFF; 0xFF = trie boundary
    uint8_t *nibs;          // [n+1] (left nibble << 4) | right nibble at depth Lp
    uint8_t *leaf_ref;      // [n][32] digest, or inline RLP (< 32 bytes)
    uint8_t *leaf_meta;     // [n] 0 = hashed, else inline length
    uint32_t *S, *E;        // [n] frontier item starting / ending at this leaf (< n leaf, else n + node id)
    const uint32_t *gap_sorted;   // [n-1]
    const uint32_t *node_start;   // [B+1]
    uint8_t *node_ref;      // [B][32]
    uint8_t *node_meta;     // [B] inline length | META_EXT | META_STORED
    uint32_t *node_l, *node_r;    // [B] leaf extent
    ushort4 *node_masks;    // [B]
    int *err;               // sticky error code
    unsigned long long *counters;  // [CNT_COUNT]
    int retain_updates;
};

struct UpdatesDev {
    uint32_t *trie_id;
    uint8_t *path_len;
    uint8_t *path_packed;
    uint16_t *state_mask, *tree_mask, *hash_mask;
    uint32_t *hash_offset;  // [n_stored] (exclusive); the engine appends the total
    uint8_t *hashes;
};

cudaError_t launch_latch_error(int *err, int *sticky, unsigned long long *counters, cudaStream_t st);
cudaError_t launch_mark_boundaries(const uint64_t *d_seg_offsets, uint64_t n_segs, uint64_t n, uint8_t *Lp, int *err,
                                   cudaStream_t st);
cudaError_t launch_lcp(const uint8_t *keys, uint64_t n, uint8_t *Lp, uint8_t *nibs, int *err, cudaStream_t st);
cudaError_t launch_iota(uint32_t *out, uint64_t n, uint32_t first, cudaStream_t st);
cudaError_t launch_gap_keys(const uint8_t *Lp, uint64_t G, uint16_t *key, uint32_t *val, uint32_t *unresolved, cudaStream_t st);
cudaError_t launch_bucket_offsets(const uint16_t *key_sorted, uint64_t G, uint32_t *bucket_off, cudaStream_t st);
cudaError_t launch_head_fix(const uint8_t *keys, uint16_t *key_sorted, const uint32_t *gap_sorted, const uint64_t *seg_offsets,
                            uint64_t n_segs, const uint32_t *unresolved, uint64_t G, cudaStream_t st);
cudaError_t launch_level_ranges(uint32_t *node_start, const uint32_t *n_nodes_p, const uint32_t *bucket_off,
                                uint32_t *level_lo, cudaStream_t st);
cudaError_t launch_leaves(const ForestDev &f, bool account, const uint8_t *values, const uint8_t *storage_roots,
                          cudaStream_t st);
cudaError_t launch_branch_level(const ForestDev &f, const uint32_t *node_order, uint32_t pos_lo, uint32_t pos_hi,
                                int d, int cls, cudaStream_t st);
cudaError_t launch_node_class_keys(const uint32_t *node_start, const uint16_t *key_sorted, const uint32_t *n_nodes_p,
                                   uint64_t max_nodes, uint8_t *keys, uint32_t *ids, uint32_t *hist, cudaStream_t st);
cudaError_t launch_segment_roots(const ForestDev &f, const uint64_t *d_seg_offsets, uint64_t n_segs, uint8_t *roots,
                                 cudaStream_t st);
cudaError_t launch_stored_flags(const ForestDev &f, uint32_t n_nodes, uint8_t *flags, uint32_t *n_hashes,
                                cudaStream_t st);
cudaError_t launch_table_order_keys(const ForestDev &f, const uint32_t *ids, uint32_t count, uint64_t *keys, cudaStream_t st);
cudaError_t launch_row_sizes(const ForestDev &f, const uint32_t *ids, uint32_t count, int packed, int storage, uint64_t *size,
                             uint32_t *key_len, cudaStream_t st);
cudaError_t launch_encode_rows(const ForestDev &f, const uint32_t *ids, uint32_t count, int packed, int storage,
                               const uint64_t *d_seg_offsets, uint64_t n_segs, const uint8_t *acct_keys,
                               const uint64_t *row_off, uint8_t *out, cudaStream_t st);
cudaError_t launch_gather_updates(const ForestDev &f, const uint32_t *stored_ids, uint32_t n_stored,
                                  const uint32_t *hash_prefix, const uint32_t *prefix_by_record,
                                  const uint64_t *d_seg_offsets, uint64_t n_segs, const UpdatesDev &out,
                                  cudaStream_t st);
cudaError_t launch_nibble_buckets(const uint8_t *keys, uint64_t n, uint64_t *offs, cudaStream_t st);
cudaError_t launch_frontier(const ForestDev &f, const uint64_t *bucket_offsets, const uint8_t *values,
                            const uint8_t *storage_roots, FrontierEntryDev *out, cudaStream_t st);
cudaError_t launch_root_from_frontier(const FrontierEntryDev *fr, uint8_t *root, cudaStream_t st);
cudaError_t launch_merge_frontiers(const FrontierEntryDev *all, int world, FrontierEntryDev *out, int *err, cudaStream_t st);
cudaError_t launch_partition_owner(const uint8_t *digests, uint64_t n, int world, uint8_t *owner, unsigned long long *counts,
                                   cudaStream_t st);
cudaError_t launch_partition_gather(const uint8_t *digests, const uint8_t *values, uint32_t vb, const uint32_t *perm, uint64_t n,
                                    uint8_t *out_d, uint8_t *out_v, cudaStream_t st);
cudaError_t launch_gather_values(const uint8_t *values, uint32_t vb, const uint32_t *perm, uint64_t n, uint8_t *out_v, cudaStream_t st);

cudaError_t launch_parent_links(const ForestDev &f, uint32_t n_nodes, uint32_t *leaf_parent, uint32_t *node_parent,
                                cudaStream_t st);
cudaError_t launch_locate(const uint8_t *keys, uint64_t n, const uint8_t *dirty_keys, uint64_t m, uint32_t *idx_out,
                          int *err, cudaStream_t st);
cudaError_t launch_mark_pending(const ForestDev &f, const uint32_t *idx, uint64_t m, const uint32_t *leaf_parent,
                                const uint32_t *node_parent, uint32_t *pending, cudaStream_t st);
cudaError_t launch_wavefront(const ForestDev &f, uint8_t *accts, uint8_t *sroots, const uint8_t *new_accts,
                             const uint8_t *new_sroots, const uint32_t *idx, uint64_t m, const uint32_t *leaf_parent,
                             const uint32_t *node_parent, uint32_t *pending, uint32_t *dirty_list, uint32_t *dirty_count,
                             uint8_t *root_out, cudaStream_t st);
cudaError_t launch_wavefront_two_stage(const ForestDev &f, uint8_t *accts, uint8_t *sroots, const uint8_t *new_accts,
                                       const uint8_t *new_sroots, const uint32_t *idx, uint64_t m,
                                       const uint32_t *leaf_parent, const uint32_t *node_parent, uint32_t *pending,
                                       uint32_t *dirty_list, uint32_t *dirty_count, uint32_t *handoff_list,
                                       uint32_t *handoff_count, uint64_t max_handoff, uint8_t *root_out, int split_depth,
                                       cudaStream_t st);
cudaError_t launch_locate_classify(const uint8_t *keys, uint64_t n, const uint8_t *dirty_keys, const uint8_t *present,
                                   uint64_t m, uint32_t *lb, uint8_t *kind, uint32_t *counts, int *err, cudaStream_t st);
cudaError_t launch_merge_marks(const uint32_t *lb, const uint8_t *kind, uint64_t m, uint32_t *ins_at, uint32_t *del,
                               uint32_t *ins_flag, cudaStream_t st);
cudaError_t launch_merge_scatter(const uint8_t *keys, const uint8_t *accts, const uint8_t *sroots, uint64_t n,
                                 const uint32_t *ins_incl, const uint32_t *del_excl, const uint32_t *del,
                                 const uint8_t *dirty_keys, const uint8_t *new_accts, const uint8_t *new_sroots,
                                 const uint32_t *lb, const uint8_t *kind, const uint32_t *ins_rank, uint64_t m, uint8_t *nkeys,
                                 uint8_t *naccts, uint8_t *nsroots, cudaStream_t st);
cudaError_t launch_stored_flags_subset(const ForestDev &f, const uint32_t *ids, uint32_t count, uint8_t *flags,
                                       uint32_t *n_hashes, cudaStream_t st);
cudaError_t launch_pick_subset(const uint32_t *ids, const uint32_t *prefix, const uint32_t *sel_pos, uint32_t n_sel,
                               uint32_t *out_ids, uint32_t *out_prefix, cudaStream_t st);

// ------------------------------------------------------------------------------------------------ ordered tries (tk_ordered.cuh)
struct OrderedLeavesDev {
    const uint8_t *key_nibs;  // [n] true key length in nibbles (the padded keys are ForestDev::keys)
    const uint32_t *item;     // [n] item (in list order) carried by the leaf at this sorted position
    const uint32_t *order;    // [n] leaf positions in the order the leaf pass visits them (longest item first)
    const uint16_t *sched_sorted;  // [n] the sorted scheduling keys (65535 - Keccak blocks of the item)
    uint32_t *n_long;         // device word: leading entries of `order` that get a warp each
    const uint8_t *values;    // concatenated pre-encoded items
    const uint64_t *val_off;  // [n+1] byte offsets of the items in `values`
    uint64_t blob_len;
};
cudaError_t launch_ordered_keys(const uint64_t *d_seg_offsets, uint64_t n_segs, uint64_t n, const uint64_t *val_off,
                                uint8_t *keys, uint8_t *key_nibs, uint32_t *item, uint16_t *sched_key, uint32_t *pos, int *err,
                                cudaStream_t st);
cudaError_t launch_ordered_leaves(const ForestDev &f, const OrderedLeavesDev &o, cudaStream_t st);

// ------------------------------------------------------------------------------------------------ mixed items (tk_items.cuh)
// The input stream of an incremental HashBuilder run (TrieNodeIter, crates/trie/trie/src/node_iter.rs:200-304): changed /
// uncovered leaves and the stored hashes of unchanged subtrees, in key order.
struct ItemLeavesDev {
    const uint8_t *key_nibs;  // [n] 64 = a leaf; 0..63 = hash of the subtree rooted at the path of that many nibbles
    const uint8_t *flags;     // [n] bit 0: the subtree's nodes are in the trie tables (children_are_in_trie -> tree mask)
    int account;              // leaf values: b200_account (72-byte rows) | U256 BE (32-byte rows); hashes: first 32 bytes of the row
};
cudaError_t launch_item_leaves(const ForestDev &f, const ItemLeavesDev &it, const uint8_t *values, const uint8_t *storage_roots,
                               cudaStream_t st);

// ------------------------------------------------------------------------------------------------ dynamic trie (tk_dtrie.cuh)
constexpr uint32_t DT_NONE = 0xFFFFFFFFu;
constexpr uint32_t DT_LEAF = 0x80000000u;
constexpr uint32_t DT_LOCKED = 0xFFFFFFFEu;  // an attach word owned by an insert run for the duration of a round (never a valid id: ids < 2^31 - 2)
constexpr uint8_t DT_DEAD = 0xFF;  // ndepth / lmeta of a freed slot
constexpr int DT_MAX_HOPS = 66;

enum : int {
    DG_UNUSED0 = 0,
    DG_NLEAVES,       // live leaves
    DG_LEAF_ALLOC,    // bump pointers
    DG_NODE_ALLOC,
    DG_LEAF_FREE,     // free-stack heights
    DG_NODE_FREE,
    DG_SEEDS,         // list lengths of the current apply
    DG_BUILT,
    DG_REMOVED,
    DG_LIST_A,
    DG_LIST_B,
    DG_NINSERT,
    DG_FREED_NOW,
    DG_WORDS = 16
};

struct DTrieDev {
    // leaves [lcap]
    uint8_t *lkey, *lval, *lsroot, *lref, *lmeta;  // lval: 72-byte account or 32-byte slot value; lsroot: accounts only
    uint32_t *lparent, *ltrie;                     // ltrie / ntrie: owning trie (nullptr = a single trie, id 0)
    uint8_t *lseed;
    // nodes [ncap]
    uint32_t *nchild;  // [ncap][16]
    uint8_t *ndepth;
    uint32_t *nparent;
    uint8_t *nref, *nmeta;
    ushort4 *nmasks;
    uint8_t *nkey;  // [ncap][32] a key of the subtree: its first ndepth nibbles are the node's path
    uint32_t *npending, *ntrie;
    uint8_t *nseed, *ncur, *nnext;
    // tries
    uint32_t *troot;     // [n_tries] child word of every trie's root
    uint8_t *top_out;    // the warp that finishes trie r stores its new root hash at top_out + top_stride * r
    uint32_t top_stride;
    uint32_t val_stride;  // 72 | 32
    int account;          // leaf encoding: rlp(TrieAccount) | rlp(U256)
    // free stacks, lists, globals
    uint32_t *leaf_free, *node_free;
    uint32_t *seeds, *built, *removed, *freed_now;
    uint32_t *unlock;  // [insert entries of the round] final value of the attach word a run owned, DT_LOCKED = none
    uint32_t *g;
    int *err;
    unsigned long long *counters;
    uint32_t lcap, ncap;
};

enum : uint8_t { DK_NOOP = 0, DK_UPDATE = 1, DK_DELETE = 2, DK_INSERT = 3, DK_TOUCH = 4 };

cudaError_t launch_dt_convert(const ForestDev &f, uint32_t n_nodes, const uint32_t *leaf_parent, const uint32_t *node_parent,
                              const uint32_t *leaf_trie, const DTrieDev &t, cudaStream_t st);
cudaError_t launch_dt_leaf_segments(const uint64_t *seg_offsets, uint64_t n_segs, uint64_t n, uint32_t *leaf_trie, cudaStream_t st);
cudaError_t launch_dt_locate(const DTrieDev &t, const uint32_t *trie_of_key, const uint8_t *keys, const uint8_t *vals,
                             const uint8_t *flags, uint64_t m, uint8_t *kind, uint32_t *leaf_of, cudaStream_t st);
cudaError_t launch_dt_update_detach(const DTrieDev &t, const uint8_t *accts, const uint8_t *sroots, uint64_t m,
                                    const uint8_t *kind, const uint32_t *leaf_of, uint32_t *touched, cudaStream_t st);
cudaError_t launch_dt_collapse_round(const DTrieDev &t, const uint32_t *list, const uint32_t *count_p, uint32_t max_count,
                                     uint8_t *defer, uint32_t *next, uint32_t *next_count, cudaStream_t st);
cudaError_t launch_dt_insert(const DTrieDev &t, const uint32_t *trie_of_key, const uint8_t *keys, const uint8_t *vals,
                             const uint8_t *sroots, const uint32_t *ins_idx, const uint32_t *n_ins_p, uint64_t max_ins,
                             uint64_t *attach, uint32_t *leaf_of, uint32_t max_per_run, uint8_t *pending, uint32_t *leftover,
                             cudaStream_t st);
cudaError_t launch_dt_rehash(const DTrieDev &t, uint32_t max_seeds, uint32_t *handoff, uint32_t *handoff_count, int split_depth,
                             bool already_marked, cudaStream_t st);
cudaError_t launch_dt_finish(const DTrieDev &t, uint32_t max_freed, cudaStream_t st);
cudaError_t launch_dt_stored_flags(const DTrieDev &t, uint32_t max_built, uint8_t *flags, uint32_t *n_hashes, cudaStream_t st);
cudaError_t launch_dt_gather_updates(const DTrieDev &t, const uint32_t *stored_ids, uint32_t n_stored,
                                     const uint32_t *hash_prefix_by_record, const UpdatesDev &out, cudaStream_t st);
cudaError_t launch_dt_removed_paths(const DTrieDev &t, uint32_t n_removed, uint8_t *path_len, uint8_t *path_packed,
                                    uint32_t *trie_id, cudaStream_t st);

cudaError_t launch_dt_wipe_list(const uint8_t *kind, const uint8_t *flags, const uint32_t *leaf_of, uint64_t m, uint32_t *tries,
                                uint32_t *count, cudaStream_t st);
cudaError_t launch_dt_wipe_begin(const DTrieDev &t, const uint32_t *tries, const uint32_t *count_p, uint32_t max_count,
                                 cudaStream_t st);
cudaError_t launch_dt_wipe_round(const DTrieDev &t, uint32_t lo, uint32_t hi, cudaStream_t st);
cudaError_t launch_dt_expand_tries(const uint64_t *seg_offsets, uint64_t m, const uint8_t *kind, const uint32_t *leaf_of,
                                   uint64_t n_entries, uint32_t *trie_of_key, cudaStream_t st);

cudaError_t launch_dt_nibble_tries(const uint8_t *keys, uint64_t m, uint32_t *trie_of_key, cudaStream_t st);
cudaError_t launch_dt_frontier(const DTrieDev &t, const uint8_t *bucket_roots, FrontierEntryDev *out, cudaStream_t st);

cudaError_t launch_dt_proof_sizes(const DTrieDev &t, const uint32_t *trie_of_target, const uint8_t *keys, uint64_t n,
                                  uint32_t *node_count, uint64_t *byte_count, cudaStream_t st);
cudaError_t launch_dt_proof_write(const DTrieDev &t, const uint32_t *trie_of_target, const uint8_t *keys, uint64_t n,
                                  const uint64_t *node_base, const uint64_t *byte_base, uint8_t *rlp, uint64_t *rlp_offset,
                                  uint8_t *node_depth, uint32_t *node_masks, cudaStream_t st);
cudaError_t launch_dt_find_leaves(const DTrieDev &t, const uint8_t *keys, uint64_t n, uint32_t *leaf_out, uint8_t *sroot_out, cudaStream_t st);
cudaError_t launch_dt_target_tries(const uint64_t *seg_offsets, uint64_t n_accounts, const uint32_t *leaf_of, uint64_t n_targets,
                                   uint32_t *trie_of_target, cudaStream_t st);
cudaError_t launch_dt_find_leaf(const DTrieDev &t, const uint8_t *key, uint32_t *out, uint64_t n_copies, cudaStream_t st);

cudaError_t launch_dt_restructure_fused(const DTrieDev &t, const uint32_t *trie_of_key, const uint8_t *keys, const uint8_t *vals,
                                        const uint8_t *flags, const uint8_t *sroots, uint32_t m, uint8_t *kind, uint32_t *leaf_of,
                                        uint32_t *list_a, uint32_t *list_b, uint8_t *defer, uint32_t *idx_a, uint32_t *idx_b,
                                        uint64_t *attach, uint8_t *pending, uint32_t max_per_run, cudaStream_t st);

}  // namespace b200
