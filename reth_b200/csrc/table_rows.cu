// table_rows.cu — byte-exact rows of reth's AccountsTrie / StoragesTrie tables from a b200_updates.
// Pure host code (no kernels): the stored-node records are ≈12 bytes per committed leaf, the device has already done
// the hashing; what is left is a key-order sort and a byte layout.
//
// Formats restated (reference paths relative to the reth workspace):
//   AccountsTrie           key StoredNibbles: one nibble per byte, variable length  crates/trie/common/src/nibbles.rs:27-66
//   PackedAccountsTrie     key PackedStoredNibbles: 32 packed bytes + count = 33 B  nibbles.rs:143-213
//   StoragesTrie           key B256 hashed address, dup value StorageTrieEntry =
//                          StoredNibblesSubKey (64 nibble bytes + count = 65 B) ‖ node  nibbles.rs:68-141, storage.rs:24-44
//   PackedStoragesTrie     dup value PackedStorageTrieEntry = 33-byte subkey ‖ node     nibbles.rs:215-300, storage.rs:70-86
//   table definitions      crates/storage/db-api/src/tables/mod.rs:484-494,542-572
//   node                   BranchNodeCompact `Compact` (alloy-trie 0.9.5, external crate): state_mask, tree_mask,
//                          hash_mask as big-endian u16, then root_hash if present (never for a stored non-root node),
//                          then the child hashes; 6 + 32·popcount(hash_mask) bytes
//   row order              MDBX key order (memcmp), duplicates by subkey — the order write_trie_updates_sorted walks
//                          (crates/storage/provider/src/providers/database/provider.rs:3125-3160,
//                          crates/trie/db/src/trie_cursor.rs:280-312)
#include "b200trie.h"
#include "pinned_pool.h"
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

struct RowsOwner {  // == engine.cu's (eng_updates.inl)
    void *block;
    int pinned;
};

namespace {

struct RowKey {
    const uint8_t *addr;  // 32-byte hashed address (storage tables) or nullptr
    const uint8_t *packed;
    uint8_t len;
};

inline bool row_less(const RowKey &a, const RowKey &b) {
    if (a.addr && b.addr && a.addr != b.addr) {
        int c = memcmp(a.addr, b.addr, 32);
        if (c) return c < 0;
    }
    int c = memcmp(a.packed, b.packed, 32);  // zero padded: equal to nibble order up to the length tiebreak
    if (c) return c < 0;
    return a.len < b.len;
}

inline void put_be16(uint8_t *p, uint16_t v) {
    p[0] = (uint8_t)(v >> 8);
    p[1] = (uint8_t)v;
}

inline uint32_t nibble_key_bytes(int32_t fmt, bool subkey, uint8_t len) {
    if (fmt == B200_KEYS_PACKED) return 33;
    return subkey ? 65 : len;
}

// nibble key in the requested format; returns bytes written
inline uint32_t put_nibble_key(uint8_t *out, int32_t fmt, bool subkey, const uint8_t *packed, uint8_t len) {
    if (fmt == B200_KEYS_PACKED) {
        uint32_t full = (len + 1u) / 2u;
        memcpy(out, packed, full);
        if (len & 1) out[full - 1] &= 0xF0;
        memset(out + full, 0, 32 - full);
        out[32] = len;
        return 33;
    }
    for (uint32_t i = 0; i < len; i++) out[i] = (i & 1) ? (packed[i >> 1] & 15) : (packed[i >> 1] >> 4);
    if (!subkey) return len;
    memset(out + len, 0, 64 - len);
    out[64] = len;
    return 65;
}

int32_t encode(const b200_updates *u, const uint8_t *acct_keys32, uint64_t n_accounts, int32_t fmt, bool storage,
               b200_rows *out) {
    if (!u || !out || (fmt != B200_KEYS_LEGACY && fmt != B200_KEYS_PACKED)) return B200_ERR_INVALID_ARG;
    if (storage && !acct_keys32 && u->n_nodes) return B200_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    const uint64_t n = u->n_nodes;
    std::vector<uint64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::vector<RowKey> keys(n);
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t len = u->path_len[i];
        if (len == 0 || len > 64) return B200_ERR_INVALID_ARG;  // the empty path is never stored (updates.rs:140-158)
        const uint8_t *addr = nullptr;
        if (storage) {
            if (u->trie_id[i] >= n_accounts) return B200_ERR_INVALID_ARG;
            addr = acct_keys32 + 32ull * u->trie_id[i];
        }
        keys[i] = RowKey{addr, u->path_packed + 32 * i, len};
        uint64_t nh = u->hash_offset[i + 1] - u->hash_offset[i];
        if (nh != (uint64_t)__builtin_popcount(u->hash_mask[i])) return B200_ERR_INVALID_ARG;
        total += (storage ? 32 : 0) + nibble_key_bytes(fmt, storage, len) + 6 + 32 * nh;
    }
    // records of a full build already arrive in table order (eng_updates.inl sorts them on the device); anything else
    // (dirty subsets, hand-made updates) is ordered here
    bool in_order = true;
    for (uint64_t i = 1; i < n && in_order; i++) in_order = !row_less(keys[i], keys[i - 1]);
    if (!in_order)
        std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return row_less(keys[a], keys[b]); });

    // one block: [row_offset (n+1) u64][key_len n u32, padded to 8][bytes]
    size_t off_bytes = (n + 1) * sizeof(uint64_t), kl_bytes = ((n * sizeof(uint32_t)) + 7) & ~size_t(7);
    uint8_t *block = (uint8_t *)malloc(off_bytes + kl_bytes + (total ? total : 1));
    if (!block) return B200_ERR_OOM;
    out->_owner = new RowsOwner{block, 0};
    out->row_offset = (uint64_t *)block;
    out->key_len = (uint32_t *)(block + off_bytes);
    out->bytes = block + off_bytes + kl_bytes;
    out->n_rows = n;
    uint8_t *p = out->bytes;
    for (uint64_t r = 0; r < n; r++) {
        uint64_t i = order[r];
        out->row_offset[r] = (uint64_t)(p - out->bytes);
        if (storage) {  // key = hashed address; value = subkey ‖ node
            memcpy(p, keys[i].addr, 32);
            p += 32;
            out->key_len[r] = 32;
            p += put_nibble_key(p, fmt, true, keys[i].packed, keys[i].len);
        } else {
            uint32_t k = put_nibble_key(p, fmt, false, keys[i].packed, keys[i].len);
            out->key_len[r] = k;
            p += k;
        }
        put_be16(p, u->state_mask[i]);
        put_be16(p + 2, u->tree_mask[i]);
        put_be16(p + 4, u->hash_mask[i]);
        p += 6;
        uint64_t lo = u->hash_offset[i], hi = u->hash_offset[i + 1];
        memcpy(p, u->hashes + 32 * lo, 32 * (hi - lo));
        p += 32 * (hi - lo);
    }
    out->row_offset[n] = (uint64_t)(p - out->bytes);
    return B200_OK;
}

}  // namespace

extern "C" {

B200_API int32_t b200_account_trie_rows(const b200_updates *account_updates, int32_t key_format, b200_rows *out) {
    return encode(account_updates, nullptr, 0, key_format, false, out);
}

B200_API int32_t b200_storage_trie_rows(const b200_updates *storage_updates, const uint8_t *acct_keys32,
                                        uint64_t n_accounts, int32_t key_format, b200_rows *out) {
    return encode(storage_updates, acct_keys32, n_accounts, key_format, true, out);
}

B200_API void b200_rows_release(b200_rows *r) {
    if (!r) return;
    if (r->_owner) {
        RowsOwner *o = static_cast<RowsOwner *>(r->_owner);
        if (o->pinned) pinned_block_free(o->block);  // rows encoded on the device (eng_updates.inl: collect_rows)
        else free(o->block);
        delete o;
    }
    memset(r, 0, sizeof *r);
}

}  // extern "C"
