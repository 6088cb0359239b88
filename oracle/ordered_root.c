/*
 * ordered_root.c — CPU restatement of reth's ordered (index-keyed) trie roots.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle.h).
 *
 *   orc_ordered_roots   OrderedTrieRootEncodedBuilder::{flush, finalize}
 *                       crates/trie/common/src/ordered_root.rs:202-216, 240-257
 *                       (empty list -> EMPTY_ROOT_HASH :241-243; key = alloy_rlp::encode_fixed_size(&index) :210,
 *                        insertion order = alloy_trie::root::adjust_index_for_rlp, alloy-trie 0.9.5 src/root.rs — a third-party
 *                        dependency not vendored in the reference: i > 0x7f -> i; i == 0x7f || i + 1 == len -> 0;
 *                        else i + 1)
 * Pinned by the golden roots of crates/ethereum/primitives/src/receipt.rs:180-245 (tests/golden/ordered_roots.json).
 */
#include "oracle.h"
#include <string.h>

static const uint8_t EMPTY_ROOT[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45,
                                       0xe6, 0x92, 0xc0, 0xf8, 0x6e, 0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c,
                                       0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

static uint64_t adjust_index_for_rlp(uint64_t i, uint64_t len) {
    if (i > 0x7f) return i;
    if (i == 0x7f || i + 1 == len) return 0;
    return i + 1;
}

/* rlp(usize) -> nibbles; returns the nibble count */
static size_t index_key_nibbles(uint64_t idx, uint8_t nib[18]) {
    uint8_t b[9];
    size_t n = 0;
    if (idx == 0) {
        b[n++] = 0x80;
    } else if (idx < 0x80) {
        b[n++] = (uint8_t)idx;
    } else {
        int bytes = 0;
        for (uint64_t t = idx; t; t >>= 8) bytes++;
        b[n++] = (uint8_t)(0x80 + bytes);
        for (int k = bytes - 1; k >= 0; k--) b[n++] = (uint8_t)(idx >> (8 * k));
    }
    for (size_t k = 0; k < n; k++) {
        nib[2 * k] = b[k] >> 4;
        nib[2 * k + 1] = b[k] & 15;
    }
    return 2 * n;
}

int orc_ordered_roots(const uint8_t *values, const uint64_t *value_offsets, const uint64_t *seg_offsets,
                      uint64_t n_lists, uint8_t *roots32) {
    for (uint64_t l = 0; l < n_lists; l++) {
        uint64_t lo = seg_offsets[l], len = seg_offsets[l + 1] - lo;
        if (len == 0) {
            memcpy(roots32 + 32 * l, EMPTY_ROOT, 32);
            continue;
        }
        orc_hb *hb = orc_hb_new(0);
        int rc = 0;
        for (uint64_t i = 0; i < len && rc == 0; i++) {
            uint64_t idx = adjust_index_for_rlp(i, len);
            uint8_t nib[18];
            size_t kn = index_key_nibbles(idx, nib);
            uint64_t a = value_offsets[lo + idx], b = value_offsets[lo + idx + 1];
            if (orc_hb_add_leaf(hb, nib, kn, values + a, (size_t)(b - a)) != 0) rc = -1;
        }
        if (rc == 0) orc_hb_root(hb, roots32 + 32 * l);
        orc_hb_free(hb);
        if (rc) return rc;
    }
    return 0;
}
