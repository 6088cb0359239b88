/*
 * state_root.c — CPU restatement of reth's root drivers over flat sorted arrays.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle.h).
 *
 *   orc_storage_roots     StorageRoot::calculate          crates/trie/trie/src/trie.rs:615-721
 *                         (empty storage -> EMPTY_ROOT_HASH :622-629; leaf value =
 *                          alloy_rlp::encode_fixed_size(U256) :668-671)
 *   orc_state_root        StateRoot::calculate, leaf path  crates/trie/trie/src/trie.rs:247-309,429-432
 *   orc_state_root_full   serial: trie.rs:160-330; threads>1: ParallelStateRoot::calculate
 *                         crates/trie/parallel/src/root.rs:81-221 (storage roots computed by a pool, account
 *                         trie folded serially on the caller thread)
 *   updates               TrieUpdates::finalize            crates/trie/common/src/updates.rs:140-158
 *                         (the empty-path entry is dropped: exclude_empty_from_pair :822-832)
 *   orc_trie_root_recursive  independent second implementation (the role `triehash` plays in
 *                         crates/trie/trie/src/test_utils.rs:9-49)
 */
#include "oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

extern void orc__stats_merge(const orc_stats *s);

static const uint8_t EMPTY_ROOT[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45,
                                       0xe6, 0x92, 0xc0, 0xf8, 0x6e, 0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c,
                                       0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

static void unpack_nibbles(const uint8_t k[32], uint8_t out[64]) {
    for (int i = 0; i < 32; i++) {
        out[2 * i] = k[i] >> 4;
        out[2 * i + 1] = k[i] & 15;
    }
}

/* ------------------------------------------------------------------ updates collection */
typedef struct {
    orc_updates u;
    uint64_t cap_nodes, cap_hashes, n_hashes;
} upd_builder;

static void ub_reserve(upd_builder *b, uint64_t more_nodes, uint64_t more_hashes) {
    if (b->u.n_nodes + more_nodes + 1 > b->cap_nodes) {
        uint64_t c = (b->cap_nodes ? b->cap_nodes * 2 : 64) + more_nodes;
        b->u.trie_id = (uint32_t *)realloc(b->u.trie_id, c * 4);
        b->u.path_len = (uint8_t *)realloc(b->u.path_len, c);
        b->u.path_packed = (uint8_t *)realloc(b->u.path_packed, c * 32);
        b->u.state_mask = (uint16_t *)realloc(b->u.state_mask, c * 2);
        b->u.tree_mask = (uint16_t *)realloc(b->u.tree_mask, c * 2);
        b->u.hash_mask = (uint16_t *)realloc(b->u.hash_mask, c * 2);
        b->u.hash_offset = (uint64_t *)realloc(b->u.hash_offset, (c + 1) * 8);
        b->cap_nodes = c;
    }
    if (b->n_hashes + more_hashes > b->cap_hashes) {
        uint64_t c = (b->cap_hashes ? b->cap_hashes * 2 : 64) + more_hashes;
        b->u.hashes = (uint8_t *)realloc(b->u.hashes, c * 32);
        b->cap_hashes = c;
    }
}

/* append the hash builder's updated_branch_nodes minus the empty path (updates.rs:147) */
static void ub_append_from_hb(upd_builder *b, orc_hb *hb, uint32_t trie_id) {
    size_t n = orc_hb_updates_len(hb);
    for (size_t i = 0; i < n; i++) {
        const orc_branch_node *bn = orc_hb_update_at(hb, i);
        if (bn->path_len == 0) continue;
        ub_reserve(b, 1, bn->n_hashes);
        uint64_t k = b->u.n_nodes;
        b->u.trie_id[k] = trie_id;
        b->u.path_len[k] = bn->path_len;
        uint8_t *pp = b->u.path_packed + 32 * k;
        memset(pp, 0, 32);
        for (int j = 0; j < bn->path_len; j++) pp[j >> 1] |= (uint8_t)(bn->path[j] << ((j & 1) ? 0 : 4));
        b->u.state_mask[k] = bn->state_mask;
        b->u.tree_mask[k] = bn->tree_mask;
        b->u.hash_mask[k] = bn->hash_mask;
        b->u.hash_offset[k] = b->n_hashes;
        memcpy(b->u.hashes + 32 * b->n_hashes, bn->hashes, 32 * (size_t)bn->n_hashes);
        b->n_hashes += bn->n_hashes;
        b->u.n_nodes = k + 1;
        b->u.hash_offset[k + 1] = b->n_hashes;
    }
}

static void ub_finish(upd_builder *b, orc_updates *out) {
    if (b->u.n_nodes == 0) {
        ub_reserve(b, 0, 0);
        b->u.hash_offset[0] = 0;
    }
    *out = b->u;
}

/* merge builder `src` into `dst` (used by the parallel driver to keep account order) */
static void ub_merge(upd_builder *dst, const upd_builder *src) {
    if (src->u.n_nodes == 0) return;
    ub_reserve(dst, src->u.n_nodes, src->n_hashes);
    uint64_t k = dst->u.n_nodes, n = src->u.n_nodes;
    memcpy(dst->u.trie_id + k, src->u.trie_id, n * 4);
    memcpy(dst->u.path_len + k, src->u.path_len, n);
    memcpy(dst->u.path_packed + 32 * k, src->u.path_packed, n * 32);
    memcpy(dst->u.state_mask + k, src->u.state_mask, n * 2);
    memcpy(dst->u.tree_mask + k, src->u.tree_mask, n * 2);
    memcpy(dst->u.hash_mask + k, src->u.hash_mask, n * 2);
    for (uint64_t i = 0; i <= n; i++) dst->u.hash_offset[k + i] = dst->n_hashes + src->u.hash_offset[i];
    memcpy(dst->u.hashes + 32 * dst->n_hashes, src->u.hashes, src->n_hashes * 32);
    dst->n_hashes += src->n_hashes;
    dst->u.n_nodes = k + n;
}

void orc_updates_free(orc_updates *u) {
    if (!u) return;
    free(u->trie_id);
    free(u->path_len);
    free(u->path_packed);
    free(u->state_mask);
    free(u->tree_mask);
    free(u->hash_mask);
    free(u->hash_offset);
    free(u->hashes);
    memset(u, 0, sizeof *u);
}

/* ------------------------------------------------------------------ one storage trie */
static int storage_root_one(const uint8_t *keys, const uint8_t *vals, uint64_t n, uint8_t root[32],
                            upd_builder *ub, uint32_t trie_id) {
    if (n == 0) { /* trie.rs:622-629 */
        memcpy(root, EMPTY_ROOT, 32);
        return 0;
    }
    orc_hb *hb = orc_hb_new(ub != NULL);
    uint8_t nib[64], val[33];
    int rc = 0;
    for (uint64_t i = 0; i < n && rc == 0; i++) {
        unpack_nibbles(keys + 32 * i, nib);
        size_t vl = orc_encode_u256(vals + 32 * i, val);
        if (vl == 1 && val[0] == 0x80) rc = -2; /* zero value == deleted slot; never a leaf */
        else if (orc_hb_add_leaf(hb, nib, 64, val, vl) != 0) rc = -1;
    }
    if (rc == 0) {
        orc_hb_root(hb, root);
        if (ub) ub_append_from_hb(ub, hb, trie_id);
    }
    orc_hb_free(hb);
    return rc;
}

typedef struct {
    const uint8_t *keys, *vals;
    const uint64_t *off;
    uint64_t n_accounts;
    uint8_t *roots;
    int want_updates;
    uint64_t next;
    int rc;
    /* per-worker */
} sr_shared;

typedef struct {
    sr_shared *sh;
    upd_builder ub;
    orc_stats stats;
} sr_worker;

#define SR_CHUNK 64

static void sr_run(sr_worker *w) {
    sr_shared *sh = w->sh;
    for (;;) {
        uint64_t lo = __atomic_fetch_add(&sh->next, SR_CHUNK, __ATOMIC_RELAXED);
        if (lo >= sh->n_accounts) break;
        uint64_t hi = lo + SR_CHUNK < sh->n_accounts ? lo + SR_CHUNK : sh->n_accounts;
        for (uint64_t a = lo; a < hi; a++) {
            uint64_t s = sh->off[a], e = sh->off[a + 1];
            int rc = storage_root_one(sh->keys + 32 * s, sh->vals + 32 * s, e - s, sh->roots + 32 * a,
                                      sh->want_updates ? &w->ub : NULL, (uint32_t)a);
            if (rc) __atomic_store_n(&sh->rc, rc, __ATOMIC_RELAXED);
        }
    }
}

static void *sr_worker_main(void *p) {
    sr_worker *w = (sr_worker *)p;
    orc_stats_reset();
    sr_run(w);
    orc_hb_free(NULL);
    orc_stats_get(&w->stats);
    return NULL;
}

static int cmp_trie_id_path(const orc_updates *u, uint64_t a, uint64_t b) {
    if (u->trie_id[a] != u->trie_id[b]) return u->trie_id[a] < u->trie_id[b] ? -1 : 1;
    /* compare nibble paths: packed bytes then length */
    uint8_t la = u->path_len[a], lb = u->path_len[b];
    uint8_t n = la < lb ? la : lb;
    const uint8_t *pa = u->path_packed + 32 * a, *pb = u->path_packed + 32 * b;
    for (uint8_t i = 0; i < n; i++) {
        uint8_t x = (i & 1) ? (pa[i >> 1] & 15) : (pa[i >> 1] >> 4);
        uint8_t y = (i & 1) ? (pb[i >> 1] & 15) : (pb[i >> 1] >> 4);
        if (x != y) return x < y ? -1 : 1;
    }
    return la < lb ? -1 : (la > lb ? 1 : 0);
}

/* canonical order: (trie_id, path). Simple merge sort over an index array, then permute. */
static void updates_sort(orc_updates *u) {
    uint64_t n = u->n_nodes;
    if (n < 2) return;
    uint64_t *idx = (uint64_t *)malloc(8 * n), *tmp = (uint64_t *)malloc(8 * n);
    for (uint64_t i = 0; i < n; i++) idx[i] = i;
    for (uint64_t w = 1; w < n; w *= 2) {
        for (uint64_t lo = 0; lo < n; lo += 2 * w) {
            uint64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            uint64_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = cmp_trie_id_path(u, idx[j], idx[i]) < 0 ? idx[j++] : idx[i++];
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        uint64_t *t = idx; idx = tmp; tmp = t;
    }
    orc_updates o;
    memset(&o, 0, sizeof o);
    o.n_nodes = n;
    o.trie_id = (uint32_t *)malloc(4 * n);
    o.path_len = (uint8_t *)malloc(n);
    o.path_packed = (uint8_t *)malloc(32 * n);
    o.state_mask = (uint16_t *)malloc(2 * n);
    o.tree_mask = (uint16_t *)malloc(2 * n);
    o.hash_mask = (uint16_t *)malloc(2 * n);
    o.hash_offset = (uint64_t *)malloc(8 * (n + 1));
    uint64_t nh = u->hash_offset[n];
    o.hashes = (uint8_t *)malloc(32 * (nh ? nh : 1));
    uint64_t h = 0;
    for (uint64_t k = 0; k < n; k++) {
        uint64_t s = idx[k];
        o.trie_id[k] = u->trie_id[s];
        o.path_len[k] = u->path_len[s];
        memcpy(o.path_packed + 32 * k, u->path_packed + 32 * s, 32);
        o.state_mask[k] = u->state_mask[s];
        o.tree_mask[k] = u->tree_mask[s];
        o.hash_mask[k] = u->hash_mask[s];
        o.hash_offset[k] = h;
        uint64_t c = u->hash_offset[s + 1] - u->hash_offset[s];
        memcpy(o.hashes + 32 * h, u->hashes + 32 * u->hash_offset[s], 32 * c);
        h += c;
    }
    o.hash_offset[n] = h;
    free(idx);
    free(tmp);
    orc_updates_free(u);
    *u = o;
}

int orc_storage_roots(const uint8_t *slot_keys32, const uint8_t *values32_be, const uint64_t *seg_offsets,
                      uint64_t n_accounts, uint8_t *roots32, orc_updates *opt_updates, int threads) {
    if (threads < 1) threads = 1;
    sr_shared sh = {slot_keys32, values32_be, seg_offsets, n_accounts, roots32, opt_updates != NULL, 0, 0};
    sr_worker *w = (sr_worker *)calloc((size_t)threads, sizeof *w);
    for (int i = 0; i < threads; i++) w[i].sh = &sh;
    if (threads == 1) {
        sr_run(&w[0]); /* inline: counters accumulate on the caller thread */
    } else {
        pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, sr_worker_main, &w[i]);
        for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
        free(t);
    }
    upd_builder all;
    memset(&all, 0, sizeof all);
    for (int i = 0; i < threads; i++) {
        if (threads > 1) orc__stats_merge(&w[i].stats);
        if (opt_updates) ub_merge(&all, &w[i].ub);
        orc_updates_free(&w[i].ub.u);
    }
    free(w);
    if (opt_updates) {
        ub_finish(&all, opt_updates);
        updates_sort(opt_updates);
    }
    return sh.rc;
}

int orc_state_root(const uint8_t *acct_keys32, const orc_account *accts, const uint8_t *storage_roots32,
                   uint64_t n, uint8_t root32[32], orc_updates *opt_updates) {
    orc_hb *hb = orc_hb_new(opt_updates != NULL);
    uint8_t nib[64], rlp[112];
    int rc = 0;
    for (uint64_t i = 0; i < n && rc == 0; i++) {
        unpack_nibbles(acct_keys32 + 32 * i, nib);
        size_t l = orc_encode_trie_account(&accts[i], storage_roots32 ? storage_roots32 + 32 * i : EMPTY_ROOT, rlp);
        if (orc_hb_add_leaf(hb, nib, 64, rlp, l) != 0) rc = -1;
    }
    if (rc == 0) {
        orc_hb_root(hb, root32);
        if (opt_updates) {
            upd_builder ub;
            memset(&ub, 0, sizeof ub);
            ub_append_from_hb(&ub, hb, 0);
            ub_finish(&ub, opt_updates);
        }
    }
    orc_hb_free(hb);
    return rc;
}

int orc_state_root_full(const uint8_t *acct_keys32, const orc_account *accts, uint64_t n_accounts,
                        const uint8_t *slot_keys32, const uint8_t *values32_be, const uint64_t *seg_offsets,
                        uint8_t root32[32], orc_updates *opt_account_updates,
                        orc_updates *opt_storage_updates, int threads) {
    uint8_t *roots = (uint8_t *)malloc(32 * (n_accounts ? n_accounts : 1));
    int rc = orc_storage_roots(slot_keys32, values32_be, seg_offsets, n_accounts, roots, opt_storage_updates,
                               threads);
    if (rc == 0) rc = orc_state_root(acct_keys32, accts, roots, n_accounts, root32, opt_account_updates);
    free(roots);
    return rc;
}

/* ------------------------------------------------------------------ independent recursive trie root */
typedef struct {
    const uint8_t *keys, *values;
    const uint64_t *voff;
} rec_ctx;

static uint8_t nibble_at(const uint8_t *k, unsigned i) { return (i & 1) ? (k[i >> 1] & 15) : (k[i >> 1] >> 4); }

static size_t rec_put_len_prefix(uint8_t *out, size_t len, uint8_t short_base, uint8_t long_base) {
    if (len < 56) {
        out[0] = (uint8_t)(short_base + len);
        return 1;
    }
    int nb = 0;
    uint8_t t[8];
    for (size_t l = len; l; l >>= 8) t[nb++] = (uint8_t)l;
    out[0] = (uint8_t)(long_base + nb);
    for (int i = 0; i < nb; i++) out[1 + i] = t[nb - 1 - i];
    return 1 + (size_t)nb;
}

static size_t rec_put_bytes(uint8_t *out, const uint8_t *d, size_t n) {
    if (n == 1 && d[0] < 0x80) {
        out[0] = d[0];
        return 1;
    }
    size_t h = rec_put_len_prefix(out, n, 0x80, 0xb7);
    memcpy(out + h, d, n);
    return h + n;
}

static size_t rec_compact_path(const uint8_t *key, unsigned from, unsigned to, int leaf, uint8_t *out) {
    unsigned n = to - from;
    size_t o = 0;
    uint8_t first = (uint8_t)((leaf ? 2 : 0) << 4);
    unsigned i = from;
    if (n & 1) {
        first |= (uint8_t)(0x10 | nibble_at(key, i));
        i++;
    }
    out[o++] = first;
    for (; i < to; i += 2) out[o++] = (uint8_t)((nibble_at(key, i) << 4) | nibble_at(key, i + 1));
    return o;
}

/* encode node for keys [lo,hi) all sharing the first `depth` nibbles; returns malloc'd rlp */
static uint8_t *rec_node(const rec_ctx *c, uint64_t lo, uint64_t hi, unsigned depth, size_t *out_len);

/* child reference: rlp if short, else hash string */
static size_t rec_ref(const rec_ctx *c, uint64_t lo, uint64_t hi, unsigned depth, uint8_t out[33]) {
    size_t l;
    uint8_t *r = rec_node(c, lo, hi, depth, &l);
    size_t n;
    if (l < 32) {
        memcpy(out, r, l);
        n = l;
    } else {
        out[0] = 0xa0;
        orc_keccak256(r, l, out + 1);
        n = 33;
    }
    free(r);
    return n;
}

static uint8_t *rec_node(const rec_ctx *c, uint64_t lo, uint64_t hi, unsigned depth, size_t *out_len) {
    const uint8_t *k0 = c->keys + 32 * lo;
    if (hi - lo == 1) {
        uint8_t path[40];
        size_t pl = rec_compact_path(k0, depth, 64, 1, path);
        const uint8_t *v = c->values + c->voff[lo];
        size_t vl = (size_t)(c->voff[lo + 1] - c->voff[lo]);
        uint8_t *body = (uint8_t *)malloc(pl + vl + 32);
        size_t b = rec_put_bytes(body, path, pl);
        b += rec_put_bytes(body + b, v, vl);
        uint8_t *out = (uint8_t *)malloc(b + 16);
        size_t h = rec_put_len_prefix(out, b, 0xc0, 0xf7);
        memcpy(out + h, body, b);
        free(body);
        *out_len = h + b;
        return out;
    }
    /* shared prefix beyond depth? first and last key bound it (sorted input) */
    const uint8_t *kl = c->keys + 32 * (hi - 1);
    unsigned shared = depth;
    while (shared < 64 && nibble_at(k0, shared) == nibble_at(kl, shared)) shared++;
    if (shared > depth) {
        uint8_t path[40], child[33];
        size_t pl = rec_compact_path(k0, depth, shared, 0, path);
        size_t cl = rec_ref(c, lo, hi, shared, child);
        uint8_t body[80];
        size_t b = rec_put_bytes(body, path, pl);
        memcpy(body + b, child, cl);
        b += cl;
        uint8_t *out = (uint8_t *)malloc(b + 16);
        size_t h = rec_put_len_prefix(out, b, 0xc0, 0xf7);
        memcpy(out + h, body, b);
        *out_len = h + b;
        return out;
    }
    uint8_t body[16 * 33 + 1];
    size_t b = 0;
    uint64_t i = lo;
    for (unsigned nib = 0; nib < 16; nib++) {
        uint64_t j = i;
        while (j < hi && nibble_at(c->keys + 32 * j, depth) == nib) j++;
        if (j == i) body[b++] = 0x80;
        else b += rec_ref(c, i, j, depth + 1, body + b);
        i = j;
    }
    body[b++] = 0x80;
    uint8_t *out = (uint8_t *)malloc(b + 16);
    size_t h = rec_put_len_prefix(out, b, 0xc0, 0xf7);
    memcpy(out + h, body, b);
    *out_len = h + b;
    return out;
}

int orc_trie_root_recursive(const uint8_t *keys32, const uint8_t *values, const uint64_t *value_offsets,
                            uint64_t n, uint8_t root32[32]) {
    if (n == 0) {
        memcpy(root32, EMPTY_ROOT, 32);
        return 0;
    }
    for (uint64_t i = 1; i < n; i++)
        if (memcmp(keys32 + 32 * (i - 1), keys32 + 32 * i, 32) >= 0) return -1;
    rec_ctx c = {keys32, values, value_offsets};
    size_t l;
    uint8_t *r = rec_node(&c, 0, n, 0, &l);
    orc_keccak256(r, l, root32);
    free(r);
    return 0;
}
