/*
 * hash_builder.c — restatement of alloy-trie 0.9.5 `HashBuilder` (external crate, not under
 * /root/reference; pinned in Cargo.lock:986-987) and its node encodings.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle.h).
 *
 * Reference anchors for behaviour:
 *   call sites   crates/trie/trie/src/trie.rs:240,251,309,432,663,668,698
 *   field set    crates/trie/common/src/hash_builder/state.rs:15-48
 *   node RLPs    crates/trie/db/tests/proof.rs:50-103,146-151 (byte-exact vectors)
 *   stored nodes crates/trie/db/tests/trie.rs:456-477,791-805 (mask vectors)
 * Algorithm text: SURVEY.md Appendix A.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

extern uint64_t orc__keccak_f_count(void);
extern uint64_t orc__bytes_hashed(void);
extern void orc__keccak_counters_reset(void);

static __thread orc_stats tl_stats;

void orc_stats_reset(void) {
    memset(&tl_stats, 0, sizeof tl_stats);
    orc__keccak_counters_reset();
}
void orc_stats_get(orc_stats *out) {
    *out = tl_stats;
    out->keccak_f += orc__keccak_f_count();
    out->rlp_bytes_hashed += orc__bytes_hashed();
}
void orc__stats_merge(const orc_stats *s) {
    tl_stats.leaves += s->leaves;
    tl_stats.branch_nodes += s->branch_nodes;
    tl_stats.extension_nodes += s->extension_nodes;
    tl_stats.hashed_nodes += s->hashed_nodes;
    tl_stats.keccak_f += s->keccak_f; /* counts carried over from worker threads */
    tl_stats.rlp_bytes_hashed += s->rlp_bytes_hashed;
}

/* ------------------------------------------------------------------ RLP helpers */
/* append RLP string header+data of `len` bytes at out, return bytes written */
static size_t rlp_put_str(uint8_t *out, const uint8_t *data, size_t len) {
    if (len == 1 && data[0] < 0x80) {
        out[0] = data[0];
        return 1;
    }
    size_t h;
    if (len < 56) {
        out[0] = (uint8_t)(0x80 + len);
        h = 1;
    } else {
        uint8_t tmp[8];
        int nb = 0;
        for (size_t l = len; l; l >>= 8) tmp[nb++] = (uint8_t)l;
        out[0] = (uint8_t)(0xb7 + nb);
        for (int i = 0; i < nb; i++) out[1 + i] = tmp[nb - 1 - i];
        h = 1 + (size_t)nb;
    }
    memcpy(out + h, data, len);
    return h + len;
}
static size_t rlp_str_len(const uint8_t *data, size_t len) {
    if (len == 1 && data[0] < 0x80) return 1;
    if (len < 56) return 1 + len;
    size_t nb = 0;
    for (size_t l = len; l; l >>= 8) nb++;
    return 1 + nb + len;
}
static size_t rlp_put_list_header(uint8_t *out, size_t payload) {
    if (payload < 56) {
        out[0] = (uint8_t)(0xc0 + payload);
        return 1;
    }
    uint8_t tmp[8];
    int nb = 0;
    for (size_t l = payload; l; l >>= 8) tmp[nb++] = (uint8_t)l;
    out[0] = (uint8_t)(0xf7 + nb);
    for (int i = 0; i < nb; i++) out[1 + i] = tmp[nb - 1 - i];
    return 1 + (size_t)nb;
}

/* hex-prefix encoding of a nibble path; returns byte length (1 + n/2) */
static size_t hex_prefix(const uint8_t *nib, size_t n, int leaf, uint8_t *out) {
    size_t o = 0, i = 0;
    uint8_t flag = leaf ? 0x20 : 0x00;
    if (n & 1) {
        out[o++] = (uint8_t)(flag | 0x10 | nib[0]);
        i = 1;
    } else {
        out[o++] = flag;
    }
    for (; i < n; i += 2) out[o++] = (uint8_t)((nib[i] << 4) | nib[i + 1]);
    return o;
}

size_t orc_encode_u256(const uint8_t v[32], uint8_t out[33]) {
    int i = 0;
    while (i < 32 && v[i] == 0) i++;
    if (i == 32) {
        out[0] = 0x80;
        return 1;
    }
    return rlp_put_str(out, v + i, (size_t)(32 - i));
}

size_t orc_encode_trie_account(const orc_account *a, const uint8_t storage_root[32], uint8_t out[112]) {
    uint8_t payload[112];
    size_t p = 0;
    uint8_t nb[8];
    int n = 0;
    for (int i = 7; i >= 0; i--) {
        uint8_t b = (uint8_t)(a->nonce >> (8 * i));
        if (n || b) nb[n++] = b;
    }
    if (n == 0) payload[p++] = 0x80;
    else p += rlp_put_str(payload + p, nb, (size_t)n);
    p += orc_encode_u256(a->balance_be, payload + p);
    payload[p++] = 0xa0;
    memcpy(payload + p, storage_root, 32);
    p += 32;
    payload[p++] = 0xa0;
    memcpy(payload + p, a->code_hash, 32);
    p += 32;
    size_t h = rlp_put_list_header(out, p);
    memcpy(out + h, payload, p);
    return h + p;
}

/* ------------------------------------------------------------------ HashBuilder */
typedef struct {
    uint8_t len;
    uint8_t b[33];
} rlpnode;

typedef struct {
    size_t off, len;
} node_span;

struct orc_hb {
    uint8_t key[72];
    size_t key_len;
    int value_kind; /* 0 none, 1 bytes, 2 hash */
    uint8_t *value;
    size_t value_len, value_cap;
    rlpnode *stack;
    size_t sp, stack_cap;
    uint16_t state_masks[72], tree_masks[72], hash_masks[72];
    size_t n_state, n_th;
    int stored_in_database;
    int retain_updates;
    orc_branch_node *upd;
    size_t n_upd, upd_cap;
    int upd_sorted;
    /* optional retention of every emitted node RLP */
    int retain_nodes;
    uint8_t *node_bytes;
    size_t node_bytes_len, node_bytes_cap;
    node_span *nodes;
    size_t n_nodes, nodes_cap;
    orc_stats stats;
    uint8_t *rlp_buf;
    size_t rlp_cap;
};

orc_hb *orc_hb_new(int retain_updates) {
    orc_hb *h = (orc_hb *)calloc(1, sizeof *h);
    h->retain_updates = retain_updates;
    h->stack_cap = 128;
    h->stack = (rlpnode *)malloc(sizeof(rlpnode) * h->stack_cap);
    h->rlp_cap = 1024;
    h->rlp_buf = (uint8_t *)malloc(h->rlp_cap);
    return h;
}
void orc_hb_free(orc_hb *h) {
    if (!h) return;
    orc__stats_merge(&h->stats);
    free(h->value);
    free(h->stack);
    free(h->upd);
    free(h->node_bytes);
    free(h->nodes);
    free(h->rlp_buf);
    free(h);
}
void orc_hb_retain_nodes(orc_hb *h, int on) { h->retain_nodes = on; }
size_t orc_hb_nodes_len(const orc_hb *h) { return h->n_nodes; }
const uint8_t *orc_hb_node_at(const orc_hb *h, size_t i, size_t *len) {
    *len = h->nodes[i].len;
    return h->node_bytes + h->nodes[i].off;
}

static void ensure_rlp(orc_hb *h, size_t need) {
    if (need > h->rlp_cap) {
        while (h->rlp_cap < need) h->rlp_cap *= 2;
        h->rlp_buf = (uint8_t *)realloc(h->rlp_buf, h->rlp_cap);
    }
}

static void record_node(orc_hb *h, const uint8_t *rlp, size_t len) {
    if (!h->retain_nodes) return;
    if (h->node_bytes_len + len > h->node_bytes_cap) {
        h->node_bytes_cap = (h->node_bytes_cap ? h->node_bytes_cap * 2 : 4096) + len;
        h->node_bytes = (uint8_t *)realloc(h->node_bytes, h->node_bytes_cap);
    }
    if (h->n_nodes == h->nodes_cap) {
        h->nodes_cap = h->nodes_cap ? h->nodes_cap * 2 : 64;
        h->nodes = (node_span *)realloc(h->nodes, sizeof(node_span) * h->nodes_cap);
    }
    memcpy(h->node_bytes + h->node_bytes_len, rlp, len);
    h->nodes[h->n_nodes].off = h->node_bytes_len;
    h->nodes[h->n_nodes].len = len;
    h->n_nodes++;
    h->node_bytes_len += len;
}

/* RlpNode::from_rlp: inline if < 32 bytes, else 0xa0 || keccak256(rlp) */
static void push_rlp(orc_hb *h, const uint8_t *rlp, size_t len) {
    if (h->sp == h->stack_cap) {
        h->stack_cap *= 2;
        h->stack = (rlpnode *)realloc(h->stack, sizeof(rlpnode) * h->stack_cap);
    }
    rlpnode *n = &h->stack[h->sp++];
    record_node(h, rlp, len);
    if (len < 32) {
        n->len = (uint8_t)len;
        memcpy(n->b, rlp, len);
    } else {
        n->len = 33;
        n->b[0] = 0xa0;
        orc_keccak256(rlp, len, n->b + 1);
        h->stats.hashed_nodes++;
    }
}
static void push_hash(orc_hb *h, const uint8_t hash[32]) {
    if (h->sp == h->stack_cap) {
        h->stack_cap *= 2;
        h->stack = (rlpnode *)realloc(h->stack, sizeof(rlpnode) * h->stack_cap);
    }
    rlpnode *n = &h->stack[h->sp++];
    n->len = 33;
    n->b[0] = 0xa0;
    memcpy(n->b + 1, hash, 32);
}

static void resize_masks(orc_hb *h, size_t new_len) {
    for (size_t i = h->n_th; i < new_len; i++) h->tree_masks[i] = h->hash_masks[i] = 0;
    h->n_th = new_len;
}
static void resize_state(orc_hb *h, size_t new_len) {
    for (size_t i = h->n_state; i < new_len; i++) h->state_masks[i] = 0;
    h->n_state = new_len;
}

static void current_root(const orc_hb *h, uint8_t out[32]) {
    static const uint8_t empty_str = 0x80;
    if (h->sp == 0) {
        orc_keccak256(&empty_str, 1, out); /* EMPTY_ROOT_HASH */
        return;
    }
    const rlpnode *n = &h->stack[h->sp - 1];
    if (n->len == 33 && n->b[0] == 0xa0) memcpy(out, n->b + 1, 32);
    else orc_keccak256(n->b, n->len, out);
}

static void store_update(orc_hb *h, const uint8_t *path, size_t path_len, uint16_t st, uint16_t tm, uint16_t hm,
                         const uint8_t (*children)[32], int n_children, int with_root) {
    if (h->n_upd == h->upd_cap) {
        h->upd_cap = h->upd_cap ? h->upd_cap * 2 : 16;
        h->upd = (orc_branch_node *)realloc(h->upd, sizeof(orc_branch_node) * h->upd_cap);
    }
    orc_branch_node *u = &h->upd[h->n_upd++];
    memset(u, 0, sizeof *u);
    memcpy(u->path, path, path_len);
    u->path_len = (uint8_t)path_len;
    u->state_mask = st;
    u->tree_mask = tm;
    u->hash_mask = hm;
    u->n_hashes = (uint8_t)n_children;
    for (int i = 0; i < n_children; i++) memcpy(u->hashes[i], children[i], 32);
    if (with_root) {
        u->has_root_hash = 1;
        current_root(h, u->root_hash);
    }
    h->upd_sorted = 0;
}

static size_t common_prefix(const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    size_t n = na < nb ? na : nb, i = 0;
    while (i < n && a[i] == b[i]) i++;
    return i;
}

static void hb_update(orc_hb *h, const uint8_t *succ, size_t succ_len) {
    int build_extensions = 0;
    uint8_t cur[72];
    size_t cur_len = h->key_len;
    memcpy(cur, h->key, cur_len);

    for (;;) {
        int preceding_exists = h->n_state != 0;
        size_t preceding_len = h->n_state ? h->n_state - 1 : 0;
        size_t cpl = common_prefix(succ, succ_len, cur, cur_len);
        size_t len = preceding_len > cpl ? preceding_len : cpl;
        /* assert(len < cur_len) */
        if (len >= cur_len) abort();

        uint8_t extra = cur[len];
        if (h->n_state <= len) resize_state(h, len + 1);
        h->state_masks[len] |= (uint16_t)(1u << extra);

        if (h->n_th < cur_len) resize_masks(h, cur_len);

        size_t len_from = len;
        if (succ_len != 0 || preceding_exists) len_from += 1;
        const uint8_t *short_key = cur + len_from;
        size_t short_len = cur_len - len_from;

        if (!build_extensions) {
            if (h->value_kind == 1) {
                /* LeafNodeRef::rlp = list[ str(hp(short,leaf)), str(value) ] */
                uint8_t hp[40];
                size_t hpl = hex_prefix(short_key, short_len, 1, hp);
                size_t payload = rlp_str_len(hp, hpl) + rlp_str_len(h->value, h->value_len);
                ensure_rlp(h, payload + 16);
                size_t o = rlp_put_list_header(h->rlp_buf, payload);
                o += rlp_put_str(h->rlp_buf + o, hp, hpl);
                o += rlp_put_str(h->rlp_buf + o, h->value, h->value_len);
                push_rlp(h, h->rlp_buf, o);
                h->stats.leaves++;
            } else {
                push_hash(h, h->value);
                if (h->stored_in_database) h->tree_masks[cur_len - 1] |= (uint16_t)(1u << cur[cur_len - 1]);
                h->hash_masks[cur_len - 1] |= (uint16_t)(1u << cur[cur_len - 1]);
                build_extensions = 1;
            }
        }

        if (build_extensions && short_len != 0) {
            /* update_masks(current, len_from) */
            if (len_from > 0) {
                uint16_t flag = (uint16_t)(1u << cur[len_from - 1]);
                h->hash_masks[len_from - 1] &= (uint16_t)~flag;
                if (h->tree_masks[cur_len - 1] != 0) h->tree_masks[len_from - 1] |= flag;
            }
            rlpnode child = h->stack[--h->sp];
            uint8_t hp[40];
            size_t hpl = hex_prefix(short_key, short_len, 0, hp);
            size_t payload = rlp_str_len(hp, hpl) + child.len;
            ensure_rlp(h, payload + 16);
            size_t o = rlp_put_list_header(h->rlp_buf, payload);
            o += rlp_put_str(h->rlp_buf + o, hp, hpl);
            memcpy(h->rlp_buf + o, child.b, child.len);
            o += child.len;
            push_rlp(h, h->rlp_buf, o);
            h->stats.extension_nodes++;
            resize_masks(h, len_from);
        }

        if (preceding_len <= cpl && succ_len != 0) return;

        if (succ_len != 0 || preceding_exists) {
            /* push_branch_node(current, len) */
            uint16_t st = h->state_masks[len];
            uint16_t hm = h->hash_masks[len];
            int nchild = __builtin_popcount(st);
            size_t first = h->sp - (size_t)nchild;
            uint8_t children[16][32];
            int n_children = 0;
            size_t payload = 1; /* value slot 0x80 */
            {
                size_t si = first;
                for (int nib = 0; nib < 16; nib++) {
                    if (st & (1u << nib)) {
                        const rlpnode *c = &h->stack[si++];
                        payload += c->len;
                        if (h->retain_updates && (hm & (1u << nib))) {
                            /* BranchNodeRef::child_hashes takes child[1..] and requires the 33-byte form;
                             * an inline branch child under a set hash bit is unreachable for
                             * keccak-derived keys — refuse rather than emulate. */
                            if (c->len != 33) abort();
                            memcpy(children[n_children++], c->b + 1, 32);
                        }
                    } else {
                        payload += 1;
                    }
                }
            }
            ensure_rlp(h, payload + 16);
            size_t o = rlp_put_list_header(h->rlp_buf, payload);
            {
                size_t si = first;
                for (int nib = 0; nib < 16; nib++) {
                    if (st & (1u << nib)) {
                        const rlpnode *c = &h->stack[si++];
                        memcpy(h->rlp_buf + o, c->b, c->len);
                        o += c->len;
                    } else {
                        h->rlp_buf[o++] = 0x80;
                    }
                }
                h->rlp_buf[o++] = 0x80;
            }
            h->sp = first;
            push_rlp(h, h->rlp_buf, o);
            h->stats.branch_nodes++;

            /* store_branch_node(current, len, children) */
            if (len > 0) h->hash_masks[len - 1] |= (uint16_t)(1u << cur[len - 1]);
            int store = h->tree_masks[len] != 0 || h->hash_masks[len] != 0;
            if (store) {
                if (len > 0) h->tree_masks[len - 1] |= (uint16_t)(1u << cur[len - 1]);
                if (h->retain_updates)
                    store_update(h, cur, len, h->state_masks[len], h->tree_masks[len], h->hash_masks[len],
                                 (const uint8_t(*)[32])children, n_children, len == 0);
            }
        }

        /* state_masks.resize(len); resize_masks(len) */
        if (h->n_state > len) h->n_state = len;
        else resize_state(h, len);
        if (h->n_th > len) h->n_th = len;
        else resize_masks(h, len);

        if (preceding_len == 0) return;

        cur_len = preceding_len;
        while (h->n_state > 0 && h->state_masks[h->n_state - 1] == 0) h->n_state--;
        build_extensions = 1;
    }
}

static int nib_cmp(const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    size_t n = na < nb ? na : nb;
    int c = memcmp(a, b, n);
    if (c) return c;
    return na < nb ? -1 : (na > nb ? 1 : 0);
}

static void set_value(orc_hb *h, int kind, const uint8_t *v, size_t len) {
    if (len > h->value_cap) {
        h->value_cap = len + 64;
        h->value = (uint8_t *)realloc(h->value, h->value_cap);
    }
    if (len) memcpy(h->value, v, len);
    h->value_len = len;
    h->value_kind = kind;
}

int orc_hb_add_leaf(orc_hb *h, const uint8_t *key, size_t key_len, const uint8_t *value, size_t vlen) {
    if (key_len > 64) return -1;
    if (nib_cmp(key, key_len, h->key, h->key_len) <= 0) return -1;
    if (h->key_len != 0) hb_update(h, key, key_len);
    memcpy(h->key, key, key_len);
    h->key_len = key_len;
    set_value(h, 1, value, vlen);
    return 0;
}

int orc_hb_add_branch(orc_hb *h, const uint8_t *key, size_t key_len, const uint8_t hash[32],
                      int stored_in_database) {
    if (key_len > 64) return -1;
    if (!(nib_cmp(key, key_len, h->key, h->key_len) > 0 || (h->key_len == 0 && key_len == 0))) return -1;
    if (h->key_len != 0) hb_update(h, key, key_len);
    else if (key_len == 0) push_hash(h, hash);
    memcpy(h->key, key, key_len);
    h->key_len = key_len;
    set_value(h, 2, hash, 32);
    h->stored_in_database = stored_in_database;
    return 0;
}

void orc_hb_root(orc_hb *h, uint8_t out[32]) {
    if (h->key_len != 0) {
        hb_update(h, NULL, 0);
        h->key_len = 0;
        h->value_kind = 0;
        h->value_len = 0;
    }
    current_root(h, out);
}

static int upd_cmp(const void *a, const void *b) {
    const orc_branch_node *x = (const orc_branch_node *)a, *y = (const orc_branch_node *)b;
    return nib_cmp(x->path, x->path_len, y->path, y->path_len);
}

static void sort_updates(orc_hb *h) {
    if (h->upd_sorted) return;
    /* map semantics: a later insert for the same path wins; mergesort keeps insertion order for ties */
    if (h->n_upd > 1) {
        size_t n = h->n_upd;
        /* stable merge sort on an index array (qsort is not stable) */
        size_t *idx = (size_t *)malloc(sizeof(size_t) * n), *tmp = (size_t *)malloc(sizeof(size_t) * n);
        for (size_t i = 0; i < n; i++) idx[i] = i;
        for (size_t w = 1; w < n; w *= 2) {
            for (size_t lo = 0; lo < n; lo += 2 * w) {
                size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
                size_t i = lo, j = mid, k = lo;
                while (i < mid && j < hi) {
                    if (upd_cmp(&h->upd[idx[j]], &h->upd[idx[i]]) < 0) tmp[k++] = idx[j++];
                    else tmp[k++] = idx[i++];
                }
                while (i < mid) tmp[k++] = idx[i++];
                while (j < hi) tmp[k++] = idx[j++];
            }
            size_t *t = idx; idx = tmp; tmp = t;
        }
        orc_branch_node *out = (orc_branch_node *)malloc(sizeof(orc_branch_node) * n);
        size_t m = 0;
        for (size_t i = 0; i < n; i++) {
            const orc_branch_node *c = &h->upd[idx[i]];
            if (m && upd_cmp(&out[m - 1], c) == 0) out[m - 1] = *c; /* later wins */
            else out[m++] = *c;
        }
        free(idx);
        free(tmp);
        free(h->upd);
        h->upd = out;
        h->n_upd = m;
        h->upd_cap = n;
    }
    h->upd_sorted = 1;
}

size_t orc_hb_updates_len(const orc_hb *h) {
    sort_updates((orc_hb *)h);
    return h->n_upd;
}
const orc_branch_node *orc_hb_update_at(orc_hb *h, size_t i) {
    sort_updates(h);
    return &h->upd[i];
}
