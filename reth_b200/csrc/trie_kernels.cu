// trie_kernels.cu — level-synchronous Merkle-Patricia-Trie commitment on sm_100a.
//
// Replaces the serial stack machine of alloy-trie's HashBuilder (driven by StateRoot::calculate,
// crates/trie/trie/src/trie.rs:247-309, and StorageRoot::calculate, :659-698) by a data-parallel
// formulation over the SORTED leaf array ("data-parallel reading"):
//
//   gap g (1..n-1) sits between leaf g-1 and leaf g;  Lp[g] = common-prefix length of the two keys in
//   nibbles (0xFF at a trie boundary).  A branch node at depth d is a maximal run of depth-d gaps with no
//   shallower gap in between; its children are the leaves / deeper branches between consecutive gaps.
//   Sorting the gaps by depth therefore yields, for every level, the list of branch nodes (CSR over gaps),
//   and levels can be hashed deepest-first with one thread per node.  S[l] / E[r] map a leaf position to the
//   frontier item that currently starts / ends there, so a node finds its <=16 children in O(1).
//
// Every node is RLP-encoded into a per-thread shared-memory strip (word-transposed: conflict-free for the
// absorb loop) and hashed with a register-resident Keccak-f[1600].  A forest of tries (all storage tries of a
// block / of the whole state) is processed in the same launches: segment boundaries are just gaps with
// Lp = 0xFF.
#include "keccak_f1600.cuh"
#include "trie_kernels.h"
#include <mutex>
#include <unordered_map>

namespace b200 {

// ------------------------------------------------------------------------------------------------ sponge strip
// Thread-private byte stream in shared memory.  Word w of thread t lives at base[w*BLOCK + t].
template <int BLOCK>
struct Strip {
    uint32_t *w;
    uint32_t prev;  // pending bytes are the top `nb` bytes of prev, stream order low->high
    uint32_t nb;
    uint32_t nw;
    __device__ __forceinline__ void init(uint32_t *smem) {
        w = smem + threadIdx.x;
        prev = 0;
        nb = 0;
        nw = 0;
    }
    __device__ __forceinline__ void byte(uint32_t b) {
        prev = (prev >> 8) | (b << 24);
        if (++nb == 4) {
            w[nw * BLOCK] = prev;
            nw++;
            nb = 0;
        }
    }
    // 4 stream bytes given as a little-endian word
    __device__ __forceinline__ void word(uint32_t x) {
        w[nw * BLOCK] = __funnelshift_rc(prev, x, 32 - 8 * nb);
        nw++;
        prev = x;
    }
    __device__ __forceinline__ void words8(const uint32_t (&x)[8]) {
#pragma unroll
        for (int i = 0; i < 8; i++) word(x[i]);
    }
    // `cnt` empty-slot markers (0x80), four at a time whatever the current byte alignment
    __device__ __forceinline__ void fill80(uint32_t cnt) {
        while (cnt >= 4) {
            word(0x80808080u);
            cnt -= 4;
        }
        while (cnt--) byte(0x80);
    }
    // bytes [b0, 32) of a 32-byte string held as 8 little-endian words
    __device__ __forceinline__ void tail32(const uint32_t (&x)[8], uint32_t b0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (4u * i >= b0) {
                word(x[i]);
            } else if (4u * i + 3 >= b0) {
                for (uint32_t b = b0 - 4u * i; b < 4; b++) byte((x[i] >> (8 * b)) & 0xff);
            }
        }
    }
    __device__ __forceinline__ uint32_t length() const { return nw * 4 + nb; }
    __device__ __forceinline__ uint32_t read_word(uint32_t i) const { return w[i * BLOCK]; }
    // Keccak pad10*1 to a multiple of the 136-byte rate; returns the number of rate blocks.
    __device__ __forceinline__ uint32_t finish() {
        uint32_t blocks = length() / 136 + 1;
        uint32_t total_words = blocks * 34;
        byte(0x01);
        while (nb != 0) byte(0);
        while (nw < total_words) {
            w[nw * BLOCK] = 0;
            nw++;
        }
        w[(total_words - 1) * BLOCK] |= 0x80000000u;
        return blocks;
    }
    __device__ __forceinline__ void reset() {
        prev = 0;
        nb = 0;
        nw = 0;
    }
};

// keccak256 of the finished strip -> 8 little-endian digest words
template <int BLOCK>
__device__ __forceinline__ void strip_keccak(const Strip<BLOCK> &s, uint32_t blocks, uint32_t (&dig)[8]) {
    uint64_t a[25];
#pragma unroll
    for (int l = 0; l < 25; l++) a[l] = 0;
    uint32_t base = 0;
    for (uint32_t b = 0; b + 1 < blocks; b++) {
#pragma unroll
        for (int l = 0; l < 17; l++)
            a[l] ^= ((uint64_t)s.read_word(base + 2 * l + 1) << 32) | s.read_word(base + 2 * l);
        keccak_f1600(a);
        base += 34;
    }
#pragma unroll
    for (int l = 0; l < 17; l++) a[l] ^= ((uint64_t)s.read_word(base + 2 * l + 1) << 32) | s.read_word(base + 2 * l);
    keccak_f1600_final(a);
#pragma unroll
    for (int l = 0; l < 4; l++) {
        dig[2 * l] = (uint32_t)a[l];
        dig[2 * l + 1] = (uint32_t)(a[l] >> 32);
    }
}

// ------------------------------------------------------------------------------------------------ helpers
static __device__ __forceinline__ void load32(const uint8_t *p, uint32_t (&x)[8]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
static __device__ __forceinline__ void load32_nc(const uint8_t *p, uint32_t (&x)[8]) {
    // plain (coherent) loads: data written by earlier kernels of the same build
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
static __device__ __forceinline__ void store32(uint8_t *p, const uint32_t (&x)[8]) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(x[0], x[1], x[2], x[3]);
    q[1] = make_uint4(x[4], x[5], x[6], x[7]);
}
// nibble i (0 = most significant) of a 32-byte big-endian key held as LE words
static __device__ __forceinline__ uint32_t key_nibble_mem(const uint8_t *key, uint32_t i) {
    uint32_t b = key[i >> 1];
    return (i & 1) ? (b & 15) : (b >> 4);
}
static __device__ __forceinline__ int depth_of(uint8_t lp) { return lp == 0xFF ? -1 : (int)lp; }

// number of leading zero BYTES of a 32-byte big-endian integer held as LE words (32 if zero)
static __device__ __forceinline__ uint32_t leading_zero_bytes(const uint32_t (&x)[8]) {
    uint32_t z = 32;
#pragma unroll
    for (int i = 7; i >= 0; i--)
        if (x[i] != 0) z = 4u * i + ((__ffs(x[i]) - 1) >> 3);
    return z;
}
static __device__ __forceinline__ uint32_t byte_at(const uint32_t (&x)[8], uint32_t j) {
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if ((j >> 2) == (uint32_t)i) w = x[i];
    return (w >> (8 * (j & 3))) & 0xff;
}

// RLP list header for a payload < 65536 bytes
template <class W>
static __device__ __forceinline__ void put_list_header(W &s, uint32_t payload) {
    if (payload < 56) {
        s.byte(0xc0 + payload);
    } else if (payload < 256) {
        s.byte(0xf8);
        s.byte(payload);
    } else {
        s.byte(0xf9);
        s.byte(payload >> 8);
        s.byte(payload & 0xff);
    }
}
static __device__ __forceinline__ uint32_t list_header_len(uint32_t payload) {
    return payload < 56 ? 1 : (payload < 256 ? 2 : 3);
}

// ------------------------------------------------------------------------------------------------ structure pass
// Lp[g] for every gap; nibs[g] = (nibble of left key at Lp) << 4 | nibble of right key.  Boundaries were
// pre-marked with 0xFF by mark_boundaries_kernel and are left alone.
__global__ void lcp_kernel(const uint8_t *__restrict__ keys, uint64_t n, uint8_t *__restrict__ Lp,
                           uint8_t *__restrict__ nibs, int *__restrict__ err) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) {
        Lp[0] = 0xFF;
        Lp[n] = 0xFF;
    }
    if (g == 0 || g >= n) return;
    if (Lp[g] == 0xFF) return;
    uint32_t a[8], b[8];
    load32(keys + 32 * (g - 1), a);
    load32(keys + 32 * g, b);
    uint32_t lcp = 64, na = 0, nbb = 0;
    bool ascending = false;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        uint32_t x = __byte_perm(a[i], 0, 0x0123), y = __byte_perm(b[i], 0, 0x0123);  // big-endian numeric
        uint32_t d = x ^ y;
        if (d != 0) {
            uint32_t nz = __clz(d) >> 2;
            lcp = 8u * i + nz;
            na = (x >> (28 - 4 * nz)) & 15;
            nbb = (y >> (28 - 4 * nz)) & 15;
            ascending = x < y;
        }
    }
    if (!ascending) {  // equal or descending keys inside one trie: flag it; later kernels of the build bail out
        atomicExch(err, B200_DEVERR_UNSORTED);
        Lp[g] = 0xFF;
        return;
    }
    Lp[g] = (uint8_t)lcp;
    nibs[g] = (uint8_t)((na << 4) | nbb);
}

__global__ void mark_boundaries_kernel(const uint64_t *__restrict__ seg_offsets, uint64_t n_segs, uint64_t n,
                                       uint8_t *__restrict__ Lp, int *__restrict__ err) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_segs) return;
    uint64_t o = seg_offsets[s];
    if (s == 0 && o != 0) atomicExch(err, B200_DEVERR_BAD_OFFSETS);
    if (s == n_segs && o != n) atomicExch(err, B200_DEVERR_BAD_OFFSETS);
    if (s < n_segs && seg_offsets[s + 1] < o) atomicExch(err, B200_DEVERR_BAD_OFFSETS);
    if (o > 0 && o < n) Lp[o] = 0xFF;
}

__global__ void iota_kernel(uint32_t *__restrict__ out, uint64_t n, uint32_t first) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first + (uint32_t)i;
}

// bucket_off[d] = first sorted position whose depth >= d, d = 0..64 (64 => number of real gaps)
__global__ void bucket_offsets_kernel(const uint8_t *__restrict__ depth_sorted, uint64_t G,
                                      uint32_t *__restrict__ bucket_off) {
    uint32_t d = threadIdx.x;
    if (d > 64) return;
    uint64_t lo = 0, hi = G;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (depth_sorted[mid] < d) lo = mid + 1;
        else hi = mid;
    }
    bucket_off[d] = (uint32_t)lo;
}

// head[j] = 1 iff sorted gap j starts a new branch node
__global__ void head_flags_kernel(const uint8_t *__restrict__ keys, const uint8_t *__restrict__ depth_sorted,
                                  const uint32_t *__restrict__ gap_sorted, const uint32_t *__restrict__ bound_rank,
                                  const uint32_t *__restrict__ G_real_p, uint64_t G, uint8_t *__restrict__ head) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= G || j >= *G_real_p) return;  // boundary gaps (0xFF) sort behind the real ones
    uint32_t d = depth_sorted[j];
    bool h = true;
    if (j > 0 && depth_sorted[j - 1] == d) {
        uint32_t gp = gap_sorted[j - 1], g = gap_sorted[j];
        bool same_seg = bound_rank == nullptr || bound_rank[gp] == bound_rank[g];
        if (same_seg) {
            if (gp + 1 == g) {
                h = false;  // the single leaf gp is a child between the two gaps
            } else {
                // leaves gp .. g-1 form one child iff they share > d nibbles
                uint32_t a[8], b[8];
                load32(keys + 32 * (uint64_t)gp, a);
                load32(keys + 32 * (uint64_t)(g - 1), b);
                uint32_t lcp = 64;
#pragma unroll
                for (int i = 7; i >= 0; i--) {
                    uint32_t x = __byte_perm(a[i] ^ b[i], 0, 0x0123);
                    if (x != 0) lcp = 8u * i + (__clz(x) >> 2);
                }
                h = !(lcp > d);
            }
        }
    }
    head[j] = h ? 1 : 0;
}

// level_lo[d] = first node id whose depth >= d  (node_start is ascending in sorted-gap position)
__global__ void level_ranges_kernel(const uint32_t *__restrict__ node_start, const uint32_t *__restrict__ n_nodes_p,
                                    const uint32_t *__restrict__ bucket_off, uint32_t *__restrict__ level_lo,
                                    uint32_t *__restrict__ node_start_sentinel_target) {
    uint32_t d = threadIdx.x;
    uint32_t B = *n_nodes_p;
    if (d == 0) node_start_sentinel_target[B] = bucket_off[64];
    if (d > 64) return;
    uint32_t target = bucket_off[d];
    uint32_t lo = 0, hi = B;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (node_start[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    level_lo[d] = lo;
    if (d == 0) level_lo[65] = B;
}

// ------------------------------------------------------------------------------------------------ leaves
// One leaf -> RlpNode.  ACCOUNT: value = rlp(TrieAccount) built on the fly (crates/trie/trie/src/trie.rs:429-432,
// crates/trie/common/src/account.rs:16-31); else value = rlp(U256) (trie.rs:668-671).
// Encodes with parent depth `pd` (suffix starts at nibble pd+1); `force_hash` for a leaf that is a whole trie.
template <class W, bool ACCOUNT>
__device__ __forceinline__ uint32_t encode_leaf(W &s, const uint32_t (&k)[8], int pd, const uint8_t *val_ptr,
                                                const uint8_t *sroot_ptr, int *err) {
    uint32_t p = (uint32_t)(pd + 1);  // first suffix nibble
    uint32_t m = 64 - p;              // suffix nibbles (1..64)
    uint32_t hp_len = 1 + (m >> 1);
    uint32_t hp_str = hp_len == 1 ? 1 : 1 + hp_len;
    uint32_t first = (p & 1) ? (0x30u | (byte_at(k, p >> 1) & 15)) : 0x20u;
    uint32_t b0 = (p + 1) >> 1;  // key bytes [b0,32) follow the flag byte

    if (!ACCOUNT) {
        uint32_t v[8];
        load32(val_ptr, v);
        uint32_t z = leading_zero_bytes(v);
        if (z == 32) {
            atomicExch(err, B200_DEVERR_ZERO_VALUE);
            z = 31;
        }
        uint32_t vb = 32 - z;
        uint32_t fb = byte_at(v, z);
        bool single = vb == 1 && fb < 0x80;
        uint32_t rlp_v = single ? 1 : 1 + vb;     // alloy_rlp::encode_fixed_size(U256)
        uint32_t val_str = single ? 1 : 1 + rlp_v;  // ... wrapped as an RLP string inside the leaf
        uint32_t payload = hp_str + val_str;
        put_list_header(s, payload);
        if (hp_len > 1) s.byte(0x80 + hp_len);
        s.byte(first);
        s.tail32(k, b0);
        if (single) {
            s.byte(fb);
        } else {
            s.byte(0x80 + rlp_v);
            s.byte(0x80 + vb);
            s.tail32(v, z);
        }
        return list_header_len(payload) + payload;
    } else {
        const uint64_t *ap = reinterpret_cast<const uint64_t *>(val_ptr);
        uint64_t nonce = __ldg(ap);
        uint32_t bal[8], code[8], sroot[8];
        {
            const uint2 *q = reinterpret_cast<const uint2 *>(val_ptr + 8);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint2 t = __ldg(q + i);
                bal[2 * i] = t.x;
                bal[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint2 t = __ldg(q + 4 + i);
                code[2 * i] = t.x;
                code[2 * i + 1] = t.y;
            }
        }
        if (sroot_ptr) {
            load32_nc(sroot_ptr, sroot);
        } else {  // EMPTY_ROOT_HASH
            sroot[0] = 0x171fe856u; sroot[1] = 0xa655cc1bu; sroot[2] = 0xe64583ffu; sroot[3] = 0x6ef8c092u;
            sroot[4] = 0x1be0485bu; sroot[5] = 0xc0ad6c99u; sroot[6] = 0xb52f6201u; sroot[7] = 0x21b463e3u;
        }
        uint32_t nn = nonce == 0 ? 0 : (8 - (__clzll((long long)nonce) >> 3));
        uint32_t nonce_rlp = (nn == 0 || (nn == 1 && nonce < 0x80)) ? 1 : 1 + nn;
        uint32_t bz = leading_zero_bytes(bal);
        uint32_t bb = 32 - bz;
        uint32_t bfb = bb ? byte_at(bal, bz) : 0;
        uint32_t bal_rlp = (bb == 0 || (bb == 1 && bfb < 0x80)) ? 1 : 1 + bb;
        uint32_t inner_payload = nonce_rlp + bal_rlp + 66;  // >= 68: two-byte list header
        uint32_t inner_total = 2 + inner_payload;           // 70..110: two-byte string header
        uint32_t payload = hp_str + 2 + inner_total;        // 73..146: two-byte list header
        s.byte(0xf8);
        s.byte(payload);
        if (hp_len > 1) s.byte(0x80 + hp_len);
        s.byte(first);
        s.tail32(k, b0);
        s.byte(0xb8);
        s.byte(inner_total);
        s.byte(0xf8);
        s.byte(inner_payload);
        if (nn == 0) {
            s.byte(0x80);
        } else {
            if (nonce_rlp > 1) s.byte(0x80 + nn);
            for (int i = (int)nn - 1; i >= 0; i--) s.byte((uint32_t)(nonce >> (8 * i)) & 0xff);
        }
        if (bb == 0) {
            s.byte(0x80);
        } else if (bal_rlp == 1) {
            s.byte(bfb);
        } else {
            s.byte(0x80 + bb);
            s.tail32(bal, bz);
        }
        s.byte(0xa0);
        s.words8(sroot);
        s.byte(0xa0);
        s.words8(code);
        return 2 + payload;
    }
}

// strip -> (ref words, meta): hashed when >= 32 bytes or forced
template <int BLOCK>
__device__ __forceinline__ uint32_t strip_to_ref(Strip<BLOCK> &s, uint32_t len, bool force_hash, uint32_t (&ref)[8],
                                                 uint32_t &hashed) {
    if (len >= 32 || force_hash) {
        uint32_t blocks = s.finish();
        strip_keccak(s, blocks, ref);
        hashed++;
        return 0;  // meta: hashed
    }
    while (s.nb != 0) s.byte(0);
#pragma unroll
    for (int i = 0; i < 8; i++) ref[i] = (uint32_t)i < s.nw ? s.read_word(i) : 0;
    return len;  // meta: inline length 1..31
}

template <int BLOCK, bool ACCOUNT>
__global__ void __launch_bounds__(BLOCK) leaf_kernel(ForestDev f, const uint8_t *__restrict__ values,
                                                     const uint8_t *__restrict__ storage_roots) {
    extern __shared__ uint32_t smem[];
    if (*(volatile int *)f.err == B200_DEVERR_UNSORTED || *(volatile int *)f.err == B200_DEVERR_BAD_OFFSETS) return;
    Strip<BLOCK> s;
    uint32_t hashed = 0;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < f.n; i += step) {
        s.init(smem);
        uint32_t k[8];
        load32(f.keys + 32 * i, k);
        int pdl = depth_of(f.Lp[i]), pdr = depth_of(f.Lp[i + 1]);
        int pd = pdl > pdr ? pdl : pdr;
        const uint8_t *vp = ACCOUNT ? values + (uint64_t)sizeof(b200_account_dev) * i : values + 32 * i;
        const uint8_t *sp = (ACCOUNT && storage_roots) ? storage_roots + 32 * i : nullptr;
        uint32_t len = encode_leaf<Strip<BLOCK>, ACCOUNT>(s, k, pd, vp, sp, f.err);
        uint32_t ref[8];
        uint32_t meta = strip_to_ref(s, len, pd < 0, ref, hashed);
        store32(f.leaf_ref + 32 * i, ref);
        f.leaf_meta[i] = (uint8_t)meta;
        f.S[i] = (uint32_t)i;
        f.E[i] = (uint32_t)i;
    }
    // one atomic per warp
    for (int o = 16; o; o >>= 1) hashed += __shfl_xor_sync(0xffffffffu, hashed, o);
    if ((threadIdx.x & 31) == 0 && hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
}

// ------------------------------------------------------------------------------------------------ branches
struct ChildInfo {
    uint32_t id;    // < n: leaf, else n + node
    uint32_t nib;
    uint32_t meta;  // low 5 bits inline length (0 = hashed), META_EXT, META_STORED
};

__device__ __forceinline__ ChildInfo fetch_child(const ForestDev &f, uint32_t j0, uint32_t c) {
    ChildInfo ci;
    if (c == 0) {
        uint32_t g = f.gap_sorted[j0];
        ci.id = f.E[g - 1];
        ci.nib = f.nibs[g] >> 4;
    } else {
        uint32_t g = f.gap_sorted[j0 + c - 1];
        ci.id = f.S[g];
        ci.nib = f.nibs[g] & 15;
    }
    ci.meta = ci.id < f.n ? f.leaf_meta[ci.id] : f.node_meta[ci.id - (uint32_t)f.n];
    return ci;
}

// hex-prefix string of key nibbles [from, to) (extension flag), as an RLP string
template <class W>
__device__ __forceinline__ uint32_t put_ext_path(W &s, const uint8_t *key, uint32_t from, uint32_t to) {
    uint32_t m = to - from;
    uint32_t hp_len = 1 + (m >> 1);
    uint32_t i = from;
    uint32_t first = 0;
    if (m & 1) {
        first = 0x10u | key_nibble_mem(key, i);
        i++;
    }
    if (hp_len > 1) s.byte(0x80 + hp_len);
    s.byte(first);  // 0x00 or 0x1n: a lone byte < 0x80 is its own RLP
    for (; i < to; i += 2) s.byte((key_nibble_mem(key, i) << 4) | key_nibble_mem(key, i + 1));
    return hp_len == 1 ? 1 : 1 + hp_len;
}

// Builds branch node v (depth d) into the strip; returns RLP length and the node's masks / extent.
template <int BLOCK>
__device__ __forceinline__ uint32_t encode_branch(Strip<BLOCK> &s, const ForestDev &f, uint32_t j0, uint32_t k,
                                                  uint32_t &state_mask, uint32_t &tree_mask, uint32_t &hash_mask,
                                                  uint32_t &l, uint32_t &r) {
    // pass 1: lengths and masks
    uint32_t payload = 17;
    state_mask = tree_mask = hash_mask = 0;
    for (uint32_t c = 0; c <= k; c++) {
        ChildInfo ci = fetch_child(f, j0, c);
        uint32_t clen = (ci.meta & META_LEN) ? (ci.meta & META_LEN) : 33;
        payload += clen - 1;
        uint32_t bit = 1u << ci.nib;
        state_mask |= bit;
        if (ci.id >= f.n) {
            if (!(ci.meta & META_EXT)) {
                hash_mask |= bit;
                if ((ci.meta & META_LEN) && f.retain_updates) atomicExch(f.err, B200_DEVERR_INLINE_HASH_CHILD);
            }
            if (ci.meta & META_STORED) tree_mask |= bit;
        }
        if (c == 0) l = ci.id < f.n ? ci.id : f.node_l[ci.id - (uint32_t)f.n];
        if (c == k) r = ci.id < f.n ? ci.id : f.node_r[ci.id - (uint32_t)f.n];
    }
    // pass 2: bytes
    put_list_header(s, payload);
    uint32_t cur = 0;
    for (uint32_t c = 0; c <= k; c++) {
        ChildInfo ci = fetch_child(f, j0, c);
        for (; cur < ci.nib; cur++) s.byte(0x80);
        const uint8_t *rp = ci.id < f.n ? f.leaf_ref + 32 * (uint64_t)ci.id
                                        : f.node_ref + 32 * (uint64_t)(ci.id - (uint32_t)f.n);
        uint32_t ref[8];
        load32_nc(rp, ref);
        uint32_t clen = ci.meta & META_LEN;
        if (clen == 0) {
            s.byte(0xa0);
            s.words8(ref);
        } else {
            for (uint32_t b = 0; b < clen; b++) s.byte(byte_at(ref, b));
        }
        cur++;
    }
    for (; cur < 16; cur++) s.byte(0x80);
    s.byte(0x80);  // value slot
    return list_header_len(payload) + payload;
}

// Wraps `child` (ref words + inline length, 0 = hashed) into an extension over key nibbles [from,to).
template <class W>
__device__ __forceinline__ uint32_t encode_extension(W &s, const uint8_t *key, uint32_t from, uint32_t to,
                                                     const uint32_t (&child)[8], uint32_t child_inline_len) {
    uint32_t m = to - from;
    uint32_t hp_len = 1 + (m >> 1);
    uint32_t path_str = hp_len == 1 ? 1 : 1 + hp_len;
    uint32_t clen = child_inline_len ? child_inline_len : 33;
    uint32_t payload = path_str + clen;
    put_list_header(s, payload);
    put_ext_path(s, key, from, to);
    if (child_inline_len == 0) {
        s.byte(0xa0);
        s.words8(child);
    } else {
        for (uint32_t b = 0; b < child_inline_len; b++) s.byte(byte_at(child, b));
    }
    return list_header_len(payload) + payload;
}

// Class-specialised variant of encode_branch: at most MAXC children, every per-child quantity lives in registers
// and all the dependent global loads of a phase (gap -> S/E -> meta -> ref) are issued back to back for the
// whole node before any of them is consumed, so one thread keeps up to MAXC requests in flight.
template <int BLOCK, int MAXC, bool COHERENT = false>
__device__ __forceinline__ uint32_t encode_branch_u(Strip<BLOCK> &s, const ForestDev &f, uint32_t j0, uint32_t k,
                                                    uint32_t &state_mask, uint32_t &tree_mask, uint32_t &hash_mask,
                                                    uint32_t &l, uint32_t &r) {
    const uint32_t n = (uint32_t)f.n;
    uint32_t g[MAXC - 1];
#pragma unroll
    for (int c = 0; c < MAXC - 1; c++) g[c] = (uint32_t)c < k ? f.gap_sorted[j0 + c] : 0u;
    uint32_t id[MAXC], nm[MAXC];  // nm = nibble | meta << 8
    id[0] = f.E[g[0] - 1];
    nm[0] = f.nibs[g[0]] >> 4;
#pragma unroll
    for (int c = 1; c < MAXC; c++) {
        id[c] = 0;
        nm[c] = 0;
        if ((uint32_t)c <= k) {
            id[c] = f.S[g[c - 1]];
            nm[c] = f.nibs[g[c - 1]] & 15;
        }
    }
#pragma unroll
    for (int c = 0; c < MAXC; c++)
        if ((uint32_t)c <= k) {
            const uint8_t *mp = id[c] < n ? f.leaf_meta + id[c] : f.node_meta + (id[c] - n);
            nm[c] |= (uint32_t)(COHERENT ? __ldcg(mp) : *mp) << 8;
        }
    uint32_t payload = 17;
    state_mask = tree_mask = hash_mask = 0;
    uint32_t last = id[0];
#pragma unroll
    for (int c = 0; c < MAXC; c++) {
        if ((uint32_t)c <= k) {
            uint32_t meta = nm[c] >> 8;
            payload += ((meta & META_LEN) ? (meta & META_LEN) : 33u) - 1;
            uint32_t bit = 1u << (nm[c] & 15);
            state_mask |= bit;
            if (id[c] >= n) {
                if (!(meta & META_EXT)) {
                    hash_mask |= bit;
                    if ((meta & META_LEN) && f.retain_updates) atomicExch(f.err, B200_DEVERR_INLINE_HASH_CHILD);
                }
                if (meta & META_STORED) tree_mask |= bit;
            }
            last = id[c];
        }
    }
    l = id[0] < n ? id[0] : f.node_l[id[0] - n];
    r = last < n ? last : f.node_r[last - n];
    put_list_header(s, payload);
    uint32_t cur = 0;
#pragma unroll
    for (int c = 0; c < MAXC; c++) {
        if ((uint32_t)c <= k) {
            const uint8_t *rp = id[c] < n ? f.leaf_ref + 32 * (uint64_t)id[c] : f.node_ref + 32 * (uint64_t)(id[c] - n);
            uint32_t ref[8];
            if (COHERENT) {
                const uint4 *q = reinterpret_cast<const uint4 *>(rp);
                uint4 x = __ldcg(q), y = __ldcg(q + 1);
                ref[0] = x.x; ref[1] = x.y; ref[2] = x.z; ref[3] = x.w;
                ref[4] = y.x; ref[5] = y.y; ref[6] = y.z; ref[7] = y.w;
            } else {
                load32_nc(rp, ref);
            }
            uint32_t nibble = nm[c] & 15;
            s.fill80(nibble - cur);
            cur = nibble;
            uint32_t clen = (nm[c] >> 8) & META_LEN;
            if (clen == 0) {
                s.byte(0xa0);
                s.words8(ref);
            } else {
                for (uint32_t b = 0; b < clen; b++) s.byte(byte_at(ref, b));
            }
            cur++;
        }
    }
    s.fill80(16 - cur + 1);  // trailing empty slots + the value slot
    return list_header_len(payload) + payload;
}

// One thread builds branch node v of depth d into its strip, hashes it and publishes it (node arrays, S/E).
template <int BLOCK, int MAXC, bool COHERENT>
__device__ __forceinline__ void thread_build_node(Strip<BLOCK> &s, uint32_t *smem, const ForestDev &f, uint32_t v, int d,
                                                  uint32_t &hashed, uint32_t &exts, uint32_t (&ref)[8]) {
    s.init(smem);
    uint32_t j0 = f.node_start[v], k = f.node_start[v + 1] - j0;
    if (k + 1 > (uint32_t)MAXC) k = MAXC - 1;  // cannot happen for well-formed input; keeps the strip in bounds
    uint32_t state_mask, tree_mask, hash_mask, l, r;
    uint32_t len = encode_branch_u<BLOCK, MAXC, COHERENT>(s, f, j0, k, state_mask, tree_mask, hash_mask, l, r);
    int pdl = depth_of(f.Lp[l]), pdr = depth_of(f.Lp[(uint64_t)r + 1]);
    int pd = pdl > pdr ? pdl : pdr;
    bool is_root = pd < 0;
    bool need_ext = pd + 1 < d;
    uint32_t meta = strip_to_ref(s, len, is_root && !need_ext, ref, hashed);
    if (need_ext) {
        s.reset();
        uint32_t elen = encode_extension(s, f.keys + 32 * (uint64_t)l, (uint32_t)(pd + 1), (uint32_t)d, ref, meta);
        meta = strip_to_ref(s, elen, is_root, ref, hashed) | META_EXT;
        exts++;
    }
    bool stored = (tree_mask | hash_mask) != 0;
    if (stored) meta |= META_STORED;
    store32(f.node_ref + 32 * (uint64_t)v, ref);
    f.node_meta[v] = (uint8_t)meta;
    f.node_l[v] = l;
    f.node_r[v] = r;
    f.node_masks[v] = make_ushort4((unsigned short)state_mask, (unsigned short)tree_mask, (unsigned short)hash_mask,
                                   (unsigned short)d);
    f.S[l] = (uint32_t)f.n + v;
    f.E[r] = (uint32_t)f.n + v;
}

// One thread per branch node of depth d.  MAXC bounds the children of every node in [pos_lo, pos_hi) (the level's
// nodes are grouped by child-count class); the strip is sized for that class, which is what sets the occupancy.
template <int BLOCK, int MAXC>
__global__ void __launch_bounds__(BLOCK) branch_kernel(ForestDev f, const uint32_t *__restrict__ node_order,
                                                       uint32_t pos_lo, uint32_t pos_hi, int d) {
    extern __shared__ uint32_t smem[];
    if (*(volatile int *)f.err != B200_DEVERR_NONE) return;  // malformed input: the structure arrays are not trustworthy
    Strip<BLOCK> s;
    uint32_t hashed = 0, exts = 0;
    const uint32_t step = gridDim.x * BLOCK;
    for (uint64_t p64 = (uint64_t)pos_lo + blockIdx.x * BLOCK + threadIdx.x; p64 < pos_hi; p64 += step) {
        uint32_t ref[8];
        thread_build_node<BLOCK, MAXC, false>(s, smem, f, __ldg(node_order + p64), d, hashed, exts, ref);
    }
    for (int o = 16; o; o >>= 1) {
        hashed += __shfl_xor_sync(0xffffffffu, hashed, o);
        exts += __shfl_xor_sync(0xffffffffu, exts, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
        if (exts) atomicAdd(&f.counters[CNT_EXT], (unsigned long long)exts);
    }
}

// ------------------------------------------------------------------------------------------------ warp-per-node
// Small levels (the top of every trie, the dirty paths of an incremental update) hold too few nodes to fill the
// machine; there the cost is the LATENCY of one node: 1-4 dependent Keccak-f on one thread is 20-40 us.  Here one
// warp builds one node: the 16 child slots are assembled by 16 lanes in parallel, and the permutation runs with
// the 25 lanes of the sponge state spread over 25 threads (theta/pi/chi as warp shuffles) — the layout the task
// statement sketches.  It is ~5x less ALU-efficient than the register-resident sponge but ~5x shorter in latency,
// so it is used only where a level fits in about one wave of warps.
__constant__ uint8_t KW_ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
__constant__ uint8_t KW_SRC[25] = {0, 6, 12, 18, 24, 3, 9, 10, 16, 22, 1, 7, 13, 19, 20, 4, 5, 11, 17, 23, 2, 8, 14, 15, 21};

static __device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
static __device__ __forceinline__ uint64_t rotl64_var(uint64_t x, uint32_t n) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n & 32) {
        uint32_t t = lo;
        lo = hi;
        hi = t;
    }
    n &= 31;
    return ((uint64_t)__funnelshift_l(lo, hi, n) << 32) | __funnelshift_l(hi, lo, n);
}

struct WarpKeccak {
    int l5, l10, l15, l20, xm1, xp1, src, n1, n2;
    uint32_t rot;
    bool lane0;
    __device__ __forceinline__ void init(int lane) {
        int i = lane % 25, x = i % 5, y = i / 5;
        l5 = (i + 5) % 25; l10 = (i + 10) % 25; l15 = (i + 15) % 25; l20 = (i + 20) % 25;
        xm1 = (x + 4) % 5; xp1 = (x + 1) % 5;
        src = KW_SRC[i]; rot = KW_ROT[i];
        n1 = 5 * y + (x + 1) % 5; n2 = 5 * y + (x + 2) % 5;
        lane0 = lane == 0;
    }
    __device__ __forceinline__ void permute(uint64_t &a) const {
#pragma unroll 1
        for (int r = 0; r < 24; r++) {
            uint64_t c = a ^ shfl64(a, l5) ^ shfl64(a, l10) ^ shfl64(a, l15) ^ shfl64(a, l20);
            uint64_t d = shfl64(c, xm1) ^ rotl64<1>(shfl64(c, xp1));
            uint64_t b = shfl64(rotl64_var(a ^ d, rot), src);
            a = b ^ (~shfl64(b, n1) & shfl64(b, n2));
            if (lane0) a ^= KECCAK_RC[r];
        }
    }
    // keccak256 of buf[0 .. blocks*136) (already padded); digest word i ends up in lane i (i < 4)
    __device__ __forceinline__ uint64_t hash(const uint8_t *buf, uint32_t blocks, int lane) const {
        uint64_t a = 0;
        const uint64_t *w = reinterpret_cast<const uint64_t *>(buf);
        for (uint32_t b = 0; b < blocks; b++) {
            if (lane < 17) a ^= w[17 * b + lane];
            permute(a);
        }
        return a;
    }
};

// byte writer over a warp's linear shared buffer (single-lane use)
struct LinBuf {
    uint8_t *p;
    uint32_t n;
    __device__ __forceinline__ void byte(uint32_t b) { p[n++] = (uint8_t)b; }
    __device__ __forceinline__ void tail32(const uint32_t (&x)[8], uint32_t b0) {
        for (uint32_t b = b0; b < 32; b++) byte(byte_at(x, b));
    }
    __device__ __forceinline__ void words8(const uint32_t (&x)[8]) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            p[n++] = (uint8_t)x[i];
            p[n++] = (uint8_t)(x[i] >> 8);
            p[n++] = (uint8_t)(x[i] >> 16);
            p[n++] = (uint8_t)(x[i] >> 24);
        }
    }
};

constexpr int WARP_BUF = 560;  // 4 rate blocks + slack, 16-byte multiple

// fetch_child with loads that bypass L1 (data produced by other SMs earlier in the SAME kernel: the wavefront)
template <bool COHERENT>
__device__ __forceinline__ ChildInfo fetch_child_c(const ForestDev &f, uint32_t j0, uint32_t c) {
    if (!COHERENT) return fetch_child(f, j0, c);
    ChildInfo ci;
    if (c == 0) {
        uint32_t g = f.gap_sorted[j0];
        ci.id = f.E[g - 1];
        ci.nib = f.nibs[g] >> 4;
    } else {
        uint32_t g = f.gap_sorted[j0 + c - 1];
        ci.id = f.S[g];
        ci.nib = f.nibs[g] & 15;
    }
    ci.meta = ci.id < f.n ? __ldcg(f.leaf_meta + ci.id) : __ldcg(f.node_meta + (ci.id - (uint32_t)f.n));
    return ci;
}

// One warp builds branch node v of depth d (all 32 lanes must call).  Returns through lane 0's stores.
template <bool COHERENT>
__device__ __forceinline__ void warp_build_node(const ForestDev &f, uint32_t v, int d, uint8_t *buf, const WarpKeccak &kw,
                                                int lane, uint32_t &hashed, uint32_t &exts, uint32_t (&out)[8]) {
    uint32_t *bufw = reinterpret_cast<uint32_t *>(buf);
    const uint32_t n = (uint32_t)f.n;
    uint32_t j0 = f.node_start[v], k = f.node_start[v + 1] - j0;
    if (k > 15) k = 15;
    // ---- lane c <= k owns child c
    const bool has = (uint32_t)lane <= k;
    ChildInfo ci{0, 0, 0};
    uint32_t ref[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t lext = 0, rext = 0;
    if (has) {
        ci = fetch_child_c<COHERENT>(f, j0, (uint32_t)lane);
        const uint8_t *rp = ci.id < n ? f.leaf_ref + 32 * (uint64_t)ci.id : f.node_ref + 32 * (uint64_t)(ci.id - n);
        if (COHERENT) {
            const uint4 *q = reinterpret_cast<const uint4 *>(rp);
            uint4 x = __ldcg(q), y = __ldcg(q + 1);
            ref[0] = x.x; ref[1] = x.y; ref[2] = x.z; ref[3] = x.w;
            ref[4] = y.x; ref[5] = y.y; ref[6] = y.z; ref[7] = y.w;
        } else {
            load32_nc(rp, ref);
        }
        if (lane == 0) lext = ci.id < n ? ci.id : f.node_l[ci.id - n];
        if ((uint32_t)lane == k) rext = ci.id < n ? ci.id : f.node_r[ci.id - n];
    }
    uint32_t clen = has ? ((ci.meta & META_LEN) ? (ci.meta & META_LEN) : 33u) : 0u;
    uint32_t bit = has ? (1u << ci.nib) : 0u;
    bool is_branch = has && ci.id >= n;
    uint32_t hbit = (is_branch && !(ci.meta & META_EXT)) ? bit : 0u;
    uint32_t tbit = (is_branch && (ci.meta & META_STORED)) ? bit : 0u;
    if (hbit && (ci.meta & META_LEN) && f.retain_updates) atomicExch(f.err, B200_DEVERR_INLINE_HASH_CHILD);
    uint32_t state_mask = __reduce_or_sync(0xffffffffu, bit);
    uint32_t hash_mask = __reduce_or_sync(0xffffffffu, hbit);
    uint32_t tree_mask = __reduce_or_sync(0xffffffffu, tbit);
    uint32_t incl = clen;  // inclusive prefix sum of child lengths
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    uint32_t children_len = __shfl_sync(0xffffffffu, incl, (int)k);
    uint32_t l = __shfl_sync(0xffffffffu, lext, 0), r = __shfl_sync(0xffffffffu, rext, (int)k);
    uint32_t payload = children_len + (15 - k) + 1;
    uint32_t hdr = list_header_len(payload), total = hdr + payload;
    uint32_t blocks = total / 136 + 1;
    for (uint32_t w = lane; w < blocks * 34; w += 32) bufw[w] = 0;
    __syncwarp();
    if (lane == 0) {
        LinBuf lb{buf, 0};
        put_list_header(lb, payload);
    }
    if (has) {  // child bytes at hdr + (lengths of earlier children) + (empty slots before this nibble)
        uint32_t off = hdr + (incl - clen) + (ci.nib - (uint32_t)lane);
        if ((ci.meta & META_LEN) == 0) {
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
    {  // empty slots: lane e < 16 owns nibble e
        uint32_t cb = __popc(state_mask & ((1u << (lane & 15)) - 1));
        uint32_t before = __shfl_sync(0xffffffffu, incl, cb ? (int)cb - 1 : 0);
        if (lane < 16 && !((state_mask >> lane) & 1)) buf[hdr + (cb ? before : 0u) + ((uint32_t)lane - cb)] = 0x80;
    }
    if (lane == 16) {
        buf[total - 1] = 0x80;  // value slot
        buf[total] |= 0x01;     // pad10*1
        buf[blocks * 136 - 1] |= 0x80;
    }
    __syncwarp();
    // ---- parent depth, extension, hash (uniform control flow)
    int pdl = depth_of(f.Lp[l]), pdr = depth_of(f.Lp[(uint64_t)r + 1]);
    int pd = pdl > pdr ? pdl : pdr;
    bool is_root = pd < 0, need_ext = pd + 1 < d;
    uint32_t meta;
    if (total >= 32 || (is_root && !need_ext)) {
        uint64_t a = kw.hash(buf, blocks, lane);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t w = shfl64(a, i);
            out[2 * i] = (uint32_t)w;
            out[2 * i + 1] = (uint32_t)(w >> 32);
        }
        meta = 0;
        hashed += lane == 0;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = bufw[i];
        meta = total;
    }
    if (need_ext) {
        __syncwarp();
        for (uint32_t w = lane; w < 34; w += 32) bufw[w] = 0;
        __syncwarp();
        uint32_t elen = 0;
        if (lane == 0) {
            LinBuf lb{buf, 0};
            elen = encode_extension(lb, f.keys + 32 * (uint64_t)l, (uint32_t)(pd + 1), (uint32_t)d, out, meta);
            buf[elen] |= 0x01;
            buf[135] |= 0x80;
        }
        elen = __shfl_sync(0xffffffffu, elen, 0);
        __syncwarp();
        if (elen >= 32 || is_root) {
            uint64_t a = kw.hash(buf, 1, lane);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint64_t w = shfl64(a, i);
                out[2 * i] = (uint32_t)w;
                out[2 * i + 1] = (uint32_t)(w >> 32);
            }
            meta = META_EXT;
            hashed += lane == 0;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = bufw[i];
            meta = elen | META_EXT;
        }
        exts += lane == 0;
    }
    if (lane == 0) {
        if ((tree_mask | hash_mask) != 0) meta |= META_STORED;
        store32(f.node_ref + 32 * (uint64_t)v, out);
        f.node_meta[v] = (uint8_t)meta;
        f.node_l[v] = l;
        f.node_r[v] = r;
        f.node_masks[v] = make_ushort4((unsigned short)state_mask, (unsigned short)tree_mask,
                                       (unsigned short)hash_mask, (unsigned short)d);
        f.S[l] = n + v;
        f.E[r] = n + v;
    }
    __syncwarp();
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) branch_warp_kernel(ForestDev f, const uint32_t *__restrict__ node_order,
                                                                uint32_t pos_lo, uint32_t pos_hi, int d) {
    __shared__ __align__(16) uint8_t sbuf[WARPS][WARP_BUF];
    if (*(volatile int *)f.err != B200_DEVERR_NONE) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpKeccak kw;
    kw.init(lane);
    uint32_t hashed = 0, exts = 0;
    const uint32_t stride = gridDim.x * WARPS;
    for (uint64_t p64 = (uint64_t)pos_lo + blockIdx.x * WARPS + warp; p64 < pos_hi; p64 += stride) {
        uint32_t out[8];
        warp_build_node<false>(f, __ldg(node_order + p64), d, sbuf[warp], kw, lane, hashed, exts, out);
    }
    if (lane == 0) {
        if (hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
        if (exts) atomicAdd(&f.counters[CNT_EXT], (unsigned long long)exts);
    }
}

// ------------------------------------------------------------------------------------------------ incremental wavefront
// pending[p] = number of dirty children of node p (dirty leaves and dirty branches), counted by walking up from
// every dirty leaf and stopping at the first ancestor somebody else already reached.
__global__ void mark_pending_kernel(ForestDev f, const uint32_t *__restrict__ idx, uint64_t m,
                                    const uint32_t *__restrict__ leaf_parent, const uint32_t *__restrict__ node_parent,
                                    uint32_t *__restrict__ pending) {
    if (*(volatile int *)f.err != B200_DEVERR_NONE) return;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    uint32_t p = leaf_parent[idx[t]];
    while (p != 0xFFFFFFFFu) {
        if (atomicAdd(&pending[p], 1u) != 0u) break;
        p = node_parent[p];
    }
}

// A warp that just finished a dirty item climbs from its parent p: whoever is the LAST dirty child to arrive at a
// node re-hashes it and goes on; everybody else retires.  Returns true iff this warp finished the root.
__device__ __forceinline__ bool warp_climb(const ForestDev &f, uint32_t p, const uint32_t *__restrict__ node_parent,
                                           uint32_t *__restrict__ pending, uint32_t *__restrict__ dirty_list,
                                           uint32_t *__restrict__ dirty_count, uint8_t *buf, const WarpKeccak &kw, int lane,
                                           uint32_t &hashed, uint32_t &exts, uint32_t (&out)[8]) {
    while (p != 0xFFFFFFFFu) {
        uint32_t last = 0;
        if (lane == 0) {
            __threadfence();  // publish what this warp wrote before announcing arrival
            last = atomicSub(&pending[p], 1u) == 1u;
            __threadfence();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) return false;
        int d = f.node_masks[p].w;
        warp_build_node<true>(f, p, d, buf, kw, lane, hashed, exts, out);
        if (lane == 0) dirty_list[atomicAdd(dirty_count, 1u)] = p;
        p = node_parent[p];
    }
    return true;
}

// One warp per dirty leaf: overwrite + re-hash the leaf, then climb: whoever is the LAST dirty child to arrive at a
// node re-hashes it and continues to its parent; everybody else retires.  The whole dirty-path re-hash of an update
// is this single launch: its latency is (levels) x (one warp-built node), with no host round trip in between.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) wavefront_kernel(ForestDev f, uint8_t *__restrict__ accts,
                                                              uint8_t *__restrict__ sroots,
                                                              const uint8_t *__restrict__ new_accts,
                                                              const uint8_t *__restrict__ new_sroots,
                                                              const uint32_t *__restrict__ idx, uint64_t m,
                                                              const uint32_t *__restrict__ leaf_parent,
                                                              const uint32_t *__restrict__ node_parent,
                                                              uint32_t *__restrict__ pending, uint32_t *__restrict__ dirty_list,
                                                              uint32_t *__restrict__ dirty_count, uint8_t *__restrict__ root_out) {
    __shared__ __align__(16) uint8_t sbuf[WARPS][WARP_BUF];
    if (*(volatile int *)f.err != B200_DEVERR_NONE) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *buf = sbuf[warp];
    uint32_t *bufw = reinterpret_cast<uint32_t *>(buf);
    WarpKeccak kw;
    kw.init(lane);
    uint32_t hashed = 0, exts = 0;
    const uint64_t t = (uint64_t)blockIdx.x * WARPS + warp;
    if (t >= m) return;
    const uint32_t i = idx[t];
    // ---- the leaf
    for (uint32_t w = lane; w < 68; w += 32) bufw[w] = 0;
    __syncwarp();
    int pdl = depth_of(f.Lp[i]), pdr = depth_of(f.Lp[(uint64_t)i + 1]);
    int pd = pdl > pdr ? pdl : pdr;
    uint32_t len = 0;
    if (lane == 0) {
        const uint64_t *src = reinterpret_cast<const uint64_t *>(new_accts + 72 * t);
        uint64_t *dst = reinterpret_cast<uint64_t *>(accts + 72 * (uint64_t)i);
#pragma unroll
        for (int w = 0; w < 9; w++) dst[w] = src[w];
        if (new_sroots && sroots) {
            uint32_t r[8];
            load32(new_sroots + 32 * t, r);
            store32(sroots + 32 * (uint64_t)i, r);
        }
        uint32_t k[8];
        load32(f.keys + 32 * (uint64_t)i, k);
        LinBuf lb{buf, 0};
        len = encode_leaf<LinBuf, true>(lb, k, pd, new_accts + 72 * t,
                                        sroots ? (new_sroots ? new_sroots + 32 * t : sroots + 32 * (uint64_t)i) : nullptr,
                                        f.err);
        buf[len] |= 0x01;
        buf[(len / 136 + 1) * 136 - 1] |= 0x80;
    }
    len = __shfl_sync(0xffffffffu, len, 0);
    __syncwarp();
    uint32_t out[8];
    {
        uint64_t a = kw.hash(buf, len / 136 + 1, lane);  // account leaves are >= 70 bytes: always hashed
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint64_t w = shfl64(a, q);
            out[2 * q] = (uint32_t)w;
            out[2 * q + 1] = (uint32_t)(w >> 32);
        }
        hashed += lane == 0;
    }
    if (lane == 0) {
        store32(f.leaf_ref + 32 * (uint64_t)i, out);
        f.leaf_meta[i] = 0;
    }
    __syncwarp();
    // ---- climb
    bool top = warp_climb(f, leaf_parent[i], node_parent, pending, dirty_list, dirty_count, buf, kw, lane, hashed, exts, out);
    if (top && lane == 0) store32(root_out, out);  // this warp re-hashed the root (or the only leaf)
    if (lane == 0) {
        if (hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
        if (exts) atomicAdd(&f.counters[CNT_EXT], (unsigned long long)exts);
    }
}

// ---- two-stage variant for large dirty sets -----------------------------------------------------------------------
// Stage A: one THREAD per dirty leaf (register-resident sponge: the ALU-efficient formulation) hashes the leaf and
// climbs through the populous deep levels (depth >= split_depth); when the next ancestor is shallower it hands the
// parent over.  Stage B (climb_kernel): one WARP per hand-over finishes the sparse upper levels with the
// latency-optimised warp-cooperative node builder.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) wavefront_thread_kernel(
    ForestDev f, uint8_t *__restrict__ accts, uint8_t *__restrict__ sroots, const uint8_t *__restrict__ new_accts,
    const uint8_t *__restrict__ new_sroots, const uint32_t *__restrict__ idx, uint64_t m,
    const uint32_t *__restrict__ leaf_parent, const uint32_t *__restrict__ node_parent, uint32_t *__restrict__ pending,
    uint32_t *__restrict__ dirty_list, uint32_t *__restrict__ dirty_count, uint32_t *__restrict__ handoff_list,
    uint32_t *__restrict__ handoff_count, uint8_t *__restrict__ root_out, int split_depth) {
    extern __shared__ uint32_t smem[];
    if (*(volatile int *)f.err != B200_DEVERR_NONE) return;
    Strip<BLOCK> s;
    uint32_t hashed = 0, exts = 0;
    uint64_t t = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (t < m) {
        const uint32_t i = idx[t];
        s.init(smem);
        {
            const uint64_t *src = reinterpret_cast<const uint64_t *>(new_accts + 72 * t);
            uint64_t *dst = reinterpret_cast<uint64_t *>(accts + 72 * (uint64_t)i);
#pragma unroll
            for (int w = 0; w < 9; w++) dst[w] = src[w];
            if (new_sroots && sroots) {
                uint32_t r[8];
                load32(new_sroots + 32 * t, r);
                store32(sroots + 32 * (uint64_t)i, r);
            }
        }
        uint32_t k[8];
        load32(f.keys + 32 * (uint64_t)i, k);
        int pdl = depth_of(f.Lp[i]), pdr = depth_of(f.Lp[(uint64_t)i + 1]);
        int pd = pdl > pdr ? pdl : pdr;
        uint32_t len = encode_leaf<Strip<BLOCK>, true>(
            s, k, pd, new_accts + 72 * t, sroots ? (new_sroots ? new_sroots + 32 * t : sroots + 32 * (uint64_t)i) : nullptr,
            f.err);
        uint32_t ref[8];
        uint32_t meta = strip_to_ref(s, len, pd < 0, ref, hashed);
        store32(f.leaf_ref + 32 * (uint64_t)i, ref);
        f.leaf_meta[i] = (uint8_t)meta;
        uint32_t p = leaf_parent[i];
        bool top = true;
        while (p != 0xFFFFFFFFu) {
            int d = f.node_masks[p].w;
            if (d < split_depth) {
                __threadfence();
                handoff_list[atomicAdd(handoff_count, 1u)] = p;
                top = false;
                break;
            }
            __threadfence();
            bool last = atomicSub(&pending[p], 1u) == 1u;
            __threadfence();
            if (!last) {
                top = false;
                break;
            }
            thread_build_node<BLOCK, 16, true>(s, smem, f, p, d, hashed, exts, ref);
            dirty_list[atomicAdd(dirty_count, 1u)] = p;
            p = node_parent[p];
        }
        if (top) store32(root_out, ref);
    }
    for (int o = 16; o; o >>= 1) {
        hashed += __shfl_xor_sync(0xffffffffu, hashed, o);
        exts += __shfl_xor_sync(0xffffffffu, exts, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
        if (exts) atomicAdd(&f.counters[CNT_EXT], (unsigned long long)exts);
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) climb_kernel(ForestDev f, const uint32_t *__restrict__ start_list,
                                                          const uint32_t *__restrict__ start_count_p,
                                                          const uint32_t *__restrict__ node_parent,
                                                          uint32_t *__restrict__ pending, uint32_t *__restrict__ dirty_list,
                                                          uint32_t *__restrict__ dirty_count, uint8_t *__restrict__ root_out) {
    __shared__ __align__(16) uint8_t sbuf[WARPS][WARP_BUF];
    if (*(volatile int *)f.err != B200_DEVERR_NONE) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpKeccak kw;
    kw.init(lane);
    uint32_t hashed = 0, exts = 0;
    const uint32_t count = *start_count_p;
    for (uint32_t e = blockIdx.x * WARPS + warp; e < count; e += gridDim.x * WARPS) {
        uint32_t out[8];
        bool top = warp_climb(f, start_list[e], node_parent, pending, dirty_list, dirty_count, sbuf[warp], kw, lane, hashed,
                              exts, out);
        if (top && lane == 0) store32(root_out, out);
    }
    if (lane == 0) {
        if (hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
        if (exts) atomicAdd(&f.counters[CNT_EXT], (unsigned long long)exts);
    }
}

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

// ------------------------------------------------------------------------------------------------ resident trie (C5)
// Parent links of a finished build: one thread per branch node tells its children who their parent is.
__global__ void parent_links_kernel(ForestDev f, uint32_t n_nodes, uint32_t *__restrict__ leaf_parent,
                                    uint32_t *__restrict__ node_parent) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    uint32_t j0 = f.node_start[v], k = f.node_start[v + 1] - j0;
    for (uint32_t c = 0; c <= k; c++) {
        ChildInfo ci = fetch_child(f, j0, c);
        if (ci.id < f.n) leaf_parent[ci.id] = v;
        else node_parent[ci.id - (uint32_t)f.n] = v;
    }
}

// Finds every dirty key in the resident sorted key array (nothing is written to the trie: if any key is missing
// the error flag makes every later kernel of the update a no-op, so the resident trie stays consistent).
__global__ void locate_kernel(const uint8_t *__restrict__ keys, uint64_t n, const uint8_t *__restrict__ dirty_keys,
                              uint64_t m, uint32_t *__restrict__ idx_out, int *__restrict__ err) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    uint32_t q[8];
    load32(dirty_keys + 32 * t, q);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = __byte_perm(q[i], 0, 0x0123);
    uint64_t lo = 0, hi = n;  // first key >= q
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint32_t kx[8];
        load32_nc(keys + 32 * mid, kx);
        bool less = false, decided = false;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t x = __byte_perm(kx[i], 0, 0x0123);
            if (!decided && x != q[i]) {
                decided = true;
                less = x < q[i];
            }
        }
        if (less) lo = mid + 1;
        else hi = mid;
    }
    bool found = false;
    if (lo < n) {
        uint32_t kx[8];
        load32_nc(keys + 32 * lo, kx);
        found = true;
#pragma unroll
        for (int i = 0; i < 8; i++) found = found && __byte_perm(kx[i], 0, 0x0123) == q[i];
    }
    if (!found) {
        atomicExch(err, B200_DEVERR_NOT_FOUND);
        idx_out[t] = 0xFFFFFFFFu;
        return;
    }
    idx_out[t] = (uint32_t)lo;
}

// ------------------------------------------------------------------------------------------------ structural updates
// lb[t] = lower bound of dirty key t in the resident keys, found[t] = exact match; classifies every entry and counts
// inserts (present && !found), deletes (!present && found) and value updates (present && found).
__global__ void locate_classify_kernel(const uint8_t *__restrict__ keys, uint64_t n, const uint8_t *__restrict__ dirty_keys,
                                       const uint8_t *__restrict__ present, uint64_t m, uint32_t *__restrict__ lb_out,
                                       uint8_t *__restrict__ kind_out /*0 noop,1 update,2 insert,3 delete*/,
                                       uint32_t *__restrict__ counts /*[0]=ins [1]=del [2]=upd*/, int *__restrict__ err) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    uint32_t q[8];
    load32(dirty_keys + 32 * t, q);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = __byte_perm(q[i], 0, 0x0123);
    if (t > 0) {  // the dirty set must be strictly ascending
        uint32_t pk[8];
        load32(dirty_keys + 32 * (t - 1), pk);
        bool less = false, decided = false;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t x = __byte_perm(pk[i], 0, 0x0123);
            if (!decided && x != q[i]) {
                decided = true;
                less = x < q[i];
            }
        }
        if (!less) atomicExch(err, B200_DEVERR_UNSORTED);
    }
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint32_t kx[8];
        load32_nc(keys + 32 * mid, kx);
        bool less = false, decided = false;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t x = __byte_perm(kx[i], 0, 0x0123);
            if (!decided && x != q[i]) {
                decided = true;
                less = x < q[i];
            }
        }
        if (less) lo = mid + 1;
        else hi = mid;
    }
    bool found = false;
    if (lo < n) {
        uint32_t kx[8];
        load32_nc(keys + 32 * lo, kx);
        found = true;
#pragma unroll
        for (int i = 0; i < 8; i++) found = found && __byte_perm(kx[i], 0, 0x0123) == q[i];
    }
    lb_out[t] = (uint32_t)lo;
    bool pres = present == nullptr || present[t] != 0;
    uint8_t kind = pres ? (found ? 1 : 2) : (found ? 3 : 0);
    kind_out[t] = kind;
    if (kind == 2) atomicAdd(&counts[0], 1u);
    if (kind == 3) atomicAdd(&counts[1], 1u);
    if (kind == 1) atomicAdd(&counts[2], 1u);
}

// ins_at[b] += 1 for every insert whose lower bound is b; del[b] = 1 for every delete; ins_flag[t] for the rank scan
__global__ void merge_marks_kernel(const uint32_t *__restrict__ lb, const uint8_t *__restrict__ kind, uint64_t m,
                                   uint32_t *__restrict__ ins_at, uint32_t *__restrict__ del, uint32_t *__restrict__ ins_flag) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    ins_flag[t] = kind[t] == 2 ? 1u : 0u;
    if (kind[t] == 2) atomicAdd(&ins_at[lb[t]], 1u);
    if (kind[t] == 3) del[lb[t]] = 1u;
}

// base element i (not deleted) moves to i + ins_incl[i] - del_excl[i]
__global__ void merge_scatter_base_kernel(const uint8_t *__restrict__ keys, const uint8_t *__restrict__ accts,
                                          const uint8_t *__restrict__ sroots, uint64_t n,
                                          const uint32_t *__restrict__ ins_incl, const uint32_t *__restrict__ del_excl,
                                          const uint32_t *__restrict__ del, uint8_t *__restrict__ nkeys,
                                          uint8_t *__restrict__ naccts, uint8_t *__restrict__ nsroots) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || del[i]) return;
    uint64_t p = i + ins_incl[i] - del_excl[i];
    uint32_t k[8];
    load32_nc(keys + 32 * i, k);
    store32(nkeys + 32 * p, k);
    const uint64_t *src = reinterpret_cast<const uint64_t *>(accts + 72 * i);
    uint64_t *dst = reinterpret_cast<uint64_t *>(naccts + 72 * p);
#pragma unroll
    for (int w = 0; w < 9; w++) dst[w] = src[w];
    if (sroots) {
        load32_nc(sroots + 32 * i, k);
        store32(nsroots + 32 * p, k);
    }
}

// dirty entries: inserts land at (lb - del_excl[lb]) + (number of inserts before them); value updates overwrite
__global__ void merge_scatter_dirty_kernel(const uint8_t *__restrict__ dirty_keys, const uint8_t *__restrict__ new_accts,
                                           const uint8_t *__restrict__ new_sroots, const uint32_t *__restrict__ lb,
                                           const uint8_t *__restrict__ kind, const uint32_t *__restrict__ ins_rank, uint64_t m,
                                           uint64_t n, const uint32_t *__restrict__ ins_incl,
                                           const uint32_t *__restrict__ del_excl, uint8_t *__restrict__ nkeys,
                                           uint8_t *__restrict__ naccts, uint8_t *__restrict__ nsroots) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    uint32_t kd = kind[t];
    if (kd != 1 && kd != 2) return;
    uint64_t b = lb[t], p;
    if (kd == 2) {
        uint64_t dels_before = b < n ? del_excl[b] : del_excl[n];
        p = b - dels_before + ins_rank[t];
    } else {
        p = b + ins_incl[b] - del_excl[b];
    }
    uint32_t k[8];
    if (kd == 2) {
        load32(dirty_keys + 32 * t, k);
        store32(nkeys + 32 * p, k);
    }
    const uint64_t *src = reinterpret_cast<const uint64_t *>(new_accts + 72 * t);
    uint64_t *dst = reinterpret_cast<uint64_t *>(naccts + 72 * p);
#pragma unroll
    for (int w = 0; w < 9; w++) dst[w] = src[w];
    if (nsroots) {
        if (new_sroots) {
            load32(new_sroots + 32 * t, k);
        } else {  // EMPTY_ROOT_HASH for an inserted account without storage information
            k[0] = 0x171fe856u; k[1] = 0xa655cc1bu; k[2] = 0xe64583ffu; k[3] = 0x6ef8c092u;
            k[4] = 0x1be0485bu; k[5] = 0xc0ad6c99u; k[6] = 0xb52f6201u; k[7] = 0x21b463e3u;
        }
        if (new_sroots || kd == 2) store32(nsroots + 32 * p, k);
    }
}

// ------------------------------------------------------------------------------------------------ multi-GPU frontier
// For each of the 16 top-nibble buckets of this rank's account shard: the bucket's node as child of a depth-0
// root branch (as_child) and as a trie of its own (as_root).  The build treated every bucket as a separate
// trie (boundary gaps), so as_root is simply the segment root; as_child re-encodes only the bucket's top node
// with parent depth 0.
template <int BLOCK, bool ACCOUNT>
__global__ void frontier_kernel(ForestDev f, const uint64_t *__restrict__ bucket_offsets /*17*/,
                                const uint8_t *__restrict__ values, const uint8_t *__restrict__ storage_roots,
                                FrontierEntryDev *__restrict__ out) {
    extern __shared__ uint32_t smem[];
    uint32_t b = threadIdx.x;
    Strip<BLOCK> s;
    s.init(smem);
    if (b >= 16 || *(volatile int *)f.err != B200_DEVERR_NONE) return;
    FrontierEntryDev e;
    for (int i = 0; i < 33; i++) e.as_child[i] = e.as_root[i] = 0;
    e.as_child_len = e.as_root_len = 0;
    uint64_t lo = bucket_offsets[b], hi = bucket_offsets[b + 1];
    if (lo < hi) {
        uint32_t item = f.S[lo];
        uint32_t ref[8], hashed = 0, meta;
        const uint8_t *rootp =
            item < f.n ? f.leaf_ref + 32 * (uint64_t)item : f.node_ref + 32 * (uint64_t)(item - (uint32_t)f.n);
        load32_nc(rootp, ref);
        e.as_root_len = 32;
        for (int i = 0; i < 32; i++) e.as_root[i] = (uint8_t)(ref[i >> 2] >> (8 * (i & 3)));
        if (item < f.n) {
            uint32_t k[8];
            load32(f.keys + 32 * (uint64_t)item, k);
            const uint8_t *vp = ACCOUNT ? values + (uint64_t)sizeof(b200_account_dev) * item : values + 32 * (uint64_t)item;
            const uint8_t *sp = (ACCOUNT && storage_roots) ? storage_roots + 32 * (uint64_t)item : nullptr;
            uint32_t len = encode_leaf<Strip<BLOCK>, ACCOUNT>(s, k, 0, vp, sp, f.err);
            meta = strip_to_ref(s, len, false, ref, hashed);
        } else {
            uint32_t v = item - (uint32_t)f.n;
            uint32_t d = f.node_masks[v].w;
            uint32_t j0 = f.node_start[v], k = f.node_start[v + 1] - j0;
            uint32_t sm, tm, hm, l, r;
            uint32_t len = encode_branch(s, f, j0, k, sm, tm, hm, l, r);
            meta = strip_to_ref(s, len, false, ref, hashed);
            if (d > 1) {
                s.reset();
                uint32_t elen = encode_extension(s, f.keys + 32 * (uint64_t)l, 1, d, ref, meta);
                meta = strip_to_ref(s, elen, false, ref, hashed);
            }
        }
        uint32_t il = meta & META_LEN;
        if (il == 0) {
            e.as_child_len = 33;
            e.as_child[0] = 0xa0;
            for (int i = 0; i < 32; i++) e.as_child[1 + i] = (uint8_t)(ref[i >> 2] >> (8 * (i & 3)));
        } else {
            e.as_child_len = (uint8_t)il;
            for (uint32_t i = 0; i < il; i++) e.as_child[i] = (uint8_t)(ref[i >> 2] >> (8 * (i & 3)));
        }
    }
    out[b] = e;
}

// Root from the gathered 16-entry frontier (single thread).
template <int BLOCK>
__global__ void root_from_frontier_kernel(const FrontierEntryDev *__restrict__ fr, uint8_t *__restrict__ root) {
    extern __shared__ uint32_t smem[];
    if (threadIdx.x != 0) return;
    Strip<BLOCK> s;
    s.init(smem);
    uint32_t nonempty = 0, only = 0;
    for (uint32_t b = 0; b < 16; b++)
        if (fr[b].as_root_len) {
            nonempty++;
            only = b;
        }
    uint32_t ref[8];
    if (nonempty == 0) {
        ref[0] = 0x171fe856u; ref[1] = 0xa655cc1bu; ref[2] = 0xe64583ffu; ref[3] = 0x6ef8c092u;
        ref[4] = 0x1be0485bu; ref[5] = 0xc0ad6c99u; ref[6] = 0xb52f6201u; ref[7] = 0x21b463e3u;
    } else if (nonempty == 1) {
        for (int i = 0; i < 8; i++) {
            const uint8_t *p = fr[only].as_root + 4 * i;
            ref[i] = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
        }
    } else {
        uint32_t payload = 1;
        for (uint32_t b = 0; b < 16; b++) payload += fr[b].as_child_len ? fr[b].as_child_len : 1;
        put_list_header(s, payload);
        for (uint32_t b = 0; b < 16; b++) {
            if (fr[b].as_child_len == 0) s.byte(0x80);
            else
                for (uint32_t i = 0; i < fr[b].as_child_len; i++) s.byte(fr[b].as_child[i]);
        }
        s.byte(0x80);
        uint32_t blocks = s.finish();
        strip_keccak(s, blocks, ref);
    }
    store32(root, ref);
}

// bucket_offsets[b] = first account whose top nibble >= b (b = 0..16)
__global__ void nibble_buckets_kernel(const uint8_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ offs) {
    uint32_t b = threadIdx.x;
    if (b > 16) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((uint32_t)(keys[32 * mid] >> 4) < b) lo = mid + 1;
        else hi = mid;
    }
    offs[b] = lo;
}

// ------------------------------------------------------------------------------------------------ launchers
static inline unsigned blocks_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

static int g_sms = 0;
static int sms() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

// grid = min(work, SM count x resident CTAs): a single full wave, grid-stride inside the kernel
template <typename K>
static unsigned persistent_grid(K kernel, int block, size_t smem, uint64_t work_items) {
    static std::mutex mu;
    static std::unordered_map<const void *, int> cache;
    int per_sm;
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find((const void *)kernel);
        if (it == cache.end()) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
            if (per_sm < 1) per_sm = 1;
            cache.emplace((const void *)kernel, per_sm);
        } else {
            per_sm = it->second;
        }
    }
    uint64_t want = (work_items + block - 1) / block;
    uint64_t cap = (uint64_t)sms() * per_sm;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

cudaError_t launch_mark_boundaries(const uint64_t *d_seg_offsets, uint64_t n_segs, uint64_t n, uint8_t *Lp, int *err,
                                   cudaStream_t st) {
    mark_boundaries_kernel<<<blocks_for(n_segs + 1, 256), 256, 0, st>>>(d_seg_offsets, n_segs, n, Lp, err);
    return cudaGetLastError();
}
cudaError_t launch_lcp(const uint8_t *keys, uint64_t n, uint8_t *Lp, uint8_t *nibs, int *err, cudaStream_t st) {
    lcp_kernel<<<blocks_for(n + 1, 256), 256, 0, st>>>(keys, n, Lp, nibs, err);
    return cudaGetLastError();
}
cudaError_t launch_iota(uint32_t *out, uint64_t n, uint32_t first, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    iota_kernel<<<blocks_for(n, 256), 256, 0, st>>>(out, n, first);
    return cudaGetLastError();
}
cudaError_t launch_bucket_offsets(const uint8_t *depth_sorted, uint64_t G, uint32_t *bucket_off, cudaStream_t st) {
    bucket_offsets_kernel<<<1, 96, 0, st>>>(depth_sorted, G, bucket_off);
    return cudaGetLastError();
}
cudaError_t launch_head_flags(const uint8_t *keys, const uint8_t *depth_sorted, const uint32_t *gap_sorted,
                              const uint32_t *bound_rank, const uint32_t *G_real_p, uint64_t G, uint8_t *head,
                              cudaStream_t st) {
    if (G == 0) return cudaSuccess;
    head_flags_kernel<<<blocks_for(G, 256), 256, 0, st>>>(keys, depth_sorted, gap_sorted, bound_rank, G_real_p, G, head);
    return cudaGetLastError();
}
cudaError_t launch_level_ranges(uint32_t *node_start, const uint32_t *n_nodes_p, const uint32_t *bucket_off,
                                uint32_t *level_lo, cudaStream_t st) {
    level_ranges_kernel<<<1, 96, 0, st>>>(node_start, n_nodes_p, bucket_off, level_lo, node_start);
    return cudaGetLastError();
}

constexpr int LEAF_BLOCK = 128;
constexpr int LEAF_WORDS_STORAGE = 34;   // <= 70 bytes -> one rate block
constexpr int LEAF_WORDS_ACCOUNT = 68;   // <= 148 bytes -> two rate blocks
constexpr int BRANCH_BLOCK = 128;
constexpr int BRANCH_WORDS = 136;        // <= 532 bytes -> four rate blocks

cudaError_t launch_leaves(const ForestDev &f, bool account, const uint8_t *values, const uint8_t *storage_roots,
                          cudaStream_t st) {
    if (f.n == 0) return cudaSuccess;
    if (account) {
        auto k = leaf_kernel<LEAF_BLOCK, true>;
        size_t smem = (size_t)LEAF_WORDS_ACCOUNT * LEAF_BLOCK * 4;
        k<<<persistent_grid(k, LEAF_BLOCK, smem, f.n), LEAF_BLOCK, smem, st>>>(f, values, storage_roots);
    } else {
        auto k = leaf_kernel<LEAF_BLOCK, false>;
        size_t smem = (size_t)LEAF_WORDS_STORAGE * LEAF_BLOCK * 4;
        k<<<persistent_grid(k, LEAF_BLOCK, smem, f.n), LEAF_BLOCK, smem, st>>>(f, values, nullptr);
    }
    return cudaGetLastError();
}

template <int MAXC, int WORDS>
static cudaError_t launch_branch_class(const ForestDev &f, const uint32_t *node_order, uint32_t pos_lo, uint32_t pos_hi,
                                       int d, cudaStream_t st) {
    auto k = branch_kernel<BRANCH_BLOCK, MAXC>;
    size_t smem = (size_t)WORDS * BRANCH_BLOCK * 4;
    k<<<persistent_grid(k, BRANCH_BLOCK, smem, pos_hi - pos_lo), BRANCH_BLOCK, smem, st>>>(f, node_order, pos_lo,
                                                                                          pos_hi, d);
    return cudaGetLastError();
}

// cls: child-count class of every node in the range (0: <=3, 1: <=7, 2: <=12, 3: <=16 children), or 3 for a
// mixed range.  The extension wrapper (<= 70 bytes) fits the smallest strip.
cudaError_t launch_branch_level(const ForestDev &f, const uint32_t *node_order, uint32_t pos_lo, uint32_t pos_hi,
                                int d, int cls, cudaStream_t st) {
    if (pos_hi <= pos_lo) return cudaSuccess;
    if (cls < 0) {  // latency path: one warp per node
        constexpr int WARPS = 4;
        uint32_t cnt = pos_hi - pos_lo;
        unsigned blocks = (cnt + WARPS - 1) / WARPS;
        unsigned cap = (unsigned)sms() * 16;
        branch_warp_kernel<WARPS><<<blocks < cap ? blocks : cap, WARPS * 32, 0, st>>>(f, node_order, pos_lo, pos_hi, d);
        return cudaGetLastError();
    }
    switch (cls) {
        case 0: return launch_branch_class<3, 34>(f, node_order, pos_lo, pos_hi, d, st);
        case 1: return launch_branch_class<7, 68>(f, node_order, pos_lo, pos_hi, d, st);
        case 2: return launch_branch_class<12, 102>(f, node_order, pos_lo, pos_hi, d, st);
        default: return launch_branch_class<16, BRANCH_WORDS>(f, node_order, pos_lo, pos_hi, d, st);
    }
}

// sort key of node v: deepest level first, then by the number of rate blocks its RLP needs when every child
// is a 33-byte hash reference (children <= 3 -> 1 block, <= 7 -> 2, <= 12 -> 3, else 4)
// hist[key] counts the nodes of every (depth, class); runs before the host knows the node count, hence the
// device-side bound.
__global__ void node_class_keys_kernel(const uint32_t *__restrict__ node_start, const uint8_t *__restrict__ depth_sorted,
                                       const uint32_t *__restrict__ n_nodes_p, uint8_t *__restrict__ keys,
                                       uint32_t *__restrict__ ids, uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n_nodes = *n_nodes_p;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_nodes; v += gridDim.x * blockDim.x) {
        uint32_t j0 = node_start[v], children = node_start[v + 1] - j0 + 1;
        uint32_t cls = children <= 3 ? 0 : (children <= 7 ? 1 : (children <= 12 ? 2 : 3));
        uint32_t key = ((63u - depth_sorted[j0]) << 2) | cls;
        keys[v] = (uint8_t)key;
        ids[v] = v;
        atomicAdd(&sh[key], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
cudaError_t launch_node_class_keys(const uint32_t *node_start, const uint8_t *depth_sorted, const uint32_t *n_nodes_p,
                                   uint64_t max_nodes, uint8_t *keys, uint32_t *ids, uint32_t *hist, cudaStream_t st) {
    if (max_nodes == 0) return cudaSuccess;
    unsigned blocks = blocks_for(max_nodes, 256);
    if (blocks > (unsigned)sms() * 8) blocks = (unsigned)sms() * 8;
    node_class_keys_kernel<<<blocks, 256, 0, st>>>(node_start, depth_sorted, n_nodes_p, keys, ids, hist);
    return cudaGetLastError();
}

cudaError_t launch_segment_roots(const ForestDev &f, const uint64_t *d_seg_offsets, uint64_t n_segs, uint8_t *roots,
                                 cudaStream_t st) {
    if (n_segs == 0) return cudaSuccess;
    segment_roots_kernel<<<blocks_for(n_segs, 256), 256, 0, st>>>(f, d_seg_offsets, n_segs, roots);
    return cudaGetLastError();
}

cudaError_t launch_stored_flags(const ForestDev &f, uint32_t n_nodes, uint8_t *flags, uint32_t *n_hashes,
                                cudaStream_t st) {
    if (n_nodes == 0) return cudaSuccess;
    stored_flags_kernel<<<blocks_for(n_nodes, 256), 256, 0, st>>>(f, n_nodes, flags, n_hashes);
    return cudaGetLastError();
}
cudaError_t launch_gather_updates(const ForestDev &f, const uint32_t *stored_ids, uint32_t n_stored,
                                  const uint32_t *hash_prefix, const uint32_t *prefix_by_record,
                                  const uint64_t *d_seg_offsets, uint64_t n_segs, const UpdatesDev &out,
                                  cudaStream_t st) {
    if (n_stored == 0) return cudaSuccess;
    gather_updates_kernel<<<blocks_for(n_stored, 128), 128, 0, st>>>(f, stored_ids, n_stored, hash_prefix,
                                                                     prefix_by_record, d_seg_offsets, n_segs, out);
    return cudaGetLastError();
}

cudaError_t launch_nibble_buckets(const uint8_t *keys, uint64_t n, uint64_t *offs, cudaStream_t st) {
    nibble_buckets_kernel<<<1, 32, 0, st>>>(keys, n, offs);
    return cudaGetLastError();
}
cudaError_t launch_frontier(const ForestDev &f, const uint64_t *bucket_offsets, const uint8_t *values,
                            const uint8_t *storage_roots, FrontierEntryDev *out, cudaStream_t st) {
    constexpr int B = 32;
    auto k = frontier_kernel<B, true>;
    size_t smem = (size_t)BRANCH_WORDS * B * 4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<1, B, smem, st>>>(f, bucket_offsets, values, storage_roots, out);
    return cudaGetLastError();
}
cudaError_t launch_root_from_frontier(const FrontierEntryDev *fr, uint8_t *root, cudaStream_t st) {
    constexpr int B = 32;
    auto k = root_from_frontier_kernel<B>;
    size_t smem = (size_t)BRANCH_WORDS * B * 4;
    k<<<1, B, smem, st>>>(fr, root);
    return cudaGetLastError();
}

// ---- resident trie launchers
cudaError_t launch_locate_classify(const uint8_t *keys, uint64_t n, const uint8_t *dirty_keys, const uint8_t *present,
                                   uint64_t m, uint32_t *lb, uint8_t *kind, uint32_t *counts, int *err, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    locate_classify_kernel<<<blocks_for(m, 128), 128, 0, st>>>(keys, n, dirty_keys, present, m, lb, kind, counts, err);
    return cudaGetLastError();
}
cudaError_t launch_merge_marks(const uint32_t *lb, const uint8_t *kind, uint64_t m, uint32_t *ins_at, uint32_t *del,
                               uint32_t *ins_flag, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    merge_marks_kernel<<<blocks_for(m, 256), 256, 0, st>>>(lb, kind, m, ins_at, del, ins_flag);
    return cudaGetLastError();
}
cudaError_t launch_merge_scatter(const uint8_t *keys, const uint8_t *accts, const uint8_t *sroots, uint64_t n,
                                 const uint32_t *ins_incl, const uint32_t *del_excl, const uint32_t *del,
                                 const uint8_t *dirty_keys, const uint8_t *new_accts, const uint8_t *new_sroots,
                                 const uint32_t *lb, const uint8_t *kind, const uint32_t *ins_rank, uint64_t m, uint8_t *nkeys,
                                 uint8_t *naccts, uint8_t *nsroots, cudaStream_t st) {
    if (n) merge_scatter_base_kernel<<<blocks_for(n, 256), 256, 0, st>>>(keys, accts, sroots, n, ins_incl, del_excl, del, nkeys,
                                                                        naccts, nsroots);
    if (m) merge_scatter_dirty_kernel<<<blocks_for(m, 256), 256, 0, st>>>(dirty_keys, new_accts, new_sroots, lb, kind, ins_rank, m,
                                                                         n, ins_incl, del_excl, nkeys, naccts, nsroots);
    return cudaGetLastError();
}
cudaError_t launch_wavefront_two_stage(const ForestDev &f, uint8_t *accts, uint8_t *sroots, const uint8_t *new_accts,
                                       const uint8_t *new_sroots, const uint32_t *idx, uint64_t m,
                                       const uint32_t *leaf_parent, const uint32_t *node_parent, uint32_t *pending,
                                       uint32_t *dirty_list, uint32_t *dirty_count, uint32_t *handoff_list,
                                       uint32_t *handoff_count, uint64_t max_handoff, uint8_t *root_out, int split_depth,
                                       cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    constexpr int TB = 64;
    auto ka = wavefront_thread_kernel<TB>;
    size_t smem = (size_t)BRANCH_WORDS * TB * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    ka<<<blocks_for(m, TB), TB, smem, st>>>(f, accts, sroots, new_accts, new_sroots, idx, m, leaf_parent, node_parent, pending,
                                            dirty_list, dirty_count, handoff_list, handoff_count, root_out, split_depth);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    constexpr int WARPS = 4;
    unsigned blocks = blocks_for(max_handoff ? max_handoff : 1, WARPS);
    unsigned cap = (unsigned)sms() * 16;
    climb_kernel<WARPS><<<blocks < cap ? blocks : cap, WARPS * 32, 0, st>>>(f, handoff_list, handoff_count, node_parent, pending,
                                                                         dirty_list, dirty_count, root_out);
    return cudaGetLastError();
}
cudaError_t launch_mark_pending(const ForestDev &f, const uint32_t *idx, uint64_t m, const uint32_t *leaf_parent,
                                const uint32_t *node_parent, uint32_t *pending, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    mark_pending_kernel<<<blocks_for(m, 128), 128, 0, st>>>(f, idx, m, leaf_parent, node_parent, pending);
    return cudaGetLastError();
}
cudaError_t launch_wavefront(const ForestDev &f, uint8_t *accts, uint8_t *sroots, const uint8_t *new_accts,
                             const uint8_t *new_sroots, const uint32_t *idx, uint64_t m, const uint32_t *leaf_parent,
                             const uint32_t *node_parent, uint32_t *pending, uint32_t *dirty_list, uint32_t *dirty_count,
                             uint8_t *root_out, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    constexpr int WARPS = 4;
    wavefront_kernel<WARPS><<<blocks_for(m, WARPS), WARPS * 32, 0, st>>>(f, accts, sroots, new_accts, new_sroots, idx, m,
                                                                        leaf_parent, node_parent, pending, dirty_list,
                                                                        dirty_count, root_out);
    return cudaGetLastError();
}
cudaError_t launch_stored_flags_subset(const ForestDev &f, const uint32_t *ids, uint32_t count, uint8_t *flags,
                                       uint32_t *n_hashes, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    stored_flags_subset_kernel<<<blocks_for(count, 256), 256, 0, st>>>(f, ids, count, flags, n_hashes);
    return cudaGetLastError();
}
cudaError_t launch_pick_subset(const uint32_t *ids, const uint32_t *prefix, const uint32_t *sel_pos, uint32_t n_sel,
                               uint32_t *out_ids, uint32_t *out_prefix, cudaStream_t st) {
    if (n_sel == 0) return cudaSuccess;
    pick_subset_kernel<<<blocks_for(n_sel, 256), 256, 0, st>>>(ids, prefix, sel_pos, n_sel, out_ids, out_prefix);
    return cudaGetLastError();
}
cudaError_t launch_parent_links(const ForestDev &f, uint32_t n_nodes, uint32_t *leaf_parent, uint32_t *node_parent,
                                cudaStream_t st) {
    if (n_nodes == 0) return cudaSuccess;
    parent_links_kernel<<<blocks_for(n_nodes, 256), 256, 0, st>>>(f, n_nodes, leaf_parent, node_parent);
    return cudaGetLastError();
}
cudaError_t launch_locate(const uint8_t *keys, uint64_t n, const uint8_t *dirty_keys, uint64_t m, uint32_t *idx_out,
                          int *err, cudaStream_t st) {
    if (m == 0) return cudaSuccess;
    locate_kernel<<<blocks_for(m, 128), 128, 0, st>>>(keys, n, dirty_keys, m, idx_out, err);
    return cudaGetLastError();
}

}  // namespace b200
