// tk_ordered.cuh — ordered (index-keyed) tries: transactions / receipts / withdrawals roots.
// Part of the single translation unit trie_kernels.cu (included inside namespace b200, after tk_launchers.cuh).
//
// reth: OrderedTrieRootEncodedBuilder (crates/trie/common/src/ordered_root.rs:146-257) feeds a HashBuilder with
// (rlp(index), pre-encoded item) in the order adjust_index_for_rlp yields (flush(), :202-216); alloy_trie's
// ordered_trie_root_with_encoder does the same for the proofs::calculate_{transaction,receipt,withdrawals}_root callers
// (crates/ethereum/primitives/src/receipt.rs:17-20, crates/ethereum/evm/src/build.rs:56-68).
//
// Here a batch of lists is one forest.  rlp(index) keys are prefix-free, so zero-padded to 32 bytes they have the same
// common prefixes and the structure passes (lcp, gaps, levels, branch kernels) run unchanged; only the leaf differs:
// its path ends at the key's true length and its value is an arbitrary byte string streamed from HBM.

// position j of the sorted key order -> list index (alloy_trie::root::adjust_index_for_rlp)
static __device__ __forceinline__ uint32_t ord_adjust_index(uint32_t j, uint32_t len) {
    if (j > 0x7f) return j;
    if (j == 0x7f || j + 1 == len) return 0;
    return j + 1;
}

// One thread per leaf position: the padded key, its true length in nibbles, the item it carries, and a scheduling key
// for the leaf pass (Keccak blocks of the item, longest first: threads of a warp then hash items of about equal length).
__global__ void ordered_keys_kernel(const uint64_t *__restrict__ seg_offsets, uint64_t n_segs, uint64_t n,
                                    const uint64_t *__restrict__ val_off, uint8_t *__restrict__ keys,
                                    uint8_t *__restrict__ key_nibs, uint32_t *__restrict__ item,
                                    uint16_t *__restrict__ sched_key, uint32_t *__restrict__ pos, int *err) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    // last segment starting at or before p (empty segments share a start: take the last of them)
    uint64_t lo = 0, hi = n_segs;  // invariant: seg_offsets[lo] <= p < seg_offsets[hi] once offsets are sane
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (seg_offsets[mid] <= p) lo = mid;
        else hi = mid;
    }
    uint64_t s0 = seg_offsets[lo], s1 = seg_offsets[lo + 1];
    uint32_t idx = 0;
    uint64_t it = p;
    if (s0 > p || s1 <= p || s1 > n) {
        atomicExch(err, B200_DEVERR_BAD_OFFSETS);
    } else {
        idx = ord_adjust_index((uint32_t)(p - s0), (uint32_t)(s1 - s0));
        it = s0 + idx;
    }
    // rlp(idx): 0 -> 0x80, < 0x80 -> itself, else 0x80 + byte count, big-endian bytes
    uint32_t w0 = 0, w1 = 0, nb;
    if (idx == 0) {
        w0 = 0x80;
        nb = 1;
    } else if (idx < 0x80) {
        w0 = idx;
        nb = 1;
    } else {
        uint32_t bytes = 4 - (__clz(idx) >> 3);
        nb = 1 + bytes;
        uint64_t acc = 0x80 + bytes;  // little-endian byte string in a register
        for (uint32_t b = 0; b < bytes; b++) acc |= (uint64_t)((idx >> (8 * (bytes - 1 - b))) & 0xff) << (8 * (b + 1));
        w0 = (uint32_t)acc;
        w1 = (uint32_t)(acc >> 32);
    }
    uint4 *q = reinterpret_cast<uint4 *>(keys + 32 * p);
    q[0] = make_uint4(w0, w1, 0, 0);
    q[1] = make_uint4(0, 0, 0, 0);
    key_nibs[p] = (uint8_t)(2 * nb);
    item[p] = (uint32_t)it;
    uint64_t blocks = (val_off[it + 1] - val_off[it] + 17) / 136;  // (offsets are validated by the leaf pass)
    sched_key[p] = (uint16_t)(65535u - (uint32_t)(blocks < 65535 ? blocks : 65535));
    pos[p] = (uint32_t)p;
}

// bytes of the big-endian representation of a length
static __device__ __forceinline__ uint32_t ord_be_len(uint32_t x) {
    return x < 0x100 ? 1 : (x < 0x10000 ? 2 : (x < 0x1000000 ? 3 : 4));
}

struct OrdPrefix {  // the bytes of a leaf in front of its value: list header, hex-prefix path, string header
    uint8_t b[24];
    uint32_t n = 0;
    __device__ __forceinline__ void put(uint32_t x) { b[n++] = (uint8_t)x; }
    __device__ __forceinline__ void put_len(uint32_t base_short, uint32_t base_long, uint32_t len) {
        if (len < 56) {
            put(base_short + len);
        } else {
            uint32_t k = ord_be_len(len);
            put(base_long + k);
            for (int i = (int)k - 1; i >= 0; i--) put((len >> (8 * i)) & 0xff);
        }
    }
};

// 8 bytes at any alignment from two aligned words; the caller guarantees [p & ~7, (p & ~7) + 16) is readable
static __device__ __forceinline__ uint64_t ord_load8(const uint8_t *p) {
    uintptr_t u = reinterpret_cast<uintptr_t>(p);
    const uint64_t *q = reinterpret_cast<const uint64_t *>(u & ~(uintptr_t)7);
    uint32_t sh = (uint32_t)(u & 7) * 8;
    uint64_t lo = __ldg(q);
    if (sh == 0) return lo;
    uint64_t hi = __ldg(q + 1);
    return (lo >> sh) | (hi << (64 - sh));
}

// ------------------------------------------------------------------------------------------------ leaf encoding
// Everything about one leaf that both leaf kernels need: where its value lies and the bytes in front of it.
struct OrdLeaf {
    OrdPrefix pre;          // list header ‖ hex-prefix path ‖ string header
    const uint8_t *vp;      // the value
    uint64_t vo;            // its offset in the blob
    uint32_t vlen, total;   // value bytes, bytes of the whole leaf RLP
    int pd;                 // depth of the branch above (-1: the leaf is a whole trie)
    bool ok;
};

// RlpNode of LeafNode{ key[pd+1 .. key_nibs), value }: the header bytes and the extent of the value.
static __device__ __forceinline__ void ord_leaf_header(const ForestDev &f, uint64_t i, const uint8_t *__restrict__ key_nibs,
                                                       const uint32_t *__restrict__ item,
                                                       const uint8_t *__restrict__ values,
                                                       const uint64_t *__restrict__ val_off, uint64_t blob_len, OrdLeaf &L) {
    uint32_t k[8];
    load32_nc(f.keys + 32 * i, k);
    int pdl = depth_of(f.Lp[i]), pdr = depth_of(f.Lp[i + 1]);
    L.pd = pdl > pdr ? pdl : pdr;
    const uint32_t kn = key_nibs[i];
    const uint32_t p = (uint32_t)(L.pd + 1);
    const uint32_t it = item[i];
    const uint64_t vo = val_off[it], ve = val_off[it + 1];
    bool bad = ve < vo || ve > blob_len || ve - vo >= (1ull << 31);
    if (p > kn) {  // (cannot happen with prefix-free keys)
        atomicExch(f.err, B200_DEVERR_CORRUPT);
        bad = true;
    } else if (bad) {
        atomicExch(f.err, B200_DEVERR_BAD_OFFSETS);
    }
    L.ok = !bad;
    if (bad) return;
    L.vo = vo;
    L.vlen = (uint32_t)(ve - vo);
    L.vp = values + vo;
    const uint32_t m = kn - p;  // path nibbles left for the leaf (0 when the key ends at the branch)
    const uint32_t hp_len = 1 + (m >> 1);
    const uint32_t hp_str = hp_len == 1 ? 1 : 1 + hp_len;
    const uint32_t first = (m & 1) ? (0x30u | (byte_at(k, p >> 1) & 15)) : 0x20u;
    const uint32_t b0 = (p + 1) >> 1;
    const bool single = L.vlen == 1 && __ldg(L.vp) < 0x80;
    const uint32_t val_hdr = single ? 0 : (L.vlen < 56 ? 1 : 1 + ord_be_len(L.vlen));
    const uint32_t payload = hp_str + val_hdr + L.vlen;
    L.pre.put_len(0xc0, 0xf7, payload);
    if (hp_len > 1) L.pre.put(0x80 + hp_len);
    L.pre.put(first);
    for (uint32_t b = b0; b < (kn >> 1); b++) L.pre.put(byte_at(k, b));
    if (!single) L.pre.put_len(0x80, 0xb7, L.vlen);
    L.total = L.pre.n + L.vlen;
}

static __device__ __forceinline__ uint32_t ord_msg_byte(const OrdLeaf &L, uint32_t j) {
    return j < L.pre.n ? L.pre.b[j] : __ldg(L.vp + (j - L.pre.n));
}

// Sponge lane `lane` (0..16) of the rate block starting at message offset `off` (`take` = message bytes in this block;
// a block with take < 136 is the last one and carries the 0x01 .. 0x80 padding): 8 bytes from two aligned loads where
// the word lies inside the value, byte by byte across the header / value seam and in the padded tail.
static __device__ __forceinline__ uint64_t ord_msg_word(const OrdLeaf &L, uint32_t off, uint32_t take, uint32_t lane,
                                                        bool aligned_blob, uint64_t blob_len) {
    const uint32_t o = off + 8 * lane, npre = L.pre.n;
    if (aligned_blob && o >= npre && o + 8 <= L.total) {
        const uint64_t word_end = ((L.vo + (o - npre)) & ~7ull) + 16;  // offsets relative to the aligned blob
        if (word_end <= blob_len) return ord_load8(L.vp + (o - npre));
    }
    uint64_t w = 0;
    for (uint32_t b = 0; b < 8; b++) {
        const uint32_t j = 8 * lane + b;  // within this block
        uint32_t x = j < take ? ord_msg_byte(L, off + j) : 0;
        if (take < 136 && j == take) x ^= 0x01;
        if (take < 136 && j == 135) x ^= 0x80;
        w |= (uint64_t)x << (8 * b);
    }
    return w;
}

constexpr uint32_t ORD_LONG_BLOCKS = 32;  // items of >= 32 rate blocks (≈4.3 KB) get a warp each

// *n_long = number of leading entries of the leaf order whose item has at least ORD_LONG_BLOCKS blocks
// (sched_sorted ascending = block count descending).
__global__ void ordered_count_long_kernel(const uint16_t *__restrict__ sched_sorted, uint64_t n, uint32_t *n_long) {
    const uint32_t limit = 65535u - ORD_LONG_BLOCKS;  // keys <= limit are long
    uint64_t lo = 0, hi = n;                           // first position with key > limit
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (sched_sorted[mid] <= limit) lo = mid + 1;
        else hi = mid;
    }
    *n_long = (uint32_t)lo;
}

// One thread per leaf (the short and medium items: order[n_long ..)).  The value is absorbed straight from HBM into the
// register-resident sponge.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) ordered_leaf_kernel(ForestDev f, const uint8_t *__restrict__ key_nibs,
                                                             const uint32_t *__restrict__ item,
                                                             const uint32_t *__restrict__ order,
                                                             const uint32_t *__restrict__ n_long_p,
                                                             const uint8_t *__restrict__ values,
                                                             const uint64_t *__restrict__ val_off, uint64_t blob_len) {
    if (*(volatile int *)f.err == B200_DEVERR_UNSORTED || *(volatile int *)f.err == B200_DEVERR_BAD_OFFSETS) return;
    const uint64_t t = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    uint32_t hashed = 0;
    if (t < f.n && t >= *n_long_p) {
        const uint64_t i = order[t];  // leaves in descending order of item length
        OrdLeaf L;
        ord_leaf_header(f, i, key_nibs, item, values, val_off, blob_len, L);
        if (L.ok) {
            uint32_t ref[8];
            uint32_t meta;
            if (L.total < 32 && L.pd >= 0) {
#pragma unroll
                for (int w = 0; w < 8; w++) ref[w] = 0;
                for (uint32_t j = 0; j < L.total; j++) {
                    uint32_t x = ord_msg_byte(L, j) << (8 * (j & 3));
#pragma unroll
                    for (int w = 0; w < 8; w++)
                        if ((j >> 2) == (uint32_t)w) ref[w] |= x;
                }
                meta = L.total;
            } else {
                const bool aligned_blob = (reinterpret_cast<uintptr_t>(values) & 7) == 0;
                uint64_t a[25];
#pragma unroll
                for (int q = 0; q < 25; q++) a[q] = 0;
                uint32_t off = 0;
                for (;;) {
                    const uint32_t take = L.total - off < 136 ? L.total - off : 136;
#pragma unroll
                    for (int lane = 0; lane < 17; lane++) a[lane] ^= ord_msg_word(L, off, take, lane, aligned_blob, blob_len);
                    keccak_f1600(a);
                    off += take;
                    if (take < 136) break;
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    ref[2 * q] = (uint32_t)a[q];
                    ref[2 * q + 1] = (uint32_t)(a[q] >> 32);
                }
                meta = 0;
                hashed = 1;
            }
            store32(f.leaf_ref + 32 * i, ref);
            f.leaf_meta[i] = (uint8_t)meta;
        }
        f.S[i] = (uint32_t)i;
        f.E[i] = (uint32_t)i;
    }
    for (int o = 16; o; o >>= 1) hashed += __shfl_xor_sync(0xffffffffu, hashed, o);
    if ((threadIdx.x & 31) == 0 && hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
}

// One warp per long item (order[0 .. n_long)): the sponge's 25 lanes spread over 25 threads (WarpKeccak, tk_warp.cuh) —
// ~5x shorter dependent chain per block than the register-resident sponge, and lanes 0..16 read the 136 bytes of a rate
// block side by side.  A 128 KB calldata transaction is 964 dependent permutations: that chain is the tail of the call.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ordered_long_leaf_kernel(ForestDev f, const uint8_t *__restrict__ key_nibs,
                                                                       const uint32_t *__restrict__ item,
                                                                       const uint32_t *__restrict__ order,
                                                                       const uint32_t *__restrict__ n_long_p,
                                                                       const uint8_t *__restrict__ values,
                                                                       const uint64_t *__restrict__ val_off,
                                                                       uint64_t blob_len) {
    if (*(volatile int *)f.err == B200_DEVERR_UNSORTED || *(volatile int *)f.err == B200_DEVERR_BAD_OFFSETS) return;
    const int lane = threadIdx.x & 31;
    const uint32_t n_long = *n_long_p;
    WarpKeccak kw;
    kw.init(lane);
    const bool aligned_blob = (reinterpret_cast<uintptr_t>(values) & 7) == 0;
    uint32_t hashed = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); t < n_long; t += (uint64_t)gridDim.x * WARPS) {
        const uint64_t i = order[t];
        OrdLeaf L;  // every lane derives the same header (a few dozen instructions; no broadcast needed)
        ord_leaf_header(f, i, key_nibs, item, values, val_off, blob_len, L);
        if (L.ok) {  // uniform across the warp
            uint64_t a = 0;
            uint32_t off = 0;
            for (;;) {
                const uint32_t take = L.total - off < 136 ? L.total - off : 136;
                if (lane < 17) a ^= ord_msg_word(L, off, take, (uint32_t)lane, aligned_blob, blob_len);
                kw.permute(a);
                off += take;
                if (take < 136) break;
            }
            uint32_t ref[8];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint64_t w = shfl64(a, q);
                ref[2 * q] = (uint32_t)w;
                ref[2 * q + 1] = (uint32_t)(w >> 32);
            }
            if (lane == 0) {
                store32(f.leaf_ref + 32 * i, ref);
                f.leaf_meta[i] = 0;
                hashed++;
            }
        }
        if (lane == 0) {
            f.S[i] = (uint32_t)i;
            f.E[i] = (uint32_t)i;
        }
    }
    if (lane == 0 && hashed) atomicAdd(&f.counters[CNT_HASHED], (unsigned long long)hashed);
}

cudaError_t launch_ordered_keys(const uint64_t *d_seg_offsets, uint64_t n_segs, uint64_t n, const uint64_t *val_off,
                                uint8_t *keys, uint8_t *key_nibs, uint32_t *item, uint16_t *sched_key, uint32_t *pos, int *err,
                                cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    ordered_keys_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_seg_offsets, n_segs, n, val_off, keys, key_nibs, item,
                                                            sched_key, pos, err);
    return cudaGetLastError();
}
cudaError_t launch_ordered_leaves(const ForestDev &f, const OrderedLeavesDev &o, cudaStream_t st) {
    if (f.n == 0) return cudaSuccess;
    constexpr int BLOCK = 128, WARPS = 4;
    ordered_count_long_kernel<<<1, 1, 0, st>>>(o.sched_sorted, f.n, o.n_long);
    auto kl = ordered_long_leaf_kernel<WARPS>;
    kl<<<persistent_grid(kl, WARPS * 32, 0, f.n * 32), WARPS * 32, 0, st>>>(f, o.key_nibs, o.item, o.order, o.n_long, o.values,
                                                                           o.val_off, o.blob_len);
    ordered_leaf_kernel<BLOCK><<<blocks_for(f.n, BLOCK), BLOCK, 0, st>>>(f, o.key_nibs, o.item, o.order, o.n_long, o.values,
                                                                         o.val_off, o.blob_len);
    return cudaGetLastError();
}
