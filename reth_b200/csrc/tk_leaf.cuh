// tk_leaf.cuh — leaf encoding (storage slot / account) and the leaf kernel.
// Part of the single translation unit trie_kernels.cu (included inside namespace b200, in this order: the later
// files use the device functions of the earlier ones).

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
