// tk_resident.cuh — resident trie: parent links, key location, structural merge kernels.
// Part of the single translation unit trie_kernels.cu (included inside namespace b200, in this order: the later
// files use the device functions of the earlier ones).

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
