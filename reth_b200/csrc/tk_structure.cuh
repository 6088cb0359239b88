// tk_structure.cuh — structure pass: gap depths, boundaries, bucket offsets, node heads, level ranges.
// Part of the single translation unit trie_kernels.cu (included inside namespace b200, in this order: the later
// files use the device functions of the earlier ones).

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

// Start of a build: the error word of the previous (possibly still unreported) build is latched into the sticky word —
// first error wins — and cleared, together with the build counters, so that this build starts clean without erasing
// what b200_sync / b200_dev_status still has to report (include/b200trie.h: violations of async calls are sticky).
__global__ void latch_error_kernel(int *__restrict__ err, int *__restrict__ sticky, unsigned long long *__restrict__ counters) {
    int e = *err;
    if (e != B200_DEVERR_NONE && *sticky == B200_DEVERR_NONE) *sticky = e;
    *err = B200_DEVERR_NONE;
    for (int i = 0; i < CNT_COUNT; i++) counters[i] = 0;
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
__global__ void head_flags_kernel(const uint8_t *__restrict__ keys, const uint8_t *__restrict__ Lp, const uint8_t *__restrict__ depth_sorted,
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
            } else if (g - gp <= 33) {
                // short span (the rule in forests of small tries): leaves gp .. g-1 share > d nibbles iff every gap between
                // them is deeper than d — a few consecutive bytes of Lp (L2-resident) instead of two random 32-byte key rows
                bool deeper = true;
                for (uint32_t q = gp + 1; q < g; q++) deeper = deeper && Lp[q] > d;
                h = !deeper;
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
