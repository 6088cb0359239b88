/*
 * keccak.c — Keccak-256 (Keccak-f[1600], rate 136 B, padding 0x01 .. 0x80), scalar C.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle.h).
 *
 * Restates what alloy-primitives 1.6.0 `keccak256` computes (external crate, not under /root/reference;
 * reth's release binary routes it to keccak-asm 0.1.6, bin/reth/Cargo.toml:85-94).  Call sites being
 * mirrored: crates/trie/common/src/key.rs:4-18 (KeccakKeyHasher), hashing_account.rs:198-202,
 * hashing_storage.rs:131-137.  Pinned by the KATs of SURVEY.md Appendix B #1-#4.
 */
#include "oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

static __thread uint64_t tl_keccak_f = 0;
static __thread uint64_t tl_bytes_hashed = 0;

uint64_t orc__keccak_f_count(void) { return tl_keccak_f; }
uint64_t orc__bytes_hashed(void) { return tl_bytes_hashed; }
void orc__keccak_counters_reset(void) { tl_keccak_f = 0; tl_bytes_hashed = 0; }

#define ROL(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

static void keccak_f1600(uint64_t a[25]) {
    uint64_t a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3], a04 = a[4];
    uint64_t a05 = a[5], a06 = a[6], a07 = a[7], a08 = a[8], a09 = a[9];
    uint64_t a10 = a[10], a11 = a[11], a12 = a[12], a13 = a[13], a14 = a[14];
    uint64_t a15 = a[15], a16 = a[16], a17 = a[17], a18 = a[18], a19 = a[19];
    uint64_t a20 = a[20], a21 = a[21], a22 = a[22], a23 = a[23], a24 = a[24];
    for (int r = 0; r < 24; r++) {
        /* theta */
        uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20;
        uint64_t c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21;
        uint64_t c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22;
        uint64_t c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23;
        uint64_t c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        uint64_t d0 = c4 ^ ROL(c1, 1), d1 = c0 ^ ROL(c2, 1), d2 = c1 ^ ROL(c3, 1);
        uint64_t d3 = c2 ^ ROL(c4, 1), d4 = c3 ^ ROL(c0, 1);
        /* theta + rho + pi: b[y][2x+3y] = rol(a[x][y] ^ d[x], r[x][y]) */
        uint64_t b00 = a00 ^ d0;
        uint64_t b10 = ROL(a01 ^ d1, 1);
        uint64_t b20 = ROL(a02 ^ d2, 62);
        uint64_t b05 = ROL(a03 ^ d3, 28);
        uint64_t b15 = ROL(a04 ^ d4, 27);
        uint64_t b16 = ROL(a05 ^ d0, 36);
        uint64_t b01 = ROL(a06 ^ d1, 44);
        uint64_t b11 = ROL(a07 ^ d2, 6);
        uint64_t b21 = ROL(a08 ^ d3, 55);
        uint64_t b06 = ROL(a09 ^ d4, 20);
        uint64_t b07 = ROL(a10 ^ d0, 3);
        uint64_t b17 = ROL(a11 ^ d1, 10);
        uint64_t b02 = ROL(a12 ^ d2, 43);
        uint64_t b12 = ROL(a13 ^ d3, 25);
        uint64_t b22 = ROL(a14 ^ d4, 39);
        uint64_t b23 = ROL(a15 ^ d0, 41);
        uint64_t b08 = ROL(a16 ^ d1, 45);
        uint64_t b18 = ROL(a17 ^ d2, 15);
        uint64_t b03 = ROL(a18 ^ d3, 21);
        uint64_t b13 = ROL(a19 ^ d4, 8);
        uint64_t b14 = ROL(a20 ^ d0, 18);
        uint64_t b24 = ROL(a21 ^ d1, 2);
        uint64_t b09 = ROL(a22 ^ d2, 61);
        uint64_t b19 = ROL(a23 ^ d3, 56);
        uint64_t b04 = ROL(a24 ^ d4, 14);
        /* chi + iota */
        a00 = b00 ^ (~b01 & b02) ^ RC[r];
        a01 = b01 ^ (~b02 & b03);
        a02 = b02 ^ (~b03 & b04);
        a03 = b03 ^ (~b04 & b00);
        a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07);
        a06 = b06 ^ (~b07 & b08);
        a07 = b07 ^ (~b08 & b09);
        a08 = b08 ^ (~b09 & b05);
        a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12);
        a11 = b11 ^ (~b12 & b13);
        a12 = b12 ^ (~b13 & b14);
        a13 = b13 ^ (~b14 & b10);
        a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17);
        a16 = b16 ^ (~b17 & b18);
        a17 = b17 ^ (~b18 & b19);
        a18 = b18 ^ (~b19 & b15);
        a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22);
        a21 = b21 ^ (~b22 & b23);
        a22 = b22 ^ (~b23 & b24);
        a23 = b23 ^ (~b24 & b20);
        a24 = b24 ^ (~b20 & b21);
    }
    a[0] = a00; a[1] = a01; a[2] = a02; a[3] = a03; a[4] = a04;
    a[5] = a05; a[6] = a06; a[7] = a07; a[8] = a08; a[9] = a09;
    a[10] = a10; a[11] = a11; a[12] = a12; a[13] = a13; a[14] = a14;
    a[15] = a15; a[16] = a16; a[17] = a17; a[18] = a18; a[19] = a19;
    a[20] = a20; a[21] = a21; a[22] = a22; a[23] = a23; a[24] = a24;
    tl_keccak_f++;
}

void orc_keccak256(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint64_t st[25];
    uint8_t blk[136];
    memset(st, 0, sizeof st);
    tl_bytes_hashed += len;
    while (len >= 136) {
        for (int i = 0; i < 17; i++) {
            uint64_t w;
            memcpy(&w, in + 8 * i, 8); /* little-endian host */
            st[i] ^= w;
        }
        keccak_f1600(st);
        in += 136;
        len -= 136;
    }
    memset(blk, 0, sizeof blk);
    if (len) memcpy(blk, in, len);
    blk[len] ^= 0x01;
    blk[135] ^= 0x80;
    for (int i = 0; i < 17; i++) {
        uint64_t w;
        memcpy(&w, blk + 8 * i, 8);
        st[i] ^= w;
    }
    keccak_f1600(st);
    memcpy(out, st, 32);
}

/* ------------------------------------------------------------------ batch drivers */
typedef struct {
    const uint8_t *in;
    uint32_t msg_len, stride;
    const uint64_t *offsets;
    uint64_t n;
    uint8_t *out;
    uint64_t next; /* atomic chunk cursor */
    int simd;      /* 8-way AVX-512 multi-buffer for the fixed-length case */
} batch_job;

extern int orc_have_avx512(void);
extern void orc_keccak256_x8(const uint8_t *in, uint32_t msg_len, uint32_t stride, uint8_t *out);

#define CHUNK 100 /* hashing_account.rs:32 WORKER_CHUNK_SIZE */

static void *batch_worker(void *p) {
    batch_job *j = (batch_job *)p;
    for (;;) {
        uint64_t lo = __atomic_fetch_add(&j->next, CHUNK, __ATOMIC_RELAXED);
        if (lo >= j->n) break;
        uint64_t hi = lo + CHUNK < j->n ? lo + CHUNK : j->n;
        if (j->offsets) {
            for (uint64_t i = lo; i < hi; i++)
                orc_keccak256(j->in + j->offsets[i], (size_t)(j->offsets[i + 1] - j->offsets[i]),
                              j->out + 32 * i);
        } else {
            uint64_t i = lo;
            if (j->simd)
                for (; i + 8 <= hi; i += 8) orc_keccak256_x8(j->in + (size_t)j->stride * i, j->msg_len, j->stride, j->out + 32 * i);
            for (; i < hi; i++) orc_keccak256(j->in + (size_t)j->stride * i, j->msg_len, j->out + 32 * i);
        }
    }
    return NULL;
}

static void run_batch(batch_job *j, int threads) {
    if (threads <= 1) {
        batch_worker(j);
        return;
    }
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, batch_worker, j);
    for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
    free(t);
}

void orc_keccak256_fixed(const uint8_t *in, uint32_t msg_len, uint32_t stride, uint64_t n, uint8_t *out32,
                         int threads) {
    batch_job j = {in, msg_len, stride, NULL, n, out32, 0, 0};
    run_batch(&j, threads);
}

/* Same result, eight sponges per AVX-512 register where the CPU has it (returns 0 and does nothing otherwise). */
int orc_keccak256_fixed_simd(const uint8_t *in, uint32_t msg_len, uint32_t stride, uint64_t n, uint8_t *out32,
                             int threads) {
    if (!orc_have_avx512() || msg_len > 135) return 0;
    batch_job j = {in, msg_len, stride, NULL, n, out32, 0, 1};
    run_batch(&j, threads);
    return 1;
}

void orc_keccak256_var(const uint8_t *data, const uint64_t *offsets, uint64_t n, uint8_t *out32,
                       int threads) {
    batch_job j = {data, 0, 0, offsets, n, out32, 0, 0};
    run_batch(&j, threads);
}
