/*
 * keccak_avx512.c — 8-way multi-buffer Keccak-256 for short fixed-length messages (AVX-512F), CPU baseline only.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle.h).
 *
 * reth itself hashes keys one at a time with scalar assembly (keccak-asm, bin/reth/Cargo.toml:85-94); this is the
 * "fastest SIMD Keccak we can write" figure BASELINE.md §2 asks to report next to the scalar port: eight sponges in
 * the eight 64-bit elements of a zmm register, vprolq for rho, vpternlogq for the 3-input XOR (0x96) and chi (0xD2).
 * Checked against the scalar oracle in tests/test_oracle_golden.py.
 */
#include "oracle.h"
#include <immintrin.h>
#include <string.h>

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

#define XOR3(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0x96)
#define CHI(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0xD2)
#define ROL(x, n) _mm512_rol_epi64(x, n)

__attribute__((target("avx512f"))) static void keccak_f1600_x8(__m512i a[25]) {
    for (int r = 0; r < 24; r++) {
        __m512i c0 = XOR3(XOR3(a[0], a[5], a[10]), a[15], a[20]);
        __m512i c1 = XOR3(XOR3(a[1], a[6], a[11]), a[16], a[21]);
        __m512i c2 = XOR3(XOR3(a[2], a[7], a[12]), a[17], a[22]);
        __m512i c3 = XOR3(XOR3(a[3], a[8], a[13]), a[18], a[23]);
        __m512i c4 = XOR3(XOR3(a[4], a[9], a[14]), a[19], a[24]);
        __m512i r0 = ROL(c1, 1), r1 = ROL(c2, 1), r2 = ROL(c3, 1), r3 = ROL(c4, 1), r4 = ROL(c0, 1);
        __m512i b00 = XOR3(a[0], c4, r0);
        __m512i b10 = ROL(XOR3(a[1], c0, r1), 1);
        __m512i b20 = ROL(XOR3(a[2], c1, r2), 62);
        __m512i b05 = ROL(XOR3(a[3], c2, r3), 28);
        __m512i b15 = ROL(XOR3(a[4], c3, r4), 27);
        __m512i b16 = ROL(XOR3(a[5], c4, r0), 36);
        __m512i b01 = ROL(XOR3(a[6], c0, r1), 44);
        __m512i b11 = ROL(XOR3(a[7], c1, r2), 6);
        __m512i b21 = ROL(XOR3(a[8], c2, r3), 55);
        __m512i b06 = ROL(XOR3(a[9], c3, r4), 20);
        __m512i b07 = ROL(XOR3(a[10], c4, r0), 3);
        __m512i b17 = ROL(XOR3(a[11], c0, r1), 10);
        __m512i b02 = ROL(XOR3(a[12], c1, r2), 43);
        __m512i b12 = ROL(XOR3(a[13], c2, r3), 25);
        __m512i b22 = ROL(XOR3(a[14], c3, r4), 39);
        __m512i b23 = ROL(XOR3(a[15], c4, r0), 41);
        __m512i b08 = ROL(XOR3(a[16], c0, r1), 45);
        __m512i b18 = ROL(XOR3(a[17], c1, r2), 15);
        __m512i b03 = ROL(XOR3(a[18], c2, r3), 21);
        __m512i b13 = ROL(XOR3(a[19], c3, r4), 8);
        __m512i b14 = ROL(XOR3(a[20], c4, r0), 18);
        __m512i b24 = ROL(XOR3(a[21], c0, r1), 2);
        __m512i b09 = ROL(XOR3(a[22], c1, r2), 61);
        __m512i b19 = ROL(XOR3(a[23], c2, r3), 56);
        __m512i b04 = ROL(XOR3(a[24], c3, r4), 14);
        a[0] = _mm512_xor_si512(CHI(b00, b01, b02), _mm512_set1_epi64((long long)RC[r]));
        a[1] = CHI(b01, b02, b03);
        a[2] = CHI(b02, b03, b04);
        a[3] = CHI(b03, b04, b00);
        a[4] = CHI(b04, b00, b01);
        a[5] = CHI(b05, b06, b07);
        a[6] = CHI(b06, b07, b08);
        a[7] = CHI(b07, b08, b09);
        a[8] = CHI(b08, b09, b05);
        a[9] = CHI(b09, b05, b06);
        a[10] = CHI(b10, b11, b12);
        a[11] = CHI(b11, b12, b13);
        a[12] = CHI(b12, b13, b14);
        a[13] = CHI(b13, b14, b10);
        a[14] = CHI(b14, b10, b11);
        a[15] = CHI(b15, b16, b17);
        a[16] = CHI(b16, b17, b18);
        a[17] = CHI(b17, b18, b19);
        a[18] = CHI(b18, b19, b15);
        a[19] = CHI(b19, b15, b16);
        a[20] = CHI(b20, b21, b22);
        a[21] = CHI(b21, b22, b23);
        a[22] = CHI(b22, b23, b24);
        a[23] = CHI(b23, b24, b20);
        a[24] = CHI(b24, b20, b21);
    }
}

int orc_have_avx512(void) { return __builtin_cpu_supports("avx512f"); }

/* eight messages of msg_len (<= 135) bytes at in + k*stride -> eight digests at out + 32*k */
__attribute__((target("avx512f"))) void orc_keccak256_x8(const uint8_t *in, uint32_t msg_len, uint32_t stride,
                                                         uint8_t *out) {
    uint64_t blk[8][17];
    for (int k = 0; k < 8; k++) {
        uint8_t *b = (uint8_t *)blk[k];
        memcpy(b, in + (size_t)k * stride, msg_len);
        memset(b + msg_len, 0, 136 - msg_len);
        b[msg_len] ^= 0x01;
        b[135] ^= 0x80;
    }
    __m512i a[25];
    for (int l = 0; l < 17; l++)
        a[l] = _mm512_set_epi64((long long)blk[7][l], (long long)blk[6][l], (long long)blk[5][l], (long long)blk[4][l],
                                (long long)blk[3][l], (long long)blk[2][l], (long long)blk[1][l], (long long)blk[0][l]);
    for (int l = 17; l < 25; l++) a[l] = _mm512_setzero_si512();
    keccak_f1600_x8(a);
    uint64_t o[4][8];
    for (int l = 0; l < 4; l++) _mm512_storeu_si512((void *)o[l], a[l]);
    for (int k = 0; k < 8; k++)
        for (int l = 0; l < 4; l++) memcpy(out + 32 * k + 8 * l, &o[l][k], 8);
}
