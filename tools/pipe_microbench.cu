// pipe_microbench.cu — issue-rate probes for the integer instructions Keccak-f is made of, on sm_100a.
// Answers: (1) LOP3 / SHF lanes per clock per SM (the ALU-pipe ceiling the keccak kernels are measured against),
// (2) whether 64-bit rotations expressed as IMAD.WIDE / IMAD.HI (FMA pipe) can be co-issued with LOP3 so that
// the rotation work leaves the ALU pipe.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__constant__ uint32_t MUL[8];

#define ITER 4096
#define ACC 8

template <int MODE>
__global__ void probe(uint32_t *out, unsigned long long *cycles) {
    uint32_t a[ACC], b[ACC];
    uint64_t w[ACC];
    for (int i = 0; i < ACC; i++) {
        a[i] = threadIdx.x * 2654435761u + i;
        b[i] = a[i] ^ 0x9e3779b9u;
        w[i] = ((uint64_t)a[i] << 32) | b[i];
    }
    uint32_t m0 = MUL[0], m1 = MUL[1];
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) {
            if (MODE == 0) {  // LOP3 only: 2 per accumulator
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xd2;" : "+r"(b[i]) : "r"(a[i]), "r"(m1));
            } else if (MODE == 1) {  // SHF only: 2 per accumulator
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(b[i]) : "r"(a[i]));
            } else if (MODE == 2) {  // IMAD.WIDE only: 2 per accumulator
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m0));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(m1));
            } else if (MODE == 3) {  // 2 LOP3 + 1 IMAD.WIDE
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xd2;" : "+r"(b[i]) : "r"(a[i]), "r"(m1));
            } else if (MODE == 4) {  // 2 LOP3 + 1 SHF (today's keccak mix)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xd2;" : "+r"(b[i]) : "r"(a[i]), "r"(m1));
            } else if (MODE == 5) {  // mul.hi.u32 only: 2 per accumulator
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m0));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m1));
            } else if (MODE == 6) {  // mad.lo.u32 only: 2 per accumulator
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m0), "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m1), "r"(a[i]));
            } else if (MODE == 7) {  // 2 LOP3 + 1 mad.lo
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"((uint32_t &)w[i]) : "r"(a[i]), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xd2;" : "+r"(b[i]) : "r"(a[i]), "r"(m1));
            } else if (MODE == 8) {  // 2 LOP3 + 2 IMAD.WIDE (fully rotation-on-FMA mix)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xd2;" : "+r"(b[i]) : "r"(a[i]), "r"(m1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(m1));
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < ACC; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ops_per_acc, int sms, int warps_per_sm) {
    int block = 256, blocks = sms * warps_per_sm * 32 / block;
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, (size_t)blocks * block * 4);
    cudaMalloc(&cyc, blocks * 8);
    probe<MODE><<<blocks, block>>>(out, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, block>>>(out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long *h = new unsigned long long[blocks];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += h[i];
    avg /= blocks;
    double ops_per_sm = (double)warps_per_sm * 32 * ITER * ACC * ops_per_acc;
    printf("%-28s warps/SM=%2d  %7.1f lane-ops/clk/SM  (%.3f ms, %.0f cyc)  err=%s\n", name, warps_per_sm,
           ops_per_sm / avg, ms, avg, cudaGetErrorString(cudaGetLastError()));
    delete[] h;
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t mul[8] = {8, 1u << 13, 3, 5, 7, 9, 11, 13};
    cudaMemcpyToSymbol(MUL, mul, sizeof mul);
    printf("SMs=%d\n", sms);
    for (int w : {8, 16, 32}) {
        run<0>("LOP3", 2, sms, w);
        run<1>("SHF", 2, sms, w);
        run<2>("IMAD.WIDE", 2, sms, w);
        run<5>("IMAD.HI (mul.hi)", 2, sms, w);
        run<6>("IMAD (mad.lo)", 2, sms, w);
        run<4>("2 LOP3 + 1 SHF", 3, sms, w);
        run<3>("2 LOP3 + 1 IMAD.WIDE", 3, sms, w);
        run<7>("2 LOP3 + 1 IMAD", 3, sms, w);
        run<8>("2 LOP3 + 2 IMAD.WIDE", 4, sms, w);
    }
    return 0;
}
