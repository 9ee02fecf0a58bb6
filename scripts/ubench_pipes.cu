// Micro-benchmark: issue throughput of the integer instructions the half-sweep kernel lives on
// (sm_100a).  Prints warp-instructions per cycle per SM sub-partition for each instruction class.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    uint32_t k0 = seed ^ 0x9E3779B9u;
    uint32_t kt = k0 + threadIdx.x * 0x9E3779B9u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {               // IMAD.WIDE.U32 (both halves consumed)
                const uint64_t p = (uint64_t)a[i] * 0xD2511F53u;
                a[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;
            } else if (MODE == 1) {        // IMAD (low 32 bits)
                a[i] = a[i] * 0xD2511F53u + k0;
            } else if (MODE == 2) {        // LOP3
                a[i] = (a[i] ^ k0) & (a[(i + 1) & 7] | 0x55555555u);
            } else if (MODE == 3) {        // IMAD.HI.U32
                a[i] = __umulhi(a[i], 0xD2511F53u) + 1u;
            } else if (MODE == 4) {        // Philox-like: WIDE + LOP3
                const uint64_t p = (uint64_t)a[i] * 0xD2511F53u;
                a[i] = (uint32_t)(p >> 32) ^ (uint32_t)p ^ k0;
            } else if (MODE == 5) {        // IADD3
                a[i] = a[i] + a[(i + 1) & 7] + k0;
            } else if (MODE == 6) {        // PRMT
                a[i] = __byte_perm(a[i], a[(i + 1) & 7], 0x7531);
            } else if (MODE == 7) {        // SHF (funnel)
                a[i] = __funnelshift_l(a[i], a[(i + 1) & 7], 8);
            } else if (MODE == 8) {        // Philox-like with a per-thread (non-uniform) key
                const uint64_t p = (uint64_t)a[i] * 0xD2511F53u;
                a[i] = (uint32_t)(p >> 32) ^ (uint32_t)p ^ kt;
            } else if (MODE == 9) {        // two WIDE per LOP3
                const uint64_t p = (uint64_t)a[i] * 0xD2511F53u;
                const uint64_t q = (uint64_t)(uint32_t)p * 0xCD9E8D57u;
                a[i] = (uint32_t)(p >> 32) ^ (uint32_t)q ^ (uint32_t)(q >> 32);
            } else if (MODE == 10) {       // WIDE only (result folded by IMAD lo)
                const uint64_t p = (uint64_t)a[i] * 0xD2511F53u;
                a[i] = (uint32_t)(p >> 32) * 3u + (uint32_t)p;
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int per_iter_instr)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    uint32_t *out;
    const int blocks = sms * 8, threads = 256, iters = 20000;
    cudaMalloc(&out, sizeof(uint32_t) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 100, 1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * threads / 32 * iters * 8 * per_iter_instr;
    const double cycles = ms * 1e-3 * khz * 1e3;     // at the max clock
    printf("%-28s %.3f warp-instr/clk/SMSP (assuming %d MHz)  %.2f ms\n", name, warp_instr / cycles / sms / 4, khz / 1000, ms);
    cudaFree(out);
}

int main()
{
    run<0>("IMAD.WIDE.U32 (+LOP3)", 2);
    run<1>("IMAD lo", 1);
    run<2>("LOP3 x2", 2);
    run<3>("IMAD.HI (+IADD)", 2);
    run<4>("WIDE + LOP3(3-in)", 2);
    run<5>("IADD3", 1);
    run<6>("PRMT", 1);
    run<7>("SHF funnel", 1);
    run<8>("WIDE + LOP3(3-in, reg key)", 2);
    run<9>("2 WIDE + LOP3", 3);
    run<10>("WIDE + IMAD", 2);
    return 0;
}
