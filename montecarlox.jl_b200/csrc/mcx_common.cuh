// mcx_common.cuh -- shared device helpers for libmcx_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcx_b200.h"

namespace mcx {

enum : uint32_t { TAG_SWEEP = 0, TAG_EXCHANGE = 1, TAG_INIT = 2, TAG_FLAT = 3 };

// ---------------------------------------------------------------------------------------------
// Philox4x32-10.  One call yields 128 bits = eight 16-bit lanes = the primary (high-half) draw
// of eight consecutive slots (RNG layout v1, include/mcx_b200.h).
// ---------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                                      uint32_t k0, uint32_t k1)
{
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// The same block, split by what its inputs depend on.  All lanes of a work item share c1, c2, c3 and the key and differ
// in c0 only, and for two rounds half of the state does not see c0 yet: PhiloxHead holds those words, folded with the
// round keys they are xored with, and the keys of the later rounds, so that a kernel forms them once per item (they live
// in uniform registers) and a block costs 18 multiplies and 19 xors per lane.  philox_tail(philox_head(c1, c2, c3, k0, k1), c0)
// == philox4x32_10(c0, c1, c2, c3, k0, k1).
struct PhiloxHead {
    uint32_t x1, x2, x3, x4;   // c3 ^ K1[0];  lo(M1 c2) ^ K0[1];  hi(M0 a) ^ K1[1];  lo(M0 a) ^ K1[2], a = hi(M1 c2) ^ c1 ^ K0[0]
    uint32_t k0[8], k1[7];     // K0[2..9], K1[3..9]
};

__host__ __device__ __forceinline__ void mul_wide(uint32_t m, uint32_t c, uint32_t &hi, uint32_t &lo)
{
    const uint64_t p = (uint64_t)m * c;                        // one IMAD.WIDE.U32
    hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
}

__host__ __device__ __forceinline__ PhiloxHead philox_head(uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    PhiloxHead H;
    uint32_t h1, l1, h0, l0;
    mul_wide(0xCD9E8D57u, c2, h1, l1);
    const uint32_t a = h1 ^ c1 ^ k0;
    mul_wide(0xD2511F53u, a, h0, l0);
    H.x1 = c3 ^ k1;
    H.x2 = l1 ^ (k0 + 0x9E3779B9u);
    H.x3 = h0 ^ (k1 + 0xBB67AE85u);
    H.x4 = l0 ^ (k1 + 2u * 0xBB67AE85u);
#pragma unroll
    for (int r = 0; r < 8; ++r) H.k0[r] = k0 + (uint32_t)(r + 2) * 0x9E3779B9u;
#pragma unroll
    for (int r = 0; r < 7; ++r) H.k1[r] = k1 + (uint32_t)(r + 3) * 0xBB67AE85u;
    return H;
}

__host__ __device__ __forceinline__ Philox4 philox_tail(const PhiloxHead &H, uint32_t c0)
{
    uint32_t h0, l0, h1, l1;
    mul_wide(0xD2511F53u, c0, h0, l0);                         // round 1: c2' = h0 ^ x1, c3' = l0
    mul_wide(0xCD9E8D57u, h0 ^ H.x1, h1, l1);                  // round 2: c0'' = h1 ^ x2, c1'' = l1, c2'' = l0 ^ x3, c3'' = lo(M0 a)
    uint32_t c0_ = h1 ^ H.x2, c1_ = l1, c2_ = l0 ^ H.x3, c3_;
    mul_wide(0xD2511F53u, c0_, h0, l0);                        // round 3
    mul_wide(0xCD9E8D57u, c2_, h1, l1);
    c0_ = h1 ^ c1_ ^ H.k0[0]; c1_ = l1; c2_ = h0 ^ H.x4; c3_ = l0;
#pragma unroll
    for (int r = 1; r < 8; ++r) {                              // rounds 4..10
        mul_wide(0xD2511F53u, c0_, h0, l0);
        mul_wide(0xCD9E8D57u, c2_, h1, l1);
        c0_ = h1 ^ c1_ ^ H.k0[r]; c1_ = l1; c2_ = h0 ^ c3_ ^ H.k1[r - 1]; c3_ = l0;
    }
    return Philox4{c0_, c1_, c2_, c3_};
}

// counter word 2 of the stream layout
__host__ __device__ __forceinline__ uint32_t ctr_word2(uint64_t t, uint32_t plane, uint32_t tag)
{
    return (uint32_t)((t >> 32) & 0xffffu) | (plane << 16) | (tag << 24);
}

__host__ __device__ __forceinline__ Philox4 stream_block(uint32_t seed_lo, uint32_t seed_hi, uint32_t chain,
                                                         uint32_t tag, uint64_t t, uint32_t blk, uint32_t plane)
{
    return philox4x32_10(blk, (uint32_t)t, ctr_word2(t, plane, tag), chain, seed_lo, seed_hi);
}

__host__ __device__ __forceinline__ uint32_t lane16(const Philox4 &p, int lane)
{
    const uint32_t w = (lane >> 1) == 0 ? p.x : (lane >> 1) == 1 ? p.y : (lane >> 1) == 2 ? p.z : p.w;
    return (w >> (16 * (lane & 1))) & 0xffffu;
}

// ---------------------------------------------------------------------------------------------
// lattice view handed to kernels
// ---------------------------------------------------------------------------------------------
struct LatView {
    uint8_t *planes;        // [nchains][2][plane_stride] one byte per spin, colour planes
    int64_t plane_stride;   // bytes between planes
    int64_t halfN;          // sites per colour plane
    int32_t Lx, Ly, Lz, half;   // half = Lx/2
    int32_t ndim, nn, model, nchains;
    // slab decomposition (k_slab.cu): this view holds rows [row_offset, row_offset + Ly) of a taller
    // lattice; the row above row 0 / below row Ly-1 lives in the plane arrays of the neighbour slabs
    // (same shape and strides).  Not a slab: up_planes == dn_planes == planes, row_offset == 0.
    uint8_t *up_planes, *dn_planes;
    int32_t row_offset;
    int32_t slab_sides;     // which neighbour GPUs this launch orders itself against: 1 = up, 2 = down (bands: one side each)
    // slabs whose neighbours are other processes / GPUs: control block in this slab's memory (k_slab.cu);
    // null otherwise.  [0], [1]: half-sweeps whose boundary rows the up / down neighbour has finished (they
    // write it); [2]: half-sweep index at attach; [3], [4]: addresses of "my" counters in the up / down
    // neighbour's block; [5]: arrival counter of the boundary CTAs (launches ordered against the up side or
    // both); [6]: raised when a wait gave up; [7]: arrival counter of launches ordered against the down side only.
    unsigned long long *slab_ctl;
    int *err;               // mcx_ctx::d_err: sticky error word of the device-side waits
};
enum { SLAB_FLAG_UP = 0, SLAB_FLAG_DN = 1, SLAB_T0 = 2, SLAB_UP_SLOT = 3, SLAB_DN_SLOT = 4, SLAB_ARRIVED = 5, SLAB_ERR = 6, SLAB_ARRIVED_DN = 7, SLAB_CTL_WORDS = 8 };

__device__ __forceinline__ uint8_t *plane_ptr(const LatView &L, int chain, int colour)
{
    return L.planes + ((int64_t)chain * 2 + colour) * L.plane_stride;
}
__device__ __forceinline__ const uint8_t *plane_ptr_of(const uint8_t *planes, const LatView &L, int chain, int colour)
{
    return planes + ((int64_t)chain * 2 + colour) * L.plane_stride;
}

// Device-side clock of a parallel-tempering handle: what changes from round to round.  The kernels of a round captured into a
// CUDA graph (mcx_pt_run) read it instead of taking the half-sweep index and the round as launch arguments, so one
// instantiated graph serves every round; the last node of a round advances it.
struct PtClock {
    unsigned long long t_base;   // half-sweep index of the round's first half-sweep
    unsigned long long round;    // exchange round
};

// per-chain accumulator block
enum { SUM_PAIR = 0, SUM_SPIN = 1, SUM_SPIN2 = 2, SUM_ACC = 3, SUM_FIELDS = 4 };

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mcx
