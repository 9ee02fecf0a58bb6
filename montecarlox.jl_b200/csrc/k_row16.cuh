// k_row16.cuh -- the 16-site thread-row update shared by the streaming kernel (k_ising2d.cu) and the
// shared-memory-resident kernel (k_resident.cu): two Philox4x32-10 blocks, packed 15-bit SWAR decision
// against the pair-threshold table, exact 32-bit redo on ties.  See the header of k_ising2d.cu.
#pragma once
#include "mcx_internal.h"

namespace mcx {
namespace {

constexpr int kTableLen = 10;                 // 2 * (nn + 1) for nn = 4
constexpr int kPairRowWords = 256;            // pair table row stride (words): address = v1 * 256 + v0
constexpr int kPairWords = 9 * kPairRowWords + kTableLen;

struct Acc {
    uint32_t flips = 0;   // number of changed sites
    int32_t s = 0;        // sum over changed sites of s        (s in {0,1})
    int32_t n = 0;        // sum over changed sites of nup
    int32_t sn = 0;       // sum over changed sites of s*nup
};

// Exact redo of one thread-row with full 32-bit draws (taken when a 15-bit tie was seen).
template <bool HEATBATH>
__device__ __noinline__ uint4 row_exact(uint4 tq, uint4 nq, const uint32_t *thi, const uint32_t *tlo,
                                        Philox4 a0, Philox4 b0, Philox4 a1, Philox4 b1)
{
    uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t nup[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        const int w = i >> 2, b = i & 3;
        const uint32_t s = (tw[w] >> (8 * b)) & 0xffu, n = (nup[w] >> (8 * b)) & 0xffu;
        const int idx = (int)(s * 5 + n);
        const uint32_t hi = lane16(i < 8 ? a0 : b0, i & 7), lo = lane16(i < 8 ? a1 : b1, i & 7);
        const uint64_t m = ((uint64_t)hi << 16) | lo;
        const uint64_t T = ((uint64_t)thi[idx] << 16) | tlo[idx];
        const uint32_t lt = m < T ? 1u : 0u;
        const uint32_t sn = HEATBATH ? lt : (s ^ lt);
        tw[w] = (tw[w] & ~(0xffu << (8 * b))) | (sn << (8 * b));
    }
    return make_uint4(tw[0], tw[1], tw[2], tw[3]);
}

__device__ __forceinline__ uint32_t shr1_fma(uint32_t w)
{
#ifdef MCX_OPT_SHF
    return w >> 1;
#else
    return __umulhi(w, 0x80000000u);      // w >> 1 on the FMA pipe (the ALU pipe is the busy one)
#endif
}

// One thread-row: 16 target sites (tq), neighbour rows U (above), C (same row, other plane),
// D (below).  PARITY 0: the in-row neighbour pair of target byte j is other-plane bytes (j-1, j);
// PARITY 1: (j, j+1).  `side` is the other-plane byte just outside the segment on that side.
template <int PARITY, bool HEATBATH, bool TRACK>
__device__ __forceinline__ uint4 update_row(const uint4 tq, const uint4 U, const uint4 C, const uint4 D,
                                            const uint32_t side, const uint32_t blk, const uint32_t t_lo,
                                            const uint32_t c2, const uint32_t c2lo, const uint32_t chain_id,
                                            const uint32_t seed_lo, const uint32_t seed_hi,
                                            const uint32_t *s_pair, const uint32_t *s_thi, const uint32_t *s_tlo,
                                            Acc &acc, const bool active)
{
#ifdef MCX_OPT_NOCOMPUTE
    // memory-pattern ceiling probe (never shipped): same loads and stores, no RNG, no decision
    return make_uint4(tq.x ^ (U.x & C.x & D.x & side), tq.y ^ (U.y & C.y & D.y), tq.z ^ (U.z & C.z & D.z), tq.w ^ (U.w & C.w & D.w & blk));
#endif
#ifdef MCX_OPT_NOPHILOX
    // cost-split probe (never shipped): everything but the generator
    const Philox4 ra = Philox4{blk * 0x9E3779B9u ^ seed_lo, blk * 0x85EBCA6Bu + t_lo, blk * 0xC2B2AE35u ^ chain_id, blk * 0x27D4EB2Fu + c2};
    const Philox4 rb = Philox4{ra.x * 0x165667B1u, ra.y * 0x9E3779B1u, ra.z * 0x85EBCA77u, ra.w * 0xC2B2AE3Du};
#else
    const Philox4 ra = philox4x32_10(blk, t_lo, c2, chain_id, seed_lo, seed_hi);
    const Philox4 rb = philox4x32_10(blk + 1, t_lo, c2, chain_id, seed_lo, seed_hi);
#endif

    uint32_t S[4];
    if (PARITY == 0) {
        S[0] = (C.x << 8) | side;
        S[1] = __funnelshift_l(C.x, C.y, 8);
        S[2] = __funnelshift_l(C.y, C.z, 8);
        S[3] = __funnelshift_l(C.z, C.w, 8);
    } else {
        S[0] = __funnelshift_r(C.x, C.y, 8);
        S[1] = __funnelshift_r(C.y, C.z, 8);
        S[2] = __funnelshift_r(C.z, C.w, 8);
        S[3] = (C.w >> 8) | (side << 24);
    }
    const uint32_t nup[4] = {U.x + D.x + C.x + S[0], U.y + D.y + C.y + S[1], U.z + D.z + C.z + S[2],
                             U.w + D.w + C.w + S[3]};
    const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};   // word k: sites 2k, 2k+1
    uint32_t nw[4];
    // r = (h15 | 0x8000) - t15 per 16-bit half: bit 15 set <=> h15 >= t15 (not surely accepted); as a
    // signed halfword r is most negative (0x8000) exactly on a tie, so one packed signed min
    // (VIMNMX3.S16x2) accumulates the tie test for the whole thread-row.
    uint32_t tie_min = 0x7fff7fffu;
    int rejected = 0;                                            // minus the number of not-accepted sites
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t idx4 = tw[w] * 20u + nup[w] * 4u;        // byte b = 4 * (5 s + nup) of site 4w+b
        const uint32_t ttA = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair) + (idx4 & 0xffffu));
        const uint32_t ttB = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair) + (idx4 >> 16));
        const uint32_t hA = (shr1_fma(rw[2 * w]) & 0x7fff7fffu) | 0x80008000u;
        const uint32_t hB = (shr1_fma(rw[2 * w + 1]) & 0x7fff7fffu) | 0x80008000u;
#ifdef MCX_OPT_ALUSUB
        const uint32_t rA = hA - ttA, rB = hB - ttB;
#else
        uint32_t rA, rB;   // subtractions issued as IMAD
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rA) : "r"(ttA), "r"(hA));
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rB) : "r"(ttB), "r"(hB));
#endif
        tie_min = __vmins2(__vmins2(tie_min, rA), rB);
        // bytes 1,3 of rA and rB with sign replication: 0xFF per site that is NOT accepted, else 0x00
        uint32_t P;   // prmt with the selector msb set replicates the byte's sign bit (__byte_perm masks that bit off)
        asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(P) : "r"(rA), "r"(rB));
        nw[w] = HEATBATH ? (~P & 0x01010101u) : (tw[w] ^ (~P & 0x01010101u));
        if (!HEATBATH && !TRACK) rejected = __dp4a((int)P, 0x01010101, rejected);
    }
    const bool tie = ((tie_min & 0x7fffu) == 0u) || ((tie_min & 0x7fff0000u) == 0u);
    if (tie) {
        // rare (2^-15 per site): settle the whole thread-row with the full 32-bit draws
        const Philox4 la = philox4x32_10(blk, t_lo, c2lo, chain_id, seed_lo, seed_hi);
        const Philox4 lb = philox4x32_10(blk + 1, t_lo, c2lo, chain_id, seed_lo, seed_hi);
        const uint4 ex = row_exact<HEATBATH>(tq, make_uint4(nup[0], nup[1], nup[2], nup[3]), s_thi, s_tlo, ra, rb, la, lb);
        nw[0] = ex.x; nw[1] = ex.y; nw[2] = ex.z; nw[3] = ex.w;
    }
    if (active) {
        if (!HEATBATH && !TRACK && !tie) {
            acc.flips += (uint32_t)(16 + rejected);            // accepted == flipped for Metropolis / Glauber
        } else {
            uint32_t fsum = 0, ssum = 0, nsum = 0, snsum = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t Fc = nw[w] ^ tw[w];               // changed sites, 0x01 per byte
                fsum += Fc;
                if (TRACK) {
                    const uint32_t SF = tw[w] & Fc;
                    ssum += SF;
                    nsum += nup[w] & (Fc * 255u);
                    snsum += nup[w] & (SF * 255u);
                }
            }
            acc.flips = __dp4a(fsum, 0x01010101u, acc.flips);
            if (TRACK) {
                acc.s = __dp4a(ssum, 0x01010101u, (uint32_t)acc.s);
                acc.n = __dp4a(nsum, 0x01010101u, (uint32_t)acc.n);
                acc.sn = __dp4a(snsum, 0x01010101u, (uint32_t)acc.sn);
            }
        }
    }
    return make_uint4(nw[0], nw[1], nw[2], nw[3]);
}

__device__ __forceinline__ uint4 ldg128(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }


// Builds the pair table of one ensemble in shared memory: entry (i1, i0) = t15[i0] | t15[i1] << 16 with
// t15 = min(T >> 17, 0x7fff) (T = 2^32, "always", ties on h15 == 0x7fff and is settled exactly).
// Call between two __syncthreads().
__device__ __forceinline__ void load_pair_table(uint32_t *s_pair, uint32_t *s_thi, uint32_t *s_tlo,
                                                const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
                                                int label)
{
    if (threadIdx.x < kTableLen) {
        s_thi[threadIdx.x] = thi_g[label * kTableLen + threadIdx.x];
        s_tlo[threadIdx.x] = tlo_g[label * kTableLen + threadIdx.x];
    }
    if (threadIdx.x < kTableLen * kTableLen) {
        const int i1 = threadIdx.x / kTableLen, i0 = threadIdx.x - i1 * kTableLen;
        const uint32_t a = min(thi_g[label * kTableLen + i0] >> 1, 0x7fffu);
        const uint32_t b = min(thi_g[label * kTableLen + i1] >> 1, 0x7fffu);
        s_pair[i1 * kPairRowWords + i0] = a | (b << 16);
    }
}

}  // namespace
}  // namespace mcx
