// k_strip.cuh -- one thread's walk down a strip of rows of a 2-D Ising colour plane: the loop the streaming kernel
// (k_ising2d.cu), the ticket-queue series kernel (k_queue.cu) and the persistent parallel-tempering rounds (k_persist.cu)
// share.  A thread owns a 16-byte column segment and takes two rows per trip; the rows of the other colour plane slide
// through a register window; the neighbour byte outside the segment comes from the adjacent lane by shuffle (edge lanes
// load it).
//
// One basic block per trip.  A row whose packed 15-bit comparison leaves a site undecided (2^-15 per site) is NOT stored;
// it is flagged and settled after the loop from global memory with the full 32-bit draws (settle_rows): rows of one colour
// do not see each other and the other plane does not change during a half-sweep, so the late row computes exactly what it
// would have computed in place.  Without a branch or a call inside the loop the compiler runs the Philox rounds of one row
// under the decision arithmetic of the other, and the loop fits 64 registers (8 CTAs of 128 threads per SM).
//
// The lane-independent part of the Philox blocks (PhiloxHead, mcx_common.cuh) is formed once per strip.
#pragma once
#include "k_row16.cuh"

namespace mcx {
namespace {

// L2ONLY: rows that other SMs rewrite during the same launch (series kernels) are read and written through L2
// (ld / st.global.cg; L1 is not coherent across SMs); the streaming kernel uses plain accesses.
template <bool L2ONLY>
__device__ __forceinline__ uint4 strip_ld128(const uint8_t *p)
{
    if (!L2ONLY) return *reinterpret_cast<const uint4 *>(p);
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <bool L2ONLY>
__device__ __forceinline__ uint32_t strip_ld8(const uint8_t *p)
{
    if (!L2ONLY) return *p;
    uint32_t v;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <bool L2ONLY>
__device__ __forceinline__ void strip_st128(uint8_t *p, const uint4 v)
{
    if (!L2ONLY) { *reinterpret_cast<uint4 *>(p) = v; return; }
    asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// where a thread's strip lies: rows [row0, row0 + rows) (row0 even, rows even and <= 32), bytes [col, col + 16) of each row
struct StripGeom {
    int row0, rows, col;
    int colL, colR;          // the byte left / right of the segment (periodic)
    bool loadL, loadR;       // that byte is not in the adjacent lane's segment: load it
    bool active;             // false: a padding lane (computes on a valid strip, stores nothing)
};

// The rows of a strip that the loop left undecided (the loop shifts two flags per trip into `ties`, row a above row b),
// redone one by one.  Returns what they add to the packed accumulators (flips, s, n, sn).  Everything comes by value: a
// reference to the kernel's LatView or Acc would put them into local memory for the whole kernel.
template <int COLOUR, bool HEATBATH, bool TRACK, bool L2ONLY>
__device__ __noinline__ uint4 settle_rows(uint8_t *tgt, const uint8_t *oth, const uint8_t *oth_up, const uint8_t *oth_dn,
                                          const int half, const int Ly, const int row_offset, uint32_t ties, const int row0, const int R,
                                          const int col, const int colL, const int colR, const uint32_t t_lo, const uint32_t c2,
                                          const uint32_t c2lo, const uint32_t chain_id, const uint32_t seed_lo,
                                          const uint32_t seed_hi, const uint32_t *s_thi, const uint32_t *s_tlo)
{
    const size_t h = (uint32_t)half;
    Acc acc;
    while (ties) {
        const int p = __ffs((int)ties) - 1;
        ties &= ties - 1;
        const int r = R - 2 - (p & ~1) + (~p & 1);
        const int row = row0 + r;
        const int parity = (r & 1) ? (COLOUR ^ 1) : COLOUR;      // row0 is even
        const uint8_t *pu = row == 0 ? oth_up + (size_t)(Ly - 1) * h : oth + (size_t)(row - 1) * h;
        const uint8_t *pd = row + 1 == Ly ? oth_dn : oth + (size_t)(row + 1) * h;
        const uint8_t *pc = oth + (size_t)row * h;
        uint8_t *pt = tgt + (size_t)row * h + col;
        const uint32_t side = strip_ld8<L2ONLY>(pc + (parity == 0 ? colL : colR));
        const uint32_t blk = (uint32_t)(((int64_t)(row + row_offset) * half + col) >> 3);
        const uint4 ex = row_settle<HEATBATH, TRACK>(parity, strip_ld128<L2ONLY>(pt), strip_ld128<L2ONLY>(pu + col),
                                                     strip_ld128<L2ONLY>(pc + col), strip_ld128<L2ONLY>(pd + col), side, blk, t_lo, c2,
                                                     c2lo, chain_id, seed_lo, seed_hi, s_thi, s_tlo, acc);
        strip_st128<L2ONLY>(pt, ex);
    }
    return make_uint4(acc.flips, (uint32_t)acc.s, (uint32_t)acc.n, (uint32_t)acc.sn);
}

// Half-sweep t of one strip: tgt / oth are the chain's target and other colour plane, oth_up / oth_dn the planes that hold
// the row above row 0 / below row Ly - 1 (the same plane unless the view is a slab or a row band), row_offset the global
// row of row 0 (Philox counters are positioned by global row).  Returns the thread's packed accumulators.
template <int COLOUR, bool HEATBATH, bool TRACK, bool L2ONLY>
__device__ __forceinline__ Acc sweep_strip(uint8_t *tgt, const uint8_t *__restrict__ oth, const uint8_t *oth_up, const uint8_t *oth_dn,
                                           const int half, const int Ly, const int row_offset, const StripGeom &g, const uint64_t t,
                                           const uint32_t chain_id, const uint32_t seed_lo, const uint32_t seed_hi,
                                           const uint32_t *s_pair, const uint32_t *s_thi, const uint32_t *s_tlo)
{
    const size_t h = (uint32_t)half, h2 = 2 * h;                 // row pitch, zero-extended once
    const uint32_t t_lo = (uint32_t)t;
    const uint32_t c2 = ctr_word2(t, 0, TAG_SWEEP), c2lo = ctr_word2(t, 1, TAG_SWEEP);
    const int row0 = g.row0, R = g.rows, col = g.col;
    const bool active = g.active;
    // even rows of the strip have parity COLOUR, odd rows COLOUR ^ 1 (row0 is even)
    const bool edgeA = COLOUR == 0 ? g.loadL : g.loadR;
    const bool edgeB = COLOUR == 0 ? g.loadR : g.loadL;
    // the edge bytes relative to the thread's own segment: row a, and row b one pitch further
    const ptrdiff_t offA = (ptrdiff_t)((COLOUR == 0 ? g.colL : g.colR) - col);
    const ptrdiff_t offB = (ptrdiff_t)((COLOUR == 0 ? g.colR : g.colL) - col) + (ptrdiff_t)h;

    // Per-thread row pointers (segment included), advanced by two pitches per trip: every address of a trip is one of them,
    // or one of them plus the pitch.
    const int rowU = row0 == 0 ? Ly - 1 : row0 - 1;
    const uint8_t *po = oth + (size_t)row0 * h + col;            // other plane, current even row
    uint8_t *pt = tgt + (size_t)row0 * h + col;                  // target plane, current even row
    uint4 U = strip_ld128<L2ONLY>((row0 == 0 ? oth_up : oth) + (size_t)rowU * h + col);
    uint4 C = strip_ld128<L2ONLY>(po);
    const bool wraps = row0 + R == Ly;                           // the row below the strip's last row is row 0 of oth_dn
    uint32_t blk = (uint32_t)(((int64_t)(row0 + row_offset) * half + col) >> 3);
    const uint32_t blk_step = (uint32_t)(half >> 3), blk_step2 = 2 * blk_step;
    const PhiloxHead H = philox_head(t_lo, c2, chain_id, seed_lo, seed_hi);
    const uint32_t pair_addr = (uint32_t)__cvta_generic_to_shared(s_pair);
    Acc acc;
    uint32_t ties = 0;                                           // rows left to settle_rows()
    // ONE prefetch per trip pulls the next trip's four new row segments into L2 (a hint: no registers wait for it): lane l
    // asks for row (l & 3) of {other + 3, other + 4, target + 2, target + 3} at the segment of lane (l & ~3), so the four
    // 128-byte lines of each row are each named by two lanes.  Past the strip the addresses fall into the next strip (or
    // the slack rows behind the last plane, mcx_lattice_create).  +1 % at L = 16384; into L1 or two rows further ahead:
    // no better (profiles/r02_call27_28_prefetch.log).
    const int pf_k = threadIdx.x & 3;
    const uint8_t *ppf = (pf_k < 2 ? oth : (const uint8_t *)tgt) + (size_t)(row0 + (pf_k == 1 ? 4 : pf_k == 2 ? 2 : 3)) * h + (col - 16 * pf_k);

    // two trips per loop iteration in the streaming kernel: the window hand-over (U = D, C = E) becomes renaming and the round
    // keys are formed once per two trips (398 against 411 instructions per trip, +1.1 % at L = 16384); the series kernels at
    // 5 CTAs/SM lose 3 % with it (profiles/r02_call44_unroll2.log)
    constexpr int kTripsPerIteration = L2ONLY ? 1 : 2;
#pragma unroll (kTripsPerIteration)
    for (int r = 0; r < R; r += 2) {
        // this trip's rows, loads first: their latency is covered by the Philox rounds below
        const uint8_t *pe = po + h2;
        if (wraps && r + 2 == R) pe = oth_dn + col;
        const uint4 E = strip_ld128<L2ONLY>(pe);
        const uint4 D = strip_ld128<L2ONLY>(po + h);
        const uint4 Ta = strip_ld128<L2ONLY>(pt), Tb = strip_ld128<L2ONLY>(pt + h);
        if (!L2ONLY) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ppf));
            ppf += h2;
        }
        uint32_t sideA = 0, sideB = 0;
        if (edgeA) sideA = strip_ld8<L2ONLY>(po + offA);
        if (edgeB) sideB = strip_ld8<L2ONLY>(po + offB);
        uint32_t sA, sB;
        if (COLOUR == 0) {
            sA = __shfl_up_sync(0xffffffffu, C.w, 1) >> 24;
            sB = __shfl_down_sync(0xffffffffu, D.x, 1) & 0xffu;
        } else {
            sA = __shfl_down_sync(0xffffffffu, C.x, 1) & 0xffu;
            sB = __shfl_up_sync(0xffffffffu, D.w, 1) >> 24;
        }
        if (edgeA) sA = sideA;
        if (edgeB) sB = sideB;
        bool tieA, tieB;
        const uint4 Na = update_row_fast<COLOUR, HEATBATH, TRACK>(philox_tail(H, blk), philox_tail(H, blk + 1), Ta, U, C, D, sA,
                                                                  pair_addr, acc, active, tieA);
        if (active && !tieA) strip_st128<L2ONLY>(pt, Na);
        const uint4 Nb = update_row_fast<COLOUR ^ 1, HEATBATH, TRACK>(philox_tail(H, blk + blk_step), philox_tail(H, blk + blk_step + 1),
                                                                      Tb, C, D, E, sB, pair_addr, acc, active, tieB);
        if (active && !tieB) strip_st128<L2ONLY>(pt + h, Nb);
        ties = (ties << 2) | (tieA ? 2u : 0u) | (tieB ? 1u : 0u);   // trip k of K: bits 2 (K - 1 - k) + 1 (row a), + 0 (row b)
        U = D; C = E;
        po = pe; pt += h2; blk += blk_step2;
    }
    if (active && ties) {
        const uint4 d = settle_rows<COLOUR, HEATBATH, TRACK, L2ONLY>(tgt, oth, oth_up, oth_dn, half, Ly, row_offset, ties, row0, R, col, g.colL,
                                                                     g.colR, t_lo, c2, c2lo, chain_id, seed_lo, seed_hi, s_thi, s_tlo);
        acc.flips += d.x; acc.s += (int32_t)d.y; acc.n += (int32_t)d.z; acc.sn += (int32_t)d.w;
    }
    return acc;
}

// per-chain sums of a strip: dspin = 2 - 4 s, dpair = -8 s nup + 16 s + 4 nup - 8 per changed site; one atomic per warp
template <bool TRACK>
__device__ __forceinline__ void strip_finish(const Acc &acc, long long *__restrict__ sums, const int chain)
{
    const int nflip = warp_sum((int)acc.flips);
    int dspin = 0, dpair = 0;
    if (TRACK) {
        const int ss = warp_sum(acc.s), nn_ = warp_sum(acc.n), sn = warp_sum(acc.sn);
        dspin = 2 * nflip - 4 * ss;
        dpair = -8 * sn + 16 * ss + 4 * nn_ - 8 * nflip;
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
        if (nflip) atomicAdd(o + SUM_ACC, (unsigned long long)(long long)nflip);
        if (TRACK) {
            if (dpair) atomicAdd(o + SUM_PAIR, (unsigned long long)(long long)dpair);
            if (dspin) atomicAdd(o + SUM_SPIN, (unsigned long long)(long long)dspin);
        }
    }
}

}  // namespace
}  // namespace mcx
