// k_ising3d.cu -- vectorised checkerboard half-sweep for 3-D Ising lattices, one byte per spin.
//
// The k_ising2d decomposition carried to three dimensions: rows are (z, y) pairs, a thread owns a 16-byte
// column segment and walks a strip of rows inside ONE z-plane, two rows per trip, with the y-neighbour rows
// of the other colour plane in a rolling register window; the two z-neighbour rows of every row are two more
// 128-bit loads.  The in-row neighbour pair of target byte j is other-plane bytes (j-1, j) or (j, j+1)
// depending on the parity of colour + y + z, which is not known at compile time here (z varies per strip).
// Randomness and decision exactly as in k_ising2d (two Philox4x32-10 blocks per thread-row, packed 15-bit
// comparison against a pair-threshold table, exact redo on ties), with #up-neighbours in 0..6:
// table index s * 7 + nup.  Bit-identical to k_sweep_rows8 and k_sweep_generic.
#include "k_bits.cuh"
#include "mcx_internal.h"

#include <cstdlib>

namespace mcx {

namespace {

constexpr int k3Threads = 128;
constexpr int k3Table = 14;                   // 2 * (nn + 1) for nn = 6
constexpr int k3RowWords = 256;               // pair table row stride (words): address = v1 * 256 + v0
constexpr int k3PairWords = (k3Table - 1) * k3RowWords + k3Table;

struct Acc3 {
    uint32_t flips = 0;
    int32_t s = 0, n = 0, sn = 0;             // over changed sites: sum s, sum nup, sum s * nup  (s in {0, 1})
};

template <bool HEATBATH>
__device__ __noinline__ uint4 row_exact3(uint4 tq, uint4 nq, const uint32_t *thi, const uint32_t *tlo,
                                         Philox4 a0, Philox4 b0, Philox4 a1, Philox4 b1)
{
    uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t nup[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        const int w = i >> 2, b = i & 3;
        const uint32_t s = (tw[w] >> (8 * b)) & 0xffu, n = (nup[w] >> (8 * b)) & 0xffu;
        const int idx = (int)(s * 7 + n);
        const uint32_t hi = lane16(i < 8 ? a0 : b0, i & 7), lo = lane16(i < 8 ? a1 : b1, i & 7);
        const uint64_t m = ((uint64_t)hi << 16) | lo;
        const uint64_t T = ((uint64_t)thi[idx] << 16) | tlo[idx];
        const uint32_t lt = m < T ? 1u : 0u;
        const uint32_t sn = HEATBATH ? lt : (s ^ lt);
        tw[w] = (tw[w] & ~(0xffu << (8 * b))) | (sn << (8 * b));
    }
    return make_uint4(tw[0], tw[1], tw[2], tw[3]);
}

// One thread-row: 16 target sites tq; U, C, D: other-plane rows y-1, y, y+1 of the same z-plane; FB: byte-wise
// sum of the other-plane rows (z-1, y) and (z+1, y).  parity 0: in-row pair (j-1, j), parity 1: (j, j+1);
// `side` is the other-plane byte just outside the segment on that side.
template <bool HEATBATH, bool TRACK>
__device__ __forceinline__ uint4 update_row3(const uint4 tq, const uint4 U, const uint4 C, const uint4 D, const uint4 FB,
                                             const uint32_t side, const int parity, const uint32_t blk, const uint32_t t_lo,
                                             const uint32_t c2, const uint32_t c2lo, const uint32_t chain_id,
                                             const uint32_t seed_lo, const uint32_t seed_hi, const uint32_t *s_pair,
                                             const uint32_t *s_thi, const uint32_t *s_tlo, Acc3 &acc, const bool active)
{
    const Philox4 ra = philox4x32_10(blk, t_lo, c2, chain_id, seed_lo, seed_hi);
    const Philox4 rb = philox4x32_10(blk + 1, t_lo, c2, chain_id, seed_lo, seed_hi);

    // S[w] = bytes of the in-row neighbour that is not C's own byte: a funnel shift over
    // {side, C.x, C.y, C.z, C.w} (parity 0, left neighbour) or {C.x, .., C.w, side} (parity 1, right neighbour)
    const uint32_t W0 = parity ? C.x : side << 24, W1 = parity ? C.y : C.x, W2 = parity ? C.z : C.y, W3 = parity ? C.w : C.z,
                   W4 = parity ? side : C.w;
    const uint32_t sh = parity ? 8u : 24u;
    const uint32_t S[4] = {__funnelshift_r(W0, W1, sh), __funnelshift_r(W1, W2, sh), __funnelshift_r(W2, W3, sh),
                           __funnelshift_r(W3, W4, sh)};
    const uint32_t nup[4] = {U.x + D.x + C.x + S[0] + FB.x, U.y + D.y + C.y + S[1] + FB.y, U.z + D.z + C.z + S[2] + FB.z,
                             U.w + D.w + C.w + S[3] + FB.w};
    const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};   // word k: sites 2k, 2k+1
    uint32_t nw[4];
    uint32_t tie_min = 0x7fff7fffu;
    int rejected = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t idx4 = tw[w] * 28u + nup[w] * 4u;        // byte b = 4 * (7 s + nup) of site 4w+b
        const uint32_t ttA = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair) + (idx4 & 0xffffu));
        const uint32_t ttB = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair) + (idx4 >> 16));
        const uint32_t hA = ((rw[2 * w] >> 1) & 0x7fff7fffu) | 0x80008000u;
        const uint32_t hB = ((rw[2 * w + 1] >> 1) & 0x7fff7fffu) | 0x80008000u;
        uint32_t rA, rB;
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rA) : "r"(ttA), "r"(hA));
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rB) : "r"(ttB), "r"(hB));
        tie_min = __vmins2(__vmins2(tie_min, rA), rB);
        uint32_t P;   // 0xFF per site that is NOT accepted
        asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(P) : "r"(rA), "r"(rB));
        nw[w] = HEATBATH ? (~P & 0x01010101u) : (tw[w] ^ (~P & 0x01010101u));
        if (!HEATBATH && !TRACK) rejected = __dp4a((int)P, 0x01010101, rejected);
    }
    const bool tie = ((tie_min & 0x7fffu) == 0u) || ((tie_min & 0x7fff0000u) == 0u);
    if (tie) {
        const Philox4 la = philox4x32_10(blk, t_lo, c2lo, chain_id, seed_lo, seed_hi);
        const Philox4 lb = philox4x32_10(blk + 1, t_lo, c2lo, chain_id, seed_lo, seed_hi);
        const uint4 ex = row_exact3<HEATBATH>(tq, make_uint4(nup[0], nup[1], nup[2], nup[3]), s_thi, s_tlo, ra, rb, la, lb);
        nw[0] = ex.x; nw[1] = ex.y; nw[2] = ex.z; nw[3] = ex.w;
    }
    if (active) {
        if (!HEATBATH && !TRACK && !tie) {
            acc.flips += (uint32_t)(16 + rejected);
        } else {
            uint32_t fsum = 0, ssum = 0, nsum = 0, snsum = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t Fc = nw[w] ^ tw[w];
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

__device__ __forceinline__ uint4 ld128_3(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <int COLOUR, bool HEATBATH, bool TRACK>
__global__ void __launch_bounds__(k3Threads, 5)
k_ising3d(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g, const int32_t *__restrict__ labels,
          long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi, uint64_t t, uint32_t first_chain, int R,
          int strips_per_plane, int blocks_per_chain, int nitems, int z0, int nz)
{
    __shared__ uint32_t s_pair[k3PairWords];
    __shared__ uint32_t s_thi[k3Table], s_tlo[k3Table];
    int cur_label = -1;

    const int half = L.half;
    const int nseg = half >> 4;
    const int64_t G = (int64_t)strips_per_plane * nz * nseg;      // this launch: z-planes [z0, z0 + nz)
    const int lane = threadIdx.x & 31;
    const uint32_t t_lo = (uint32_t)t;
    const uint32_t c2 = ctr_word2(t, 0, TAG_SWEEP), c2lo = ctr_word2(t, 1, TAG_SWEEP);
    const int64_t plane_rows = (int64_t)L.Ly * half;             // bytes of one z-plane of a colour plane

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chain = item / blocks_per_chain;
        const int label = labels[chain];
        if (label != cur_label) {
            __syncthreads();
            for (int i = threadIdx.x; i < k3Table; i += k3Threads) {
                s_thi[i] = thi_g[label * k3Table + i];
                s_tlo[i] = tlo_g[label * k3Table + i];
            }
            for (int i = threadIdx.x; i < k3Table * k3Table; i += k3Threads) {
                const int i1 = i / k3Table, i0 = i - i1 * k3Table;
                const uint32_t a = min(thi_g[label * k3Table + i0] >> 1, 0x7fffu);
                const uint32_t b = min(thi_g[label * k3Table + i1] >> 1, 0x7fffu);
                s_pair[i1 * k3RowWords + i0] = a | (b << 16);
            }
            __syncthreads();
            cur_label = label;
        }
        const int64_t g0 = (int64_t)(item - chain * blocks_per_chain) * k3Threads + threadIdx.x;
        const bool active = g0 < G;
        const int64_t g = active ? g0 : G - 1;
        const int sidx = (int)(g / nseg);                         // strip index over all planes
        const int seg = (int)(g - (int64_t)sidx * nseg);
        const int zl = sidx / strips_per_plane;
        const int z = z0 + zl;
        const int y0 = (sidx - zl * strips_per_plane) * R;         // even
        const uint32_t chain_id = first_chain + (uint32_t)chain;
        const int pa = (COLOUR + z) & 1;                          // in-row pairing of the strip's even rows

        uint8_t *tgt = plane_ptr(L, chain, COLOUR) + (int64_t)z * plane_rows;
        const uint8_t *__restrict__ oth = plane_ptr(L, chain, COLOUR ^ 1) + (int64_t)z * plane_rows;
        const uint8_t *__restrict__ othF = plane_ptr(L, chain, COLOUR ^ 1) + (int64_t)(z == 0 ? L.Lz - 1 : z - 1) * plane_rows;
        const uint8_t *__restrict__ othB = plane_ptr(L, chain, COLOUR ^ 1) + (int64_t)(z == L.Lz - 1 ? 0 : z + 1) * plane_rows;
        const int col = seg << 4;
        const int colL = (seg == 0 ? half : col) - 1;
        const int colR = (seg == nseg - 1) ? 0 : col + 16;
        const bool loadL = (lane == 0) || (seg == 0);
        const bool loadR = (lane == 31) || (seg == nseg - 1);
        // row a (even y) pairs with the left byte when pa == 0, row b (odd y) with the other side
        const bool edgeA = pa == 0 ? loadL : loadR, edgeB = pa == 0 ? loadR : loadL;
        const int colA = pa == 0 ? colL : colR, colB = pa == 0 ? colR : colL;

        const int yU = y0 == 0 ? L.Ly - 1 : y0 - 1;
        const uint8_t *po = oth + (int64_t)y0 * half;
        uint8_t *pt = tgt + (int64_t)y0 * half + col;
        int64_t zoff = (int64_t)y0 * half + col;                  // offset of row y inside a z-plane, this segment
        uint4 U = ld128_3(oth + (int64_t)yU * half + col);
        uint4 C = ld128_3(po + col);
        uint32_t blk = (uint32_t)((((int64_t)z * L.Ly + y0) * half + col) >> 3);
        const uint32_t blk_step = (uint32_t)(half >> 3);
        Acc3 acc;

#pragma unroll 1
        for (int r = 0; r < R; r += 2) {
            const int y = y0 + r;
            const uint8_t *pe = (y + 2 == L.Ly) ? oth : po + 2 * (int64_t)half;
            const uint4 E = ld128_3(pe + col);
            const uint4 D = ld128_3(po + half + col);
            const uint4 Ta = ld128_3(pt), Tb = ld128_3(pt + half);
            const uint4 FBa = add4(ld128_3(othF + zoff), ld128_3(othB + zoff));
            const uint4 FBb = add4(ld128_3(othF + zoff + half), ld128_3(othB + zoff + half));
            uint32_t sideA = 0, sideB = 0;
            if (edgeA) sideA = po[colA];
            if (edgeB) sideB = po[half + colB];
            // side bytes from the neighbouring lanes: left neighbour = previous lane's last byte, right = next lane's first
            const uint32_t cl = __shfl_up_sync(0xffffffffu, C.w, 1) >> 24, cr = __shfl_down_sync(0xffffffffu, C.x, 1) & 0xffu;
            const uint32_t dl = __shfl_up_sync(0xffffffffu, D.w, 1) >> 24, dr = __shfl_down_sync(0xffffffffu, D.x, 1) & 0xffu;
            uint32_t sA = pa == 0 ? cl : cr, sB = pa == 0 ? dr : dl;
            if (edgeA) sA = sideA;
            if (edgeB) sB = sideB;
            const uint4 Na = update_row3<HEATBATH, TRACK>(Ta, U, C, D, FBa, sA, pa, blk, t_lo, c2, c2lo, chain_id, seed_lo, seed_hi,
                                                          s_pair, s_thi, s_tlo, acc, active);
            if (active) *reinterpret_cast<uint4 *>(pt) = Na;
            asm volatile("" ::: "memory");
            const uint4 Nb = update_row3<HEATBATH, TRACK>(Tb, C, D, E, FBb, sB, pa ^ 1, blk + blk_step, t_lo, c2, c2lo, chain_id,
                                                          seed_lo, seed_hi, s_pair, s_thi, s_tlo, acc, active);
            if (active) *reinterpret_cast<uint4 *>(pt + half) = Nb;
            U = D; C = E;
            po += 2 * (int64_t)half; pt += 2 * (int64_t)half; zoff += 2 * (int64_t)half; blk += 2 * blk_step;
        }

        // per-chain sums over changed sites (nn = 6): dspin = 2 - 4 s, dpair = -8 s nup + 24 s + 4 nup - 12
        const int nflip = warp_sum((int)acc.flips);
        int dspin = 0, dpair = 0;
        if (TRACK) {
            const int ss = warp_sum(acc.s), nn_ = warp_sum(acc.n), sn = warp_sum(acc.sn);
            dspin = 2 * nflip - 4 * ss;
            dpair = -8 * sn + 24 * ss + 4 * nn_ - 12 * nflip;
        }
        if (lane == 0) {
            unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
            if (nflip) atomicAdd(o + SUM_ACC, (unsigned long long)(long long)nflip);
            if (TRACK) {
                if (dpair) atomicAdd(o + SUM_PAIR, (unsigned long long)(long long)dpair);
                if (dspin) atomicAdd(o + SUM_SPIN, (unsigned long long)(long long)dspin);
            }
        }
    }
}


// The same half-sweep on one-bit-per-spin planes (MCX_STORAGE_BIT, layout in k_bits.cuh): a trip loads one word of
// the target plane (both rows), one new word of the other plane and one word each of the z - 1 / z + 1 planes, expands
// them to 0/1 bytes and runs update_row3 unchanged.
template <int COLOUR, bool HEATBATH, bool TRACK>
__global__ void __launch_bounds__(k3Threads, 5)
k_ising3d_bits(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g, const int32_t *__restrict__ labels,
               long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi, uint64_t t, uint32_t first_chain, int R,
               int strips_per_plane, int blocks_per_chain, int nitems, int z0, int nz)
{
    __shared__ uint32_t s_pair[k3PairWords];
    __shared__ uint32_t s_thi[k3Table], s_tlo[k3Table];
    int cur_label = -1;

    const int half = L.half;
    const int nseg = half >> 4;
    const int64_t G = (int64_t)strips_per_plane * nz * nseg;      // this launch: z-planes [z0, z0 + nz)
    const int lane = threadIdx.x & 31;
    const uint32_t t_lo = (uint32_t)t;
    const uint32_t c2 = ctr_word2(t, 0, TAG_SWEEP), c2lo = ctr_word2(t, 1, TAG_SWEEP);
    const int64_t plane_words = (int64_t)(L.Ly >> 1) * nseg;      // words of one z-plane of a colour plane

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chain = item / blocks_per_chain;
        const int label = labels[chain];
        if (label != cur_label) {
            __syncthreads();
            for (int i = threadIdx.x; i < k3Table; i += k3Threads) {
                s_thi[i] = thi_g[label * k3Table + i];
                s_tlo[i] = tlo_g[label * k3Table + i];
            }
            for (int i = threadIdx.x; i < k3Table * k3Table; i += k3Threads) {
                const int i1 = i / k3Table, i0 = i - i1 * k3Table;
                const uint32_t a = min(thi_g[label * k3Table + i0] >> 1, 0x7fffu);
                const uint32_t b = min(thi_g[label * k3Table + i1] >> 1, 0x7fffu);
                s_pair[i1 * k3RowWords + i0] = a | (b << 16);
            }
            __syncthreads();
            cur_label = label;
        }
        const int64_t g0 = (int64_t)(item - chain * blocks_per_chain) * k3Threads + threadIdx.x;
        const bool active = g0 < G;
        const int64_t g = active ? g0 : G - 1;
        const int sidx = (int)(g / nseg);
        const int seg = (int)(g - (int64_t)sidx * nseg);
        const int zl = sidx / strips_per_plane;
        const int z = z0 + zl;
        const int y0 = (sidx - zl * strips_per_plane) * R;         // even
        const uint32_t chain_id = first_chain + (uint32_t)chain;
        const int pa = (COLOUR + z) & 1;

        uint32_t *tgt = reinterpret_cast<uint32_t *>(plane_ptr(L, chain, COLOUR)) + (int64_t)z * plane_words;
        const uint32_t *__restrict__ othp = reinterpret_cast<const uint32_t *>(plane_ptr(L, chain, COLOUR ^ 1));
        const uint32_t *__restrict__ oth = othp + (int64_t)z * plane_words;
        const uint32_t *__restrict__ othF = othp + (int64_t)(z == 0 ? L.Lz - 1 : z - 1) * plane_words;
        const uint32_t *__restrict__ othB = othp + (int64_t)(z == L.Lz - 1 ? 0 : z + 1) * plane_words;
        const int col = seg << 4;
        const int colL = (seg == 0 ? half : col) - 1;
        const int colR = (seg == nseg - 1) ? 0 : col + 16;
        const bool loadL = (lane == 0) || (seg == 0);
        const bool loadR = (lane == 31) || (seg == nseg - 1);
        const bool edgeA = pa == 0 ? loadL : loadR, edgeB = pa == 0 ? loadR : loadL;
        const int colA = pa == 0 ? colL : colR, colB = pa == 0 ? colR : colL;
        const int segA = colA >> 4, segB = colB >> 4;
        const int bitA = bit_pos(0, colA & 15), bitB = bit_pos(1, colB & 15);

        const int yU = y0 == 0 ? L.Ly - 1 : y0 - 1;               // odd: second row of its word
        int64_t woff = (int64_t)(y0 >> 1) * nseg;                 // word offset of this trip's row pair inside a z-plane
        uint4 U = bits_expand(oth[(int64_t)(yU >> 1) * nseg + seg], 1);
        uint32_t Wc = oth[woff + seg];
        uint4 C = bits_expand(Wc, 0);
        uint32_t blk = (uint32_t)((((int64_t)z * L.Ly + y0) * half + col) >> 3);
        const uint32_t blk_step = (uint32_t)(half >> 3);
        Acc3 acc;

#pragma unroll 1
        for (int r = 0; r < R; r += 2) {
            const int y = y0 + r;
            const uint32_t We = oth[((y + 2 == L.Ly) ? 0 : woff + nseg) + seg];
            const uint32_t Tw = tgt[woff + seg];
            const uint32_t Fw = othF[woff + seg], Bw = othB[woff + seg];
            uint32_t sideA = 0, sideB = 0;
            if (edgeA) sideA = (oth[woff + segA] >> bitA) & 1u;
            if (edgeB) sideB = (oth[woff + segB] >> bitB) & 1u;
            const uint4 D = bits_expand(Wc, 1);
            const uint32_t cl = __shfl_up_sync(0xffffffffu, C.w, 1) >> 24, cr = __shfl_down_sync(0xffffffffu, C.x, 1) & 0xffu;
            const uint32_t dl = __shfl_up_sync(0xffffffffu, D.w, 1) >> 24, dr = __shfl_down_sync(0xffffffffu, D.x, 1) & 0xffu;
            uint32_t sA = pa == 0 ? cl : cr, sB = pa == 0 ? dr : dl;
            if (edgeA) sA = sideA;
            if (edgeB) sB = sideB;
            const uint4 Na = update_row3<HEATBATH, TRACK>(bits_expand(Tw, 0), U, C, D, add4(bits_expand(Fw, 0), bits_expand(Bw, 0)), sA, pa,
                                                          blk, t_lo, c2, c2lo, chain_id, seed_lo, seed_hi, s_pair, s_thi, s_tlo, acc, active);
            uint32_t out = bits_compress(Na);
            const uint4 E = bits_expand(We, 0);
            const uint4 Nb = update_row3<HEATBATH, TRACK>(bits_expand(Tw, 1), C, D, E, add4(bits_expand(Fw, 1), bits_expand(Bw, 1)), sB,
                                                          pa ^ 1, blk + blk_step, t_lo, c2, c2lo, chain_id, seed_lo, seed_hi, s_pair, s_thi,
                                                          s_tlo, acc, active);
            out += bits_compress(Nb) << 4;
            if (active) tgt[woff + seg] = out;
            U = D; C = E; Wc = We;
            woff += nseg; blk += 2 * blk_step;
        }

        const int nflip = warp_sum((int)acc.flips);
        int dspin = 0, dpair = 0;
        if (TRACK) {
            const int ss = warp_sum(acc.s), nn_ = warp_sum(acc.n), sn = warp_sum(acc.sn);
            dspin = 2 * nflip - 4 * ss;
            dpair = -8 * sn + 24 * ss + 4 * nn_ - 12 * nflip;
        }
        if (lane == 0) {
            unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
            if (nflip) atomicAdd(o + SUM_ACC, (unsigned long long)(long long)nflip);
            if (TRACK) {
                if (dpair) atomicAdd(o + SUM_PAIR, (unsigned long long)(long long)dpair);
                if (dspin) atomicAdd(o + SUM_SPIN, (unsigned long long)(long long)dspin);
            }
        }
    }
}


template <int COLOUR, bool HEATBATH, bool TRACK>
void launch_3d(mcx_lattice *lat, uint64_t t)
{
    // a chain sub-range (launch_sweeps_ising2d_grouped) is the same launch on shifted base pointers
    LatView L = lat->view;
    const int c0 = g_launch_range.chain0, nch = g_launch_range.nchains < 0 ? lat->nchains : g_launch_range.nchains;
    L.planes += (int64_t)c0 * 2 * L.plane_stride;
    L.nchains = nch;
    cudaStream_t stream = g_launch_range.use_stream ? g_launch_range.stream : lat->ctx->stream;
    const int nseg = L.half >> 4;
    const int64_t ctas = (int64_t)lat->ctx->sm_count * 5;
    // a band of z-planes (launch_sweeps_ising2d_banded): the same launch over planes [z0, z0 + nz)
    const bool band = g_launch_range.nrows > 0;
    const int z0 = band ? g_launch_range.row0 : 0, nz = band ? g_launch_range.nrows : L.Lz;
    // strips never cross a z-plane: the tallest even divisor of Ly up to 16 that still gives ~4 items per resident CTA
    // (counted over the whole lattice, so that all bands of a half-sweep use the same strips)
    int R = 2;
    for (int r = 16; r >= 2; r -= 2) {
        if (L.Ly % r != 0) continue;
        const int64_t items = ((int64_t)(L.Ly / r) * L.Lz * nseg + k3Threads - 1) / k3Threads * lat->nchains;
        R = r;
        if (items >= 4 * ctas || r <= 4) break;
    }
    const int strips_per_plane = L.Ly / R;
    const int64_t G = (int64_t)strips_per_plane * nz * nseg;
    const int blocks_per_chain = (int)((G + k3Threads - 1) / k3Threads);
    const int nitems = (int)((int64_t)blocks_per_chain * nch);
    auto kern = lat->storage == MCX_STORAGE_BIT ? k_ising3d_bits<COLOUR, HEATBATH, TRACK> : k_ising3d<COLOUR, HEATBATH, TRACK>;
    static thread_local int resident_of[2] = {0, 0};
    int &resident = resident_of[lat->storage == MCX_STORAGE_BIT ? 1 : 0];
    if (!resident) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, k3Threads, 0);
        if (resident < 1) resident = 1;
    }
    int grid = lat->ctx->sm_count * resident;
    if (grid > nitems) grid = nitems;
    kern<<<grid, k3Threads, 0, stream>>>(L, lat->d_thi, lat->d_tlo, lat->d_labels + c0, lat->d_sums + (int64_t)c0 * SUM_FIELDS,
                                        (uint32_t)lat->seed, (uint32_t)(lat->seed >> 32), t, lat->first_chain + (uint32_t)c0, R,
                                        strips_per_plane, blocks_per_chain, nitems, z0, nz);
    lat->ctx->launches++;
}

}  // namespace

// false: not applicable (shape or table layout), nothing launched
bool launch_sweep_ising3d(mcx_lattice *lat, int colour, uint64_t t)
{
    if (lat->ndim != 3 || lat->model != MCX_ISING || lat->view.Lx % 32 != 0) return false;
    const bool bits = lat->storage == MCX_STORAGE_BIT;          // bit planes have no other kernel: always here
    if (lat->view.Ly % 2 != 0 || lat->table_len != k3Table || (knobs().ising3d == 0 && !bits)) return false;
    if (!bits && (int64_t)(lat->view.Ly / 2) * lat->view.Lz * (lat->view.half >> 4) < 96) return false;     // tiny: rows-of-8 kernel
    const bool track = lat->track_sums, hb = lat->rule == MCX_HEATBATH;
    if (colour == 0) {
        if (hb) { if (track) launch_3d<0, true, true>(lat, t); else launch_3d<0, true, false>(lat, t); }
        else    { if (track) launch_3d<0, false, true>(lat, t); else launch_3d<0, false, false>(lat, t); }
    } else {
        if (hb) { if (track) launch_3d<1, true, true>(lat, t); else launch_3d<1, true, false>(lat, t); }
        else    { if (track) launch_3d<1, false, true>(lat, t); else launch_3d<1, false, false>(lat, t); }
    }
    return true;
}

}  // namespace mcx
