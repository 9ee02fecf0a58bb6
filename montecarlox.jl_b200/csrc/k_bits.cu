// k_bits.cu -- MCX_STORAGE_BIT: layout conversion, init, full recompute (2-D and 3-D) and the 2-D Ising half-sweep on
// one-bit-per-spin colour planes (layout in k_bits.cuh).  The 3-D half-sweep is in k_ising3d.cu next to its int8 twin.
//
// The half-sweep is the k_ising2d decomposition (one thread = a 16-site column segment walking a strip of rows, two
// rows per trip, other-plane rows in a rolling register window) with the loads and stores replaced: a trip loads one
// 32-bit word of the target plane (both rows) and one new word of the other plane, expands them to the 0/1-byte
// working form, runs the int8 kernel's own update_row (same Philox positions, same packed 15-bit decision, same exact
// redo) and writes one word back.  0.375 B per attempt instead of 3; an L = 16384 lattice is 2 x 16 MiB and stays in L2.
#include "k_bits.cuh"
#include "k_strip.cuh"

namespace mcx {

namespace {

constexpr int kBitThreads = 128;

// The rows of a strip that the loop left undecided (flags as in k_strip.cuh), redone one by one: the row's 16 bits are
// cut out of its word, settled with the full 32-bit draws and merged back (only this thread writes the word).
template <int COLOUR, bool HEATBATH, bool TRACK>
__device__ __noinline__ uint4 settle_rows_bits(uint8_t *tgt, const uint8_t *oth, const uint8_t *oth_up, const uint8_t *oth_dn,
                                               const int half, const int Ly, const int row_offset, uint32_t ties, const int row0,
                                               const int R, const int seg, const int colL, const int colR, const uint32_t t_lo,
                                               const uint32_t c2, const uint32_t c2lo, const uint32_t chain_id, const uint32_t seed_lo,
                                               const uint32_t seed_hi, const uint32_t *s_thi, const uint32_t *s_tlo)
{
    const int nseg = half >> 4, col = seg << 4;
    Acc acc;
    while (ties) {
        const int p = __ffs((int)ties) - 1;
        ties &= ties - 1;
        const int r = R - 2 - (p & ~1) + (~p & 1);
        const int row = row0 + r;
        const int rho = r & 1;                                   // row0 is even: the row's place in its word
        const int parity = rho ? (COLOUR ^ 1) : COLOUR;
        const uint4 U = row == 0 ? bits_expand(*bits_word(oth_up, Ly - 1, nseg, seg), 1) : bits_expand(*bits_word(oth, row - 1, nseg, seg), rho ^ 1);
        const uint4 D = row + 1 == Ly ? bits_expand(*bits_word(oth_dn, 0, nseg, seg), 0) : bits_expand(*bits_word(oth, row + 1, nseg, seg), rho ^ 1);
        const uint4 C = bits_expand(*bits_word(oth, row, nseg, seg), rho);
        const uint32_t side = bits_site(oth, row, nseg, parity == 0 ? colL : colR);
        uint32_t *ptw = reinterpret_cast<uint32_t *>(tgt) + (int64_t)(row >> 1) * nseg + seg;
        const uint32_t Tw = *ptw;
        const uint32_t blk = (uint32_t)(((int64_t)(row + row_offset) * half + col) >> 3);
        const uint4 ex = row_settle<HEATBATH, TRACK>(parity, bits_expand(Tw, rho), U, C, D, side, blk, t_lo, c2, c2lo, chain_id, seed_lo,
                                                     seed_hi, s_thi, s_tlo, acc);
        *ptw = (Tw & (rho ? 0x0f0f0f0fu : 0xf0f0f0f0u)) | (bits_compress(ex) << (4 * rho));
    }
    return make_uint4(acc.flips, (uint32_t)acc.s, (uint32_t)acc.n, (uint32_t)acc.sn);
}

template <int COLOUR, bool HEATBATH, bool TRACK, bool FULL, bool SLAB>
__global__ void __launch_bounds__(kBitThreads, 8)
k_ising2d_bits(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
               const int32_t *__restrict__ labels, long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi,
               uint64_t t, uint32_t first_chain, int R, int nstrips, int blocks_per_chain, int nitems)
{
    __shared__ uint32_t s_pair[kPairWords];
    __shared__ uint32_t s_thi[kTableLen], s_tlo[kTableLen];
    int cur_label = -1;

    const int half = L.half;
    const int nseg = half >> 4;
    const uint32_t G = (uint32_t)nstrips * (uint32_t)nseg;       // thread-items of a chain: < 2^31
    const int lane = threadIdx.x & 31;
    const uint32_t t_lo = (uint32_t)t;
    const uint32_t c2 = ctr_word2(t, 0, TAG_SWEEP);
    const uint32_t c2lo = ctr_word2(t, 1, TAG_SWEEP);
    const uint32_t pair_addr = (uint32_t)__cvta_generic_to_shared(s_pair);

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chain = item / blocks_per_chain;
        const int label = labels[chain];
        if (label != cur_label) {
            __syncthreads();
            load_pair_table(s_pair, s_thi, s_tlo, thi_g, tlo_g, label);
            __syncthreads();
            cur_label = label;
        }
        const uint32_t g0 = (uint32_t)(item - chain * blocks_per_chain) * kBitThreads + threadIdx.x;
        const bool active = FULL ? true : g0 < G;
        const uint32_t g = active ? g0 : G - 1;
        const int strip = (int)(g / (uint32_t)nseg);
        const int seg = (int)(g - (uint32_t)strip * (uint32_t)nseg);
        const int row0 = strip * R;                               // even
        const uint32_t chain_id = first_chain + (uint32_t)chain;

        uint8_t *tgt = plane_ptr(L, chain, COLOUR);
        const uint8_t *__restrict__ oth = plane_ptr(L, chain, COLOUR ^ 1);
        const int col = seg << 4;
        const int colL = (seg == 0 ? half : col) - 1;             // site left of the segment (periodic)
        const int colR = (seg == nseg - 1) ? 0 : col + 16;        // site right of the segment
        const bool loadL = (lane == 0) || (seg == 0);
        const bool loadR = (lane == 31) || (seg == nseg - 1);
        // even rows of this strip have parity COLOUR, odd rows COLOUR ^ 1 (row0 is even)
        const bool edgeA = COLOUR == 0 ? loadL : loadR;
        const bool edgeB = COLOUR == 0 ? loadR : loadL;
        const int colA = COLOUR == 0 ? colL : colR;
        const int colB = COLOUR == 0 ? colR : colL;
        const int segA = colA >> 4, segB = colB >> 4;
        const int bitA = bit_pos(0, colA & 15), bitB = bit_pos(1, colB & 15);

        // the row above row 0 / below row Ly - 1: the own plane (periodic) or, for a row band, the rows around the band
        const uint8_t *oth_dn = SLAB ? plane_ptr_of(L.dn_planes, L, chain, COLOUR ^ 1) : oth;
        const uint8_t *oth_up = SLAB ? plane_ptr_of(L.up_planes, L, chain, COLOUR ^ 1) : oth;
        const int rowU = row0 == 0 ? L.Ly - 1 : row0 - 1;         // odd: the second row of its word
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(oth) + (int64_t)(row0 >> 1) * nseg;   // other plane, this trip's row pair
        uint32_t *ptw = reinterpret_cast<uint32_t *>(tgt) + (int64_t)(row0 >> 1) * nseg + seg;        // target word of this trip
        uint4 U = bits_expand(*bits_word(SLAB && row0 == 0 ? oth_up : oth, rowU, nseg, seg), 1);
        uint32_t Wc = pw[seg];                                    // rows row0 (C) and row0 + 1 (D)
        uint4 C = bits_expand(Wc, 0);
        uint32_t blk = (uint32_t)(((int64_t)(row0 + (SLAB ? L.row_offset : 0)) * half + col) >> 3);
        const uint32_t blk_step = (uint32_t)(half >> 3);
        const PhiloxHead H = philox_head(t_lo, c2, chain_id, seed_lo, seed_hi);
        Acc acc;
        uint32_t ties = 0;                                        // rows left to settle_rows_bits() (R <= 32)

        // one basic block per trip, as in k_strip.cuh: a row with an undecided site keeps its bits and is settled after the strip
#pragma unroll 1
        for (int r = 0; r < R; r += 2) {
            const int row = row0 + r;
            // the next row pair of the other plane; wraps only at the very last row of the lattice / band
            const uint32_t *pe = (row + 2 == L.Ly) ? reinterpret_cast<const uint32_t *>(oth_dn) : pw + nseg;
            const uint32_t We = pe[seg];
            const uint32_t Tw = *ptw;
            uint32_t sideA = 0, sideB = 0;
            if (edgeA) sideA = (pw[segA] >> bitA) & 1u;
            if (edgeB) sideB = (pw[segB] >> bitB) & 1u;
            const uint4 D = bits_expand(Wc, 1);
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
            const uint4 Na = update_row_fast<COLOUR, HEATBATH, TRACK>(philox_tail(H, blk), philox_tail(H, blk + 1), bits_expand(Tw, 0), U, C, D, sA,
                                                                      pair_addr, acc, active, tieA);
            uint32_t out = tieA ? (Tw & 0x0f0f0f0fu) : bits_compress(Na);
            const uint4 E = bits_expand(We, 0);
            const uint4 Nb = update_row_fast<COLOUR ^ 1, HEATBATH, TRACK>(philox_tail(H, blk + blk_step), philox_tail(H, blk + blk_step + 1),
                                                                          bits_expand(Tw, 1), C, D, E, sB, pair_addr, acc, active, tieB);
            out |= tieB ? (Tw & 0xf0f0f0f0u) : (bits_compress(Nb) << 4);
            if (active) *ptw = out;
            ties = (ties << 2) | (tieA ? 2u : 0u) | (tieB ? 1u : 0u);
            U = D; C = E; Wc = We;
            pw += nseg; ptw += nseg; blk += 2 * blk_step;
        }
        if (active && ties) {
            const uint4 d = settle_rows_bits<COLOUR, HEATBATH, TRACK>(tgt, oth, oth_up, oth_dn, half, L.Ly, SLAB ? L.row_offset : 0, ties, row0, R,
                                                                      seg, colL, colR, t_lo, c2, c2lo, chain_id, seed_lo, seed_hi, s_thi, s_tlo);
            acc.flips += d.x; acc.s += (int32_t)d.y; acc.n += (int32_t)d.z; acc.sn += (int32_t)d.w;
        }
        strip_finish<TRACK>(acc, sums, chain);
    }
}

// ---------------------------------------------------------------------------------------------
// layout conversion, init, recompute: one thread per plane word = (row pair m, segment), any of 2-D / 3-D
// (rows = Ly * Lz; Ly is even, so a row pair never straddles two z-planes)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int even_x_colour(const LatView &L, int64_t row)
{
    const int y = (int)(row % L.Ly), z = (int)(row / L.Ly);
    return (y + z) & 1;                                           // colour of the sites with even x in this row
}

__global__ void __launch_bounds__(256)
k_pack_bits(LatView L, const int8_t *__restrict__ staging, int64_t N, int64_t words_per_chain)
{
    const int chain = blockIdx.y;
    const int nseg = L.half >> 4;
    uint32_t *p0 = reinterpret_cast<uint32_t *>(plane_ptr(L, chain, 0)), *p1 = reinterpret_cast<uint32_t *>(plane_ptr(L, chain, 1));
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < words_per_chain; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = g / nseg;
        const int seg = (int)(g - m * nseg);
        uint32_t out[2] = {0u, 0u};
#pragma unroll
        for (int rho = 0; rho < 2; ++rho) {
            const int64_t row = 2 * m + rho;
            const uint4 *src = reinterpret_cast<const uint4 *>(staging + (int64_t)chain * N + row * L.Lx + seg * 32);
            const uint4 a = src[0], b = src[1];
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint4 ev, od;
            uint32_t *e = &ev.x, *o = &od.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t e0 = (~w[2 * k] >> 7) & 0x01010101u, e1 = (~w[2 * k + 1] >> 7) & 0x01010101u;   // +1 -> 1, -1 -> 0
                e[k] = __byte_perm(e0, e1, 0x6420);
                o[k] = __byte_perm(e0, e1, 0x7531);
            }
            const int ce = even_x_colour(L, row);
            out[ce] |= bits_compress(ev) << (4 * rho);
            out[ce ^ 1] |= bits_compress(od) << (4 * rho);
        }
        p0[g] = out[0];
        p1[g] = out[1];
    }
}

__global__ void __launch_bounds__(256)
k_unpack_bits(LatView L, int8_t *__restrict__ staging, int64_t N, int64_t words_per_chain)
{
    const int chain = blockIdx.y;
    const int nseg = L.half >> 4;
    const uint32_t *p0 = reinterpret_cast<const uint32_t *>(plane_ptr(L, chain, 0)), *p1 = reinterpret_cast<const uint32_t *>(plane_ptr(L, chain, 1));
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < words_per_chain; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = g / nseg;
        const int seg = (int)(g - m * nseg);
        const uint32_t in[2] = {p0[g], p1[g]};
#pragma unroll
        for (int rho = 0; rho < 2; ++rho) {
            const int64_t row = 2 * m + rho;
            const int ce = even_x_colour(L, row);
            const uint4 ev = bits_expand(in[ce], rho), od = bits_expand(in[ce ^ 1], rho);
            const uint32_t e[4] = {ev.x, ev.y, ev.z, ev.w}, o[4] = {od.x, od.y, od.z, od.w};
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t lo = __byte_perm(e[k], o[k], 0x5140), hi = __byte_perm(e[k], o[k], 0x7362);
                w[2 * k] = (lo ^ 0x01010101u) * 0xFEu + 0x01010101u;        // 1 -> 0x01 (+1), 0 -> 0xFF (-1)
                w[2 * k + 1] = (hi ^ 0x01010101u) * 0xFEu + 0x01010101u;
            }
            uint4 *dst = reinterpret_cast<uint4 *>(staging + (int64_t)chain * N + row * L.Lx + seg * 32);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

// init!(sys, mode; rng): the 32 x-sites a thread holds per row are exactly one word of the INIT stream
__global__ void __launch_bounds__(256)
k_init_bits(LatView L, int mode, uint32_t seed_lo, uint32_t seed_hi, uint32_t first_chain, int64_t words_per_chain)
{
    const int chain = blockIdx.y;
    const int nseg = L.half >> 4;
    uint32_t *p0 = reinterpret_cast<uint32_t *>(plane_ptr(L, chain, 0)), *p1 = reinterpret_cast<uint32_t *>(plane_ptr(L, chain, 1));
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < words_per_chain; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = g / nseg;
        const int seg = (int)(g - m * nseg);
        uint32_t out[2] = {0u, 0u};
#pragma unroll
        for (int rho = 0; rho < 2; ++rho) {
            const int64_t row = 2 * m + rho;
            uint32_t bits = mode == MCX_INIT_UP ? 0xffffffffu : 0u;
            if (mode == MCX_INIT_RANDOM) {
                const int64_t i = (row + L.row_offset) * L.Lx + seg * 32;
                const Philox4 p = stream_block(seed_lo, seed_hi, first_chain + chain, TAG_INIT, 0, (uint32_t)(i >> 7), 0);
                const int w = (int)((i >> 5) & 3);
                bits = w == 0 ? p.x : w == 1 ? p.y : w == 2 ? p.z : p.w;
            }
            uint32_t ev = 0, od = 0;          // site j = k of the even-x / odd-x plane at its transposed position
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                ev |= ((bits >> (2 * k)) & 1u) << bit_pos(rho, k);
                od |= ((bits >> (2 * k + 1)) & 1u) << bit_pos(rho, k);
            }
            const int ce = even_x_colour(L, row);
            out[ce] |= ev;
            out[ce ^ 1] |= od;
        }
        p0[g] = out[0];
        p1[g] = out[1];
    }
}

// host bit buffers (site i of a chain = bit i & 7 of byte i >> 3, 1 = up) <-> the int8 staging buffer
__global__ void __launch_bounds__(256)
k_hostbits_to_staging(const uint32_t *__restrict__ bits, int8_t *__restrict__ staging, int64_t nwords)
{
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nwords; g += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = bits[g];
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t n = (v >> (4 * k)) & 0xfu;
            const uint32_t e = (n * 0x00204081u) & 0x01010101u;             // bit b of the nibble -> byte b
            w[k] = (e ^ 0x01010101u) * 0xFEu + 0x01010101u;                 // 1 -> +1, 0 -> -1
        }
        uint4 *dst = reinterpret_cast<uint4 *>(staging + g * 32);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

__global__ void __launch_bounds__(256)
k_staging_to_hostbits(const int8_t *__restrict__ staging, uint32_t *__restrict__ bits, int64_t nwords)
{
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nwords; g += (int64_t)gridDim.x * blockDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(staging + g * 32);
        const uint4 a = src[0], b = src[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t e = (~w[k] >> 7) & 0x01010101u;                  // +1 -> 1, -1 -> 0 per byte
            v |= (((e * 0x10204080u) >> 28) & 0xfu) << (4 * k);             // byte b -> bit b of a nibble
        }
        bits[g] = v;
    }
}

// _recompute_cached! on bit planes, nn = 4 or 6: with e in {0, 1},
//   sum_<ij> s_i s_j = sum over colour-0 sites of (2 e0 - 1)(2 nup - nn) = 4 sum(e0 nup) - 2 sum(nup) - 2 nn sum(e0) + nn N/2
//   sum s = 2 (sum e0 + sum e1) - N
__global__ void __launch_bounds__(256)
k_recompute_bits(LatView L, long long *__restrict__ sums, int64_t words_per_chain)
{
    const int chain = blockIdx.y;
    const int half = L.half, nseg = half >> 4;
    const uint8_t *__restrict__ q0 = plane_ptr(L, chain, 0);
    const uint8_t *__restrict__ q1 = plane_ptr(L, chain, 1);
    long long e0n = 0, nsum = 0, e0 = 0, e1 = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < words_per_chain; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = g / nseg;
        const int seg = (int)(g - m * nseg), col = seg << 4;
        const uint32_t T0 = reinterpret_cast<const uint32_t *>(q0)[g], C1 = reinterpret_cast<const uint32_t *>(q1)[g];
#pragma unroll
        for (int rho = 0; rho < 2; ++rho) {
            const int64_t row = 2 * m + rho;
            const int y = (int)(row % L.Ly), z = (int)(row / L.Ly);
            const int64_t zr = (int64_t)z * L.Ly;
            const int64_t ru = zr + (y == 0 ? L.Ly - 1 : y - 1), rd = zr + (y == L.Ly - 1 ? 0 : y + 1);
            const uint4 T = bits_expand(T0, rho), C = bits_expand(C1, rho);
            const uint4 U = bits_expand(*bits_word(q1, ru, nseg, seg), (int)(ru & 1));
            const uint4 D = bits_expand(*bits_word(q1, rd, nseg, seg), (int)(rd & 1));
            uint32_t S[4];
            if (((y + z) & 1) == 0) {       // colour-0 sites sit at x = 2j: in-row neighbours j - 1, j
                const uint32_t side = bits_site(q1, row, nseg, (seg == 0 ? half : col) - 1);
                S[0] = (C.x << 8) | side;
                S[1] = __funnelshift_l(C.x, C.y, 8); S[2] = __funnelshift_l(C.y, C.z, 8); S[3] = __funnelshift_l(C.z, C.w, 8);
            } else {
                const uint32_t side = bits_site(q1, row, nseg, seg == nseg - 1 ? 0 : col + 16);
                S[0] = __funnelshift_r(C.x, C.y, 8); S[1] = __funnelshift_r(C.y, C.z, 8); S[2] = __funnelshift_r(C.z, C.w, 8);
                S[3] = (C.w >> 8) | (side << 24);
            }
            uint32_t nup[4] = {U.x + D.x + C.x + S[0], U.y + D.y + C.y + S[1], U.z + D.z + C.z + S[2], U.w + D.w + C.w + S[3]};
            if (L.ndim == 3) {
                const int64_t rf = (int64_t)(z == 0 ? L.Lz - 1 : z - 1) * L.Ly + y, rb = (int64_t)(z == L.Lz - 1 ? 0 : z + 1) * L.Ly + y;
                const uint4 F = bits_expand(*bits_word(q1, rf, nseg, seg), (int)(rf & 1));
                const uint4 B = bits_expand(*bits_word(q1, rb, nseg, seg), (int)(rb & 1));
                nup[0] += F.x + B.x; nup[1] += F.y + B.y; nup[2] += F.z + B.z; nup[3] += F.w + B.w;
            }
            const uint32_t tw[4] = {T.x, T.y, T.z, T.w}, cw[4] = {C.x, C.y, C.z, C.w};
            uint32_t a = 0, b = 0, c = 0, d = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                a = __dp4a(nup[w] & (tw[w] * 255u), 0x01010101u, a);
                b = __dp4a(nup[w], 0x01010101u, b);
                c = __dp4a(tw[w], 0x01010101u, c);
                d = __dp4a(cw[w], 0x01010101u, d);
            }
            e0n += a; nsum += b; e0 += c; e1 += d;
        }
    }
    e0n = warp_sum_ll(e0n); nsum = warp_sum_ll(nsum); e0 = warp_sum_ll(e0); e1 = warp_sum_ll(e1);
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
        long long pair = 4 * e0n - 2 * nsum - 2 * L.nn * e0, spin = 2 * (e0 + e1);
        if (blockIdx.x == 0 && threadIdx.x == 0) { pair += (long long)L.nn * L.halfN; spin -= 2 * L.halfN; }
        atomicAdd(o + SUM_PAIR, (unsigned long long)pair);
        atomicAdd(o + SUM_SPIN, (unsigned long long)spin);
    }
}

template <int COLOUR, bool HEATBATH, bool TRACK>
void launch_bits_t(mcx_lattice *lat, uint64_t t)
{
    // chain sub-ranges and row bands exactly as launch_v of k_ising2d.cu, with rows of half / 8 bytes
    LatView L = lat->view;
    const int c0 = g_launch_range.chain0, nch = g_launch_range.nchains < 0 ? lat->nchains : g_launch_range.nchains;
    L.planes += (int64_t)c0 * 2 * L.plane_stride;
    L.up_planes += (int64_t)c0 * 2 * L.plane_stride;
    L.dn_planes += (int64_t)c0 * 2 * L.plane_stride;
    L.nchains = nch;
    const bool band = g_launch_range.nrows > 0;
    if (band) {
        const int Ly = lat->view.Ly, y0 = g_launch_range.row0, nr = g_launch_range.nrows;   // y0, nr even
        const int64_t rowbytes = L.half >> 3;
        L.up_planes = L.planes + ((int64_t)((y0 - 1 + Ly) % Ly) - (nr - 1)) * rowbytes;     // its row nr - 1 is row y0 - 1
        L.dn_planes = L.planes + (int64_t)((y0 + nr) % Ly) * rowbytes;                      // its row 0 is row y0 + nr
        L.planes += (int64_t)y0 * rowbytes;
        L.Ly = nr;
        L.row_offset += y0;
    }
    cudaStream_t stream = g_launch_range.use_stream ? g_launch_range.stream : lat->ctx->stream;
    int R = band ? g_launch_range.R : 16;
    if (!band) {
        // 16-row strips unless that leaves fewer than ~4 work items per resident CTA (small batches)
        const int64_t ctas = (int64_t)lat->ctx->sm_count * 6;
        for (; R > 2; R >>= 1) {
            if (L.Ly % R != 0) continue;
            const int64_t items = ((int64_t)(L.Ly / R) * (L.half >> 4) + kBitThreads - 1) / kBitThreads * nch;
            if (items >= 4 * ctas || R <= 4) break;
        }
        while (L.Ly % R != 0) R -= 2;
        if (knobs().rows_per_strip > 0 && knobs().rows_per_strip <= 32 && L.Ly % knobs().rows_per_strip == 0 && knobs().rows_per_strip % 2 == 0) R = knobs().rows_per_strip;   // <= 32: the kernel flags left-over rows in 32 bits
    }
    const int nstrips = L.Ly / R;
    const int nseg = L.half >> 4;
    const int64_t G = (int64_t)nstrips * nseg;
    const int blocks_per_chain = (int)((G + kBitThreads - 1) / kBitThreads);
    const int nitems = (int)((int64_t)blocks_per_chain * nch);
    const bool full = G % kBitThreads == 0 && knobs().full != 0;
    auto kern = band ? (full ? k_ising2d_bits<COLOUR, HEATBATH, TRACK, true, true> : k_ising2d_bits<COLOUR, HEATBATH, TRACK, false, true>)
                     : (full ? k_ising2d_bits<COLOUR, HEATBATH, TRACK, true, false> : k_ising2d_bits<COLOUR, HEATBATH, TRACK, false, false>);
    static thread_local int resident = 0;
    if (!resident) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_ising2d_bits<COLOUR, HEATBATH, TRACK, false, false>, kBitThreads, 0);
        if (resident < 1) resident = 1;
    }
    const int ctas_per_sm = knobs().ctas_per_sm >= 0 ? knobs().ctas_per_sm : resident;
    int grid = lat->ctx->sm_count * ctas_per_sm;
    if (grid > nitems) grid = nitems;
    kern<<<grid, kBitThreads, 0, stream>>>(L, lat->d_thi, lat->d_tlo, lat->d_labels + c0, lat->d_sums + (int64_t)c0 * SUM_FIELDS,
                                          (uint32_t)lat->seed, (uint32_t)(lat->seed >> 32), t, lat->first_chain + (uint32_t)c0, R,
                                          nstrips, blocks_per_chain, nitems);
    lat->ctx->launches++;
}

template <int COLOUR>
void launch_bits_c(mcx_lattice *lat, uint64_t t)
{
    const bool track = lat->track_sums;
    if (lat->rule == MCX_HEATBATH) {
        if (track) launch_bits_t<COLOUR, true, true>(lat, t); else launch_bits_t<COLOUR, true, false>(lat, t);
    } else {
        if (track) launch_bits_t<COLOUR, false, true>(lat, t); else launch_bits_t<COLOUR, false, false>(lat, t);
    }
}

dim3 word_grid(const mcx_lattice *lat, int64_t words)
{
    int64_t blocks = (words + 255) / 256;
    const int64_t cap = (int64_t)lat->ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    return dim3((unsigned)blocks, (unsigned)lat->nchains);
}

}  // namespace

bool bits_shape_ok(int model, int ndim, const int32_t *dims)
{
    return model == MCX_ISING && (ndim == 2 || ndim == 3) && dims[0] % 32 == 0 && dims[1] % 2 == 0;
}

void launch_pack_bits(mcx_lattice *lat)
{
    const int64_t words = lat->view.halfN >> 5;
    k_pack_bits<<<word_grid(lat, words), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_staging, lat->N, words);
    lat->ctx->launches++;
}

void launch_unpack_bits(mcx_lattice *lat)
{
    const int64_t words = lat->view.halfN >> 5;
    k_unpack_bits<<<word_grid(lat, words), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_staging, lat->N, words);
    lat->ctx->launches++;
}

void launch_init_bits(mcx_lattice *lat, int mode, uint64_t seed)
{
    const int64_t words = lat->view.halfN >> 5;
    k_init_bits<<<word_grid(lat, words), 256, 0, lat->ctx->stream>>>(lat->view, mode, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                    lat->first_chain, words);
    lat->ctx->launches++;
}

void launch_recompute_bits(mcx_lattice *lat)
{
    const int64_t words = lat->view.halfN >> 5;
    k_recompute_bits<<<word_grid(lat, words), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_sums, words);
    lat->ctx->launches++;
}

void launch_hostbits_to_staging(mcx_lattice *lat, const void *d_bits, cudaStream_t stream)
{
    const int64_t nwords = lat->N * lat->nchains / 32;
    int64_t blocks = (nwords + 255) / 256;
    const int64_t cap = (int64_t)lat->ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_hostbits_to_staging<<<(unsigned)blocks, 256, 0, stream>>>((const uint32_t *)d_bits, lat->d_staging, nwords);
    lat->ctx->launches++;
}

void launch_staging_to_hostbits(mcx_lattice *lat, void *d_bits, cudaStream_t stream)
{
    const int64_t nwords = lat->N * lat->nchains / 32;
    int64_t blocks = (nwords + 255) / 256;
    const int64_t cap = (int64_t)lat->ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_staging_to_hostbits<<<(unsigned)blocks, 256, 0, stream>>>(lat->d_staging, (uint32_t *)d_bits, nwords);
    lat->ctx->launches++;
}

// half-sweep of a 2-D bit lattice (whole batch, a chain group or a row band: g_launch_range)
void launch_half_sweep_bits2d(mcx_lattice *lat, int colour, uint64_t t)
{
    if (colour == 0) launch_bits_c<0>(lat, t); else launch_bits_c<1>(lat, t);
}

}  // namespace mcx
