// k_ising2d.cu -- vectorised checkerboard half-sweep for 2-D Ising lattices, one byte per spin.
//
// Work decomposition.  The target colour plane is a [Ly][Lx/2] byte matrix.  One thread owns a
// 16-byte column segment (one 128-bit load/store per row) and walks down a strip of R rows,
// keeping the three neighbour rows of the *other* colour plane (up, centre, down) in a rolling
// register window, so every byte of the other plane is read once per strip (+2 halo rows) and
// every byte of the target plane is read once and written once: 3 bytes per attempt, the
// algorithmic minimum.  The left/right neighbour byte that falls outside the
// thread's segment comes from the adjacent lane by warp shuffle (edge lanes load one byte).
//
// Randomness.  A thread-row needs 16 draws = two Philox4x32-10 blocks (eight 16-bit lanes each).
// The 16-bit lane is the HIGH half of the 32-bit draw; the decision m < T is settled by the high
// halves unless they tie (probability 2^-16 per site), in which case the row is redone exactly
// with the low halves from plane 1.  Results are bit-identical to k_sweep_generic and to the
// oracle for every shape both accept.
//
// Persistent grid: gridDim.x = SMs x resident CTAs, items handed out round-robin.
#include "mcx_internal.h"

#include <cstdlib>

namespace mcx {

namespace {

constexpr int kThreads = 128;
constexpr int kTableLen = 10;   // 2 * (nn + 1) for nn = 4

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int b) { return __byte_perm(w, 0, 0x4440 | b); }

struct Acc {
    uint32_t flips = 0;   // number of changed sites
    int32_t s = 0;        // sum over changed sites of s        (s in {0,1})
    int32_t n = 0;        // sum over changed sites of nup
    int32_t sn = 0;       // sum over changed sites of s*nup
};

// Exact redo of one thread-row with full 32-bit draws (taken when a high-half tie was seen).
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

template <bool HEATBATH, bool TRACK>
__global__ void __launch_bounds__(kThreads)
k_ising2d(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
          const int32_t *__restrict__ labels, int n_labels, long long *__restrict__ sums, uint32_t seed_lo,
          uint32_t seed_hi, uint64_t t, int colour, uint32_t first_chain, int R, int nstrips,
          int blocks_per_chain, int nitems)
{
    extern __shared__ uint32_t smem[];
    uint32_t *s_thi = smem;                               // [n_labels][10]
    uint32_t *s_tlo = smem + n_labels * kTableLen;        // [n_labels][10]
    for (int i = threadIdx.x; i < n_labels * kTableLen; i += kThreads) {
        s_thi[i] = thi_g[i];
        s_tlo[i] = tlo_g[i];
    }
    __syncthreads();

    const int half = L.half;
    const int nseg = half >> 4;
    const int64_t G = (int64_t)nstrips * nseg;
    const int lane = threadIdx.x & 31;
    const uint32_t c2 = ctr_word2(t, 0, TAG_SWEEP);
    const uint32_t c2lo = ctr_word2(t, 1, TAG_SWEEP);

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chain = item / blocks_per_chain;
        const int64_t g0 = (int64_t)(item - chain * blocks_per_chain) * kThreads + threadIdx.x;
        const bool active = g0 < G;
        const int64_t g = active ? g0 : G - 1;
        const int strip = (int)(g / nseg);
        const int seg = (int)(g - (int64_t)strip * nseg);
        const int row0 = strip * R;
        const uint32_t chain_id = first_chain + (uint32_t)chain;
        const uint32_t *thi = s_thi + labels[chain] * kTableLen;
        const uint32_t *tlo = s_tlo + labels[chain] * kTableLen;

        uint8_t *tgt = plane_ptr(L, chain, colour);
        const uint8_t *__restrict__ oth = plane_ptr(L, chain, colour ^ 1);
        const int col = seg << 4;
        const int colL = (seg == 0 ? half : col) - 1;             // byte left of the segment (periodic)
        const int colR = (seg == nseg - 1) ? 0 : col + 16;        // byte right of the segment
        const bool loadL = (lane == 0) || (seg == 0);
        const bool loadR = (lane == 31) || (seg == nseg - 1);

        auto row_ptr = [&](const uint8_t *base, int row) { return base + (int64_t)row * half; };
        auto down_of = [&](int row) { return row == L.Ly - 1 ? 0 : row + 1; };
        // side byte of `row` for the edge lanes (parity p: 0 -> left byte, 1 -> right byte)
        auto edge_side = [&](int row) -> uint32_t {
            const int p = (colour + row) & 1;
            if (p == 0 ? loadL : loadR) return row_ptr(oth, row)[p == 0 ? colL : colR];
            return 0u;
        };
        const int rowU = row0 == 0 ? L.Ly - 1 : row0 - 1;
        uint4 U = *reinterpret_cast<const uint4 *>(row_ptr(oth, rowU) + col);
        uint4 C = *reinterpret_cast<const uint4 *>(row_ptr(oth, row0) + col);
        uint4 D = *reinterpret_cast<const uint4 *>(row_ptr(oth, down_of(row0)) + col);
        uint4 Tq = *reinterpret_cast<const uint4 *>(row_ptr(tgt, row0) + col);
        uint32_t side_edge = edge_side(row0);
        Acc acc;

#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const int row = row0 + r;
            // prefetch the next row's inputs before working on this one
            uint4 Dn = D, Tn = Tq;
            uint32_t side_edge_n = 0;
            if (r + 1 < R) {
                Dn = *reinterpret_cast<const uint4 *>(row_ptr(oth, down_of(row + 1)) + col);
                Tn = *reinterpret_cast<const uint4 *>(row_ptr(tgt, row + 1) + col);
                side_edge_n = edge_side(row + 1);
            }
            const int p = (colour + row) & 1;   // x offset of the target sites in this row

            // the two Philox blocks of this thread-row
            const uint32_t blk = (uint32_t)(((int64_t)row * half + col) >> 3);
            const Philox4 ra = philox4x32_10(blk, (uint32_t)t, c2, chain_id, seed_lo, seed_hi);
            const Philox4 rb = philox4x32_10(blk + 1, (uint32_t)t, c2, chain_id, seed_lo, seed_hi);

            // neighbour in the same row: other-plane byte j-1 (p == 0) or j+1 (p == 1)
            uint32_t S[4];
            if (p == 0) {
                uint32_t side = __shfl_up_sync(0xffffffffu, C.w, 1) >> 24;
                if (loadL) side = side_edge;
                S[0] = (C.x << 8) | side;
                S[1] = __funnelshift_l(C.x, C.y, 8);
                S[2] = __funnelshift_l(C.y, C.z, 8);
                S[3] = __funnelshift_l(C.z, C.w, 8);
            } else {
                uint32_t side = __shfl_down_sync(0xffffffffu, C.x, 1) & 0xffu;
                if (loadR) side = side_edge;
                S[0] = __funnelshift_r(C.x, C.y, 8);
                S[1] = __funnelshift_r(C.y, C.z, 8);
                S[2] = __funnelshift_r(C.z, C.w, 8);
                S[3] = (C.w >> 8) | (side << 24);
            }
            const uint32_t nup[4] = {U.x + D.x + C.x + S[0], U.y + D.y + C.y + S[1], U.z + D.z + C.z + S[2],
                                     U.w + D.w + C.w + S[3]};
            const uint32_t tw[4] = {Tq.x, Tq.y, Tq.z, Tq.w};
            uint32_t nw[4];
            bool tie = false;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t rlo = (w & 1) ? ((w < 2) ? ra.z : rb.z) : ((w < 2) ? ra.x : rb.x);
                const uint32_t rhi = (w & 1) ? ((w < 2) ? ra.w : rb.w) : ((w < 2) ? ra.y : rb.y);
                const uint32_t idx4 = tw[w] * 20u + nup[w] * 4u;     // byte b = 4 * (5 s + nup)
                uint32_t F = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t thr = *reinterpret_cast<const uint32_t *>(
                        reinterpret_cast<const char *>(thi) + byte_of(idx4, b));
                    const uint32_t rr = (b < 2) ? rlo : rhi;
                    const uint32_t h = (b & 1) ? (rr >> 16) : (rr & 0xffffu);
                    if (h < thr) F |= 1u << (8 * b);
                    tie |= (h == thr);
                }
                nw[w] = HEATBATH ? F : (tw[w] ^ F);
            }
            if (tie) {
                // rare: settle with the low halves (plane 1) for the whole thread-row
                const Philox4 la = philox4x32_10(blk, (uint32_t)t, c2lo, chain_id, seed_lo, seed_hi);
                const Philox4 lb = philox4x32_10(blk + 1, (uint32_t)t, c2lo, chain_id, seed_lo, seed_hi);
                const uint4 ex = row_exact<HEATBATH>(Tq, make_uint4(nup[0], nup[1], nup[2], nup[3]), thi, tlo, ra, rb,
                                                     la, lb);
                nw[0] = ex.x; nw[1] = ex.y; nw[2] = ex.z; nw[3] = ex.w;
            }
            if (active) {
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
                *reinterpret_cast<uint4 *>(tgt + (int64_t)row * half + col) = make_uint4(nw[0], nw[1], nw[2], nw[3]);
            }
            U = C; C = D; D = Dn; Tq = Tn; side_edge = side_edge_n;
        }

        // per-chain sums: dspin = 2 - 4 s, dpair = -8 s nup + 16 s + 4 nup - 8 per changed site
        int nflip = warp_sum((int)acc.flips);
        int dspin = 0, dpair = 0;
        if (TRACK) {
            const int ss = warp_sum(acc.s), nn_ = warp_sum(acc.n), sn = warp_sum(acc.sn);
            dspin = 2 * nflip - 4 * ss;
            dpair = -8 * sn + 16 * ss + 4 * nn_ - 8 * nflip;
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

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int pick_rows_per_strip(int Ly, int want)
{
    // largest even divisor of Ly that is <= want
    for (int r = want; r >= 2; --r)
        if ((r % 2) == 0 && (Ly % r) == 0) return r;
    return 2;
}

template <bool HEATBATH, bool TRACK>
void launch_t(mcx_lattice *lat, int colour, uint64_t t)
{
    const LatView &L = lat->view;
    const int R = pick_rows_per_strip(L.Ly, env_int("MCX_ROWS_PER_STRIP", 16));
    const int nstrips = L.Ly / R;
    const int nseg = L.half >> 4;
    const int64_t G = (int64_t)nstrips * nseg;
    const int blocks_per_chain = (int)((G + kThreads - 1) / kThreads);
    const int64_t nitems64 = (int64_t)blocks_per_chain * lat->nchains;
    const int nitems = (int)nitems64;
    const size_t smem = (size_t)lat->n_labels * kTableLen * 2 * sizeof(uint32_t);
    auto kern = k_ising2d<HEATBATH, TRACK>;
    static thread_local int resident = 0;
    if (!resident) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kThreads, smem);
        if (resident < 1) resident = 1;
    }
    const int ctas_per_sm = env_int("MCX_CTAS_PER_SM", resident);
    int grid = lat->ctx->sm_count * ctas_per_sm;
    if (grid > nitems) grid = nitems;
    kern<<<grid, kThreads, smem, lat->ctx->stream>>>(L, lat->d_thi, lat->d_tlo, lat->d_labels, lat->n_labels,
                                                    lat->d_sums, (uint32_t)lat->seed, (uint32_t)(lat->seed >> 32), t,
                                                    colour, lat->first_chain, R, nstrips, blocks_per_chain, nitems);
    lat->ctx->launches++;
}

}  // namespace

bool launch_sweep_ising2d(mcx_lattice *lat, int colour, uint64_t t)
{
    if (!lat->fast2d || lat->model != MCX_ISING || lat->storage != MCX_STORAGE_INT8) return false;
    if ((size_t)lat->n_labels * kTableLen * 8 > 40 * 1024) return false;
    const bool track = lat->track_sums;
    if (lat->rule == MCX_HEATBATH) {
        if (track) launch_t<true, true>(lat, colour, t); else launch_t<true, false>(lat, colour, t);
    } else {
        if (track) launch_t<false, true>(lat, colour, t); else launch_t<false, false>(lat, colour, t);
    }
    return true;
}

}  // namespace mcx
