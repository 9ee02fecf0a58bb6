// k_bc3d.cu -- vectorised checkerboard half-sweep for 3-D Blume-Capel lattices (Metropolis / Glauber / heat bath).
//
// The k_bc2d rule (spin_flip!(sys::AbstractBlumeCapel, alg), SpinSystems/src/blume_capel.jl:41-85: Bool draw, _propose_state,
// acceptance draw against the host-built integer threshold T[(e * 2 + b) * 13 + raw]; heat bath: one draw against
// T0[raw], T1[raw]) on the k_ising3d decomposition: a thread owns a 16-byte column segment and walks a strip of rows inside
// one z-plane, two rows per trip, y-neighbour rows of the other colour plane in a rolling register window, the two
// z-neighbour rows of every row as two more 128-bit loads.  raw = sum of the six neighbours' encodings (0..12), so the
// tables have 78 (26) entries and the pair-threshold table 78 rows of 512 bytes (byte address = v1 * 512 + v0 * 4, 39.7 KB
// of shared memory).  Packed 15-bit decisions, exact 32-bit redo on ties; bit-identical to k_sweep_rows8 / k_sweep_generic.
#include "mcx_internal.h"

namespace mcx {

namespace {

constexpr int kB3Threads = 128;
constexpr int kB3Raw = 13;                         // 2 nn + 1 values of the neighbour sum
constexpr int kB3Table = 6 * kB3Raw;               // (e * 2 + b) * 13 + raw
constexpr int kB3HbTable = 2 * kB3Raw;             // T0[raw], T1[raw]
constexpr int kB3RowWords = 128;
constexpr int kB3PairWords = (kB3Table - 1) * kB3RowWords + kB3Table;
constexpr int kB3HbT1Words = kB3Raw * kB3RowWords; // the T1 pair table starts after thirteen rows of the T0 pair table

struct Bc3Acc {
    uint32_t nacc = 0;
    int32_t e = 0, p = 0, e1 = 0, p1 = 0, en = 0, pn = 0;     // as BcAcc of k_bc2d.cu
};

__device__ __forceinline__ uint32_t bc3_prop4(uint32_t e, uint32_t b)
{
    const uint32_t e0 = e & 0x01010101u, e1 = (e >> 1) & 0x01010101u;
    const uint32_t lower = (e0 | e1) ^ 0x01010101u, upper = 0x02020202u - e1;
    const uint32_t bm = b * 255u;
    return (lower & bm) | (upper & ~bm);
}

template <bool HB>
__device__ __noinline__ uint4 bc3_row_exact(uint4 tq, uint4 nq, uint4 bq, const uint32_t *thi, const uint32_t *tlo, Philox4 ah,
                                            Philox4 bh, Philox4 al, Philox4 bl)
{
    uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t nr[4] = {nq.x, nq.y, nq.z, nq.w}, bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        const int w = i >> 2, k = i & 3;
        const uint32_t e = (tw[w] >> (8 * k)) & 0xffu, n = (nr[w] >> (8 * k)) & 0xffu, b = (bb[w] >> (8 * k)) & 0xffu;
        const uint32_t hi = lane16(i < 8 ? ah : bh, i & 7), lo = lane16(i < 8 ? al : bl, i & 7);
        const uint64_t m = ((uint64_t)hi << 16) | lo;
        uint32_t out = e;
        if (HB) {
            const uint64_t T0 = ((uint64_t)thi[n] << 16) | tlo[n], T1 = ((uint64_t)thi[kB3Raw + n] << 16) | tlo[kB3Raw + n];
            out = m < T0 ? 0u : m < T1 ? 1u : 2u;
        } else {
            const int idx = (int)((e * 2 + b) * kB3Raw + n);
            const uint64_t T = ((uint64_t)thi[idx] << 16) | tlo[idx];
            if (m < T) out = e == 0 ? (b ? 1u : 2u) : e == 1 ? (b ? 0u : 2u) : (b ? 0u : 1u);
        }
        tw[w] = (tw[w] & ~(0xffu << (8 * k))) | (out << (8 * k));
    }
    return make_uint4(tw[0], tw[1], tw[2], tw[3]);
}

// One thread-row: 16 target sites tq; U, C, D: other-plane rows y-1, y, y+1 of the same z-plane; FB: byte-wise sum of the
// other-plane rows (z-1, y) and (z+1, y).  parity 0: in-row pair (j-1, j), parity 1: (j, j+1).
template <bool TRACK, bool HB>
__device__ __forceinline__ uint4 bc3_update_row(const uint4 tq, const uint4 U, const uint4 C, const uint4 D, const uint4 FB,
                                                const uint32_t side, const int parity, const uint32_t blk, const uint32_t t_lo,
                                                const uint32_t c2p0, const uint32_t c2p1, const uint32_t c2p2, const uint32_t c2p3,
                                                const uint32_t chain_id, const uint32_t seed_lo, const uint32_t seed_hi,
                                                const uint32_t *s_pair, const uint32_t *s_thi, const uint32_t *s_tlo, Bc3Acc &acc,
                                                const bool active)
{
    uint32_t B4[4] = {0u, 0u, 0u, 0u};
    if (!HB) {
        const Philox4 pa = philox4x32_10(blk, t_lo, c2p0, chain_id, seed_lo, seed_hi);
        const Philox4 pb = philox4x32_10(blk + 1, t_lo, c2p0, chain_id, seed_lo, seed_hi);
        B4[0] = (__byte_perm(pa.x, pa.y, 0x7531) >> 7) & 0x01010101u;
        B4[1] = (__byte_perm(pa.z, pa.w, 0x7531) >> 7) & 0x01010101u;
        B4[2] = (__byte_perm(pb.x, pb.y, 0x7531) >> 7) & 0x01010101u;
        B4[3] = (__byte_perm(pb.z, pb.w, 0x7531) >> 7) & 0x01010101u;
    }
    const Philox4 ra = philox4x32_10(blk, t_lo, HB ? c2p0 : c2p2, chain_id, seed_lo, seed_hi);
    const Philox4 rb = philox4x32_10(blk + 1, t_lo, HB ? c2p0 : c2p2, chain_id, seed_lo, seed_hi);

    const uint32_t W0 = parity ? C.x : side << 24, W1 = parity ? C.y : C.x, W2 = parity ? C.z : C.y, W3 = parity ? C.w : C.z,
                   W4 = parity ? side : C.w;
    const uint32_t sh = parity ? 8u : 24u;
    const uint32_t S[4] = {__funnelshift_r(W0, W1, sh), __funnelshift_r(W1, W2, sh), __funnelshift_r(W2, W3, sh),
                           __funnelshift_r(W3, W4, sh)};
    const uint32_t raw[4] = {U.x + D.x + C.x + S[0] + FB.x, U.y + D.y + C.y + S[1] + FB.y, U.z + D.z + C.z + S[2] + FB.z,
                             U.w + D.w + C.w + S[3] + FB.w};
    const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    uint32_t nw[4];
    uint32_t tie_min = 0x7fff7fffu;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t X = HB ? raw[w] : tw[w] * 26u + B4[w] * 13u + raw[w];          // byte = table index (<= 77)
        const uint32_t A = (X & 0x00ff00ffu) * 4u + ((X >> 8) & 0x00ff00ffu) * 512u;   // halfword = v1 * 512 + v0 * 4
        const uint32_t ttA = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair) + (A & 0xffffu));
        const uint32_t ttB = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair) + (A >> 16));
        const uint32_t hA = ((rw[2 * w] >> 1) & 0x7fff7fffu) | 0x80008000u;
        const uint32_t hB = ((rw[2 * w + 1] >> 1) & 0x7fff7fffu) | 0x80008000u;
        uint32_t rA, rB;
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rA) : "r"(ttA), "r"(hA));
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(rB) : "r"(ttB), "r"(hB));
        tie_min = __vmins2(__vmins2(tie_min, rA), rB);
        uint32_t P;   // 0xFF per site that is NOT accepted (heat bath: whose draw is not below T0)
        asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(P) : "r"(rA), "r"(rB));
        if (HB) {
            const uint32_t uuA = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair + kB3HbT1Words) + (A & 0xffffu));
            const uint32_t uuB = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair + kB3HbT1Words) + (A >> 16));
            uint32_t qA, qB, Q;
            asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(qA) : "r"(uuA), "r"(hA));
            asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(qB) : "r"(uuB), "r"(hB));
            tie_min = __vmins2(__vmins2(tie_min, qA), qB);
            asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(Q) : "r"(qA), "r"(qB));
            nw[w] = (P & 0x01010101u) + (Q & 0x01010101u);
        } else {
            nw[w] = (tw[w] & P) | (bc3_prop4(tw[w], B4[w]) & ~P);
        }
    }
    const bool tie = ((tie_min & 0x7fffu) == 0u) || ((tie_min & 0x7fff0000u) == 0u);
    if (tie) {
        const Philox4 la = philox4x32_10(blk, t_lo, HB ? c2p1 : c2p3, chain_id, seed_lo, seed_hi);
        const Philox4 lb = philox4x32_10(blk + 1, t_lo, HB ? c2p1 : c2p3, chain_id, seed_lo, seed_hi);
        const uint4 ex = bc3_row_exact<HB>(tq, make_uint4(raw[0], raw[1], raw[2], raw[3]), make_uint4(B4[0], B4[1], B4[2], B4[3]),
                                           s_thi, s_tlo, ra, rb, la, lb);
        nw[0] = ex.x; nw[1] = ex.y; nw[2] = ex.z; nw[3] = ex.w;
    }
    if (active) {
        uint32_t cnt = 0, se = 0, sp = 0, se1 = 0, sp1 = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t x = nw[w] ^ tw[w];
            const uint32_t ch = (x | (x >> 1)) & 0x01010101u, chm = ch * 255u;
            cnt += ch;
            if (TRACK) {
                const uint32_t eA = tw[w] & chm, pA = nw[w] & chm;
                se += eA; sp += pA; se1 += eA & 0x01010101u; sp1 += pA & 0x01010101u;
                acc.en = __dp4a(eA, raw[w], (uint32_t)acc.en);
                acc.pn = __dp4a(pA, raw[w], (uint32_t)acc.pn);
            }
        }
        acc.nacc = __dp4a(cnt, 0x01010101u, acc.nacc);
        if (TRACK) {
            acc.e = __dp4a(se, 0x01010101u, (uint32_t)acc.e);
            acc.p = __dp4a(sp, 0x01010101u, (uint32_t)acc.p);
            acc.e1 = __dp4a(se1, 0x01010101u, (uint32_t)acc.e1);
            acc.p1 = __dp4a(sp1, 0x01010101u, (uint32_t)acc.p1);
        }
    }
    return make_uint4(nw[0], nw[1], nw[2], nw[3]);
}

__device__ __forceinline__ uint4 b3_ld(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint4 b3_add(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <int COLOUR, bool TRACK, bool HB>
__global__ void __launch_bounds__(kB3Threads, 4)
k_bc3d(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g, const int32_t *__restrict__ labels,
       long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi, uint64_t t, uint32_t first_chain, int R, int strips_per_plane,
       int blocks_per_chain, int nitems)
{
    __shared__ uint32_t s_pair[kB3PairWords];
    __shared__ uint32_t s_thi[kB3Table], s_tlo[kB3Table];
    int cur_label = -1;
    constexpr int kTab = HB ? kB3HbTable : kB3Table;

    const int half = L.half;
    const int nseg = half >> 4;
    const int64_t G = (int64_t)strips_per_plane * L.Lz * nseg;
    const int lane = threadIdx.x & 31;
    const uint32_t t_lo = (uint32_t)t;
    const uint32_t c2p0 = ctr_word2(t, 0, TAG_SWEEP), c2p1 = ctr_word2(t, 1, TAG_SWEEP), c2p2 = ctr_word2(t, 2, TAG_SWEEP),
                   c2p3 = ctr_word2(t, 3, TAG_SWEEP);
    const int64_t plane_rows = (int64_t)L.Ly * half;

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chain = item / blocks_per_chain;
        const int label = labels[chain];
        if (label != cur_label) {
            __syncthreads();
            for (int i = threadIdx.x; i < kTab; i += kB3Threads) {
                s_thi[i] = thi_g[label * kTab + i];
                s_tlo[i] = tlo_g[label * kTab + i];
            }
            if (HB) {
                for (int i = threadIdx.x; i < 2 * kB3Raw * kB3Raw; i += kB3Threads) {
                    const int k = i / (kB3Raw * kB3Raw), j = i - k * kB3Raw * kB3Raw, v1 = j / kB3Raw, v0 = j - v1 * kB3Raw;
                    const uint32_t a = min(thi_g[label * kTab + k * kB3Raw + v0] >> 1, 0x7fffu);
                    const uint32_t b = min(thi_g[label * kTab + k * kB3Raw + v1] >> 1, 0x7fffu);
                    s_pair[k * kB3HbT1Words + v1 * kB3RowWords + v0] = a | (b << 16);
                }
            } else {
                for (int i = threadIdx.x; i < kB3Table * kB3Table; i += kB3Threads) {
                    const int v1 = i / kB3Table, v0 = i - v1 * kB3Table;
                    const uint32_t a = min(thi_g[label * kB3Table + v0] >> 1, 0x7fffu);
                    const uint32_t b = min(thi_g[label * kB3Table + v1] >> 1, 0x7fffu);
                    s_pair[v1 * kB3RowWords + v0] = a | (b << 16);
                }
            }
            __syncthreads();
            cur_label = label;
        }
        const int64_t g0 = (int64_t)(item - chain * blocks_per_chain) * kB3Threads + threadIdx.x;
        const bool active = g0 < G;
        const int64_t g = active ? g0 : G - 1;
        const int sidx = (int)(g / nseg);
        const int seg = (int)(g - (int64_t)sidx * nseg);
        const int z = sidx / strips_per_plane;
        const int y0 = (sidx - z * strips_per_plane) * R;         // even
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
        const bool edgeA = pa == 0 ? loadL : loadR, edgeB = pa == 0 ? loadR : loadL;
        const int colA = pa == 0 ? colL : colR, colB = pa == 0 ? colR : colL;

        const int yU = y0 == 0 ? L.Ly - 1 : y0 - 1;
        const uint8_t *po = oth + (int64_t)y0 * half;
        uint8_t *pt = tgt + (int64_t)y0 * half + col;
        int64_t zoff = (int64_t)y0 * half + col;
        uint4 U = b3_ld(oth + (int64_t)yU * half + col);
        uint4 C = b3_ld(po + col);
        uint32_t blk = (uint32_t)((((int64_t)z * L.Ly + y0) * half + col) >> 3);
        const uint32_t blk_step = (uint32_t)(half >> 3);
        Bc3Acc acc;

#pragma unroll 1
        for (int r = 0; r < R; r += 2) {
            const int y = y0 + r;
            const uint8_t *pe = (y + 2 == L.Ly) ? oth : po + 2 * (int64_t)half;
            const uint4 E = b3_ld(pe + col);
            const uint4 D = b3_ld(po + half + col);
            const uint4 Ta = b3_ld(pt), Tb = b3_ld(pt + half);
            const uint4 FBa = b3_add(b3_ld(othF + zoff), b3_ld(othB + zoff));
            const uint4 FBb = b3_add(b3_ld(othF + zoff + half), b3_ld(othB + zoff + half));
            uint32_t sideA = 0, sideB = 0;
            if (edgeA) sideA = po[colA];
            if (edgeB) sideB = po[half + colB];
            const uint32_t cl = __shfl_up_sync(0xffffffffu, C.w, 1) >> 24, cr = __shfl_down_sync(0xffffffffu, C.x, 1) & 0xffu;
            const uint32_t dl = __shfl_up_sync(0xffffffffu, D.w, 1) >> 24, dr = __shfl_down_sync(0xffffffffu, D.x, 1) & 0xffu;
            uint32_t sA = pa == 0 ? cl : cr, sB = pa == 0 ? dr : dl;
            if (edgeA) sA = sideA;
            if (edgeB) sB = sideB;
            const uint4 Na = bc3_update_row<TRACK, HB>(Ta, U, C, D, FBa, sA, pa, blk, t_lo, c2p0, c2p1, c2p2, c2p3, chain_id, seed_lo,
                                                       seed_hi, s_pair, s_thi, s_tlo, acc, active);
            if (active) *reinterpret_cast<uint4 *>(pt) = Na;
            asm volatile("" ::: "memory");
            const uint4 Nb = bc3_update_row<TRACK, HB>(Tb, C, D, E, FBb, sB, pa ^ 1, blk + blk_step, t_lo, c2p0, c2p1, c2p2, c2p3, chain_id,
                                                       seed_lo, seed_hi, s_pair, s_thi, s_tlo, acc, active);
            if (active) *reinterpret_cast<uint4 *>(pt + half) = Nb;
            U = D; C = E;
            po += 2 * (int64_t)half; pt += 2 * (int64_t)half; zoff += 2 * (int64_t)half; blk += 2 * blk_step;
        }

        // per-chain sums over the changed sites, with s = e - 1 and nbr = raw - 6
        const int nacc = warp_sum((int)acc.nacc);
        int dspin = 0, dspin2 = 0, dpair = 0;
        if (TRACK) {
            const int se = warp_sum(acc.e), sp = warp_sum(acc.p), se1 = warp_sum(acc.e1), sp1 = warp_sum(acc.p1);
            const int sen = warp_sum(acc.en), spn = warp_sum(acc.pn);
            dspin = sp - se;
            dspin2 = se1 - sp1;
            dpair = (spn - sen) - 6 * (sp - se);
        }
        if (lane == 0) {
            unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
            if (nacc) atomicAdd(o + SUM_ACC, (unsigned long long)(long long)nacc);
            if (TRACK) {
                if (dpair) atomicAdd(o + SUM_PAIR, (unsigned long long)(long long)dpair);
                if (dspin) atomicAdd(o + SUM_SPIN, (unsigned long long)(long long)dspin);
                if (dspin2) atomicAdd(o + SUM_SPIN2, (unsigned long long)(long long)dspin2);
            }
        }
    }
}

template <int COLOUR, bool TRACK, bool HB>
void launch_b3(mcx_lattice *lat, uint64_t t)
{
    LatView L = lat->view;
    const int c0 = g_launch_range.chain0, nch = g_launch_range.nchains < 0 ? lat->nchains : g_launch_range.nchains;
    L.planes += (int64_t)c0 * 2 * L.plane_stride;
    L.nchains = nch;
    cudaStream_t stream = g_launch_range.use_stream ? g_launch_range.stream : lat->ctx->stream;
    const int nseg = L.half >> 4;
    auto kern = k_bc3d<COLOUR, TRACK, HB>;
    static thread_local int resident = 0;
    if (!resident) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kB3Threads, 0);
        if (resident < 1) resident = 1;
    }
    const int64_t ctas = (int64_t)lat->ctx->sm_count * resident;
    int R = 2;
    for (int r = 16; r >= 2; r -= 2) {
        if (L.Ly % r != 0) continue;
        const int64_t items = ((int64_t)(L.Ly / r) * L.Lz * nseg + kB3Threads - 1) / kB3Threads * lat->nchains;
        R = r;
        if (items >= 4 * ctas || r <= 4) break;
    }
    const int strips_per_plane = L.Ly / R;
    const int64_t G = (int64_t)strips_per_plane * L.Lz * nseg;
    const int blocks_per_chain = (int)((G + kB3Threads - 1) / kB3Threads);
    const int nitems = (int)((int64_t)blocks_per_chain * nch);
    int grid = (int)ctas;
    if (grid > nitems) grid = nitems;
    kern<<<grid, kB3Threads, 0, stream>>>(L, lat->d_thi, lat->d_tlo, lat->d_labels + c0, lat->d_sums + (int64_t)c0 * SUM_FIELDS,
                                         (uint32_t)lat->seed, (uint32_t)(lat->seed >> 32), t, lat->first_chain + (uint32_t)c0, R,
                                         strips_per_plane, blocks_per_chain, nitems);
    lat->ctx->launches++;
}

}  // namespace

// false: not applicable (shape, storage or table layout), nothing launched
bool launch_sweep_bc3d(mcx_lattice *lat, int colour, uint64_t t)
{
    if (lat->ndim != 3 || lat->model != MCX_BLUME_CAPEL || lat->storage != MCX_STORAGE_INT8 || lat->view.Lx % 32 != 0) return false;
    const bool hb = lat->rule == MCX_HEATBATH;
    if (lat->view.Ly % 2 != 0 || lat->table_len != (hb ? kB3HbTable : kB3Table) || knobs().bc2d == 0) return false;
    if ((int64_t)(lat->view.Ly / 2) * lat->view.Lz * (lat->view.half >> 4) < 96) return false;     // tiny: rows-of-8 kernel
    const bool track = lat->track_sums;
    if (hb) {
        if (colour == 0) { if (track) launch_b3<0, true, true>(lat, t); else launch_b3<0, false, true>(lat, t); }
        else             { if (track) launch_b3<1, true, true>(lat, t); else launch_b3<1, false, true>(lat, t); }
    } else {
        if (colour == 0) { if (track) launch_b3<0, true, false>(lat, t); else launch_b3<0, false, false>(lat, t); }
        else             { if (track) launch_b3<1, true, false>(lat, t); else launch_b3<1, false, false>(lat, t); }
    }
    return true;
}

}  // namespace mcx
