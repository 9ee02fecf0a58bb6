// k_bc2d.cu -- vectorised checkerboard half-sweep for 2-D Blume-Capel lattices (Metropolis / Glauber).
//
// Same decomposition as k_ising2d.cu (16-byte column segments, strips of rows, two rows per trip, rolling
// register window of the other colour plane, warp-shuffle side bytes, persistent grid) with the
// Blume-Capel rule of spin_flip!(sys::AbstractBlumeCapel, alg) (SpinSystems/src/blume_capel.jl:41-59):
//   b      = rand(rng, Bool)                  draw 0 of the site (Philox plane 0, bit 15 of its lane)
//   s_new  = _propose_state(s, b)             blume_capel.jl:21-30
//   accept = rand(rng) < p(s, s_new, sum)     draw 1 (planes 2 / 3) against the host-built integer threshold
//                                             T[(e * 2 + b) * 9 + raw], e = s + 1, raw = sum of the four
//                                             neighbours' encodings (0..8)
// so a thread-row of 16 sites takes four Philox4x32-10 blocks (two for the Bool draws, two for the
// acceptance draws).  The decision is the packed 15-bit comparison of k_ising2d: two sites per instruction
// against a pair-threshold table in shared memory (54 rows of 256 B: byte address = v1 * 256 + v0 * 4),
// ties (2^-15 per site) redone exactly with the 32-bit draws.  Bit-identical to k_sweep_rows8 and
// k_sweep_generic.
//
// Heat bath (blume_capel.jl:61-85): ONE Float64 draw per site (planes 0 / 1) against two thresholds that depend on the
// neighbour sum only, new = m < T0[raw] ? -1 : m < T1[raw] ? 0 : +1 -- two Philox blocks per thread-row instead of four,
// two packed comparisons per pair of sites against two 9 x 9 pair tables, new encoding = [m >= T0] + [m >= T1].
#include "mcx_internal.h"

#include <cstdlib>

namespace mcx {

namespace {

constexpr int kBcThreads = 128;
constexpr int kBcTable = 54;
constexpr int kBcRowWords = 64;
constexpr int kBcPairWords = (kBcTable - 1) * kBcRowWords + kBcTable;

struct BcAcc {
    uint32_t nacc = 0;                 // changed (= accepted) sites
    int32_t e = 0, p = 0;              // sums of old / new encodings over changed sites
    int32_t e1 = 0, p1 = 0;            // how many of them were / became 1 (spin 0)
    int32_t en = 0, pn = 0;            // sums of old / new encoding times raw
};

__device__ __forceinline__ uint32_t bc_prop4(uint32_t e, uint32_t b)
{
    // _propose_state on four sites: b ? lower : upper of the two other states, with
    // lower = (e == 0 ? 1 : 0), upper = (e == 2 ? 1 : 2)
    const uint32_t e0 = e & 0x01010101u, e1 = (e >> 1) & 0x01010101u;
    const uint32_t lower = (e0 | e1) ^ 0x01010101u, upper = 0x02020202u - e1;
    const uint32_t bm = b * 255u;
    return (lower & bm) | (upper & ~bm);
}

// exact redo of one thread-row with the full 32-bit acceptance draws
__device__ __noinline__ uint4 bc_row_exact(uint4 tq, uint4 nq, uint4 bq, const uint32_t *thi, const uint32_t *tlo,
                                           Philox4 a2, Philox4 b2, Philox4 a3, Philox4 b3)
{
    uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t nr[4] = {nq.x, nq.y, nq.z, nq.w}, bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        const int w = i >> 2, k = i & 3;
        const uint32_t e = (tw[w] >> (8 * k)) & 0xffu, n = (nr[w] >> (8 * k)) & 0xffu, b = (bb[w] >> (8 * k)) & 0xffu;
        const int idx = (int)((e * 2 + b) * 9 + n);
        const uint32_t hi = lane16(i < 8 ? a2 : b2, i & 7), lo = lane16(i < 8 ? a3 : b3, i & 7);
        const uint64_t m = ((uint64_t)hi << 16) | lo;
        const uint64_t T = ((uint64_t)thi[idx] << 16) | tlo[idx];
        if (m < T) {
            const uint32_t prop = e == 0 ? (b ? 1u : 2u) : e == 1 ? (b ? 0u : 2u) : (b ? 0u : 1u);
            tw[w] = (tw[w] & ~(0xffu << (8 * k))) | (prop << (8 * k));
        }
    }
    return make_uint4(tw[0], tw[1], tw[2], tw[3]);
}


// heat bath: exact redo of one thread-row with the full 32-bit draws
__device__ __noinline__ uint4 bc_row_exact_hb(uint4 nq, const uint32_t *thi, const uint32_t *tlo, Philox4 a0, Philox4 b0, Philox4 a1,
                                              Philox4 b1)
{
    uint32_t out[4] = {0u, 0u, 0u, 0u};
    const uint32_t nr[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        const int w = i >> 2, k = i & 3;
        const uint32_t n = (nr[w] >> (8 * k)) & 0xffu;
        const uint32_t hi = lane16(i < 8 ? a0 : b0, i & 7), lo = lane16(i < 8 ? a1 : b1, i & 7);
        const uint64_t m = ((uint64_t)hi << 16) | lo;
        const uint64_t T0 = ((uint64_t)thi[n] << 16) | tlo[n], T1 = ((uint64_t)thi[9 + n] << 16) | tlo[9 + n];
        out[w] |= (m < T0 ? 0u : m < T1 ? 1u : 2u) << (8 * k);
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

constexpr int kHbTable = 18;                  // 2 * (2 nn + 1): thresholds T0[raw], T1[raw]
constexpr int kHbT1Words = 9 * kBcRowWords;   // the T1 pair table starts after nine rows of the T0 pair table

template <int PARITY, bool TRACK, bool HB>
__device__ __forceinline__ uint4 bc_update_row(const uint4 tq, const uint4 U, const uint4 C, const uint4 D, const uint32_t side,
                                               const uint32_t blk, const uint32_t t_lo, const uint32_t c2p0, const uint32_t c2p1,
                                               const uint32_t c2p2, const uint32_t c2p3, const uint32_t chain_id, const uint32_t seed_lo,
                                               const uint32_t seed_hi, const uint32_t *s_pair, const uint32_t *s_thi,
                                               const uint32_t *s_tlo, BcAcc &acc, const bool active)
{
    // Bool draws: bit 15 of every 16-bit lane of plane 0, as one byte per site
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
    const uint32_t raw[4] = {U.x + D.x + C.x + S[0], U.y + D.y + C.y + S[1], U.z + D.z + C.z + S[2], U.w + D.w + C.w + S[3]};
    const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};   // word k: sites 2k, 2k+1
    uint32_t nw[4];
    uint32_t tie_min = 0x7fff7fffu;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t X = HB ? raw[w] : tw[w] * 18u + B4[w] * 9u + raw[w];   // byte = table index: raw, or (e * 2 + b) * 9 + raw (<= 53)
        const uint32_t A = X + 3u * (X & 0x00ff00ffu);              // even bytes * 4: halfword = v1 * 256 + v0 * 4
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
            const uint32_t uuA = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair + kHbT1Words) + (A & 0xffffu));
            const uint32_t uuB = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pair + kHbT1Words) + (A >> 16));
            uint32_t qA, qB, Q;
            asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(qA) : "r"(uuA), "r"(hA));
            asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(qB) : "r"(uuB), "r"(hB));
            tie_min = __vmins2(__vmins2(tie_min, qA), qB);
            asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(Q) : "r"(qA), "r"(qB));
            nw[w] = (P & 0x01010101u) + (Q & 0x01010101u);         // new encoding = [m >= T0] + [m >= T1]
        } else {
            nw[w] = (tw[w] & P) | (bc_prop4(tw[w], B4[w]) & ~P);
        }
    }
    const bool tie = ((tie_min & 0x7fffu) == 0u) || ((tie_min & 0x7fff0000u) == 0u);
    if (tie) {
        const Philox4 la = philox4x32_10(blk, t_lo, HB ? c2p1 : c2p3, chain_id, seed_lo, seed_hi);
        const Philox4 lb = philox4x32_10(blk + 1, t_lo, HB ? c2p1 : c2p3, chain_id, seed_lo, seed_hi);
        const uint4 ex = HB ? bc_row_exact_hb(make_uint4(raw[0], raw[1], raw[2], raw[3]), s_thi, s_tlo, ra, rb, la, lb)
                            : bc_row_exact(tq, make_uint4(raw[0], raw[1], raw[2], raw[3]), make_uint4(B4[0], B4[1], B4[2], B4[3]),
                                           s_thi, s_tlo, ra, rb, la, lb);
        nw[0] = ex.x; nw[1] = ex.y; nw[2] = ex.z; nw[3] = ex.w;
    }
    if (active) {
        uint32_t cnt = 0, se = 0, sp = 0, se1 = 0, sp1 = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t x = nw[w] ^ tw[w];                                   // non-zero byte: changed site
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

__device__ __forceinline__ uint4 bc_ldg128(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }

template <int COLOUR, bool TRACK, bool HB>
__global__ void __launch_bounds__(kBcThreads, 5)
k_bc2d(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g, const int32_t *__restrict__ labels,
       long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi, uint64_t t, uint32_t first_chain, int R, int nstrips,
       int blocks_per_chain, int nitems)
{
    __shared__ uint32_t s_pair[kBcPairWords];
    __shared__ uint32_t s_thi[kBcTable], s_tlo[kBcTable];
    int cur_label = -1;

    const int half = L.half;
    const int nseg = half >> 4;
    const int64_t G = (int64_t)nstrips * nseg;
    const int lane = threadIdx.x & 31;
    const uint32_t t_lo = (uint32_t)t;
    const uint32_t c2p0 = ctr_word2(t, 0, TAG_SWEEP), c2p1 = ctr_word2(t, 1, TAG_SWEEP), c2p2 = ctr_word2(t, 2, TAG_SWEEP),
                   c2p3 = ctr_word2(t, 3, TAG_SWEEP);
    constexpr int kTab = HB ? kHbTable : kBcTable;

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chain = item / blocks_per_chain;
        const int label = labels[chain];
        if (label != cur_label) {
            // pair table of this chain's ensemble: word (v1, v0) = t15[v0] | t15[v1] << 16, t15 = min(T >> 17, 0x7fff)
            __syncthreads();
            for (int i = threadIdx.x; i < kTab; i += kBcThreads) {
                s_thi[i] = thi_g[label * kTab + i];
                s_tlo[i] = tlo_g[label * kTab + i];
            }
            if (HB) {
                // two 9 x 9 pair tables: T0[raw] (word offset 0) and T1[raw] (word offset kHbT1Words)
                for (int i = threadIdx.x; i < 2 * 81; i += kBcThreads) {
                    const int k = i / 81, j = i - k * 81, v1 = j / 9, v0 = j - v1 * 9;
                    const uint32_t a = min(thi_g[label * kTab + k * 9 + v0] >> 1, 0x7fffu);
                    const uint32_t b = min(thi_g[label * kTab + k * 9 + v1] >> 1, 0x7fffu);
                    s_pair[k * kHbT1Words + v1 * kBcRowWords + v0] = a | (b << 16);
                }
            } else {
                for (int i = threadIdx.x; i < kBcTable * kBcTable; i += kBcThreads) {
                    const int v1 = i / kBcTable, v0 = i - v1 * kBcTable;
                    const uint32_t a = min(thi_g[label * kBcTable + v0] >> 1, 0x7fffu);
                    const uint32_t b = min(thi_g[label * kBcTable + v1] >> 1, 0x7fffu);
                    s_pair[v1 * kBcRowWords + v0] = a | (b << 16);
                }
            }
            __syncthreads();
            cur_label = label;
        }
        const int64_t g0 = (int64_t)(item - chain * blocks_per_chain) * kBcThreads + threadIdx.x;
        const bool active = g0 < G;
        const int64_t g = active ? g0 : G - 1;
        const int strip = (int)(g / nseg);
        const int seg = (int)(g - (int64_t)strip * nseg);
        const int row0 = strip * R;                               // even
        const uint32_t chain_id = first_chain + (uint32_t)chain;

        uint8_t *tgt = plane_ptr(L, chain, COLOUR);
        const uint8_t *__restrict__ oth = plane_ptr(L, chain, COLOUR ^ 1);
        const int col = seg << 4;
        const int colL = (seg == 0 ? half : col) - 1;
        const int colR = (seg == nseg - 1) ? 0 : col + 16;
        const bool loadL = (lane == 0) || (seg == 0);
        const bool loadR = (lane == 31) || (seg == nseg - 1);
        const bool edgeA = COLOUR == 0 ? loadL : loadR;
        const bool edgeB = COLOUR == 0 ? loadR : loadL;
        const int colA = COLOUR == 0 ? colL : colR;
        const int colB = COLOUR == 0 ? colR : colL;

        const int rowU = row0 == 0 ? L.Ly - 1 : row0 - 1;
        const uint8_t *po = oth + (int64_t)row0 * half;
        uint8_t *pt = tgt + (int64_t)row0 * half + col;
        uint4 U = bc_ldg128(oth + (int64_t)rowU * half + col);
        uint4 C = bc_ldg128(po + col);
        uint32_t blk = (uint32_t)(((int64_t)row0 * half + col) >> 3);
        const uint32_t blk_step = (uint32_t)(half >> 3);
        BcAcc acc;

#pragma unroll 1
        for (int r = 0; r < R; r += 2) {
            const int row = row0 + r;
            const uint8_t *pe = (row + 2 == L.Ly) ? oth : po + 2 * (int64_t)half;
            const uint4 E = bc_ldg128(pe + col);
            const uint4 D = bc_ldg128(po + half + col);
            const uint4 Ta = bc_ldg128(pt), Tb = bc_ldg128(pt + half);
            uint32_t sideA = 0, sideB = 0;
            if (edgeA) sideA = po[colA];
            if (edgeB) sideB = po[half + colB];
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
            const uint4 Na = bc_update_row<COLOUR, TRACK, HB>(Ta, U, C, D, sA, blk, t_lo, c2p0, c2p1, c2p2, c2p3, chain_id, seed_lo, seed_hi,
                                                         s_pair, s_thi, s_tlo, acc, active);
            if (active) *reinterpret_cast<uint4 *>(pt) = Na;
            asm volatile("" ::: "memory");
            const uint4 Nb = bc_update_row<COLOUR ^ 1, TRACK, HB>(Tb, C, D, E, sB, blk + blk_step, t_lo, c2p0, c2p1, c2p2, c2p3, chain_id,
                                                             seed_lo, seed_hi, s_pair, s_thi, s_tlo, acc, active);
            if (active) *reinterpret_cast<uint4 *>(pt + half) = Nb;
            U = D; C = E;
            po += 2 * (int64_t)half; pt += 2 * (int64_t)half; blk += 2 * blk_step;
        }

        // per-chain sums over the changed sites, with s = e - 1 and nbr = raw - 4:
        //   dspin = sum(p - e), dspin2 = sum((p-1)^2 - (e-1)^2) = #(e == 1) - #(p == 1), dpair = sum((p - e) * nbr)
        const int nacc = warp_sum((int)acc.nacc);
        int dspin = 0, dspin2 = 0, dpair = 0;
        if (TRACK) {
            const int se = warp_sum(acc.e), sp = warp_sum(acc.p), se1 = warp_sum(acc.e1), sp1 = warp_sum(acc.p1);
            const int sen = warp_sum(acc.en), spn = warp_sum(acc.pn);
            dspin = sp - se;
            dspin2 = se1 - sp1;
            dpair = (spn - sen) - 4 * (sp - se);
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
void launch_bc(mcx_lattice *lat, uint64_t t)
{
    // a chain sub-range (launch_sweeps_ising2d_grouped) is the same launch on shifted base pointers
    LatView L = lat->view;
    const int c0 = g_launch_range.chain0, nch = g_launch_range.nchains < 0 ? lat->nchains : g_launch_range.nchains;
    L.planes += (int64_t)c0 * 2 * L.plane_stride;
    L.nchains = nch;
    cudaStream_t stream = g_launch_range.use_stream ? g_launch_range.stream : lat->ctx->stream;
    const int nseg = L.half >> 4;
    const int64_t ctas = (int64_t)lat->ctx->sm_count * 5;
    // strip height as in k_ising2d: 16 rows unless the batch would leave fewer than ~4 items per resident CTA
    int R = 16;
    for (; R > 4; R >>= 1) {
        if (L.Ly % R != 0) continue;
        const int64_t items = ((int64_t)(L.Ly / R) * nseg + kBcThreads - 1) / kBcThreads * lat->nchains;
        if (items >= 4 * ctas) break;
    }
    while (R > 2 && L.Ly % R != 0) R -= 2;
    const int nstrips = L.Ly / R;
    const int64_t G = (int64_t)nstrips * nseg;
    const int blocks_per_chain = (int)((G + kBcThreads - 1) / kBcThreads);
    const int nitems = (int)((int64_t)blocks_per_chain * nch);
    auto kern = k_bc2d<COLOUR, TRACK, HB>;
    static thread_local int resident = 0;
    if (!resident) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kBcThreads, 0);
        if (resident < 1) resident = 1;
    }
    int grid = lat->ctx->sm_count * resident;
    if (grid > nitems) grid = nitems;
    kern<<<grid, kBcThreads, 0, stream>>>(L, lat->d_thi, lat->d_tlo, lat->d_labels + c0, lat->d_sums + (int64_t)c0 * SUM_FIELDS,
                                         (uint32_t)lat->seed, (uint32_t)(lat->seed >> 32), t, lat->first_chain + (uint32_t)c0, R, nstrips,
                                         blocks_per_chain, nitems);
    lat->ctx->launches++;
}

}  // namespace

// false: not applicable (shape, rule or table layout), nothing launched
bool launch_sweep_bc2d(mcx_lattice *lat, int colour, uint64_t t)
{
    if (!lat->fast2d || lat->model != MCX_BLUME_CAPEL || lat->storage != MCX_STORAGE_INT8 || lat->slab) return false;
    const bool hb = lat->rule == MCX_HEATBATH;
    if (lat->table_len != (hb ? kHbTable : kBcTable) || knobs().bc2d == 0) return false;
    if ((int64_t)(lat->view.Ly / 2) * (lat->view.half >> 4) < 96) return false;       // tiny lattices: rows-of-8 kernel
    const bool track = lat->track_sums;
    if (hb) {
        if (colour == 0) { if (track) launch_bc<0, true, true>(lat, t); else launch_bc<0, false, true>(lat, t); }
        else             { if (track) launch_bc<1, true, true>(lat, t); else launch_bc<1, false, true>(lat, t); }
    } else {
        if (colour == 0) { if (track) launch_bc<0, true, false>(lat, t); else launch_bc<0, false, false>(lat, t); }
        else             { if (track) launch_bc<1, true, false>(lat, t); else launch_bc<1, false, false>(lat, t); }
    }
    return true;
}

}  // namespace mcx
