// k_flat.cu -- flat-histogram chains: multicanonical and Wang-Landau sweeps.
//
// The acceptance of these ensembles depends on the chain's GLOBAL observable
// (spin_flip!(sys, alg::AbstractImportanceSampling), SpinSystems/src/ising.jl:25-33 passes
// E_new, E_old), so a chain is serial by construction and the parallelism is across chains only
// (finding 4).  One WARP owns one chain and visits every site once per sweep in
// checkerboard order (colour 0 slots ascending, then colour 1; FLAT stream): same-colour sites are not
// neighbours, so the proposals, neighbour sums and random draws of a batch of 128 sites are prepared
// by the 32 lanes in parallel and only the short acceptance recurrence runs serially on lane 0.  The
// kernel works on the lattice's colour planes directly.  Per attempt (restating importance_sampling.jl:69-85,
// ensembles/multicanonical.jl:25-30, algorithms/wang_landau.jl:29-37, binned_object.jl:22-24):
//     log_ratio = lw[bin(x_new)] - lw[bin(x_old)]
//     accepted  = log_ratio > 0 || rand < exp(log_ratio)
//     muca: histogram[bin(x_vis)] += 1        WL: lw[bin(x_vis)] -= logf
// Histograms are privatised per block in shared memory (integer counters) when they fit and
// merged into the global histogram with one atomic per non-empty bin at the end of the launch.
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "mcx_internal.h"

namespace mcx {

namespace {

constexpr int kWarps = 4;                 // chains per block (one warp each)
constexpr int kBatch = 128;               // same-colour slots prepared in parallel per step (4 per lane)
constexpr int kSmemBins = 8192;           // 32 KB of uint32 counters

struct FlatParams {
    LatView L;
    long long *sums;            // [nchains][SUM_FIELDS]
    double *logweight;          // muca: [nbins]; WL: [nchains][nbins]
    unsigned long long *hist;   // [nbins]
    int *error;
    int64_t start, step, nbins;
    double beta_pair, logf, J;
    uint32_t seed_lo, seed_hi, first_chain;
    uint64_t sweep0;
    int nsweeps, policy, step_shift;
    int wl_spec;                // Wang-Landau decisions at once: -1 adaptive (serial / 8 / 32), 0 serial only, 8, 32
    int win_half, win_len;      // shared-memory window of the log-weight table around the chain's bin (WIN kernels):
                                // win_half = the farthest a batch can move the bin, win_len = 4 * win_half + 1 entries
};

// rand < exp(log_ratio) for rand = m * 2^-32, m = hi << 16 | lo, decided exactly but cheaply:
//  1. a float estimate of p with a rigorous error margin settles all but ~2^-16 of the attempts from
//     the high half alone (no double-precision exp, no second Philox block);
//  2. otherwise the double-precision exp decides on the high half, and only if (hi, hi + 1) brackets
//     p * 2^16 is the low half fetched and the full 32-bit comparison made.
// The result equals (double)m * 2^-32 < exp(log_ratio) in every case.
template <class LoFn>
__device__ __forceinline__ bool draw_less_exp(float hif, double log_ratio, LoFn lo_fn)
{
    const float lf = (float)log_ratio;
    const float p16 = __expf(lf) * 65536.0f;
    const float eps = 3.0e-6f + fabsf(lf) * 1.0e-6f;       // > 4x the worst-case relative error of p16
    if (hif + 1.0f <= p16 * (1.0f - eps)) return true;
    if (hif > 0.0f && hif >= p16 * (1.0f + eps)) return false;   // hi == 0: a denormal p could still win
    const double p = exp(log_ratio), pd16 = p * 65536.0, hid = (double)hif;
    if (hid + 1.0 <= pd16) return true;
    if (hid >= pd16) return false;
    const uint32_t lo = lo_fn();
    return (hid * 65536.0 + (double)lo) * (1.0 / 4294967296.0) < p;
}

// div(x - start, step) truncating toward zero like Julia's div (binned_object.jl:22-24)
__device__ __forceinline__ int64_t bin_of(int64_t x, int64_t start, int64_t step, int shift)
{
    const int64_t d = x - start;
    if (shift >= 0) return (d + ((d >> 63) & (step - 1))) >> shift;
    return d / step;
}

// physical neighbour-spin sum of the site in slot q of colour `colour` (all neighbours sit in `oth`)
__device__ __forceinline__ int neighbour_sum(const LatView &L, const uint8_t *oth, uint32_t q, int colour)
{
    const uint32_t half = (uint32_t)L.half;
    const uint32_t row = q / half, j = q - row * half;
    const uint32_t y = row % (uint32_t)L.Ly, z = row / (uint32_t)L.Ly;
    const int x = (int)(2 * j + ((colour + y + z) & 1));
    const int64_t rb = (int64_t)row * half;
    const int xl = x == 0 ? L.Lx - 1 : x - 1, xr = x == L.Lx - 1 ? 0 : x + 1;
    int raw = oth[rb + (xl >> 1)] + oth[rb + (xr >> 1)];
    if (L.ndim > 1) {
        const uint32_t yu = y == 0 ? L.Ly - 1 : y - 1, yd = y == (uint32_t)L.Ly - 1 ? 0 : y + 1;
        raw += oth[((int64_t)z * L.Ly + yu) * half + j] + oth[((int64_t)z * L.Ly + yd) * half + j];
    }
    if (L.ndim > 2) {
        const uint32_t zu = z == 0 ? L.Lz - 1 : z - 1, zd = z == (uint32_t)L.Lz - 1 ? 0 : z + 1;
        raw += oth[((int64_t)zu * L.Ly + y) * half + j] + oth[((int64_t)zd * L.Ly + y) * half + j];
    }
    return L.model == MCX_ISING ? 2 * raw - L.nn : raw - L.nn;
}

// One warp per chain.  Per batch of 128 same-colour slots:
//   phase 1 (32 lanes): Philox blocks -> 16-bit draws; per site the proposal and its deltas
//                       (no two sites of one colour are neighbours, so these do not depend on the
//                        decisions inside the batch);
//   phase 2 (lane 0):   the serial recurrence in the global observable: bin lookup, log-ratio,
//                       _accept!, record_visit! / Wang-Landau update, running sums;
//   phase 3 (32 lanes): accepted sites are written back.
//
// WIN: the serial recurrence reads (Wang-Landau: and writes) log-weights of bins next to the chain's current one -- an
// attempt moves the bin by at most a few -- so each warp keeps a window of the table, 4 x the farthest one batch can
// travel, in shared memory: the dependent chain of an attempt then waits for a shared-memory load (~30 cycles) instead
// of an L2 / HBM round trip per decision.  The window is recentred (dirty entries written back first) only when the next
// batch could leave it; values are the table's own doubles, so trajectories are unchanged.
template <int OBS, int KIND, bool SMEM_HIST, bool WIN>
__global__ void __launch_bounds__(kWarps * 32) k_flat_warp(FlatParams P)
{
    extern __shared__ double s_dyn[];
    double *s_win = s_dyn;                                                              // [kWarps][win_len]
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_dyn + (WIN ? kWarps * P.win_len : 0));
    __shared__ int32_t s_d[kWarps][kBatch];      // packed deltas of each prepared site
    __shared__ float s_hi[kWarps][kBatch];       // high half of the site's Float64 draw, as a float (exact)
    __shared__ uint8_t s_b[kWarps][kBatch];      // Bool draw (Blume-Capel proposal)
    __shared__ uint8_t s_acc[kWarps][kBatch];    // accepted flag of each site of the batch
    if (SMEM_HIST) {
        for (int i = threadIdx.x; i < P.nbins; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kWarps + w;
    if (c < P.L.nchains) {
        const LatView &L = P.L;
        long long *st = P.sums + (int64_t)c * SUM_FIELDS;
        // chain state lives in lane 0's registers
        long long pair = st[SUM_PAIR], spin = st[SUM_SPIN], spin2 = L.model == MCX_ISING ? 2 * L.halfN : st[SUM_SPIN2], nacc = 0;
        double *lw = P.logweight + (KIND == MCX_FLAT_WANG_LANDAU ? (int64_t)c * P.nbins : 0);
        const uint32_t chain_id = P.first_chain + (uint32_t)c;
        const int shift = P.step_shift;
        int64_t io = bin_of(OBS == MCX_OBS_ENERGY ? -pair : spin2, P.start, P.step, shift);
        int dead = io < 0 || io >= P.nbins;
        if (dead && lane == 0) atomicExch(P.error, 1);
        double lw_old = dead ? 0.0 : lw[io];
        double *win = s_win + (WIN ? w * P.win_len : 0);
        int64_t wbase = 0;                       // bin of win[0]
        bool win_valid = false;
        int dlo = 0x7fffffff, dhi = -1;          // window entries lane 0 has written since the last load (Wang-Landau)
        auto win_flush = [&]() {
            if (WIN && KIND == MCX_FLAT_WANG_LANDAU && win_valid) {
                __syncwarp();
                const int lo = __shfl_sync(0xffffffffu, dlo, 0), hi = __shfl_sync(0xffffffffu, dhi, 0);
                for (int k = lo + lane; k <= hi; k += 32) lw[wbase + k] = win[k];
                dlo = 0x7fffffff; dhi = -1;
                __syncwarp();
            }
        };
        auto LWR = [&](int64_t b) -> double { return WIN ? win[b - wbase] : lw[b]; };
        auto LWW = [&](int64_t b, double v) {
            if (WIN) {
                const int k = (int)(b - wbase);
                win[k] = v; dlo = min(dlo, k); dhi = max(dhi, k);
            } else lw[b] = v;
        };
        // Boltzmann part of the two-component observable: H1 = J * sum_pair_interactions (a Float64 in the
        // reference, blume_capel.jl:123) carried as a running double; exact for integer-valued J
        double Ho1 = P.J * (P.J * (double)pair);
        int64_t run_bin = io;
        unsigned long long run_cnt = 0;
        bool speculate = true;      // multicanonical: decide 32 attempts at once (phase 2); adapted per batch
        // Wang-Landau: decide `wl_width` consecutive attempts at once (0: the serial loop on lane 0); adapted per batch
        int wl_width = KIND == MCX_FLAT_WANG_LANDAU ? (P.wl_spec < 0 ? 8 : P.wl_spec) : 0;
        const uint32_t halfN = (uint32_t)L.halfN;
        constexpr uint32_t PF = OBS == MCX_OBS_ENERGY ? 0 : 2;     // plane of the Float64 draw's high half

        for (int sw = 0; sw < P.nsweeps && !dead; ++sw)
            for (int colour = 0; colour < 2 && !dead; ++colour) {
                const uint64_t t = 2 * (P.sweep0 + (uint64_t)sw) + (uint64_t)colour;
                uint8_t *tgt = plane_ptr(L, c, colour);
                const uint8_t *oth = plane_ptr(L, c, colour ^ 1);
                for (uint32_t qb = 0; qb < halfN && !dead; qb += kBatch) {
                    const int cnt = (int)min((uint32_t)kBatch, halfN - qb);
                    // ---- phase 1a: draws.  lanes 0-15: the Float64 plane; lanes 16-31: the Bool plane
                    {
                        const uint32_t plane = lane < 16 ? PF : 0;
                        const int bl = lane & 15;
                        if (lane < 16 || OBS != MCX_OBS_ENERGY) {
                            const Philox4 r = stream_block(P.seed_lo, P.seed_hi, chain_id, TAG_FLAT, t, (qb >> 3) + bl, plane);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t v = lane16(r, k);
                                if (lane < 16) s_hi[w][8 * bl + k] = (float)v;
                                else s_b[w][8 * bl + k] = (uint8_t)(v >> 15);
                            }
                        }
                    }
                    __syncwarp();
                    // ---- phase 1b: proposal and deltas of every site of the batch
#pragma unroll
                    for (int j = 0; j < kBatch / 32; ++j) {
                        const int idx = 32 * j + lane;
                        if (idx < cnt) {
                            const uint32_t q = qb + idx;
                            const int nb = neighbour_sum(L, oth, q, colour);
                            const int enc = tgt[q];
                            int32_t d;
                            if (OBS == MCX_OBS_ENERGY) {
                                // flip_changes / delta_energy (ising.jl:187-198), integer path J = 1, h = 0
                                const int s = 2 * enc - 1;
                                const int dE = 2 * s * nb;                       // = -dpair
                                d = (dE & 0xff) | ((s & 0xff) << 8);
                            } else {
                                // _propose_state + propose_changes (blume_capel.jl:21-30,235-241)
                                const int s = enc - 1, b = s_b[w][idx];
                                const int s_new = s == -1 ? (b ? 0 : 1) : s == 0 ? (b ? -1 : 1) : (b ? -1 : 0);
                                const int dspin = s_new - s, dspin2 = s_new * s_new - s * s, dpair = dspin * nb;
                                d = (dspin & 0xff) | ((dspin2 & 0xff) << 8) | ((dpair & 0xff) << 16) | ((s_new + 1) << 24);
                            }
                            s_d[w][idx] = d;
                        }
                    }
                    __syncwarp();
                    // ---- phase 2: the chain's serial recurrence.
                    // Multicanonical weights do not change during a sweep, so a rejected attempt leaves the chain
                    // state untouched: the 32 lanes decide 32 consecutive attempts against the CURRENT state at
                    // once, the first accepted one (ballot) is applied, and only the attempts after it are
                    // decided again.  Every decision is taken with the exact state the serial loop would have, so
                    // the trajectory is unchanged; the cost drops from one dependent evaluation per attempt to
                    // one per ACCEPTED attempt (+ 1 per 32).  Wang-Landau changes lw at every attempt: serial loop.
                    const long long nacc_before = nacc;
                    if (WIN && (!win_valid || io - P.win_half < wbase || io + P.win_half >= wbase + P.win_len)) {
                        win_flush();
                        wbase = io - P.win_len / 2;
                        for (int k = lane; k < P.win_len; k += 32) {
                            const int64_t b = wbase + k;
                            win[k] = b >= 0 && b < P.nbins ? __ldcg(lw + b) : 0.0;
                        }
                        win_valid = true;
                    }
                    // a decision loop that stops early (BoundsError, policy 0) must not leave the previous batch's
                    // flags behind for phase 3: attempts that are never decided are not accepted
#pragma unroll
                    for (int j = 0; j < kBatch / 32; ++j) s_acc[w][32 * j + lane] = 0;
                    __syncwarp();
                    if (KIND == MCX_FLAT_MUCA && speculate) {
                        for (int base = 0; base < cnt && !dead; base += 32) {
                            const int idx = base + lane;
                            const bool valid = idx < cnt;
                            const int32_t d = valid ? s_d[w][idx] : 0;
                            const float hif = valid ? s_hi[w][idx] : 0.0f;
                            int dpair, dspin, dspin2;
                            if (OBS == MCX_OBS_ENERGY) {
                                const int dE = (int8_t)(d & 0xff), s = (int8_t)((d >> 8) & 0xff);
                                dpair = -dE; dspin = -2 * s; dspin2 = 0;
                            } else {
                                dspin = (int8_t)(d & 0xff); dspin2 = (int8_t)((d >> 8) & 0xff); dpair = (int8_t)((d >> 16) & 0xff);
                            }
                            bool mine_accepted = false;
                            int start = 0;                                   // attempts [0, start) of this group are decided
                            const int gend = min(32, cnt - base);
                            while (start < gend) {
                                // decision of my attempt against the current state (lanes before `start` idle)
                                const bool live = valid && lane >= start;
                                const int64_t x_new = OBS == MCX_OBS_ENERGY ? -pair - dpair : spin2 + dspin2;
                                const int64_t in = bin_of(x_new, P.start, P.step, shift);
                                const bool inside = in >= 0 && in < P.nbins;
                                bool acc_i = false;
                                double lw_new = lw_old, Hn1 = Ho1;
                                if (live && inside) {
                                    lw_new = in == io ? lw_old : LWR(in);
                                    double log_ratio;
                                    if (OBS == MCX_OBS_ENERGY) {
                                        log_ratio = lw_new - lw_old;
                                    } else {
                                        Hn1 = Ho1 + P.J * (P.J * (double)dpair);
                                        log_ratio = (-P.beta_pair * Hn1 + lw_new) - (-P.beta_pair * Ho1 + lw_old);
                                    }
                                    if (log_ratio > 0) acc_i = true;
                                    else acc_i = draw_less_exp(hif, log_ratio, [&]() {
                                        const uint32_t q = qb + idx;
                                        const Philox4 rl = stream_block(P.seed_lo, P.seed_hi, chain_id, TAG_FLAT, t, q >> 3, PF + 1);
                                        return lane16(rl, (int)(q & 7));
                                    });
                                }
                                const bool oob_i = live && !inside && P.policy == 0;        // BoundsError when reached
                                const unsigned ev = __ballot_sync(0xffffffffu, acc_i || oob_i);
                                const int k = ev ? __ffs(ev) - 1 : gend;                    // first attempt that changes anything
                                // attempts [start, k) are rejected: they visit the current bin
                                const int nrej = k - start;
                                if (nrej > 0 && lane == 0) {
                                    if (SMEM_HIST) atomicAdd(&s_hist[io], (uint32_t)nrej);
                                    else if (io == run_bin) run_cnt += nrej;
                                    else {
                                        if (run_cnt) atomicAdd(P.hist + run_bin, run_cnt);
                                        run_bin = io; run_cnt = nrej;
                                    }
                                }
                                if (k >= gend) break;
                                if (__shfl_sync(0xffffffffu, (int)oob_i, k)) {
                                    if (lane == 0) atomicExch(P.error, 1);
                                    dead = 1;
                                    break;
                                }
                                // attempt k is accepted: its values become the chain state on every lane
                                pair += __shfl_sync(0xffffffffu, dpair, k);
                                spin += __shfl_sync(0xffffffffu, dspin, k);
                                spin2 += __shfl_sync(0xffffffffu, dspin2, k);
                                nacc += 1;
                                io = __shfl_sync(0xffffffffu, in, k);
                                lw_old = __shfl_sync(0xffffffffu, lw_new, k);
                                Ho1 = __shfl_sync(0xffffffffu, Hn1, k);
                                if (lane == k) mine_accepted = true;
                                if (lane == 0) {                                            // record_visit! at the new bin
                                    if (SMEM_HIST) atomicAdd(&s_hist[io], 1u);
                                    else if (io == run_bin) run_cnt += 1;
                                    else {
                                        if (run_cnt) atomicAdd(P.hist + run_bin, run_cnt);
                                        run_bin = io; run_cnt = 1;
                                    }
                                }
                                start = k + 1;
                            }
                            if (valid) s_acc[w][idx] = (uint8_t)mine_accepted;
                        }
                    } else if (KIND == MCX_FLAT_WANG_LANDAU && wl_width > 0) {
                        // Wang-Landau changes lw at every attempt, but a REJECTED attempt changes it predictably: it
                        // visits the current bin, lw[io] -= logf (wang_landau.jl:33-35), and leaves everything else
                        // alone.  So the lanes decide G consecutive attempts at once, lane j assuming that the j
                        // attempts before it were rejected: it replays their j subtractions (one rounding each, as
                        // the serial loop does) to get the current bin's weight it would see.  The first accepted
                        // attempt (ballot) is applied and the attempts after it are decided again.  Every decision
                        // uses exactly the values the serial loop would have; the table of a chain is private, and only
                        // bin io changes between two accepted attempts, so the other lanes' loads lw[in] are current.
                        const int G = wl_width;
                        for (int base = 0; base < cnt && !dead; base += G) {
                            const int idx = base + lane;
                            const bool valid = lane < G && idx < cnt;
                            const int32_t d = valid ? s_d[w][idx] : 0;
                            const float hif = valid ? s_hi[w][idx] : 0.0f;
                            int dpair, dspin, dspin2;
                            if (OBS == MCX_OBS_ENERGY) {
                                const int dE = (int8_t)(d & 0xff), s = (int8_t)((d >> 8) & 0xff);
                                dpair = -dE; dspin = -2 * s; dspin2 = 0;
                            } else {
                                dspin = (int8_t)(d & 0xff); dspin2 = (int8_t)((d >> 8) & 0xff); dpair = (int8_t)((d >> 16) & 0xff);
                            }
                            bool mine_accepted = false;
                            int start = 0;                                   // attempts [0, start) of this group are decided
                            const int gend = min(G, cnt - base);
                            while (start < gend) {
                                const bool live = valid && lane >= start;
                                double lw_cur = lw_old;                      // lw[io] after the rejections before my attempt
                                for (int i = start; i < lane && live; ++i) lw_cur -= P.logf;
                                const int64_t x_new = OBS == MCX_OBS_ENERGY ? -pair - dpair : spin2 + dspin2;
                                const int64_t in = bin_of(x_new, P.start, P.step, shift);
                                const bool inside = in >= 0 && in < P.nbins;
                                bool acc_i = false;
                                double lw_new = lw_cur, Hn1 = Ho1;
                                if (live && inside) {
                                    lw_new = in == io ? lw_cur : LWR(in);
                                    double log_ratio;
                                    if (OBS == MCX_OBS_ENERGY) {
                                        log_ratio = lw_new - lw_cur;
                                    } else {
                                        Hn1 = Ho1 + P.J * (P.J * (double)dpair);
                                        log_ratio = (-P.beta_pair * Hn1 + lw_new) - (-P.beta_pair * Ho1 + lw_cur);
                                    }
                                    if (log_ratio > 0) acc_i = true;
                                    else acc_i = draw_less_exp(hif, log_ratio, [&]() {
                                        const uint32_t q = qb + idx;
                                        const Philox4 rl = stream_block(P.seed_lo, P.seed_hi, chain_id, TAG_FLAT, t, q >> 3, PF + 1);
                                        return lane16(rl, (int)(q & 7));
                                    });
                                }
                                const bool oob_i = live && !inside && P.policy == 0;        // BoundsError when reached
                                const unsigned ev = __ballot_sync(0xffffffffu, acc_i || oob_i);
                                const int k = ev ? __ffs(ev) - 1 : gend;                    // first attempt that is not a rejection
                                // attempts [start, k) are rejected: lw[io] after their visits is what lane k started
                                // from, or one more subtraction after the last lane if nothing else happened
                                const int nrej = k - start;
                                const double lw_last = __shfl_sync(0xffffffffu, lw_cur, min(k, gend - 1));
                                const double lw_after = k < gend ? lw_last : lw_last - P.logf;
                                if (k >= gend) {
                                    lw_old = lw_after;
                                    if (lane == 0) LWW(io, lw_old);
                                    break;
                                }
                                if (__shfl_sync(0xffffffffu, (int)oob_i, k)) {
                                    lw_old = lw_after;
                                    if (lane == 0) {
                                        if (nrej > 0) LWW(io, lw_old);
                                        atomicExch(P.error, 1);
                                    }
                                    dead = 1;
                                    break;
                                }
                                // attempt k is accepted: the chain moves to bin in_k and visits it
                                const int64_t in_k = __shfl_sync(0xffffffffu, in, k);
                                const double lw_new_k = __shfl_sync(0xffffffffu, lw_new, k);
                                pair += __shfl_sync(0xffffffffu, dpair, k);
                                spin += __shfl_sync(0xffffffffu, dspin, k);
                                spin2 += __shfl_sync(0xffffffffu, dspin2, k);
                                nacc += 1;
                                Ho1 = __shfl_sync(0xffffffffu, Hn1, k);
                                if (lane == 0 && nrej > 0) LWW(io, lw_after);
                                io = in_k;
                                lw_old = lw_new_k - P.logf;                 // lw_old = lw_new; lw_old -= logf
                                if (lane == 0) LWW(io, lw_old);
                                if (lane == k) mine_accepted = true;
                                __syncwarp();                               // lane 0's stores before the next round's loads
                                start = k + 1;
                            }
                            if (valid) s_acc[w][idx] = (uint8_t)mine_accepted;
                        }
                    } else
                    if (lane == 0) {
                        for (int idx = 0; idx < cnt; ++idx) {
                            const int32_t d = s_d[w][idx];
                            int dpair, dspin, dspin2;
                            int64_t x_new;
                            if (OBS == MCX_OBS_ENERGY) {
                                const int dE = (int8_t)(d & 0xff), s = (int8_t)((d >> 8) & 0xff);
                                dpair = -dE; dspin = -2 * s; dspin2 = 0;
                                x_new = -pair + dE;
                            } else {
                                dspin = (int8_t)(d & 0xff); dspin2 = (int8_t)((d >> 8) & 0xff); dpair = (int8_t)((d >> 16) & 0xff);
                                x_new = spin2 + dspin2;          // H = (J*sum_pair, sum_spins2), muca_BlumeCapel.jl:81-89
                            }
                            // _binindex for integer bins: div(x - start, step) + 1 (binned_object.jl:22-24); 0-based here
                            const int64_t in = bin_of(x_new, P.start, P.step, shift);
                            const bool inside = in >= 0 && in < P.nbins;
                            if (!inside && P.policy == 0) {                          // BoundsError
                                atomicExch(P.error, 1);
                                dead = 1;
                                break;
                            }
                            bool accepted = false;
                            double lw_new = lw_old, Hn1 = Ho1;
                            if (inside) {
                                lw_new = in == io ? lw_old : LWR(in);
                                double log_ratio;
                                if (OBS == MCX_OBS_ENERGY) {
                                    log_ratio = lw_new - lw_old;
                                } else {
                                    Hn1 = Ho1 + P.J * (P.J * (double)dpair);
                                    log_ratio = (-P.beta_pair * Hn1 + lw_new) - (-P.beta_pair * Ho1 + lw_old);
                                }
                                // _accept! (importance_sampling.jl:80-85)
                                if (log_ratio > 0) accepted = true;
                                else accepted = draw_less_exp(s_hi[w][idx], log_ratio, [&]() {
                                    const uint32_t q = qb + idx;
                                    const Philox4 rl = stream_block(P.seed_lo, P.seed_hi, chain_id, TAG_FLAT, t, q >> 3, PF + 1);
                                    return lane16(rl, (int)(q & 7));
                                });
                            }
                            if (accepted) {
                                pair += dpair; spin += dspin; spin2 += dspin2; nacc += 1;
                                io = in; lw_old = lw_new; Ho1 = Hn1;
                            }
                            s_acc[w][idx] = (uint8_t)accepted;
                            // record_visit! / Wang-Landau update at the visited bin (= io after the move)
                            if (KIND == MCX_FLAT_MUCA) {
                                if (SMEM_HIST) atomicAdd(&s_hist[io], 1u);
                                else if (io == run_bin) run_cnt += 1;
                                else {
                                    if (run_cnt) atomicAdd(P.hist + run_bin, run_cnt);
                                    run_bin = io; run_cnt = 1;
                                }
                            } else {
                                lw_old -= P.logf;
                                LWW(io, lw_old);
                            }
                        }
                    }
                    dead = __shfl_sync(0xffffffffu, dead, 0);
                    if (KIND == MCX_FLAT_WANG_LANDAU) {
                        if (wl_width == 0) {     // the serial loop ran on lane 0: every lane needs the chain state
                            pair = __shfl_sync(0xffffffffu, pair, 0); spin = __shfl_sync(0xffffffffu, spin, 0);
                            spin2 = __shfl_sync(0xffffffffu, spin2, 0); nacc = __shfl_sync(0xffffffffu, nacc, 0);
                            io = __shfl_sync(0xffffffffu, io, 0); lw_old = __shfl_sync(0xffffffffu, lw_old, 0);
                            Ho1 = __shfl_sync(0xffffffffu, Ho1, 0);
                        }
                        if (P.wl_spec < 0) {
                            // a round of group decisions costs about 2.3 (8 lanes) or 3.3 (32 lanes) serial attempts
                            // (profiles/r01_wl_group_decisions.md): acceptance of this batch > 40 % -> serial loop,
                            // > 15 % -> 8 at once, else 32
                            const long long a = nacc - nacc_before;
                            wl_width = a * 5 > (long long)cnt * 2 ? 0 : a * 20 > (long long)cnt * 3 ? 8 : 32;
                        }
                    }
                    if (KIND == MCX_FLAT_MUCA) {
                        if (!speculate) {        // the serial loop ran on lane 0: every lane needs the chain state
                            pair = __shfl_sync(0xffffffffu, pair, 0); spin = __shfl_sync(0xffffffffu, spin, 0);
                            spin2 = __shfl_sync(0xffffffffu, spin2, 0); nacc = __shfl_sync(0xffffffffu, nacc, 0);
                            io = __shfl_sync(0xffffffffu, io, 0); lw_old = __shfl_sync(0xffffffffu, lw_old, 0);
                            Ho1 = __shfl_sync(0xffffffffu, Ho1, 0);
                        }
                        // deciding 32 attempts at once pays while fewer than ~45 % of them are accepted
                        speculate = (nacc - nacc_before) * 20 < (long long)cnt * 9;
                    }
                    __syncwarp();
                    // ---- phase 3: write the accepted sites back
#pragma unroll
                    for (int j = 0; j < kBatch / 32; ++j) {
                        const int idx = 32 * j + lane;
                        if (idx < cnt && s_acc[w][idx]) {
                            const uint32_t q = qb + idx;
                            if (OBS == MCX_OBS_ENERGY) tgt[q] ^= 1;
                            else tgt[q] = (uint8_t)(s_d[w][idx] >> 24);
                        }
                    }
                    __syncwarp();
                }
            }
        win_flush();
        if (lane == 0) {
            if (KIND == MCX_FLAT_MUCA && !SMEM_HIST && run_cnt) atomicAdd(P.hist + run_bin, run_cnt);
            st[SUM_PAIR] = pair; st[SUM_SPIN] = spin;
            if (L.model != MCX_ISING) st[SUM_SPIN2] = spin2;
            st[SUM_ACC] += nacc;
        }
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < P.nbins; i += blockDim.x)
            if (s_hist[i]) atomicAdd(P.hist + i, (unsigned long long)s_hist[i]);
    }
}

// update!(ens::MulticanonicalEnsemble; mode=:simple): lw -= (h > 0 ? log(h) : 0) (multicanonical.jl:32-44)
__global__ void k_muca_update(double *__restrict__ lw, const unsigned long long *__restrict__ hist, int64_t nbins)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbins) return;
    const unsigned long long hh = hist[i];
    if (hh > 0) lw[i] -= log((double)hh);
}

template <int OBS, int KIND, bool SMEM_HIST, bool WIN>
void launch_sweep_k(mcx_flat *f, const FlatParams &P, size_t smem)
{
    mcx_lattice *lat = f->lat;
    const int blocks = (lat->nchains + kWarps - 1) / kWarps;
    auto kern = k_flat_warp<OBS, KIND, SMEM_HIST, WIN>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, kWarps * 32, smem, lat->ctx->stream>>>(P);
    lat->ctx->launches++;
}

template <int OBS, int KIND>
void launch_sweep_t(mcx_flat *f, FlatParams &P)
{
    const bool smem_hist = KIND == MCX_FLAT_MUCA && f->nbins <= kSmemBins;
    // window of the log-weight table per warp: an attempt moves the bin by at most `maxstep` bins
    // (|dE| <= 2 nn for the Ising energy, |d sum s^2| <= 1), a batch of kBatch attempts by kBatch * maxstep
    const int64_t max_dx = OBS == MCX_OBS_ENERGY ? 2 * f->lat->nn : 1;
    const int64_t maxstep = (max_dx + f->step - 1) / f->step;
    P.win_half = (int)(kBatch * (maxstep < 1 ? 1 : maxstep));
    P.win_len = 4 * P.win_half + 1;
    const size_t hist_bytes = smem_hist ? (size_t)f->nbins * sizeof(uint32_t) : 0;
    const size_t win_bytes = (size_t)kWarps * P.win_len * sizeof(double);
    const bool win = knobs().flat_window != 0 && win_bytes + hist_bytes <= 96 * 1024;
    if (!win) P.win_half = P.win_len = 0;
    if (smem_hist) {
        if (win) launch_sweep_k<OBS, KIND, true, true>(f, P, win_bytes + hist_bytes);
        else launch_sweep_k<OBS, KIND, true, false>(f, P, hist_bytes);
    } else {
        if (win) launch_sweep_k<OBS, KIND, false, true>(f, P, win_bytes);
        else launch_sweep_k<OBS, KIND, false, false>(f, P, 0);
    }
}

}  // namespace

void launch_flat_load(mcx_flat *) {}     // the warp-per-chain kernel works on the colour planes directly
void launch_flat_store(mcx_flat *) {}

void launch_flat_sweep(mcx_flat *f, uint64_t sweep0, int nsweeps)
{
    mcx_lattice *lat = f->lat;
    FlatParams P;
    P.L = lat->view; P.sums = lat->d_sums; P.logweight = f->d_logweight;
    P.hist = (unsigned long long *)f->d_histogram; P.error = f->d_error;
    P.start = f->start; P.step = f->step; P.nbins = f->nbins;
    P.beta_pair = f->beta_pair; P.logf = f->logf; P.J = lat->J;
    P.seed_lo = (uint32_t)lat->seed; P.seed_hi = (uint32_t)(lat->seed >> 32); P.first_chain = lat->first_chain;
    P.sweep0 = sweep0; P.nsweeps = nsweeps; P.policy = f->policy;
    P.wl_spec = knobs().wl_spec == 0 || knobs().wl_spec == 8 || knobs().wl_spec == 32 ? knobs().wl_spec : -1;
    P.step_shift = -1;
    for (int sh = 0; sh < 62; ++sh)
        if (f->step == ((int64_t)1 << sh)) P.step_shift = sh;
    if (f->observable == MCX_OBS_ENERGY) {
        if (f->kind == MCX_FLAT_MUCA) launch_sweep_t<MCX_OBS_ENERGY, MCX_FLAT_MUCA>(f, P);
        else launch_sweep_t<MCX_OBS_ENERGY, MCX_FLAT_WANG_LANDAU>(f, P);
    } else {
        if (f->kind == MCX_FLAT_MUCA) launch_sweep_t<MCX_OBS_SPIN2_WITH_PAIR_BOLTZMANN, MCX_FLAT_MUCA>(f, P);
        else launch_sweep_t<MCX_OBS_SPIN2_WITH_PAIR_BOLTZMANN, MCX_FLAT_WANG_LANDAU>(f, P);
    }
}

void launch_flat_update(mcx_flat *f)
{
    mcx_lattice *lat = f->lat;
    k_muca_update<<<(unsigned)((f->nbins + 255) / 256), 256, 0, lat->ctx->stream>>>(
        f->d_logweight, (const unsigned long long *)f->d_histogram, f->nbins);
    lat->ctx->launches++;
}

}  // namespace mcx

// ------------------------------------------------------------------------------------ C ABI (flat)
using namespace mcx;

int32_t mcx_set_error(int32_t code, const char *msg);   // mcx_api.cu

#define FREQ(cond, code, msg) do { if (!(cond)) return mcx_set_error(code, msg); } while (0)
#define FCUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return mcx_set_error(MCX_ERR_CUDA, cudaGetErrorString(e__)); } while (0)

extern "C" {

int32_t mcx_flat_create(mcx_lattice *lat, int32_t kind, int32_t observable, int64_t bin_start, int64_t bin_step,
                        int64_t nbins, double beta_pair, int32_t out_of_range_policy, mcx_flat **out)
{
    FREQ(lat && out, MCX_ERR_ARGUMENT, "NULL argument");
    FREQ(kind == MCX_FLAT_MUCA || kind == MCX_FLAT_WANG_LANDAU, MCX_ERR_ARGUMENT, "unknown flat-histogram kind");
    FREQ(observable == MCX_OBS_ENERGY || observable == MCX_OBS_SPIN2_WITH_PAIR_BOLTZMANN, MCX_ERR_ARGUMENT, "unknown observable");
    FREQ(bin_step > 0 && nbins >= 1, MCX_ERR_ARGUMENT, "bins need step > 0 and at least one bin");
    FREQ(out_of_range_policy == 0 || out_of_range_policy == 1, MCX_ERR_ARGUMENT, "policy must be 0 (BoundsError) or 1 (reject)");
    FREQ(lat->view.halfN < ((int64_t)1 << 31), MCX_ERR_UNSUPPORTED, "flat-histogram chains support up to 2^32 sites per lattice");
    FREQ(lat->storage == MCX_STORAGE_INT8, MCX_ERR_UNSUPPORTED, "flat-histogram chains visit single sites: they need int8 planes");
    if (observable == MCX_OBS_ENERGY)
        FREQ(lat->model == MCX_ISING && lat->J == 1.0 && lat->h == 0.0, MCX_ERR_UNSUPPORTED,
             "the energy observable is the integer path: Ising with J = 1, h = 0");
    else
        FREQ(lat->model == MCX_BLUME_CAPEL, MCX_ERR_UNSUPPORTED, "the (pair, spin^2) observable needs a Blume-Capel lattice");
    FCUDA(cudaSetDevice(lat->ctx->device));
    mcx_flat *f = new (std::nothrow) mcx_flat();
    FREQ(f, MCX_ERR_STATE, "out of host memory");
    memset(f, 0, sizeof(*f));
    f->lat = lat; f->kind = kind; f->observable = observable; f->start = bin_start; f->step = bin_step; f->nbins = nbins;
    f->beta_pair = beta_pair; f->logf = 1.0; f->policy = out_of_range_policy;
    f->ntables = kind == MCX_FLAT_WANG_LANDAU ? lat->nchains : 1;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&f->d_logweight, sizeof(double) * (size_t)nbins * f->ntables)) != cudaSuccess ||
        (e = cudaMalloc((void **)&f->d_histogram, sizeof(unsigned long long) * (size_t)nbins)) != cudaSuccess ||
        (e = cudaMalloc((void **)&f->d_error, sizeof(int))) != cudaSuccess) {
        mcx_flat_destroy(f);
        return mcx_set_error(MCX_ERR_CUDA, cudaGetErrorString(e));
    }
    FCUDA(cudaMemset(f->d_logweight, 0, sizeof(double) * (size_t)nbins * f->ntables));
    FCUDA(cudaMemset(f->d_histogram, 0, sizeof(unsigned long long) * (size_t)nbins));
    FCUDA(cudaMemset(f->d_error, 0, sizeof(int)));
    *out = f;
    return MCX_OK;
}

int32_t mcx_flat_destroy(mcx_flat *f)
{
    if (!f) return MCX_OK;
    cudaSetDevice(f->lat->ctx->device);
    cudaStreamSynchronize(f->lat->ctx->stream);
    cudaFree(f->d_logweight); cudaFree(f->d_histogram); cudaFree(f->d_error);
    delete f;
    return MCX_OK;
}

static int32_t flat_check_error(mcx_flat *f)
{
    int err = 0;
    FCUDA(cudaMemcpyAsync(&err, f->d_error, sizeof(int), cudaMemcpyDeviceToHost, f->lat->ctx->stream));
    FCUDA(cudaStreamSynchronize(f->lat->ctx->stream));
    if (err) {
        cudaMemset(f->d_error, 0, sizeof(int));
        return mcx_set_error(MCX_ERR_BOUNDS, "BoundsError: a chain left the binned range of the log-weight table");
    }
    return MCX_OK;
}

int32_t mcx_flat_set_logweight(mcx_flat *f, const double *logweight)
{
    FREQ(f && logweight, MCX_ERR_ARGUMENT, "NULL argument");
    FCUDA(cudaSetDevice(f->lat->ctx->device));
    FCUDA(cudaMemcpyAsync(f->d_logweight, logweight, sizeof(double) * (size_t)f->nbins * f->ntables, cudaMemcpyHostToDevice,
                          f->lat->ctx->stream));
    FCUDA(cudaStreamSynchronize(f->lat->ctx->stream));
    return MCX_OK;
}

int32_t mcx_flat_get_logweight(mcx_flat *f, double *logweight)
{
    FREQ(f && logweight, MCX_ERR_ARGUMENT, "NULL argument");
    FCUDA(cudaSetDevice(f->lat->ctx->device));
    FCUDA(cudaMemcpyAsync(logweight, f->d_logweight, sizeof(double) * (size_t)f->nbins * f->ntables, cudaMemcpyDeviceToHost,
                          f->lat->ctx->stream));
    return flat_check_error(f);
}

int32_t mcx_flat_get_histogram(mcx_flat *f, double *histogram)
{
    FREQ(f && histogram, MCX_ERR_ARGUMENT, "NULL argument");
    FCUDA(cudaSetDevice(f->lat->ctx->device));
    std::vector<unsigned long long> h((size_t)f->nbins);
    FCUDA(cudaMemcpyAsync(h.data(), f->d_histogram, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost,
                          f->lat->ctx->stream));
    int32_t st = flat_check_error(f);
    for (size_t i = 0; i < h.size(); ++i) histogram[i] = (double)h[i];   // reference histograms are Float64 counts
    return st;
}

int32_t mcx_flat_reset_histogram(mcx_flat *f)
{
    FREQ(f, MCX_ERR_ARGUMENT, "flat is NULL");
    FCUDA(cudaSetDevice(f->lat->ctx->device));
    FCUDA(cudaMemsetAsync(f->d_histogram, 0, sizeof(unsigned long long) * (size_t)f->nbins, f->lat->ctx->stream));
    return MCX_OK;
}

int32_t mcx_flat_set_logf(mcx_flat *f, double logf)
{
    FREQ(f, MCX_ERR_ARGUMENT, "flat is NULL");
    f->logf = logf;
    return MCX_OK;
}

int32_t mcx_flat_get_logf(mcx_flat *f, double *logf)
{
    FREQ(f && logf, MCX_ERR_ARGUMENT, "NULL argument");
    *logf = f->logf;
    return MCX_OK;
}

int32_t mcx_flat_sweep(mcx_flat *f, int64_t nsweeps)
{
    FREQ(f, MCX_ERR_ARGUMENT, "flat is NULL");
    FREQ(nsweeps >= 0, MCX_ERR_ARGUMENT, "nsweeps must be >= 0");
    mcx_lattice *lat = f->lat;
    FCUDA(cudaSetDevice(lat->ctx->device));
    knobs_refresh();
    if (lat->sums_dirty) { launch_recompute(lat); lat->sums_dirty = false; }
    // shared-memory counters are 32-bit: bound the attempts one block can record per launch
    const double per_sweep = 4.0 * (double)lat->N;
    int64_t chunk = (int64_t)(4.0e9 / per_sweep);
    if (chunk < 1) chunk = 1;
    for (int64_t done = 0; done < nsweeps; done += chunk) {
        const int n = (int)((nsweeps - done) < chunk ? (nsweeps - done) : chunk);
        launch_flat_sweep(f, lat->sweep + (uint64_t)done, n);
    }
    lat->sweep += (uint64_t)nsweeps;
    lat->steps += nsweeps * lat->N;
    FCUDA(cudaGetLastError());
    return MCX_OK;
}

int32_t mcx_flat_update(mcx_flat *f)
{
    FREQ(f, MCX_ERR_ARGUMENT, "flat is NULL");
    FCUDA(cudaSetDevice(f->lat->ctx->device));
    if (f->kind == MCX_FLAT_MUCA) launch_flat_update(f);
    else f->logf *= 0.5;                                   // update!(ens::WangLandauEnsemble; power=0.5)
    FCUDA(cudaGetLastError());
    return MCX_OK;
}

int32_t mcx_flat_device_histogram(mcx_flat *f, void **device_ptr, int64_t *nbins)
{
    FREQ(f && device_ptr, MCX_ERR_ARGUMENT, "NULL argument");
    *device_ptr = f->d_histogram;
    if (nbins) *nbins = f->nbins;
    return MCX_OK;
}

int32_t mcx_flat_device_logweight(mcx_flat *f, void **device_ptr, int64_t *nbins)
{
    FREQ(f && device_ptr, MCX_ERR_ARGUMENT, "NULL argument");
    *device_ptr = f->d_logweight;
    if (nbins) *nbins = f->nbins * f->ntables;
    return MCX_OK;
}

}  // extern "C"
