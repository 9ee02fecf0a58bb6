// k_persist.cu -- whole parallel-tempering runs in ONE launch: nrounds x (sweeps, energies to all ranks, replica
// exchange, labels), with nothing queued by the host between rounds (replica_exchange.jl:158-178 driven by the loop
// of docs/src/examples/spin_systems/pt_Ising2D.jl:52-57).
//
// Decomposition.  Every WARP of a co-resident grid owns fixed work items (round-robin, item = warp + j * warps) and
// walks the phases (half-sweeps) itself.  An item is 32 thread-items -- a 16-byte column segment of a strip of rows
// each, the thread-row update of k_row16.cuh -- so a warp is the unit of scheduling: it keeps a private copy of its
// chain's pair-threshold table in shared memory (rebuilt only when an exchange moved the chain's label) and nothing in
// the kernel is a CTA-wide barrier.  No launch boundary separates the half-sweeps, only data dependencies: an item of
// phase p may start once the neighbour items have finished phase p - 1 (they wrote the rows it reads and have read
// the rows it overwrites); one progress word per item, published with a release store, polled by a lane per
// neighbour.  A waiting item only ever waits for an earlier phase and every warp walks the phases in order, so the
// earliest unfinished (phase, item) can always run: no deadlock as long as the grid is co-resident (sized from the
// occupancy calculator).  Rows written by other SMs during the launch are read with ld.global.cg (L2).
// For row-aligned widths (a warp's 32 thread-items are one 512-byte row segment: Lx % 1024 == 0) the strips have
// individual even heights, chosen so that the items are a whole multiple of the warps: a round that ends in an
// exchange is a barrier, and a partly filled last wave of items would idle most of the machine.
//
// A round is `sweeps_per_round` sweeps (2 S phases) and, unless the sums are tracked per flip (S < 3), one more phase
// whose items re-evaluate sum_<ij> s_i s_j and sum s of their rows.  Completion is counted per chain and then per
// rank (nchains + 1 counters, so no single word takes thousands of atomics at once); the warp that completes the rank
// closes the round: energies of the rank's replicas into every rank's buffer over NVLink + arrival counter (the fused
// all-gather of k_pt.cu), wait for all ranks' counters, the pair decisions of k_pt_exchange (same EXCHANGE stream, same
// float expression), labels, and one release flag.  Each CTA has ONE poller of that flag (the others watch shared
// memory): thousands of warps polling one L2 word delayed the closing warp's own accesses by tens of microseconds.
// Trajectories, ladder state and counters are bit-identical to mcx_sweep + mcx_pt_publish + mcx_pt_exchange per round
// (tests/test_gpu_pt_persistent.py), which are held to the oracle.
#include "k_strip.cuh"

namespace mcx {

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
#ifndef MCX_PERSIST_MINB
#define MCX_PERSIST_MINB 5
#endif
constexpr int kMinBlocks = MCX_PERSIST_MINB;                 // 96 registers; 6 CTAs / 80 registers was measured slower for the series kernels
// shared memory of one warp: the pair table, the exact 32-bit thresholds, padding to a multiple of 128 bytes
constexpr int kWarpTableWords = (kPairWords + 2 * kTableLen + 31) / 32 * 32;

// control words, zeroed per launch, 128 bytes apart (pollers of one must not queue up in front of the other's atomics).
// Q_DONE: chains whose last phase of a round is finished (cumulative over the launch); Q_EXCHANGED: rounds whose
// exchange is decided and whose labels are final.  They are followed by one completion counter per chain (16 words
// apart) and the progress words.
enum { Q_DONE = 0, Q_EXCHANGED = 16, Q_WORDS = 32, Q_CHAIN_STRIDE = 4 };

// device view of the ladder for the rounds closed inside the kernel
struct PtDev {
    int n, nlocal, first_slot, nranks, rank;
    int stage0;                        // stage of the first round of this launch
    int recompute;                     // 1: the last phase of a round re-evaluates the sums (untracked sweeps)
    int peers;
    unsigned long long round0;         // exchange round of the first round of this launch
    const double *betas;
    double *x;                         // this rank's energy buffer(s): [n], or [2][n] by round parity with peers
    int32_t *index, *slot_of, *labels;
    long long *steps, *accepted;
    double *const *peer_x;
    unsigned long long *const *peer_arrived;
    const unsigned long long *arrived;
    int *pt_err;
    double J, h;
    long long pair0, spin0;            // constant terms of the recomputed sums: 4 (N/2) and -N
};

__device__ __forceinline__ uint4 ld_cg128(const uint8_t *p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg8(const uint8_t *p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int32_t ld_cg32(const int32_t *p)
{
    int32_t v;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
__device__ __forceinline__ void raise_error(int *err, int code)
{
    if (err) { *(volatile int *)err = code; __threadfence_system(); }
}

// decomposition of an item's 32 thread-items shared by the update and the recompute phase
struct ItemGeom {
    int strip, seg, col, colL, colR, row0, rows;
    bool active, loadL, loadR;
};
__device__ __forceinline__ ItemGeom item_geom(const LatView &L, const int item, const int R, const int nstrips)
{
    ItemGeom g;
    const int half = L.half, nseg = half >> 4;
    const int G = nstrips * nseg;                             // thread-items per chain and phase: < 2^31 (rounds_plan)
    const int lane = threadIdx.x & 31;
    const int g0 = item * 32 + lane;
    g.active = g0 < G;
    const int gi = g.active ? g0 : G - 1;
    g.strip = gi / nseg;
    g.seg = gi - g.strip * nseg;
    g.row0 = g.strip * R;                                     // even
    g.rows = min(R, L.Ly - g.row0);                           // the last strip may be shorter (even)
    g.col = g.seg << 4;
    g.colL = (g.seg == 0 ? half : g.col) - 1;                 // byte left of the segment (periodic)
    g.colR = (g.seg == nseg - 1) ? 0 : g.col + 16;            // byte right of the segment
    g.loadL = (lane == 0) || (g.seg == 0);
    g.loadR = (lane == 31) || (g.seg == nseg - 1);
    return g;
}

// per-chain sums of a finished item: dspin = 2 - 4 s, dpair = -8 s nup + 16 s + 4 nup - 8 per changed site
template <bool TRACK>
__device__ __forceinline__ void item_finish(const Acc &acc, long long *__restrict__ sums, const int chain)
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

// one item of one half-sweep: 32 thread-items, each a 16-byte column segment of a strip of rows (sweep_strip, k_strip.cuh:
// the streaming kernel's loop with loads and stores through L2)
template <int COLOUR, bool HEATBATH, bool TRACK>
__device__ __forceinline__ void strip_item(const LatView &L, const int chain, const ItemGeom &g, const uint64_t t,
                                           const uint32_t *s_pair, const uint32_t *s_thi, const uint32_t *s_tlo,
                                           long long *__restrict__ sums, const uint32_t seed_lo, const uint32_t seed_hi,
                                           const uint32_t first_chain)
{
    uint8_t *tgt = plane_ptr(L, chain, COLOUR);
    const uint8_t *oth = plane_ptr(L, chain, COLOUR ^ 1);
    StripGeom sg;
    sg.row0 = g.row0; sg.rows = g.rows; sg.col = g.col; sg.colL = g.colL; sg.colR = g.colR;
    sg.loadL = g.loadL; sg.loadR = g.loadR; sg.active = g.active;
    const Acc acc = sweep_strip<COLOUR, HEATBATH, TRACK, true>(tgt, oth, oth, oth, L.half, L.Ly, 0, sg, t, first_chain + (uint32_t)chain,
                                                               seed_lo, seed_hi, s_pair, s_thi, s_tlo);
    item_finish<TRACK>(acc, sums, chain);
}

// the recompute phase of a round: _recompute_cached! (ising.jl:500-504) of the item's rows, both colour planes read
// through L2.  With e in {0,1}: sum_<ij> s_i s_j = 4 sum(e0 nup) - 2 sum(nup) - 8 sum(e0) + 4 (N/2) over the colour-0
// sites, sum s = 2 (sum e0 + sum e1) - N (k_recompute2d); the constants are the values the sums were reset to.
__device__ __noinline__ void strip_item_recompute(const uint8_t *p0, const uint8_t *p1, const int half, const int Ly, const ItemGeom g,
                                                  unsigned long long *o)
{
    int e0n = 0, nsum = 0, e0 = 0, e1 = 0;
    if (g.active) {
        for (int r = 0; r < g.rows; ++r) {
            const int row = g.row0 + r;
            const int ru = row == 0 ? Ly - 1 : row - 1, rd = row == Ly - 1 ? 0 : row + 1;
            const uint4 T = ld_cg128(p0 + (int64_t)row * half + g.col);
            const uint4 C = ld_cg128(p1 + (int64_t)row * half + g.col);
            const uint4 U = ld_cg128(p1 + (int64_t)ru * half + g.col);
            const uint4 D = ld_cg128(p1 + (int64_t)rd * half + g.col);
            uint32_t S[4];
            if ((row & 1) == 0) {       // colour-0 sites of an even row sit at x = 2j: neighbours j-1, j
                const uint32_t side = ld_cg8(p1 + (int64_t)row * half + g.colL);
                S[0] = (C.x << 8) | side;
                S[1] = __funnelshift_l(C.x, C.y, 8); S[2] = __funnelshift_l(C.y, C.z, 8); S[3] = __funnelshift_l(C.z, C.w, 8);
            } else {
                const uint32_t side = ld_cg8(p1 + (int64_t)row * half + g.colR);
                S[0] = __funnelshift_r(C.x, C.y, 8); S[1] = __funnelshift_r(C.y, C.z, 8); S[2] = __funnelshift_r(C.z, C.w, 8);
                S[3] = (C.w >> 8) | (side << 24);
            }
            const uint32_t tw[4] = {T.x, T.y, T.z, T.w}, cw[4] = {C.x, C.y, C.z, C.w};
            const uint32_t nup[4] = {U.x + D.x + C.x + S[0], U.y + D.y + C.y + S[1], U.z + D.z + C.z + S[2], U.w + D.w + C.w + S[3]};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                e0n = __dp4a(nup[w] & (tw[w] * 255u), 0x01010101u, (uint32_t)e0n);
                nsum = __dp4a(nup[w], 0x01010101u, (uint32_t)nsum);
                e0 = __dp4a(tw[w], 0x01010101u, (uint32_t)e0);
                e1 = __dp4a(cw[w], 0x01010101u, (uint32_t)e1);
            }
        }
    }
    e0n = warp_sum(e0n); nsum = warp_sum(nsum); e0 = warp_sum(e0); e1 = warp_sum(e1);
    if ((threadIdx.x & 31) == 0) {
        const long long pair = 4ll * e0n - 2ll * nsum - 8ll * e0, spin = 2ll * (e0 + e1);
        if (pair) atomicAdd(o + SUM_PAIR, (unsigned long long)pair);
        if (spin) atomicAdd(o + SUM_SPIN, (unsigned long long)spin);
    }
}

// The warp that finished the last item of round `r` (relative to the launch) closes it: energies to every rank,
// update!(rx, xs) (replica_exchange.jl:158-178; body of k_pt_exchange), labels, release of round r + 1.
// All ladder state is read and written through L2 (volatile): consecutive rounds are closed by different SMs.
__device__ __noinline__ void pt_close_round(const PtDev *__restrict__ Pd, const uint32_t r, long long *sums, const uint32_t seed_lo,
                                            const uint32_t seed_hi, unsigned long long *ctl, int *err)
{
    const PtDev P = *Pd;                                   // constant during the launch
    const int lane = threadIdx.x & 31;
    const unsigned long long round = P.round0 + r;
    const int stage = (P.stage0 + (int)r) & 1;
    const int off = P.peers ? (int)(round & 1) * P.n : 0;
    for (int c = lane; c < P.nlocal; c += 32) {
        volatile long long *s = sums + (int64_t)c * SUM_FIELDS;
        const long long pair = s[SUM_PAIR], spin = s[SUM_SPIN];
        double e = -(P.J * (double)pair);                    // energy(sys), ising.jl:175-178
        if (P.h != 0.0) e -= P.h * (double)spin;
        if (P.peers) {
            for (int k = 0; k < P.nranks; ++k) ((volatile double *)P.peer_x[k])[off + P.first_slot + c] = e;
        } else {
            ((volatile double *)P.x)[off + P.first_slot + c] = e;
        }
        if (P.recompute) { s[SUM_PAIR] = P.pair0; s[SUM_SPIN] = P.spin0; }   // the next round's recompute phase adds to these
    }
    if (P.peers) {
        __threadfence_system();
        __syncwarp();
        for (int k = lane; k < P.nranks; k += 32) {
            *(volatile unsigned long long *)(P.peer_arrived[k] + P.rank) = round + 1;
            __threadfence_system();
        }
        // ... and wait for every rank's energies of this round (20 s: give up, never hang)
        for (int k = lane; k < P.nranks; k += 32) {
            const volatile unsigned long long *a = P.arrived + k;
            const unsigned long long t0 = now_ns();
            while (*a < round + 1) {
                __nanosleep(100);
                if (now_ns() - t0 > 20000000000ull) { *P.pt_err = 1; raise_error(err, ASYNC_ERR_PT_PEERS); break; }
            }
        }
        __threadfence_system();
    } else {
        __threadfence();
    }
    __syncwarp();
    const volatile double *x = P.x + off;
    volatile int32_t *index = P.index, *slot_of = P.slot_of, *labels = P.labels;
    volatile long long *steps = P.steps, *accepted = P.accepted;
    // 0-based pair k joins ladder indices k and k+1; stage 0 takes k = 0,2,4,.. (reference first=1)
    for (int k = (stage & 1) + 2 * lane; k < P.n - 1; k += 64) {
        const int ri = slot_of[k], rj = slot_of[k + 1];
        steps[k] += 1;
        // u = rand(algorithm(rx, ri).rng): EXCHANGE stream of slot ri at this round, 53 bits
        const Philox4 p = stream_block(seed_lo, seed_hi, (uint32_t)ri, TAG_EXCHANGE, round, 0, 0);
        const uint64_t w = ((uint64_t)p.y << 32) | p.x;
        const double u = (double)(w >> 11) * (1.0 / 9007199254740992.0);
        const double bi = P.betas[k], bj = P.betas[k + 1];
        const double xi = x[ri], xj = x[rj];
        // exchange_log_ratio (:110-113) with logweight(E) = -beta*E (ensembles/boltzmann.jl:28)
        const double lr = ((-bi * xj) - (-bi * xi)) + ((-bj * xi) - (-bj * xj));
        const bool acc = (lr > 0) || (u < exp(lr));
        if (acc) {
            accepted[k] += 1;
            index[ri] = k + 1; index[rj] = k;
            slot_of[k] = rj; slot_of[k + 1] = ri;
            if (ri >= P.first_slot && ri < P.first_slot + P.nlocal) labels[ri - P.first_slot] = k + 1;
            if (rj >= P.first_slot && rj < P.first_slot + P.nlocal) labels[rj - P.first_slot] = k;
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        *(volatile unsigned long long *)(ctl + Q_EXCHANGED) = (unsigned long long)r + 1ull;
    }
}

// pair table of one ensemble into the warp's private shared memory (load_pair_table of k_row16.cuh, 32 lanes)
__device__ __forceinline__ void load_pair_table_warp(uint32_t *s_pair, uint32_t *s_thi, uint32_t *s_tlo,
                                                     const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g, int label)
{
    const int lane = threadIdx.x & 31;
    if (lane < kTableLen) {
        s_thi[lane] = thi_g[label * kTableLen + lane];
        s_tlo[lane] = tlo_g[label * kTableLen + lane];
    }
    for (int e = lane; e < kTableLen * kTableLen; e += 32) {
        const int i1 = e / kTableLen, i0 = e - i1 * kTableLen;
        const uint32_t a = min(thi_g[label * kTableLen + i0] >> 1, 0x7fffu);
        const uint32_t b = min(thi_g[label * kTableLen + i1] >> 1, 0x7fffu);
        s_pair[i1 * kPairRowWords + i0] = a | (b << 16);
    }
}

enum { GEOM_UNIFORM = 0, GEOM_ROWS = 1 };

// row-aligned geometry: item = (strip, seg block); strip s covers rows [2 floor(s Ly/2 / nstrips), 2 floor((s+1) Ly/2 / nstrips))
__device__ __forceinline__ int strip_row(const int s, const int Ly, const int nstrips)
{
    return 2 * (int)(((unsigned)s * (unsigned)(Ly >> 1)) / (unsigned)nstrips);
}
__device__ __forceinline__ ItemGeom item_geom_rows(const LatView &L, const int item, const int nstrips, const int sbn)
{
    ItemGeom g;
    const int half = L.half, nseg = half >> 4;
    const int lane = threadIdx.x & 31;
    g.strip = item / sbn;
    const int sb = item - g.strip * sbn;
    g.seg = sb * 32 + lane;
    g.active = true;                                          // nseg % 32 == 0
    g.row0 = strip_row(g.strip, L.Ly, nstrips);
    g.rows = strip_row(g.strip + 1, L.Ly, nstrips) - g.row0;
    g.col = g.seg << 4;
    g.colL = (g.seg == 0 ? half : g.col) - 1;
    g.colR = (g.seg == nseg - 1) ? 0 : g.col + 16;
    g.loadL = (lane == 0) || (g.seg == 0);
    g.loadR = (lane == 31) || (g.seg == nseg - 1);
    return g;
}

template <bool HEATBATH, bool TRACK>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_ising2d_rounds(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
                 const int32_t *labels, long long *sums, uint32_t seed_lo, uint32_t seed_hi,
                 uint64_t t0, uint32_t phases /* per round: half-sweeps (+ 1 recompute phase) */, uint32_t nrounds, uint32_t first_chain,
                 int R, int nstrips, int ipc /* items per chain and phase */, int geom, unsigned long long *ctl,
                 uint32_t *progress /* [nchains][ipc]: phases of the launch finished by the item */,
                 const PtDev *__restrict__ Pd, int recompute)
{
    extern __shared__ uint32_t s_tables[];                     // [kWarps][kWarpTableWords]
    __shared__ unsigned int s_round;                           // rounds this CTA has seen released
    __shared__ int s_poller;                                   // a warp of this CTA is polling the global flag
    const int lane = threadIdx.x & 31;
    uint32_t *s_pair = s_tables + (threadIdx.x >> 5) * kWarpTableWords;
    uint32_t *s_thi = s_pair + kPairWords, *s_tlo = s_thi + kTableLen;
    const int nseg = L.half >> 4;
    const int sbn = (nseg + 31) >> 5;                          // 32-segment blocks per row
    const int nitems = L.nchains * ipc;
    const int W = (int)gridDim.x * kWarps, w = (int)blockIdx.x * kWarps + (int)(threadIdx.x >> 5);
    const uint32_t sweep_phases = recompute ? phases - 1 : phases;
    const int per = (nseg + 31) / 32 + 1;                      // uniform geometry: upper bound of items touching one strip
    int cur_label = -1, cur_chain = -1;
    uint32_t hh = 0;                                           // phases since the start of the launch
    if (threadIdx.x == 0) { s_round = 0; s_poller = 0; }
    __syncthreads();                                           // the only CTA-wide barrier of the kernel

    for (uint32_t r = 0; r < nrounds; ++r) {
        if (r > 0) {
            // the previous round's exchange must be decided (labels final) before this round reads a label
            if (lane == 0) {
                const unsigned long long t_begin = now_ns();
                while (*(volatile unsigned int *)&s_round < r) {
                    if (atomicCAS(&s_poller, 0, 1) == 0) {     // this CTA's poller for now
                        while (ld_acquire64(ctl + Q_EXCHANGED) < r) {
                            __nanosleep(200);
                            if (now_ns() - t_begin > 25000000000ull) { raise_error(L.err, ASYNC_ERR_PT_ROUND); break; }
                        }
                        *(volatile unsigned int *)&s_round = r;
                        __threadfence_block();
                        atomicExch(&s_poller, 0);
                    } else {
                        __nanosleep(100);
                    }
                }
                __threadfence();                               // acquire side for the labels read below
            }
            __syncwarp();
            cur_chain = -1;
        }
        for (uint32_t ph = 0; ph < phases; ++ph, ++hh) {
            for (int it = w; it < nitems; it += W) {
                const int chain = it / ipc, item = it - chain * ipc;
                uint32_t *prog = progress + (int64_t)chain * ipc;
                const ItemGeom g = geom == GEOM_ROWS ? item_geom_rows(L, item, nstrips, sbn) : item_geom(L, item, R, nstrips);
                if (hh > 0) {
                    // the neighbour items must have finished phase hh - 1
                    int dep = -1;
                    if (geom == GEOM_ROWS) {
                        const int strip = g.strip, sb = item - strip * sbn;
                        const int ndeps = sbn > 1 ? 4 : 2;
                        if (lane == 0) dep = (strip == 0 ? nstrips - 1 : strip - 1) * sbn + sb;
                        else if (lane == 1) dep = (strip == nstrips - 1 ? 0 : strip + 1) * sbn + sb;
                        else if (lane == 2) dep = strip * sbn + (sb == 0 ? sbn - 1 : sb - 1);
                        else if (lane == 3) dep = strip * sbn + (sb == sbn - 1 ? 0 : sb + 1);
                        if (lane >= ndeps) dep = -1;
                        if (dep >= 0) {
                            const unsigned long long t_begin = now_ns();
                            while (ld_acquire(prog + dep) < hh) {
                                __nanosleep(32);
                                if (now_ns() - t_begin > 10000000000ull) { raise_error(L.err, ASYNC_ERR_QUEUE_DEP); break; }
                            }
                        }
                    } else {
                        const int s_lo = (item * 32) / nseg;
                        int s_hi = (item * 32 + 31) / nseg;
                        if (s_hi > nstrips - 1) s_hi = nstrips - 1;
                        const int nspan = min(s_hi - s_lo + 3, nstrips);          // strips s_lo - 1 .. s_hi + 1, periodic
                        for (int k = lane; k < nspan * per; k += 32) {
                            const int sidx = k / per, j = k - sidx * per;
                            int strip = s_lo - 1 + sidx;
                            strip = strip < 0 ? strip + nstrips : strip >= nstrips ? strip - nstrips : strip;
                            const int first = (strip * nseg) >> 5;
                            const int last = (strip * nseg + nseg - 1) >> 5;
                            dep = first + j;
                            if (dep <= last) {
                                const unsigned long long t_begin = now_ns();
                                while (ld_acquire(prog + dep) < hh) {
                                    __nanosleep(32);
                                    if (now_ns() - t_begin > 10000000000ull) { raise_error(L.err, ASYNC_ERR_QUEUE_DEP); break; }
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                if (ph < sweep_phases) {
                    if (chain != cur_chain) {
                        const int label = ld_cg32(labels + chain);     // rewritten between rounds by the closing warp
                        cur_chain = chain;
                        if (label != cur_label) {
                            __syncwarp();
                            load_pair_table_warp(s_pair, s_thi, s_tlo, thi_g, tlo_g, label);
                            cur_label = label;
                            __syncwarp();
                        }
                    }
                    const uint64_t t = t0 + (uint64_t)r * sweep_phases + ph;
                    if (t & 1) strip_item<1, HEATBATH, TRACK>(L, chain, g, t, s_pair, s_thi, s_tlo, sums, seed_lo, seed_hi, first_chain);
                    else strip_item<0, HEATBATH, TRACK>(L, chain, g, t, s_pair, s_thi, s_tlo, sums, seed_lo, seed_hi, first_chain);
                } else {
                    strip_item_recompute(plane_ptr(L, chain, 0), plane_ptr(L, chain, 1), L.half, L.Ly, g,
                                         (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS));
                }
                __syncwarp();                                  // all rows of the item stored, all sums added
                const bool last_phase = ph == phases - 1;
                int closer = 0;
                if (lane == 0) {
                    __threadfence();                           // ... and visible before the progress word says so
                    *(volatile uint32_t *)(prog + item) = hh + 1u;
                    if (last_phase) {
                        // items of a chain, then chains of the rank
                        const unsigned long long mine = atomicAdd(ctl + Q_WORDS + (size_t)chain * Q_CHAIN_STRIDE, 1ull) + 1ull;
                        if (mine == (unsigned long long)ipc * ((unsigned long long)r + 1ull)) {
                            __threadfence();
                            const unsigned long long done = atomicAdd(ctl + Q_DONE, 1ull) + 1ull;
                            closer = done == (unsigned long long)L.nchains * ((unsigned long long)r + 1ull);
                            if (closer) __threadfence();       // every item's sums precede its count
                        }
                    }
                }
                if (last_phase) {
                    closer = __shfl_sync(0xffffffffu, closer, 0);
                    if (closer) pt_close_round(Pd, r, sums, seed_lo, seed_hi, ctl, L.err);
                }
            }
        }
    }
}

// sets the control words and progress words to zero and, with a recompute phase, the pair / spin sums to the
// constant terms the phase adds to (replaces a cudaMemsetAsync: one launch either way)
__global__ void k_rounds_prepare(unsigned long long *ctl, size_t words64, long long *sums, int nchains, int reset_sums,
                                long long pair0, long long spin0)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t k = i; k < words64; k += (size_t)gridDim.x * blockDim.x) ctl[k] = 0ull;
    if (reset_sums && i < (size_t)nchains) {
        sums[i * SUM_FIELDS + SUM_PAIR] = pair0;
        sums[i * SUM_FIELDS + SUM_SPIN] = spin0;
    }
}

struct RoundsPlan {
    int R, nstrips, ipc, geom;
    int64_t per_phase;     // items per phase over all chains
    int64_t workers_max;   // warps of a full grid
    int64_t grid;          // CTAs to launch
};

constexpr size_t kRoundsSmem = sizeof(uint32_t) * kWarps * kWarpTableWords;

// decomposition for the static kernel
template <typename K>
bool rounds_plan(const mcx_lattice *lat, K kern, RoundsPlan &q)
{
    const LatView &L = lat->view;
    const int nseg = L.half >> 4;
    static thread_local K prepared = nullptr;
    if (prepared != kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        prepared = kern;
    }
    int resident = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kThreads, kRoundsSmem) != cudaSuccess || resident < 1) {
        cudaGetLastError();
        return false;
    }
    int64_t grid_max = (int64_t)lat->ctx->sm_count * resident;
    if (knobs().queue_grid > 0 && knobs().queue_grid < grid_max) grid_max = knobs().queue_grid;
    q.workers_max = grid_max * kWarps;
    if (nseg % 32 == 0) {
        // strips of individual even heights: as many items as warps, or a whole multiple when 16-row strips give more
        q.geom = GEOM_ROWS;
        const int sbn = nseg / 32;
        const int64_t per_strip = (int64_t)lat->nchains * sbn;                 // items that one more strip per chain adds
        int64_t nstrips;
        if (knobs().queue_rows > 0) {
            nstrips = (L.Ly + knobs().queue_rows - 1) / knobs().queue_rows;
        } else {
            const int64_t n16 = per_strip * ((L.Ly + 15) / 16);
            const int64_t m = n16 <= q.workers_max ? 1 : (n16 + q.workers_max - 1) / q.workers_max;
            nstrips = m * q.workers_max / per_strip;
        }
        if (nstrips < (L.Ly + 31) / 32) nstrips = (L.Ly + 31) / 32;           // sweep_strip flags left-over rows in 32 bits
        if (nstrips > L.Ly / 2) nstrips = L.Ly / 2;
        if (nstrips < 1) nstrips = 1;
        q.nstrips = (int)nstrips;
        q.R = (int)((L.Ly + nstrips - 1) / nstrips + 1) & ~1;                  // tallest strip (reported only)
        q.ipc = q.nstrips * sbn;
    } else {
        q.geom = GEOM_UNIFORM;
        int r = knobs().queue_rows > 0 ? knobs().queue_rows : 16;
        if (r > L.Ly) r = L.Ly;
        if (r > 32) r = 32;                                    // sweep_strip flags left-over rows in 32 bits
        r &= ~1;
        while (r > 2 && L.Ly % r != 0) r -= 2;
        if (r < 2) r = 2;
        if (knobs().queue_rows <= 0) {
            // shorter strips while they add items for idle warps
            while (r > 2 && ((int64_t)(L.Ly / r) * nseg + 31) / 32 * lat->nchains < q.workers_max) {
                int r2 = r - 2;
                while (r2 > 2 && L.Ly % r2 != 0) r2 -= 2;
                if (r2 < 2 || L.Ly % r2 != 0) break;
                r = r2;
                if (r <= 4) break;
            }
        }
        q.R = r;
        q.nstrips = L.Ly / r;
        if (q.nstrips < 3) return false;                       // the dependency span assumes distinct neighbours
        const int64_t Gt = (int64_t)q.nstrips * nseg;
        if (Gt + 64 >= ((int64_t)1 << 31)) return false;
        q.ipc = (int)((Gt + 31) / 32);
    }
    q.per_phase = (int64_t)q.ipc * lat->nchains;
    if (q.per_phase >= ((int64_t)1 << 31)) return false;
    q.grid = (q.per_phase + kWarps - 1) / kWarps;
    if (q.grid > grid_max) q.grid = grid_max;
    return true;
}

}  // namespace

// mcx_pt_run as one launch: nrounds x (sweeps_per_round sweeps, energies to all ranks, exchange).  The caller has
// made sure the energies can reach all ranks without the host (one rank, or peers attached).  Advances lat->sweep,
// lat->steps, pt->stage and pt->round.  MCX_PT_PERSIST=1: whenever the shape allows; =0: never; unset: when the rounds are
// short (an exchange at least every 16 sweeps: rounds queued from the host are bound by its launch rate) AND the batch
// is small (at most one and a half 16-row items per resident warp and half-sweep, e.g. the 32 or 64 replicas of
// 1024 x 1024 a rank holds at 8 or 4 GPUs).  Measured on one B200 (profiles/r02_pt_persistent.md), exchange after every
// sweep: 32 replicas 23094 rounds/s against 21007 queued from the host, 64 replicas 12600 against 11822; but 256
// replicas 3402 against 4665 -- big batches keep the streaming kernel's chain-group launches.
bool launch_pt_rounds_persistent(mcx_pt *pt, int64_t nrounds, int64_t S)
{
    mcx_lattice *lat = pt->lat;
    const int want = knobs().pt_persist;
    if (want == 0 || nrounds < 1 || S < 1 || lat->rule < 0) return false;
    if (lat->storage != MCX_STORAGE_INT8 || lat->slab || !lat->fast2d || lat->model != MCX_ISING) return false;
    if (knobs().variant >= 0 || knobs().rows_per_strip >= 0 || knobs().force_generic > 0) return false;
    if (want < 0 && S > 16) return false;
    // Across ranks the rounds queued by the host are faster: measured on 4 B200s with 64 replicas of 1024 x 1024 per rank
    // (profiles/r02_pt_multi_gpu.md), exchange after every sweep 13084 PT sweeps/s against 10659 through this launch, every
    // 200 sweeps 21187 against 17663 -- one warp closing the round behind two system-scope fences, with every other warp
    // of the rank parked, costs more than the publish / exchange kernels it replaces.  MCX_PT_PERSIST=1 still forces it.
    if (want < 0 && pt->peers && pt->nranks > 1) return false;
    mcx_ctx *ctx = lat->ctx;
    // short rounds: keep the sums current per flip; long rounds: sweep without bookkeeping, one recompute phase per round
    const bool track = S < 3, recompute = !track;
    const bool hb = lat->rule == MCX_HEATBATH;
    auto kern = hb ? (track ? k_ising2d_rounds<true, true> : k_ising2d_rounds<true, false>)
                   : (track ? k_ising2d_rounds<false, true> : k_ising2d_rounds<false, false>);
    RoundsPlan q;
    if (!rounds_plan(lat, kern, q)) return false;
    if (want < 0) {
        const int64_t natural = (int64_t)lat->nchains * ((lat->view.Ly + 15) / 16) * (((lat->view.half >> 4) + 31) / 32);
        if (2 * natural > 3 * q.workers_max) return false;
    }
    const int64_t phases = 2 * S + (recompute ? 1 : 0);
    if (phases >= ((int64_t)1 << 31)) return false;
    const size_t ctl_words = Q_WORDS + (size_t)lat->nchains * Q_CHAIN_STRIDE;
    const size_t need = sizeof(unsigned long long) * ctl_words + ((sizeof(uint32_t) * (size_t)q.per_phase + 7) & ~(size_t)7);
    if (lat->queue_bytes < need) {
        cudaFree(lat->d_queue);
        lat->d_queue = nullptr; lat->queue_bytes = 0;
        if (cudaMalloc((void **)&lat->d_queue, need) != cudaSuccess) { cudaGetLastError(); return false; }
        lat->queue_bytes = need;
    }
    if (!pt->d_dev && cudaMalloc(&pt->d_dev, sizeof(PtDev)) != cudaSuccess) { cudaGetLastError(); pt->d_dev = nullptr; return false; }
    if (track && lat->sums_dirty) { launch_recompute(lat); lat->sums_dirty = false; }
    lat->track_sums = track;
    PtDev P = {};
    P.n = pt->n; P.nlocal = lat->nchains; P.first_slot = pt->first_slot;
    P.nranks = pt->peers ? pt->nranks : 1; P.rank = pt->peers ? pt->rank : 0;
    P.recompute = recompute ? 1 : 0; P.peers = pt->peers ? 1 : 0;
    P.betas = pt->d_betas; P.x = pt->d_x; P.index = pt->d_index; P.slot_of = pt->d_slot_of; P.labels = lat->d_labels;
    P.steps = pt->d_steps; P.accepted = pt->d_accepted;
    P.peer_x = pt->d_peer_x; P.peer_arrived = pt->d_peer_arrived; P.arrived = pt->d_arrived; P.pt_err = pt->d_err;
    P.J = lat->J; P.h = lat->h;
    P.pair0 = 4 * lat->view.halfN; P.spin0 = -2 * lat->view.halfN;
    unsigned long long *ctl = (unsigned long long *)lat->d_queue;
    uint32_t *progress = (uint32_t *)(ctl + ctl_words);
    LatView L = lat->view;
    L.err = ctx->d_err;
    const int64_t max_rounds = ((int64_t)1 << 31) / phases;    // progress words count the phases of a launch in 32 bits
    for (int64_t done = 0; done < nrounds;) {
        const int64_t chunk = nrounds - done < max_rounds ? nrounds - done : max_rounds;
        P.stage0 = pt->stage; P.round0 = pt->round;
        if (cudaMemcpyAsync(pt->d_dev, &P, sizeof(P), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            cudaGetLastError();
            if (done == 0) return false;
            break;
        }
        k_rounds_prepare<<<64, 256, 0, ctx->stream>>>(ctl, lat->queue_bytes / 8, lat->d_sums, lat->nchains, recompute ? 1 : 0, P.pair0, P.spin0);
        kern<<<(unsigned)q.grid, kThreads, kRoundsSmem, ctx->stream>>>(L, lat->d_thi, lat->d_tlo, lat->d_labels, lat->d_sums, (uint32_t)lat->seed,
                                                                       (uint32_t)(lat->seed >> 32), 2 * lat->sweep, (uint32_t)phases, (uint32_t)chunk,
                                                                       lat->first_chain, q.R, q.nstrips, q.ipc, q.geom, ctl, progress,
                                                                       (const PtDev *)pt->d_dev, recompute ? 1 : 0);
        ctx->launches += 2;                                    // k_rounds_prepare + the rounds
        lat->sweep += (uint64_t)(chunk * S);
        lat->steps += chunk * S * lat->N;
        pt->stage = (pt->stage + (int)(chunk & 1)) & 1;
        pt->round += (uint64_t)chunk;
        done += chunk;
    }
    // after a recompute phase the closing warp has reset the sums for a next round that never came
    lat->sums_dirty = recompute;
    pt->persist_R = q.R;
    return true;
}

}  // namespace mcx
