// k_queue.cu -- a whole series of sweeps of a batch of 2-D Ising lattices in ONE launch.
//
// The streaming kernel (k_ising2d.cu) needs one launch per half-sweep: small batches (a parallel-tempering
// rank's share of the replicas) then spend a third of their time in launch ramps and
// tails, because a half-sweep is only ~1000 CTA-items for ~900 resident CTAs.  Here the work items of ALL
// half-sweeps of the series are numbered in one sequence, ticket = (half-sweep, chain, item), and a
// persistent grid takes tickets from a global counter.  Nothing separates the half-sweeps but data
// dependencies: item i of half-sweep h reads the other colour's rows one above / below its own and
// overwrites rows that its neighbours read during half-sweep h - 1, so it may start once the items covering
// its strips and the strips around them have finished half-sweep h - 1 (one progress word per item, released
// with a fence + store, polled by a few threads of the CTA).  Tickets are taken in order, so an item only
// ever waits for lower tickets, which are already running: no deadlock, no co-residency requirement, and the
// SMs never drain until the series ends.  Rows written by other SMs during the same launch are read with
// ld.global.cg (L2; L1 is not coherent across SMs).  The per-site rule, the Philox counters and the sums are
// the streaming kernel's (update_row, k_row16.cuh): trajectories are bit-identical.
#include "k_strip.cuh"

namespace mcx {

namespace {

constexpr int kThreads = 128;
// resident CTAs per SM the kernel is compiled for.  5 (96 registers); 6 compiles to 80 registers without spills but was
// measured 6-11 % slower at every batch size (37 M more warp instructions per 20 sweeps of 32 x 1024^2:
// profiles/r02_queue_kernel.md)
#ifndef MCX_QUEUE_MINB
#define MCX_QUEUE_MINB 5
#endif

enum { Q_TICKET = 0, Q_WORDS = 2 };

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one CTA-item of one half-sweep: 128 thread-items, each a 16-byte column segment of a strip of R rows (sweep_strip,
// k_strip.cuh: the streaming kernel's loop with loads and stores through L2)
template <int COLOUR, bool HEATBATH, bool TRACK>
__device__ __forceinline__ void queue_item(const LatView &L, const int chain, const int item, const uint64_t t,
                                           const uint32_t *s_pair, const uint32_t *s_thi, const uint32_t *s_tlo,
                                           long long *__restrict__ sums, const uint32_t seed_lo, const uint32_t seed_hi,
                                           const uint32_t first_chain, const int R, const int nstrips)
{
    const int half = L.half;
    const int nseg = half >> 4;
    const uint32_t G = (uint32_t)nstrips * (uint32_t)nseg;       // < 2^31 (the launcher checks)
    const int lane = threadIdx.x & 31;
    const uint32_t g0 = (uint32_t)item * kThreads + threadIdx.x;
    const bool active = g0 < G;
    const uint32_t g = active ? g0 : G - 1;
    const int strip = (int)(g / (uint32_t)nseg);
    const int seg = (int)(g - (uint32_t)strip * (uint32_t)nseg);
    uint8_t *tgt = plane_ptr(L, chain, COLOUR);
    const uint8_t *oth = plane_ptr(L, chain, COLOUR ^ 1);
    StripGeom sg;
    sg.row0 = strip * R; sg.rows = R; sg.col = seg << 4;
    sg.colL = (seg == 0 ? half : sg.col) - 1;
    sg.colR = (seg == nseg - 1) ? 0 : sg.col + 16;
    sg.loadL = (lane == 0) || (seg == 0);
    sg.loadR = (lane == 31) || (seg == nseg - 1);
    sg.active = active;
    const Acc acc = sweep_strip<COLOUR, HEATBATH, TRACK, true>(tgt, oth, oth, oth, half, L.Ly, 0, sg, t, first_chain + (uint32_t)chain,
                                                               seed_lo, seed_hi, s_pair, s_thi, s_tlo);
    strip_finish<TRACK>(acc, sums, chain);
}

template <bool HEATBATH, bool TRACK>
__global__ void __launch_bounds__(kThreads, MCX_QUEUE_MINB)
k_ising2d_queue(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
                const int32_t *__restrict__ labels, long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi,
                uint64_t t0, uint64_t nhalf, uint32_t first_chain, int R, int nstrips, int ipc /* items per chain and half-sweep */,
                unsigned long long *ctl, uint32_t *progress /* [nchains][ipc]: half-sweeps of the series finished by the item */)
{
    __shared__ uint32_t s_pair[kPairWords];
    __shared__ uint32_t s_thi[kTableLen], s_tlo[kTableLen];
    __shared__ unsigned long long s_ticket;
    int cur_label = -1;
    const int nseg = L.half >> 4;
    const uint64_t per_half = (uint64_t)L.nchains * (uint64_t)ipc;
    const uint64_t total = nhalf * per_half;

    if (threadIdx.x == 0) s_ticket = atomicAdd(ctl + Q_TICKET, 1ull);
    __syncthreads();
    unsigned long long ticket = s_ticket;
    while (ticket < total) {
        __syncthreads();                                       // every thread holds `ticket`: the slot may be refilled
        if (threadIdx.x == 0) s_ticket = atomicAdd(ctl + Q_TICKET, 1ull);   // next ticket; its latency hides behind this item
        const uint64_t h = ticket / per_half;
        const uint32_t rem = (uint32_t)(ticket - h * per_half);
        const int chain = (int)(rem / (uint32_t)ipc);
        const int item = (int)(rem - (uint32_t)chain * (uint32_t)ipc);
        uint32_t *prog = progress + (int64_t)chain * ipc;
        if (h > 0) {
            // the items that cover my strips and the strips just above / below them must have finished half-sweep h - 1
            const int s_lo = (int)(((int64_t)item * kThreads) / nseg);
            int s_hi = (int)(((int64_t)item * kThreads + kThreads - 1) / nseg);
            if (s_hi > nstrips - 1) s_hi = nstrips - 1;
            const int nspan = min(s_hi - s_lo + 3, nstrips);              // strips s_lo - 1 .. s_hi + 1, periodic
            // walk the strips of the span; each contributes the items [first, last] that hold its thread-items
            for (int k = threadIdx.x; k < nspan * ((nseg + kThreads - 1) / kThreads + 1); k += kThreads) {
                const int per = (nseg + kThreads - 1) / kThreads + 1;     // upper bound of items touching one strip
                const int sidx = k / per, j = k - sidx * per;
                int strip = s_lo - 1 + sidx;
                strip = strip < 0 ? strip + nstrips : strip >= nstrips ? strip - nstrips : strip;
                const int first = (int)(((int64_t)strip * nseg) / kThreads);
                const int last = (int)(((int64_t)strip * nseg + nseg - 1) / kThreads);
                const int dep = first + j;
                if (dep <= last) {
                    unsigned long long t_begin, t_now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
                    while (ld_acquire(prog + dep) < (uint32_t)h) {
                        __nanosleep(64);
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
                        if (t_now - t_begin > 10000000000ull) {          // 10 s: give up, never hang; the host's next call fails
                            if (L.err) { *(volatile int *)L.err = ASYNC_ERR_QUEUE_DEP; __threadfence_system(); }
                            break;
                        }
                    }
                }
            }
        }
        const int label = labels[chain];
        __syncthreads();                                       // dependencies seen by the whole CTA; tables free to change
        if (label != cur_label) {
            load_pair_table(s_pair, s_thi, s_tlo, thi_g, tlo_g, label);
            cur_label = label;
            __syncthreads();
        }
        const uint64_t t = t0 + h;
        if (t & 1) queue_item<1, HEATBATH, TRACK>(L, chain, item, t, s_pair, s_thi, s_tlo, sums, seed_lo, seed_hi, first_chain, R, nstrips);
        else queue_item<0, HEATBATH, TRACK>(L, chain, item, t, s_pair, s_thi, s_tlo, sums, seed_lo, seed_hi, first_chain, R, nstrips);
        __syncthreads();                                       // all rows of the item stored
        if (threadIdx.x == 0) {
            __threadfence();                                   // ... and visible before the progress word says so
            *(volatile uint32_t *)(prog + item) = (uint32_t)h + 1u;
        }
        ticket = s_ticket;
    }
}

}  // namespace

// nsweeps whole sweeps of lat->sweep .. in one launch (advances lat->sweep); false: not applicable, nothing launched.
// MCX_QUEUE=1: whenever the shape allows; =0: never; unset: small batches only -- a half-sweep of between half
// and one and a half work items per resident CTA (e.g. the 32 replicas of 1024 x 1024 a parallel-tempering rank
// holds at 8 GPUs: 1207 attempts/ns against 1018 with one launch per half-sweep and chain group,
// profiles/r01_queue_kernel.md); bigger batches are faster with the chain-group launches.
bool launch_sweeps_ising2d_queue(mcx_lattice *lat, int64_t nsweeps)
{
    const int want = knobs().queue;
    if (want == 0 || nsweeps < 1) return false;
    if (lat->storage != MCX_STORAGE_INT8 || lat->slab || !lat->fast2d || lat->model != MCX_ISING) return false;
    if (knobs().variant >= 0 || knobs().rows_per_strip >= 0 || knobs().force_generic > 0) return false;
    if (want < 0 && (nsweeps < 4 || knobs().groups == 0)) return false;
    const bool single = lat->nchains == 1;                      // one mid-size lattice: see the policy below
    if (want < 0 && single && knobs().bands >= 0) return false; // MCX_BANDS set: the caller chose (or excluded) the band launches
    mcx_ctx *ctx = lat->ctx;
    const LatView &L = lat->view;
    const int nseg = L.half >> 4;
    const bool hb = lat->rule == MCX_HEATBATH, track = lat->track_sums;
    auto kern = hb ? (track ? k_ising2d_queue<true, true> : k_ising2d_queue<true, false>)
                   : (track ? k_ising2d_queue<false, true> : k_ising2d_queue<false, false>);
    int resident = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kThreads, 0) != cudaSuccess || resident < 1) return false;
    const int64_t grid_max = (int64_t)ctx->sm_count * resident;
    // strip height: 16 rows if a half-sweep then still has an item for every resident CTA, else 8 (measured at 32
    // replicas of 1024 x 1024: 8 rows 1207 attempts/ns, 16 rows 1010, 4 rows 1018); MCX_QUEUE_ROWS overrides
    // A single lattice of 1024 ... 8192 rows is launch- and tail-bound with one launch (or eight band launches) per
    // half-sweep; as ONE series launch with strips short enough to give every resident CTA an item per half-sweep it runs
    // at 1391 instead of 946 attempts/ns (L = 8192, 16-row strips), 963 / 598 (4096, 4 rows), 339 / 213 (2048, 2 rows),
    // 118 / 67 (1024, 2 rows): profiles/r02_queue_single.md.  L = 16384 stays with the row bands (1522 against 1250).
    int R = knobs().queue_rows > 0 ? (knobs().queue_rows > 32 ? 32 : knobs().queue_rows) : 16, ipc = 0;   // <= 32: sweep_strip's flags
    const int r_min = single ? 2 : 8;
    for (;; R >>= 1) {
        int r = R;
        while (r > 2 && (L.Ly % r != 0 || r % 2 != 0)) --r;
        const int64_t Gt = (int64_t)(L.Ly / r) * nseg;
        ipc = (int)((Gt + kThreads - 1) / kThreads);
        if ((int64_t)ipc * lat->nchains >= grid_max || R <= r_min || knobs().queue_rows > 0) { R = r; break; }
    }
    const int nstrips = L.Ly / R;
    if (nstrips < 3) return false;                             // the dependency span assumes distinct neighbours
    const int64_t per_half = (int64_t)ipc * lat->nchains;
    if (want < 0 && ((single ? 8 : 2) * per_half < grid_max || 2 * per_half > 3 * grid_max)) return false;
    if (per_half >= ((int64_t)1 << 31)) return false;
    // control block + progress words, sized for this lattice, zeroed per series
    const size_t need = sizeof(unsigned long long) * Q_WORDS + sizeof(uint32_t) * (size_t)per_half;
    if (lat->queue_bytes < need) {
        cudaFree(lat->d_queue);
        lat->d_queue = nullptr; lat->queue_bytes = 0;
        if (cudaMalloc((void **)&lat->d_queue, need) != cudaSuccess) { cudaGetLastError(); return false; }
        lat->queue_bytes = need;
    }
    unsigned long long *ctl = (unsigned long long *)lat->d_queue;
    uint32_t *progress = (uint32_t *)(ctl + Q_WORDS);
    LatView Lk = L;
    Lk.err = ctx->d_err;                                       // a dependency wait that gives up raises the context's error word
    int64_t grid = grid_max;
    if (grid > per_half * 2) grid = per_half * 2;
    // 2^32 half-sweeps of progress per launch at most; chunk long series
    for (int64_t done = 0; done < nsweeps;) {
        const int64_t chunk = nsweeps - done < 500000000 ? nsweeps - done : 500000000;
        cudaMemsetAsync(lat->d_queue, 0, need, ctx->stream);
        kern<<<(unsigned)grid, kThreads, 0, ctx->stream>>>(Lk, lat->d_thi, lat->d_tlo, lat->d_labels, lat->d_sums, (uint32_t)lat->seed,
                                                           (uint32_t)(lat->seed >> 32), 2 * (lat->sweep + (uint64_t)done),
                                                           2 * (uint64_t)chunk, lat->first_chain, R, nstrips, ipc, ctl, progress);
        ctx->launches++;
        done += chunk;
    }
    lat->sweep += (uint64_t)nsweeps;
    return true;
}

}  // namespace mcx
