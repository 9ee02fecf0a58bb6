// k_ising2d.cu -- vectorised checkerboard half-sweep for 2-D Ising lattices, one byte per spin.
//
// Work decomposition.  The target colour plane is a [Ly][Lx/2] byte matrix.  One thread owns a
// 16-byte column segment (one 128-bit load/store per row) and walks down a strip of R rows, two
// rows per loop trip, keeping the neighbour rows of the *other* colour plane in a rolling register
// window, so every byte of the other plane is read once per strip (+2 halo rows) and every byte of
// the target plane is read once and written once: 3 bytes per attempt, the algorithmic minimum
// (SURVEY.md 8d).  The left/right neighbour byte that falls outside the thread's segment comes from
// the adjacent lane by warp shuffle (edge lanes load one byte, prefetched a trip ahead).
//
// Randomness and the decision.  A thread-row needs 16 draws = two Philox4x32-10 blocks (eight
// 16-bit lanes each, the HIGH halves of the 32-bit draws).  The decision m < T is first tried on
// the top 15 bits, two sites per instruction: with hs = (w >> 1) & 0x7fff7fff (two 15-bit values)
// and tt the pair of 15-bit thresholds of the two sites (ONE shared-memory load from a pair table
// indexed by both sites' (spin, #up-neighbours) codes),
//     r  = (hs | 0x80008000) - tt           bit 15 / 31:  h15 >= t15   (not surely accepted)
//     r2 = r - 0x00010001                   bit 15 / 31:  h15 >  t15   (surely rejected)
// and a site is undecided (probability 2^-15) iff bit(r) & ~bit(r2).  Any undecided site sends the
// thread-row to an exact redo with the full 32-bit draws (low halves from Philox plane 1).  The
// result is bit-identical to k_sweep_generic and to the oracle.
//
// Persistent grid: gridDim.x = SMs x resident CTAs, items handed out round-robin.
#include "k_strip.cuh"

#include <cstdlib>

namespace mcx {

thread_local LaunchRange g_launch_range;
thread_local const unsigned long long *g_t_clock = nullptr;

namespace {

constexpr int kThreads = 128;
#ifndef MCX_MINB
#define MCX_MINB 8
#endif
constexpr int kMinBlocks = MCX_MINB;   // CTAs per SM the register budget is held to: 8 -> 64 registers (6: 1571, 7: 1554, 8: 1634, 9: 1581, 10: 1530 attempts/ns on one box)

// Cross-GPU ordering of a slab's half-sweep t (k_slab.cu): its boundary strips may start once both
// neighbours have finished the boundary strips of their half-sweep t - 1 ...
__device__ __forceinline__ void slab_wait(unsigned long long *ctl, uint64_t t, int sides, int *err)
{
    const volatile unsigned long long *f = ctl;
    const unsigned long long need = t - f[SLAB_T0];
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (((sides & 1) && f[SLAB_FLAG_UP] < need) || ((sides & 2) && f[SLAB_FLAG_DN] < need)) {
        __nanosleep(100);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) {                                 // 20 s: give up rather than hang the GPU
            ctl[SLAB_ERR] = 1;
            if (err) { *(volatile int *)err = ASYNC_ERR_SLAB; __threadfence_system(); }   // seen by the host's next call
            break;
        }
    }
    __threadfence_system();
}
// ... and the last of the nb boundary CTAs to finish tells the neighbour(s).
__device__ __forceinline__ void slab_signal(unsigned long long *ctl, uint64_t t, unsigned nb, int sides)
{
    __threadfence();
    unsigned *counter = (unsigned *)(ctl + (sides == 2 ? SLAB_ARRIVED_DN : SLAB_ARRIVED));
    const unsigned arrived = atomicAdd(counter, 1u);
    if (arrived + 1 == nb) {
        *(volatile unsigned *)counter = 0;
        const unsigned long long value = t - ((volatile unsigned long long *)ctl)[SLAB_T0] + 1;
        __threadfence_system();
        if (sides & 1) *(volatile unsigned long long *)ctl[SLAB_UP_SLOT] = value;
        if (sides & 2) *(volatile unsigned long long *)ctl[SLAB_DN_SLOT] = value;
        __threadfence_system();
    }
}

#ifdef MCX_OPT_TRACE
// timing probe (scripts/trace_ctas.py): per CTA {start ns, end ns, SM id, items done} of the last launch,
// followed by the start time of each of its first 8 items
__device__ unsigned long long g_trace[12 * 4096];
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
#endif

template <int COLOUR, bool HEATBATH, bool TRACK, bool FULL, bool SLAB>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_ising2d(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
          const int32_t *__restrict__ labels, long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi,
          uint64_t t, uint32_t first_chain, int R, int nstrips, int blocks_per_chain, int nitems, const unsigned long long *t_clock)
{
    // a launch replayed from a CUDA graph (mcx_pt_run): t counts from the device clock of the round (PtClock::t_base)
    if (t_clock) t += *(const volatile unsigned long long *)t_clock;
    __shared__ uint32_t s_pair[kPairWords];
    __shared__ uint32_t s_thi[kTableLen], s_tlo[kTableLen];
    int cur_label = -1;

    const int half = L.half;
    const int nseg = half >> 4;
    const int64_t G = (int64_t)nstrips * nseg;
    const int lane = threadIdx.x & 31;
#ifdef MCX_OPT_TRACE
    const unsigned long long trace_t0 = globaltimer_ns();
    int trace_items = 0;
#endif

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
#ifdef MCX_OPT_TRACE
        if (threadIdx.x == 0 && blockIdx.x < 4096 && trace_items < 8) g_trace[12 * blockIdx.x + 4 + trace_items] = globaltimer_ns();
        ++trace_items;
#endif
        const int chain = item / blocks_per_chain;
        const int label = labels[chain];
        if (label != cur_label) {
            __syncthreads();
            load_pair_table(s_pair, s_thi, s_tlo, thi_g, tlo_g, label);
            __syncthreads();
            cur_label = label;
        }
        // thread-items of a chain are numbered in 32 bits (the launcher checks G < 2^31): one 32-bit division per item
        const uint32_t g0 = (uint32_t)(item - chain * blocks_per_chain) * kThreads + threadIdx.x;
        const bool active = FULL ? true : g0 < (uint32_t)G;       // FULL: G % kThreads == 0, no idle lanes
        const uint32_t g = active ? g0 : (uint32_t)G - 1;
        int strip = (int)(g / (uint32_t)nseg);
        const int seg = (int)(g - (uint32_t)strip * (uint32_t)nseg);
        // Slab of a taller lattice: the two strips that touch the neighbour slabs come first (strip order
        // rotated by one), so that their rows are final -- and the neighbours told so -- while the interior
        // is still being swept.  Only the CTAs holding them wait for the neighbours' previous half-sweep.
        bool boundary_item = false;
        if (SLAB) {
            strip = strip == 0 ? nstrips - 1 : strip - 1;
            boundary_item = L.slab_ctl != nullptr && (int64_t)item * kThreads < 2 * (int64_t)nseg;
            if (boundary_item) {
                if (threadIdx.x == 0) slab_wait(L.slab_ctl, t, L.slab_sides, L.err);
                __syncthreads();
            }
        }
        // the row above row 0 / below row Ly - 1: the own plane (periodic) or, for a slab or a row band, the neighbours' planes
        uint8_t *tgt = plane_ptr(L, chain, COLOUR);
        const uint8_t *oth = plane_ptr(L, chain, COLOUR ^ 1);
        const uint8_t *oth_dn = SLAB ? plane_ptr_of(L.dn_planes, L, chain, COLOUR ^ 1) : oth;
        const uint8_t *oth_up = SLAB ? plane_ptr_of(L.up_planes, L, chain, COLOUR ^ 1) : oth;
        StripGeom sg;
        sg.row0 = strip * R; sg.rows = R; sg.col = seg << 4;
        sg.colL = (seg == 0 ? half : sg.col) - 1;
        sg.colR = (seg == nseg - 1) ? 0 : sg.col + 16;
        sg.loadL = (lane == 0) || (seg == 0);
        sg.loadR = (lane == 31) || (seg == nseg - 1);
        sg.active = active;
        const Acc acc = sweep_strip<COLOUR, HEATBATH, TRACK, false>(tgt, oth, oth_up, oth_dn, half, L.Ly, SLAB ? L.row_offset : 0, sg, t,
                                                                    first_chain + (uint32_t)chain, seed_lo, seed_hi, s_pair, s_thi, s_tlo);
        strip_finish<TRACK>(acc, sums, chain);
        if (SLAB && boundary_item) {
            __syncthreads();
            if (threadIdx.x == 0) slab_signal(L.slab_ctl, t, (unsigned)((2 * (int64_t)nseg + kThreads - 1) / kThreads), L.slab_sides);
        }
    }
#ifdef MCX_OPT_TRACE
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < 4096) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_trace[12 * blockIdx.x + 0] = trace_t0;
        g_trace[12 * blockIdx.x + 1] = globaltimer_ns();
        g_trace[12 * blockIdx.x + 2] = smid;
        g_trace[12 * blockIdx.x + 3] = (unsigned long long)trace_items;
    }
#endif
}

// _recompute_cached! for row-aligned 2-D Ising planes: one thread per 16-byte segment of the colour-0
// plane.  With e in {0,1}: sum_<ij> s_i s_j = sum over colour-0 sites of (2 e0 - 1)(2 nup - 4)
//   = 4 sum(e0 nup) - 2 sum(nup) - 8 sum(e0) + 4 (N/2),   sum s = 2 (sum e0 + sum e1) - N.
__global__ void __launch_bounds__(256)
k_recompute2d(LatView L, long long *__restrict__ sums, int64_t segs_per_chain)
{
    const int chain = blockIdx.y;
    const int half = L.half, nseg = half >> 4;
    const uint8_t *__restrict__ p0 = plane_ptr(L, chain, 0);
    const uint8_t *__restrict__ p1 = plane_ptr(L, chain, 1);
    long long e0n = 0, nsum = 0, e0 = 0, e1 = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < segs_per_chain;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(g / nseg), seg = (int)(g - (int64_t)row * nseg), col = seg << 4;
        const int ru = row == 0 ? L.Ly - 1 : row - 1, rd = row == L.Ly - 1 ? 0 : row + 1;
        const uint4 T = ldg128(p0 + (int64_t)row * half + col);
        const uint4 C = ldg128(p1 + (int64_t)row * half + col);
        const uint4 U = ldg128((row == 0 ? plane_ptr_of(L.up_planes, L, chain, 1) : p1) + (int64_t)ru * half + col);
        const uint4 D = ldg128((row == L.Ly - 1 ? plane_ptr_of(L.dn_planes, L, chain, 1) : p1) + (int64_t)rd * half + col);
        uint32_t S[4];
        if ((row & 1) == 0) {       // colour-0 sites of an even row sit at x = 2j: neighbours j-1, j
            const uint32_t side = p1[(int64_t)row * half + ((seg == 0 ? half : col) - 1)];
            S[0] = (C.x << 8) | side;
            S[1] = __funnelshift_l(C.x, C.y, 8); S[2] = __funnelshift_l(C.y, C.z, 8); S[3] = __funnelshift_l(C.z, C.w, 8);
        } else {
            const uint32_t side = p1[(int64_t)row * half + (seg == nseg - 1 ? 0 : col + 16)];
            S[0] = __funnelshift_r(C.x, C.y, 8); S[1] = __funnelshift_r(C.y, C.z, 8); S[2] = __funnelshift_r(C.z, C.w, 8);
            S[3] = (C.w >> 8) | (side << 24);
        }
        const uint32_t tw[4] = {T.x, T.y, T.z, T.w}, cw[4] = {C.x, C.y, C.z, C.w};
        const uint32_t nup[4] = {U.x + D.x + C.x + S[0], U.y + D.y + C.y + S[1], U.z + D.z + C.z + S[2], U.w + D.w + C.w + S[3]};
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
    e0n = warp_sum_ll(e0n); nsum = warp_sum_ll(nsum); e0 = warp_sum_ll(e0); e1 = warp_sum_ll(e1);
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
        // the constant terms (+4 N/2 and -N) are added once by block 0 / warp 0
        long long pair = 4 * e0n - 2 * nsum - 8 * e0, spin = 2 * (e0 + e1);
        if (blockIdx.x == 0 && threadIdx.x == 0) { pair += 4 * L.halfN; spin -= 2 * L.halfN; }
        atomicAdd(o + SUM_PAIR, (unsigned long long)pair);
        atomicAdd(o + SUM_SPIN, (unsigned long long)spin);
    }
}

// Layout conversion for row-aligned 2-D Ising lattices: one thread moves 32 consecutive x-sites of a
// row (two 128-bit words of host-order spins) to/from 16 bytes of each colour plane.
// Even x of row y belongs to colour y & 1, odd x to the other colour, both at plane byte x >> 1.
__global__ void __launch_bounds__(256)
k_pack2d(LatView L, const int8_t *__restrict__ staging, int64_t N, int64_t segs_per_chain)
{
    const int chain = blockIdx.y;
    const int nseg = L.half >> 4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < segs_per_chain; g += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(g / nseg), seg = (int)(g - (int64_t)row * nseg);
        const uint4 *src = reinterpret_cast<const uint4 *>(staging + (int64_t)chain * N + (int64_t)row * L.Lx + seg * 32);
        const uint4 a = src[0], b = src[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t ev[4], od[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t e0 = (~w[2 * k] >> 7) & 0x01010101u, e1 = (~w[2 * k + 1] >> 7) & 0x01010101u;   // +1 -> 1, -1 -> 0
            ev[k] = __byte_perm(e0, e1, 0x6420);
            od[k] = __byte_perm(e0, e1, 0x7531);
        }
        const int ce = row & 1;
        *reinterpret_cast<uint4 *>(plane_ptr(L, chain, ce) + (int64_t)row * L.half + seg * 16) = make_uint4(ev[0], ev[1], ev[2], ev[3]);
        *reinterpret_cast<uint4 *>(plane_ptr(L, chain, ce ^ 1) + (int64_t)row * L.half + seg * 16) = make_uint4(od[0], od[1], od[2], od[3]);
    }
}

__global__ void __launch_bounds__(256)
k_unpack2d(LatView L, int8_t *__restrict__ staging, int64_t N, int64_t segs_per_chain)
{
    const int chain = blockIdx.y;
    const int nseg = L.half >> 4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < segs_per_chain; g += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(g / nseg), seg = (int)(g - (int64_t)row * nseg);
        const int ce = row & 1;
        const uint4 e = ldg128(plane_ptr(L, chain, ce) + (int64_t)row * L.half + seg * 16);
        const uint4 o = ldg128(plane_ptr(L, chain, ce ^ 1) + (int64_t)row * L.half + seg * 16);
        const uint32_t ev[4] = {e.x, e.y, e.z, e.w}, od[4] = {o.x, o.y, o.z, o.w};
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t lo = __byte_perm(ev[k], od[k], 0x5140), hi = __byte_perm(ev[k], od[k], 0x7362);
            w[2 * k] = (lo ^ 0x01010101u) * 0xFEu + 0x01010101u;        // 1 -> 0x01 (+1), 0 -> 0xFF (-1)
            w[2 * k + 1] = (hi ^ 0x01010101u) * 0xFEu + 0x01010101u;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(staging + (int64_t)chain * N + (int64_t)row * L.Lx + seg * 32);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// init!(sys, mode; rng): the 32 sites of a thread are exactly one word of the INIT stream
__global__ void __launch_bounds__(256)
k_init2d(LatView L, int mode, uint32_t seed_lo, uint32_t seed_hi, uint32_t first_chain, int64_t segs_per_chain)
{
    const int chain = blockIdx.y;
    const int nseg = L.half >> 4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < segs_per_chain; g += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(g / nseg), seg = (int)(g - (int64_t)row * nseg);
        uint32_t bits = mode == MCX_INIT_UP ? 0xffffffffu : 0u;
        if (mode == MCX_INIT_RANDOM) {
            const int64_t i = (int64_t)(row + L.row_offset) * L.Lx + seg * 32;
            const Philox4 p = stream_block(seed_lo, seed_hi, first_chain + chain, TAG_INIT, 0, (uint32_t)(i >> 7), 0);
            const int w = (int)((i >> 5) & 3);
            bits = w == 0 ? p.x : w == 1 ? p.y : w == 2 ? p.z : p.w;
        }
        uint32_t ev[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            ev[k >> 2] |= ((bits >> (2 * k)) & 1u) << (8 * (k & 3));
            od[k >> 2] |= ((bits >> (2 * k + 1)) & 1u) << (8 * (k & 3));
        }
        const int ce = row & 1;
        *reinterpret_cast<uint4 *>(plane_ptr(L, chain, ce) + (int64_t)row * L.half + seg * 16) = make_uint4(ev[0], ev[1], ev[2], ev[3]);
        *reinterpret_cast<uint4 *>(plane_ptr(L, chain, ce ^ 1) + (int64_t)row * L.half + seg * 16) = make_uint4(od[0], od[1], od[2], od[3]);
    }
}


int pick_rows_per_strip(int Ly, int want);

// Strip height: 16 rows unless that leaves fewer than ~4 work items per resident CTA (batched small
// lattices, replicas sharded over many GPUs); then shorter strips keep the persistent grid balanced.
int auto_rows_per_strip(const mcx_lattice *lat)
{
    const int forced = knobs().rows_per_strip > 0 ? knobs().rows_per_strip : 0;
    if (forced > 0) return pick_rows_per_strip(lat->view.Ly, forced > 32 ? 32 : forced);   // the kernel flags left-over rows in 32 bits
    const int64_t nseg = lat->view.half >> 4;
    const int64_t ctas = (int64_t)lat->ctx->sm_count * 6;
    int R = 16;
    for (; R > 4; R >>= 1) {
        const int r = pick_rows_per_strip(lat->view.Ly, R);
        const int64_t items = ((int64_t)(lat->view.Ly / r) * nseg + kThreads - 1) / kThreads * lat->nchains;
        if (items >= 4 * ctas) break;
    }
    return pick_rows_per_strip(lat->view.Ly, R);
}

int pick_rows_per_strip(int Ly, int want)
{
    // largest even divisor of Ly that is <= want
    for (int r = want; r >= 2; --r)
        if ((r % 2) == 0 && (Ly % r) == 0) return r;
    return 2;
}



template <int COLOUR, bool HEATBATH, bool TRACK>
void launch_t(mcx_lattice *lat, uint64_t t)
{
    // a chain sub-range is the same launch on shifted base pointers: every per-chain array is indexed from them
    LatView L = lat->view;
    const int c0 = g_launch_range.chain0, nch = g_launch_range.nchains < 0 ? lat->nchains : g_launch_range.nchains;
    L.planes += (int64_t)c0 * 2 * L.plane_stride;
    L.up_planes += (int64_t)c0 * 2 * L.plane_stride;
    L.dn_planes += (int64_t)c0 * 2 * L.plane_stride;
    L.nchains = nch;
    // a row band is a slab of the lattice whose "neighbour slabs" are the rows around it in the same planes:
    // the SLAB kernel variant runs it unchanged (k_slab.cu explains the three base pointers)
    const bool band = g_launch_range.nrows > 0;
    if (band) {
        const int Ly = lat->view.Ly, y0 = g_launch_range.row0, nr = g_launch_range.nrows;
        const int64_t half = L.half;
        const bool top = y0 == 0, bottom = y0 + nr == Ly;
        // the band's "neighbour slabs": the rows around it in the same planes, except that the first / last band
        // of a slab of a taller lattice keeps the slab's own neighbour on that side (and orders itself against it)
        uint8_t *up = lat->slab && top ? L.up_planes + (int64_t)(Ly - nr) * half
                                       : L.planes + ((int64_t)((y0 - 1 + Ly) % Ly) - (nr - 1)) * half;   // its row nr - 1 is row y0 - 1
        uint8_t *dn = lat->slab && bottom ? L.dn_planes : L.planes + (int64_t)((y0 + nr) % Ly) * half;   // its row 0 is row y0 + nr
        L.up_planes = up; L.dn_planes = dn;
        L.planes += (int64_t)y0 * half;
        L.Ly = nr;
        L.row_offset += y0;
        L.slab_sides = (top ? 1 : 0) | (bottom ? 2 : 0);
        if (!L.slab_sides) L.slab_ctl = nullptr;
    }
    cudaStream_t stream = g_launch_range.use_stream ? g_launch_range.stream : lat->ctx->stream;
    const int R = band ? g_launch_range.R : auto_rows_per_strip(lat);
    const int nstrips = L.Ly / R;
    const int nseg = L.half >> 4;
    const int64_t G = (int64_t)nstrips * nseg;
    // G < 2^31 always: 2^31 thread-items of 16 x R >= 32 bytes per plane are beyond the memory of the device
    const int blocks_per_chain = (int)((G + kThreads - 1) / kThreads);
    const int nitems = (int)((int64_t)blocks_per_chain * nch);
    // the shipped variant also exists with the idle-lane predicate compiled out (+2.8 %, r01_tune_allactive.log)
    // and with the halo rows taken from the neighbour slabs (k_slab.cu)
    const bool slab = lat->slab != nullptr || band;
    const bool full = G % kThreads == 0 && knobs().full != 0;
    auto kern = slab ? (full ? k_ising2d<COLOUR, HEATBATH, TRACK, true, true> : k_ising2d<COLOUR, HEATBATH, TRACK, false, true>)
                     : (full ? k_ising2d<COLOUR, HEATBATH, TRACK, true, false> : k_ising2d<COLOUR, HEATBATH, TRACK, false, false>);
    static thread_local int resident = 0;
    if (!resident) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_ising2d<COLOUR, HEATBATH, TRACK, false, false>, kThreads, 0);
        if (resident < 1) resident = 1;
    }
    const int ctas_per_sm = knobs().ctas_per_sm >= 0 ? knobs().ctas_per_sm : resident;
    // persistent grid: every SM gets its full complement of CTAs (trimming the grid so that all CTAs
    // run the same number of items was measured 8 % slower: SMs with fewer CTAs do not finish sooner)
    int grid = lat->ctx->sm_count * ctas_per_sm;
    if (grid > nitems) grid = nitems;
    kern<<<grid, kThreads, 0, stream>>>(L, lat->d_thi, lat->d_tlo, lat->d_labels + c0, lat->d_sums + (int64_t)c0 * SUM_FIELDS,
                                       (uint32_t)lat->seed, (uint32_t)(lat->seed >> 32), t, lat->first_chain + (uint32_t)c0, R,
                                       nstrips, blocks_per_chain, nitems, g_t_clock);
    lat->ctx->launches++;
}

template <int COLOUR>
void launch_c(mcx_lattice *lat, uint64_t t)
{
    if (lat->storage == MCX_STORAGE_BIT) { launch_half_sweep_bits2d(lat, COLOUR, t); return; }   // k_bits.cu, same launch ranges
    const bool track = lat->track_sums;
    if (lat->rule == MCX_HEATBATH) {
        if (track) launch_t<COLOUR, true, true>(lat, t); else launch_t<COLOUR, true, false>(lat, t);
    } else {
        if (track) launch_t<COLOUR, false, true>(lat, t); else launch_t<COLOUR, false, false>(lat, t);
    }
}

}  // namespace

static dim3 seg_grid(const mcx_lattice *lat, int64_t segs)
{
    int64_t blocks = (segs + 255) / 256;
    const int64_t cap = (int64_t)lat->ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    return dim3((unsigned)blocks, (unsigned)lat->nchains);
}

static bool is_fast2d(const mcx_lattice *lat)
{
    return lat->fast2d && lat->model == MCX_ISING && lat->storage == MCX_STORAGE_INT8;
}

bool launch_pack_ising2d(mcx_lattice *lat)
{
    if (!is_fast2d(lat)) return false;
    const int64_t segs = (int64_t)lat->view.Ly * (lat->view.half >> 4);
    k_pack2d<<<seg_grid(lat, segs), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_staging, lat->N, segs);
    lat->ctx->launches++;
    return true;
}

bool launch_unpack_ising2d(mcx_lattice *lat)
{
    if (!is_fast2d(lat)) return false;
    const int64_t segs = (int64_t)lat->view.Ly * (lat->view.half >> 4);
    k_unpack2d<<<seg_grid(lat, segs), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_staging, lat->N, segs);
    lat->ctx->launches++;
    return true;
}

bool launch_init_ising2d(mcx_lattice *lat, int mode, uint64_t seed)
{
    if (!is_fast2d(lat)) return false;
    const int64_t segs = (int64_t)lat->view.Ly * (lat->view.half >> 4);
    k_init2d<<<seg_grid(lat, segs), 256, 0, lat->ctx->stream>>>(lat->view, mode, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                lat->first_chain, segs);
    lat->ctx->launches++;
    return true;
}

bool launch_recompute_ising2d(mcx_lattice *lat)
{
    if (!lat->fast2d || lat->model != MCX_ISING || lat->storage != MCX_STORAGE_INT8) return false;
    const int64_t segs = (int64_t)lat->view.Ly * (lat->view.half >> 4);
    int64_t blocks = (segs + 255) / 256;
    const int64_t cap = (int64_t)lat->ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_recompute2d<<<dim3((unsigned)blocks, (unsigned)lat->nchains), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_sums, segs);
    lat->ctx->launches++;
    return true;
}

static bool aux_streams(mcx_ctx *ctx)
{
    if (ctx->aux_ready) return true;
    for (int i = 0; i < 16; ++i) {
        if (cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->aux_join[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
    }
    if (cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
    ctx->aux_ready = true;
    return true;
}

bool launch_sweeps_ising2d_grouped(mcx_lattice *lat, int64_t nsweeps)
{
    if (lat->slab) return false;
    // which vectorised half-sweep serves this lattice (each launcher re-checks its own conditions)
    const bool d3 = lat->ndim == 3 && lat->model == MCX_ISING && lat->view.Lx % 32 == 0 && knobs().ising3d != 0;
    const bool bc = lat->fast2d && lat->model == MCX_BLUME_CAPEL;
    if (!d3 && !lat->fast2d) return false;
    if (bc && knobs().bc2d == 0) return false;
    if (knobs().variant >= 0 || knobs().rows_per_strip >= 0) return false;
    const int groups_env = knobs().groups;
    if (groups_env == 0 || groups_env == 1) return false;
    const int R = d3 ? 16 : auto_rows_per_strip(lat);
    const int64_t G = (int64_t)((lat->view.Ly + R - 1) / R) * (d3 ? lat->view.Lz : 1) * (lat->view.half >> 4);
    if (G < 96) return false;                                          // rows-of-8 territory
    const int64_t items = (G + kThreads - 1) / kThreads * lat->nchains;
    const int64_t ctas = (int64_t)lat->ctx->sm_count * 6;
    int groups = groups_env > 1 ? groups_env : 4;
    if (groups > 8) groups = 8;
    if (groups > lat->nchains) groups = lat->nchains;
    if (groups_env < 0)
        while (groups > 1 && items / groups < ctas / 4) --groups;     // a group's launch should still fill a good part of the SMs
    if (groups < 2) return false;
    mcx_ctx *ctx = lat->ctx;
    if (!aux_streams(ctx)) return false;
    cudaEventRecord(ctx->aux_fork, ctx->stream);
    for (int g = 0; g < groups; ++g) cudaStreamWaitEvent(ctx->aux[g], ctx->aux_fork, 0);
    // group-major would serialise the groups on the host side; interleave so that their launches alternate
    for (int64_t s = 0; s < nsweeps; ++s)
        for (int colour = 0; colour < 2; ++colour)
            for (int g = 0; g < groups; ++g) {
                const int c0 = (int)((int64_t)lat->nchains * g / groups), c1 = (int)((int64_t)lat->nchains * (g + 1) / groups);
                g_launch_range.chain0 = c0; g_launch_range.nchains = c1 - c0; g_launch_range.stream = ctx->aux[g]; g_launch_range.use_stream = true;
                const uint64_t t = 2 * (lat->sweep + (uint64_t)s) + (uint64_t)colour;
                if (d3) launch_sweep_ising3d(lat, colour, t);
                else if (bc) launch_sweep_bc2d(lat, colour, t);
                else if (colour == 0) launch_c<0>(lat, t); else launch_c<1>(lat, t);
            }
    g_launch_range = LaunchRange();
    for (int g = 0; g < groups; ++g) {
        cudaEventRecord(ctx->aux_join[g], ctx->aux[g]);
        cudaStreamWaitEvent(ctx->stream, ctx->aux_join[g], 0);
    }
    return true;
}

bool launch_sweeps_ising2d_banded(mcx_lattice *lat, int64_t nsweeps)
{
    // one Ising lattice: 2-D in bands of rows, 3-D in bands of z-planes (k_ising3d.cu runs a z-range per launch)
    const bool d3 = lat->ndim == 3 && lat->model == MCX_ISING && lat->view.Lx % 32 == 0 && lat->table_len == 14 && knobs().ising3d != 0 &&
                    !lat->slab;
    if ((!d3 && (!lat->fast2d || lat->model != MCX_ISING)) || lat->nchains != 1) return false;
    if (lat->slab && !(lat->slab->attached && lat->slab->remote)) return false;      // in-process slabs advance in lockstep
    if (knobs().variant >= 0 || knobs().rows_per_strip >= 0) return false;
    const int bands_env = knobs().bands;
    if (bands_env == 0 || bands_env == 1) return false;
    // 8 bands (the hardware queues a process gets by default) of 16-row strips, each at least ~100 CTA items:
    // measured +8 % at L = 8192, +17 % at 16384, +10 % at 32768; 2 bands gain nothing (each waits for the other),
    // 4 bands gain less, shorter strips or smaller bands lose (profiles/r01_bands_groups.md).  MCX_BANDS forces a count.
    const int Ly = d3 ? lat->view.Lz : lat->view.Ly;      // the banded dimension
    const bool forced = bands_env > 1;
    int bands = forced ? (bands_env > 16 ? 16 : bands_env) : 8;
    int R = knobs().band_rows > 0 ? knobs().band_rows : 16;            // MCX_BAND_ROWS: tuning hook
    if (!d3 && !forced && knobs().band_rows <= 0 && Ly % (32 * 16) == 0 &&
        ((int64_t)(Ly / 16 / 32) * (lat->view.half >> 4) + kThreads - 1) / kThreads >= 100) {
        // big lattices at 8 CTAs/SM: 16 bands of 32-row strips (a CTA's set-up is spread over twice the rows, and twice
        // the launches fill each other's tails): 1643 against 1601 attempts/ns at L = 16384 (profiles/r02_launch_shapes.md)
        bands = 16; R = 32;
    }
    if (d3) {
        // a band must leave its neighbours a plane of their own: at least two planes per band
        if (Ly % bands != 0 || Ly / bands < 2 || lat->view.Ly % 2 != 0) return false;
        if (lat->storage != MCX_STORAGE_BIT && (int64_t)(lat->view.Ly / 2) * lat->view.Lz * (lat->view.half >> 4) < 96) return false;   // rows-of-8 territory
        const int64_t items_per_band = ((int64_t)(Ly / bands) * (lat->view.Ly / 4) * (lat->view.half >> 4) + kThreads - 1) / kThreads;
        if (!forced && items_per_band < 512) return false;      // 512^3: +16 %; 256^3 (128 items per band) loses to the launch rate: profiles/r02_bands3d.md
    } else {
        if (Ly % (R * bands) != 0) return false;
        const int64_t items_per_band = ((int64_t)(Ly / bands / R) * (lat->view.half >> 4) + kThreads - 1) / kThreads;
        if (!forced && items_per_band < 100) bands = 0;
    }
    if (!bands) return false;
    const int nr = Ly / bands;
    mcx_ctx *ctx = lat->ctx;
    if (!aux_streams(ctx)) return false;
    cudaEventRecord(ctx->aux_fork, ctx->stream);
    for (int b = 0; b < bands; ++b) cudaStreamWaitEvent(ctx->aux[b], ctx->aux_fork, 0);
    for (int64_t s = 0; s < nsweeps; ++s)
        for (int colour = 0; colour < 2; ++colour) {
            const bool first = s == 0 && colour == 0;
            // half-sweep h of band b needs half-sweep h - 1 of bands b - 1, b, b + 1 (b: stream order).  All waits of
            // this half-sweep are queued before any of its records, so they refer to the previous half-sweep's.
            if (!first)
                for (int b = 0; b < bands; ++b) {
                    cudaStreamWaitEvent(ctx->aux[b], ctx->aux_join[(b + bands - 1) % bands], 0);
                    cudaStreamWaitEvent(ctx->aux[b], ctx->aux_join[(b + 1) % bands], 0);
                }
            for (int b = 0; b < bands; ++b) {
                g_launch_range = LaunchRange();
                g_launch_range.row0 = b * nr; g_launch_range.nrows = nr; g_launch_range.R = R; g_launch_range.stream = ctx->aux[b]; g_launch_range.use_stream = true;
                const uint64_t t = 2 * (lat->sweep + (uint64_t)s) + (uint64_t)colour;
                if (d3) launch_sweep_ising3d(lat, colour, t);
                else if (colour == 0) launch_c<0>(lat, t); else launch_c<1>(lat, t);
                cudaEventRecord(ctx->aux_join[b], ctx->aux[b]);
            }
        }
    g_launch_range = LaunchRange();
    for (int b = 0; b < bands; ++b) cudaStreamWaitEvent(ctx->stream, ctx->aux_join[b], 0);
    return true;
}

__global__ void k_tclock_set(unsigned long long *clock, unsigned long long v) { *clock = v; }
__global__ void k_tclock_add(unsigned long long *clock, unsigned long long dv) { *clock += dv; }

// A long series of sweeps of one big lattice costs the host 64 - 128 stream operations per sweep (launches, event records and
// waits of the row bands) -- as long as a sweep takes on the device, and worse when eight ranks share a host.  The band
// launches of 32 sweeps are therefore captured once into a CUDA graph whose kernels add a device clock to their
// half-sweep index (the last node advances it), and a series is that graph replayed.  Same launches, same trajectories.
int64_t launch_sweeps_ising2d_banded_graph(mcx_lattice *lat, int64_t nsweeps)
{
    // sweeps per replay: a replay ends with all bands joined, which costs the overlap of one half-sweep's tail -- 3 % of the
    // rate at 4 sweeps per replay (1650 against 1704 attempts/ns at L = 16384), < 0.5 % at 32; MCX_SWEEP_GRAPH=n sets it
    const Knobs &k = knobs();
    const int kSweeps = k.sweep_graph > 1 ? (k.sweep_graph > 256 ? 256 : k.sweep_graph) : 32;
    if (k.sweep_graph == 0 || nsweeps < 2 * kSweeps + 1 || lat->sweep_graph_K < 0) return 0;   // < 0: this context's stream cannot be captured
    if (!lat->fast2d || lat->model != MCX_ISING || lat->storage != MCX_STORAGE_INT8 || lat->slab || lat->nchains != 1) return 0;
    if (k.variant >= 0 || k.rows_per_strip >= 0 || k.force_generic > 0 || k.bands == 0 || k.bands == 1) return 0;
    mcx_ctx *ctx = lat->ctx;
    if (!lat->d_tclock && cudaMalloc((void **)&lat->d_tclock, sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        lat->d_tclock = nullptr;
        return 0;
    }
    mcx_lattice::SweepGraphKey key;                             // what the captured launch arguments depend on
    memset(&key, 0, sizeof(key));
    key.seed = lat->seed; key.first_chain = lat->first_chain; key.rule = lat->rule; key.track = lat->track_sums ? 1 : 0;
    key.bands = k.bands; key.band_rows = k.band_rows * 1024 + kSweeps;
    key.planes = lat->view.planes; key.thi = lat->d_thi; key.labels = lat->d_labels; key.sums = lat->d_sums;
    if (lat->sweep_graph && memcmp(&key, &lat->sweep_graph_key, sizeof(key)) != 0) {
        cudaGraphExecDestroy(lat->sweep_graph);
        lat->sweep_graph = nullptr;
    }
    // `done` sweeps have been queued by this call; lat->sweep itself is advanced by the caller with the return value
    int64_t done = 0;
    if (!lat->sweep_graph) {
        // one sweep launch by launch first: the auxiliary streams and events exist before the capture
        if (!launch_sweeps_ising2d_banded(lat, 1)) return 0;
        done = 1; nsweeps -= 1;
        const uint64_t launches0 = ctx->launches, sweep0 = lat->sweep;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            lat->sweep_graph_K = -1;                           // e.g. the legacy default stream: do not try again on this handle
            return done;
        }
        lat->sweep = 0;                                        // captured launch arguments are relative to the clock
        g_t_clock = lat->d_tclock;
        const bool ok = launch_sweeps_ising2d_banded(lat, kSweeps);
        g_t_clock = nullptr;
        lat->sweep = sweep0;
        if (ok) k_tclock_add<<<1, 1, 0, ctx->stream>>>(lat->d_tclock, 2ull * kSweeps);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        lat->sweep_graph_launches = ctx->launches - launches0 + 1;
        ctx->launches = launches0;                             // captured, not executed
        if (!ok || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return done;
        }
        const cudaError_t ei = cudaGraphInstantiate(&lat->sweep_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { cudaGetLastError(); lat->sweep_graph = nullptr; return done; }
        lat->sweep_graph_key = key;
        lat->sweep_graph_K = kSweeps;
    }
    const int64_t replays = nsweeps / lat->sweep_graph_K;
    if (replays < 1) return done;
    k_tclock_set<<<1, 1, 0, ctx->stream>>>(lat->d_tclock, 2 * (lat->sweep + (uint64_t)done));
    ctx->launches++;
    for (int64_t g = 0; g < replays; ++g) {
        if (cudaGraphLaunch(lat->sweep_graph, ctx->stream) != cudaSuccess) { cudaGetLastError(); return done + g * lat->sweep_graph_K; }
        ctx->launches += lat->sweep_graph_launches;
    }
    return done + replays * lat->sweep_graph_K;
}

bool launch_sweep_ising2d(mcx_lattice *lat, int colour, uint64_t t)
{
    if (!lat->fast2d || lat->model != MCX_ISING) return false;
    if (lat->storage == MCX_STORAGE_BIT) { launch_half_sweep_bits2d(lat, colour, t); return true; }   // the only 2-D kernel on bit planes
    // small lattices: a chain yields fewer than one CTA of 16-byte segments x strips, so most lanes of
    // this kernel would idle; the rows-of-8 kernel (8 sites per thread) fills the machine instead
    {
        const int R = pick_rows_per_strip(lat->view.Ly, knobs().rows_per_strip >= 0 ? (knobs().rows_per_strip > 32 ? 32 : knobs().rows_per_strip) : 16);   // small-lattice test uses 16
        const int64_t G = (int64_t)(lat->view.Ly / R) * (lat->view.half >> 4);
        if (G < 96 && !lat->slab && knobs().variant < 0 && knobs().rows_per_strip < 0) return false;
    }
    if (colour == 0) launch_c<0>(lat, t); else launch_c<1>(lat, t);
    return true;
}

}  // namespace mcx

#ifdef MCX_OPT_TRACE
extern "C" int mcx_debug_trace(unsigned long long *out, int nctas)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, mcx::g_trace, sizeof(unsigned long long) * 12 * (size_t)nctas);
}
#endif
