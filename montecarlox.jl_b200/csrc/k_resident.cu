// k_resident.cu -- whole sweep series of small 2-D Ising lattices inside shared memory.
//
// A lattice of up to ~1.5 MB (both colour planes) is split by rows over the CTAs of ONE thread-block
// cluster (1, 2, 4 or 8 CTAs = SMs) and stays in their shared memory for all nsweeps sweeps of a
// mcx_sweep call: HBM is touched once to load and once to store it.  Each CTA keeps its rows of both
// planes plus one halo row above and below per plane; after a thread updates a row on the CTA's edge it
// also stores it into the neighbour CTA's halo row through distributed shared memory, and one cluster
// barrier per half-sweep publishes everything.  No kernel-launch boundary, no grid-wide tail: this is
// what parallel tempering needs when a GPU holds only a few dozen replicas (strong scaling of
// BASELINE.json configs[2]; profiles/r01_cta_timeline.md shows the ~7 us per-launch tail it removes).
//
// The per-row work is the same update_row as the streaming kernel (k_row16.cuh), with the same
// positioned Philox counters, so trajectories are bit-identical to k_ising2d and k_sweep_generic.
#include "k_row16.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace mcx {

namespace {

constexpr int kResMinWork = 256;                        // thread-rows a CTA must have per half-sweep to be worth it
constexpr int64_t kResMaxSweepsPerLaunch = 1 << 14;     // keeps the per-thread int32 accumulators far from overflow


__device__ __forceinline__ uint4 lds128(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }

// One colour of one sweep over this CTA's rows.  s_tgt / s_oth: plane bases in shared memory, local row
// lr in [-1, rows_cta] at byte offset (lr + 1) * half.  rem_up / rem_dn: the halo rows of plane COLOUR in
// the CTAs holding the rows above / below (generic pointers into distributed shared memory).
template <int COLOUR, bool HEATBATH, bool TRACK, int NT>
__device__ __forceinline__ void resident_half_sweep(uint8_t *s_tgt, const uint8_t *s_oth, uint8_t *rem_up, uint8_t *rem_dn,
                                                    const int half, const int nseg, const int rows_cta, const int R,
                                                    const int G, const int row_begin, const uint64_t t,
                                                    const uint32_t chain_id, const uint32_t seed_lo, const uint32_t seed_hi,
                                                    const uint32_t *s_pair, const uint32_t *s_thi, const uint32_t *s_tlo,
                                                    Acc &acc)
{
    const int lane = threadIdx.x & 31;
    const uint32_t t_lo = (uint32_t)t;
    const uint32_t c2 = ctr_word2(t, 0, TAG_SWEEP);
    const uint32_t c2lo = ctr_word2(t, 1, TAG_SWEEP);
    const int Gpad = (G + 31) & ~31;                              // whole warps enter the loop (shuffles)

    for (int g0 = threadIdx.x; g0 < Gpad; g0 += NT) {
        const bool active = g0 < G;
        const int g = active ? g0 : G - 1;
        const int strip = g / nseg;
        const int seg = g - strip * nseg;
        const int lr0 = strip * R;                                // even local row
        const int col = seg << 4;
        const int colL = (seg == 0 ? half : col) - 1;
        const int colR = (seg == nseg - 1) ? 0 : col + 16;
        const bool loadL = (lane == 0) || (seg == 0);
        const bool loadR = (lane == 31) || (seg == nseg - 1);
        const bool edgeA = COLOUR == 0 ? loadL : loadR;
        const bool edgeB = COLOUR == 0 ? loadR : loadL;
        const int colA = COLOUR == 0 ? colL : colR;
        const int colB = COLOUR == 0 ? colR : colL;

        const uint8_t *po = s_oth + (lr0 + 1) * half;             // other plane, current even row
        uint8_t *pt = s_tgt + (lr0 + 1) * half + col;             // target plane, current even row
        uint4 U = lds128(po - half + col);
        uint4 C = lds128(po + col);
        uint32_t blk = (uint32_t)(((int64_t)(row_begin + lr0) * half + col) >> 3);
        const uint32_t blk_step = (uint32_t)(half >> 3);

#pragma unroll 1
        for (int r = 0; r < R; r += 2) {
            const uint4 D = lds128(po + half + col);
            const uint4 E = lds128(po + 2 * half + col);
            const uint4 Ta = lds128(pt), Tb = lds128(pt + half);
            uint32_t sA, sB;
            if (COLOUR == 0) {
                sA = __shfl_up_sync(0xffffffffu, C.w, 1) >> 24;
                sB = __shfl_down_sync(0xffffffffu, D.x, 1) & 0xffu;
            } else {
                sA = __shfl_down_sync(0xffffffffu, C.x, 1) & 0xffu;
                sB = __shfl_up_sync(0xffffffffu, D.w, 1) >> 24;
            }
            if (edgeA) sA = po[colA];
            if (edgeB) sB = po[half + colB];
            const uint4 Na = update_row<COLOUR, HEATBATH, TRACK>(Ta, U, C, D, sA, blk, t_lo, c2, c2lo, chain_id, seed_lo,
                                                                 seed_hi, s_pair, s_thi, s_tlo, acc, active);
            if (active) {
                *reinterpret_cast<uint4 *>(pt) = Na;
                if (lr0 + r == 0) *reinterpret_cast<uint4 *>(rem_up + col) = Na;              // my first row
            }
            const uint4 Nb = update_row<COLOUR ^ 1, HEATBATH, TRACK>(Tb, C, D, E, sB, blk + blk_step, t_lo, c2, c2lo, chain_id,
                                                                     seed_lo, seed_hi, s_pair, s_thi, s_tlo, acc, active);
            if (active) {
                *reinterpret_cast<uint4 *>(pt + half) = Nb;
                if (lr0 + r + 2 == rows_cta) *reinterpret_cast<uint4 *>(rem_dn + col) = Nb;   // my last row
            }
            U = D; C = E;
            po += 2 * half; pt += 2 * half; blk += 2 * blk_step;
        }
    }
}

extern __shared__ __align__(16) uint8_t s_dyn[];

template <bool HEATBATH, bool TRACK, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
k_ising2d_resident(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
                   const int32_t *__restrict__ labels, long long *__restrict__ sums, uint32_t seed_lo, uint32_t seed_hi,
                   uint64_t t0, int nsweeps, uint32_t first_chain, int rows_cta, int R)
{
    __shared__ uint32_t s_pair[kPairWords];
    __shared__ uint32_t s_thi[kTableLen], s_tlo[kTableLen];

    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int nclusters = gridDim.x / csize;
    const int cluster_id = blockIdx.x / csize;

    const int half = L.half;
    const int nseg = half >> 4;
    const int plane_bytes = (rows_cta + 2) * half;
    const int G = (rows_cta / R) * nseg;
    const int row_begin = rank * rows_cta;
    const int lane = threadIdx.x & 31;

    uint8_t *s_plane[2] = {s_dyn, s_dyn + plane_bytes};
    // halo rows of the neighbours: the CTA above (rank - 1) gets my first row as ITS bottom halo (local row
    // rows_cta), the CTA below (rank + 1) gets my last row as ITS top halo (local row -1).  Periodic.
    uint8_t *up_base = cluster.map_shared_rank(s_dyn, (rank + csize - 1) % csize);
    uint8_t *dn_base = cluster.map_shared_rank(s_dyn, (rank + 1) % csize);
    uint8_t *rem_up[2] = {up_base + (rows_cta + 1) * half, up_base + plane_bytes + (rows_cta + 1) * half};
    uint8_t *rem_dn[2] = {dn_base, dn_base + plane_bytes};

    const int segs_all = (rows_cta + 2) * nseg;                   // rows -1 .. rows_cta
    const int segs_own = rows_cta * nseg;

    for (int chain = cluster_id; chain < L.nchains; chain += nclusters) {
        const int label = labels[chain];
        __syncthreads();
        load_pair_table(s_pair, s_thi, s_tlo, thi_g, tlo_g, label);
        // rows row_begin - 1 .. row_begin + rows_cta of both planes (halo rows wrap around the lattice)
        for (int p = 0; p < 2; ++p) {
            const uint8_t *src = plane_ptr(L, chain, p);
            for (int i = threadIdx.x; i < segs_all; i += NT) {
                const int lr = i / nseg - 1, seg = i - (lr + 1) * nseg;
                int row = row_begin + lr;
                row = row < 0 ? row + L.Ly : (row >= L.Ly ? row - L.Ly : row);
                *reinterpret_cast<uint4 *>(s_plane[p] + (lr + 1) * half + (seg << 4)) =
                    *reinterpret_cast<const uint4 *>(src + (int64_t)row * half + (seg << 4));
            }
        }
        cluster.sync();                                           // every CTA of the cluster holds its rows

        const uint32_t chain_id = first_chain + (uint32_t)chain;
        Acc acc;
        for (int s = 0; s < nsweeps; ++s) {
            const uint64_t t = t0 + 2 * (uint64_t)s;
            resident_half_sweep<0, HEATBATH, TRACK, NT>(s_plane[0], s_plane[1], rem_up[0], rem_dn[0], half, nseg, rows_cta, R, G,
                                                    row_begin, t, chain_id, seed_lo, seed_hi, s_pair, s_thi, s_tlo, acc);
            cluster.sync();
            resident_half_sweep<1, HEATBATH, TRACK, NT>(s_plane[1], s_plane[0], rem_up[1], rem_dn[1], half, nseg, rows_cta, R, G,
                                                    row_begin, t + 1, chain_id, seed_lo, seed_hi, s_pair, s_thi, s_tlo, acc);
            cluster.sync();
        }

        for (int p = 0; p < 2; ++p) {
            uint8_t *dst = plane_ptr(L, chain, p);
            for (int i = threadIdx.x; i < segs_own; i += NT) {
                const int lr = i / nseg, seg = i - lr * nseg;
                *reinterpret_cast<uint4 *>(dst + (int64_t)(row_begin + lr) * half + (seg << 4)) =
                    *reinterpret_cast<const uint4 *>(s_plane[p] + (lr + 1) * half + (seg << 4));
            }
        }
        // per-chain sums: dspin = 2 - 4 s, dpair = -8 s nup + 16 s + 4 nup - 8 per changed site
        const long long nflip = warp_sum_ll((long long)acc.flips);
        long long dspin = 0, dpair = 0;
        if (TRACK) {
            const long long ss = warp_sum_ll(acc.s), nn_ = warp_sum_ll(acc.n), sn = warp_sum_ll(acc.sn);
            dspin = 2 * nflip - 4 * ss;
            dpair = -8 * sn + 16 * ss + 4 * nn_ - 8 * nflip;
        }
        if (lane == 0) {
            unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
            if (nflip) atomicAdd(o + SUM_ACC, (unsigned long long)nflip);
            if (TRACK) {
                if (dpair) atomicAdd(o + SUM_PAIR, (unsigned long long)dpair);
                if (dspin) atomicAdd(o + SUM_SPIN, (unsigned long long)dspin);
            }
        }
    }
    cluster.sync();      // no CTA leaves while a neighbour may still store into its shared memory
}

struct ResidentPlan {
    int csize, rows_cta, R, nclusters;
    size_t smem;
};

template <bool HEATBATH, bool TRACK, int NT>
bool plan_and_launch(mcx_lattice *lat, int64_t nsweeps, bool dry_run)
{
    const LatView &L = lat->view;
    const int half = L.half, nseg = half >> 4;
    auto kern = k_ising2d_resident<HEATBATH, TRACK, NT>;
    static thread_local bool attr_set = false;
    const size_t smem_cap = 200 * 1024;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set = true;
    }
    // candidates: cluster sizes whose row share is even, fits shared memory and gives a CTA enough work;
    // take the smallest, then widen while the batch cannot fill the SMs
    ResidentPlan plan{0, 0, 0, 0, 0};
    const int forced = knobs().resident_cluster > 0 ? knobs().resident_cluster : 0;
    for (int c = 1; c <= 8; c *= 2) {
        if (L.Ly % (2 * c) != 0) break;
        const int rows = L.Ly / c;
        const size_t smem = 2 * (size_t)(rows + 2) * half;
        if (smem > smem_cap) continue;
        if (forced) {
            if (c != forced) continue;
        } else {
            if (plan.csize && ((rows / 2) * nseg < kResMinWork ||                       // too little work per CTA
                               (int64_t)lat->nchains * plan.csize >= lat->ctx->sm_count))   // SMs already busy
                break;
        }
        plan.csize = c; plan.rows_cta = rows; plan.smem = smem;
    }
    if (!plan.csize) return false;
    // strip height: the largest one that deals the CTA's thread-rows out in whole rounds of NT threads
    plan.R = 2;
    for (int r = 16; r >= 2; r -= 2)
        if (plan.rows_cta % r == 0 && ((plan.rows_cta / r) * nseg) % NT == 0) { plan.R = r; break; }
    if (const int fr = knobs().resident_rows > 0 ? knobs().resident_rows : 0) if (fr % 2 == 0 && plan.rows_cta % fr == 0) plan.R = fr;

    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)plan.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = lat->ctx->stream;
    cfg.gridDim = dim3((unsigned)plan.csize);                     // placeholder for the occupancy query
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return false;
    }
    plan.nclusters = lat->nchains < max_clusters ? lat->nchains : max_clusters;
    // 8-CTA clusters (lattices near 1 MiB) leave SMs idle (14 such clusters fit a B200) and run the batch in
    // waves: measured slower than the streaming kernel unless the batch is one partial wave of >= 2 lattices
    if (!forced && knobs().resident != 1 && plan.csize == 8 &&
        (lat->nchains < 2 || lat->nchains > max_clusters))
        return false;
    if (dry_run) return true;
    cfg.gridDim = dim3((unsigned)(plan.nclusters * plan.csize));
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, L, (const uint32_t *)lat->d_thi, (const uint32_t *)lat->d_tlo,
                                             (const int32_t *)lat->d_labels, lat->d_sums, (uint32_t)lat->seed,
                                             (uint32_t)(lat->seed >> 32), (uint64_t)(2 * lat->sweep), (int)nsweeps,
                                             lat->first_chain, plan.rows_cta, plan.R);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    lat->ctx->launches++;
    return true;
}

}  // namespace

// Runs nsweeps whole sweeps in one launch if the lattice qualifies (2-D Ising, int8, Lx % 32 == 0, fits the
// shared memory of a cluster); returns false (nothing launched) otherwise.  MCX_RESIDENT=0 disables it,
// =1 forces it whenever it fits.  Default policy (scripts/bench_small.py, profiles/r01_resident_small.md):
// series of >= 2 sweeps over batches of at most 32 Mi sites -- there the streaming kernel is bound by
// launch latency and by the tail of its persistent grid, and this kernel is 1.2-4x faster; larger batches
// stream faster (1200 vs ~850 attempts/ns).
bool launch_sweeps_resident(mcx_lattice *lat, int64_t nsweeps)
{
    if (!lat->fast2d || lat->model != MCX_ISING || lat->storage != MCX_STORAGE_INT8 || lat->slab) return false;
    if (nsweeps < 1 || nsweeps > kResMaxSweepsPerLaunch) return false;
    const int mode = knobs().resident;
    if (mode == 0) return false;
    // the tuning hooks of the streaming kernel select that kernel
    if (mode != 1 && (knobs().variant >= 0 || knobs().rows_per_strip >= 0 || knobs().full >= 0)) return false;
    if (mode != 1) {
        const int64_t kResidentMaxSites = (int64_t)32 << 20;
        if (nsweeps < 2 || (int64_t)lat->nchains * lat->N > kResidentMaxSites) return false;
    }
    const bool heatbath = lat->rule == MCX_HEATBATH, track = lat->track_sums;
    // CTA width: enough threads for the thread-rows one CTA can have in flight (16-byte segments x row pairs
    // of the smallest cluster share), so that tiny lattices run many narrow CTAs per SM instead of one wide one
    int64_t work = (int64_t)(lat->view.Ly / 2) * (lat->view.half >> 4);
    const int nt = knobs().resident_threads > 0 ? knobs().resident_threads : (work <= 128 ? 128 : work <= 256 ? 256 : 512);
#define MCX_RES_DISPATCH(NT)                                                                                     \
    (heatbath ? (track ? plan_and_launch<true, true, NT>(lat, nsweeps, false) : plan_and_launch<true, false, NT>(lat, nsweeps, false)) \
              : (track ? plan_and_launch<false, true, NT>(lat, nsweeps, false) : plan_and_launch<false, false, NT>(lat, nsweeps, false)))
    if (nt == 128) return MCX_RES_DISPATCH(128);
    if (nt == 256) return MCX_RES_DISPATCH(256);
    return MCX_RES_DISPATCH(512);
#undef MCX_RES_DISPATCH
}

}  // namespace mcx
