// k_rows8.cu -- checkerboard half-sweep for any model / rule / 2-D or 3-D lattice whose rows hold a
// multiple of eight plane bytes (Lx % 16 == 0): one thread per Philox block = eight consecutive
// slots of one row, moved with 64-bit loads and stores.  The neighbour rows (y +- 1, z +- 1 and the
// same row of the other colour plane) are read as 64-bit words too and summed byte-wise; only the
// one in-row neighbour byte outside the group is a byte load.  Same table-driven rule as
// k_sweep_generic (k_site.cuh), same trajectories; used for Blume-Capel and 3-D lattices, and for
// 2-D Ising rows that are not a multiple of 32 sites.  Rule tables live in shared memory.
#include "k_site.cuh"
#include "mcx_internal.h"

namespace mcx {

namespace {

constexpr int kThreads8 = 128;
constexpr int kMaxTable = 80;   // Blume-Capel Metropolis in 3-D: 6 * 13 = 78 entries

__device__ __forceinline__ uint64_t ld64(const uint8_t *p) { return *reinterpret_cast<const uint64_t *>(p); }

template <int MODEL, int RULE, int NDIM>
__global__ void __launch_bounds__(kThreads8)
k_sweep_rows8(LatView L, const uint32_t *__restrict__ thi_g, const uint32_t *__restrict__ tlo_g,
              const int32_t *__restrict__ labels, int table_len, long long *__restrict__ sums, uint32_t seed_lo,
              uint32_t seed_hi, uint64_t t, int colour, uint32_t first_chain, int64_t nblk)
{
    __shared__ uint32_t s_thi[kMaxTable], s_tlo[kMaxTable];
    const int chain = blockIdx.y;
    const int tb = labels[chain] * table_len;
    for (int i = threadIdx.x; i < table_len; i += blockDim.x) {
        s_thi[i] = thi_g[tb + i];
        s_tlo[i] = tlo_g[tb + i];
    }
    __syncthreads();
    const uint32_t chain_id = first_chain + (uint32_t)chain;
    constexpr int nn = 2 * NDIM;
    const int half = L.half, groups = half >> 3;
    uint8_t *tgt = plane_ptr(L, chain, colour);
    const uint8_t *__restrict__ oth = plane_ptr(L, chain, colour ^ 1);
    SiteAcc acc;
    for (int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; blk < nblk; blk += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(blk / groups);
        const int col = (int)(blk - (int64_t)row * groups) << 3;
        const int y = NDIM == 2 ? row : row % L.Ly;
        const int z = NDIM == 2 ? 0 : row / L.Ly;
        const int p = (colour + y + z) & 1;            // x offset of the target sites in this row
        const int64_t rb = (int64_t)row * half;
        const int yu = y == 0 ? L.Ly - 1 : y - 1, yd = y == L.Ly - 1 ? 0 : y + 1;
        const uint64_t T = ld64(tgt + rb + col);
        const uint64_t Cc = ld64(oth + rb + col);
        uint64_t raw = Cc + ld64(oth + ((int64_t)z * L.Ly + yu) * half + col) + ld64(oth + ((int64_t)z * L.Ly + yd) * half + col);
        if (NDIM == 3) {
            const int zu = z == 0 ? L.Lz - 1 : z - 1, zd = z == L.Lz - 1 ? 0 : z + 1;
            raw += ld64(oth + ((int64_t)zu * L.Ly + y) * half + col) + ld64(oth + ((int64_t)zd * L.Ly + y) * half + col);
        }
        if (p == 0) {
            const uint64_t side = oth[rb + (col == 0 ? half : col) - 1];
            raw += (Cc << 8) | side;
        } else {
            const uint64_t side = oth[rb + (col + 8 == half ? 0 : col + 8)];
            raw += (Cc >> 8) | (side << 56);
        }
        const Philox4 r0 = stream_block(seed_lo, seed_hi, chain_id, TAG_SWEEP, t, (uint32_t)blk, 0);
        Philox4 r2 = r0;
        if (MODEL == MCX_BLUME_CAPEL && RULE != MCX_HEATBATH)
            r2 = stream_block(seed_lo, seed_hi, chain_id, TAG_SWEEP, t, (uint32_t)blk, 2);
        // fast pass: branch-free 16-bit decisions; any tie of the high halves (2^-16 per site)
        // sends the whole group of eight through the exact rule below
        uint64_t Tn = 0;
        bool tie = false;
        SiteAcc fast;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int so = (int)((T >> (8 * k)) & 0xff), rk = (int)((raw >> (8 * k)) & 0xff);
            int sn;
            if (MODEL == MCX_ISING) {
                const uint32_t h = lane16(r0, k), th = s_thi[so * (nn + 1) + rk];
                const int lt = h < th;
                tie |= h == th;
                sn = RULE == MCX_HEATBATH ? lt : (so ^ lt);
                const int ch = sn != so, sgn = 2 * so - 1;
                fast.dpair += ch * (-2 * sgn * (2 * rk - nn));
                fast.dspin += ch * (-2 * sgn);
                fast.nacc += ch;
            } else if (RULE == MCX_HEATBATH) {
                const uint32_t h = lane16(r0, k), t0 = s_thi[rk], t1 = s_thi[(2 * nn + 1) + rk];
                tie |= (h == t0) | (h == t1);
                sn = h < t0 ? 0 : (h < t1 ? 1 : 2);
                const int d = sn - so;
                fast.dpair += d * (rk - nn);
                fast.dspin += d;
                fast.dspin2 += (sn - 1) * (sn - 1) - (so - 1) * (so - 1);
                fast.nacc += d != 0;
            } else {
                const int b = (int)(lane16(r0, k) >> 15);
                const int prop = so == 0 ? (b ? 1 : 2) : so == 1 ? (b ? 0 : 2) : (b ? 0 : 1);
                const uint32_t h = lane16(r2, k), th = s_thi[(so * 2 + b) * (2 * nn + 1) + rk];
                tie |= h == th;
                sn = h < th ? prop : so;
                const int d = sn - so;
                fast.dpair += d * (rk - nn);
                fast.dspin += d;
                fast.dspin2 += (sn - 1) * (sn - 1) - (so - 1) * (so - 1);
                fast.nacc += d != 0;
            }
            Tn |= (uint64_t)sn << (8 * k);
        }
        if (tie) {
            Tn = 0;
            fast = SiteAcc();
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                const int so = (int)((T >> (8 * k)) & 0xff), rk = (int)((raw >> (8 * k)) & 0xff);
                const int sn = site_update<MODEL, RULE>(so, rk, nn, r0, r2, k, s_thi, s_tlo, [&](uint32_t plane) {
                    const Philox4 rl = stream_block(seed_lo, seed_hi, chain_id, TAG_SWEEP, t, (uint32_t)blk, plane);
                    return lane16(rl, k);
                }, fast);
                Tn |= (uint64_t)sn << (8 * k);
            }
        }
        acc.dpair += fast.dpair; acc.dspin += fast.dspin; acc.dspin2 += fast.dspin2; acc.nacc += fast.nacc;
        if (Tn != T) *reinterpret_cast<uint64_t *>(tgt + rb + col) = Tn;
    }
    flush_site_acc<MODEL>(acc, sums + (int64_t)chain * SUM_FIELDS);
}

template <int MODEL, int RULE, int NDIM>
void launch_rows8_t(mcx_lattice *lat, int colour, uint64_t t)
{
    const int64_t nblk = lat->view.halfN >> 3;
    int64_t blocks = (nblk + kThreads8 - 1) / kThreads8;
    const int64_t cap = (int64_t)lat->ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    k_sweep_rows8<MODEL, RULE, NDIM><<<dim3((unsigned)blocks, (unsigned)lat->nchains), kThreads8, 0, lat->ctx->stream>>>(
        lat->view, lat->d_thi, lat->d_tlo, lat->d_labels, lat->table_len, lat->d_sums, (uint32_t)lat->seed,
        (uint32_t)(lat->seed >> 32), t, colour, lat->first_chain, nblk);
    lat->ctx->launches++;
}

template <int MODEL, int NDIM>
void launch_rows8_m(mcx_lattice *lat, int colour, uint64_t t)
{
    // Glauber shares the Metropolis instantiation: the two differ only in their tables
    if (lat->rule == MCX_HEATBATH) launch_rows8_t<MODEL, MCX_HEATBATH, NDIM>(lat, colour, t);
    else launch_rows8_t<MODEL, MCX_METROPOLIS, NDIM>(lat, colour, t);
}

}  // namespace

bool launch_sweep_rows8(mcx_lattice *lat, int colour, uint64_t t)
{
    if (lat->storage != MCX_STORAGE_INT8 || (lat->view.half & 7) != 0 || lat->ndim < 2 || lat->table_len > kMaxTable) return false;
    if (lat->model == MCX_ISING) {
        if (lat->ndim == 2) launch_rows8_m<MCX_ISING, 2>(lat, colour, t); else launch_rows8_m<MCX_ISING, 3>(lat, colour, t);
    } else {
        if (lat->ndim == 2) launch_rows8_m<MCX_BLUME_CAPEL, 2>(lat, colour, t); else launch_rows8_m<MCX_BLUME_CAPEL, 3>(lat, colour, t);
    }
    return true;
}

}  // namespace mcx
