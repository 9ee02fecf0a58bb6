// k_generic.cu -- layout conversion, init, full recompute and the shape-generic checkerboard sweep.
//
// The generic sweep handles every (model, rule, ndim, even dims) combination with one thread per
// Philox block (eight consecutive slots of the target colour plane).  Row-aligned 2-D lattices use
// the vectorised kernels of k_ising2d.cu instead; both produce identical trajectories because the
// random stream is addressed by site, not by thread (RNG layout v1).
//
// Per-site rules restate (through host-built integer tables, include/mcx_b200.h):
//   Ising   spin_flip! ising.jl:35-58, flip_changes :187-192/:484-489, modify! :200-205/:493-498
//   BC      spin_flip! blume_capel.jl:52-85, _propose_state :21-30, propose_changes :235-241
//   accept! metropolis.jl:14-17,121-127; _accept! importance_sampling.jl:80-85
#include "k_site.cuh"
#include "mcx_internal.h"

namespace mcx {

// ------------------------------------------------------------------ coordinates
struct Site {
    int x, y, z;
    int64_t row;   // y + Ly*z
};

__device__ __forceinline__ Site site_of_slot(const LatView &L, int64_t q, int colour)
{
    Site s;
    s.row = q / L.half;
    const int j = (int)(q - s.row * L.half);
    s.y = (int)(s.row % L.Ly);
    s.z = (int)(s.row / L.Ly);
    s.x = 2 * j + ((colour + s.y + s.z) & 1);
    return s;
}

// sum of the raw encodings of the 2*ndim neighbours (all live in the other colour plane)
__device__ __forceinline__ int neighbour_raw_sum(const LatView &L, const uint8_t *__restrict__ oth, const Site &s)
{
    const int64_t rowbase = s.row * L.half;
    const int xl = s.x == 0 ? L.Lx - 1 : s.x - 1;
    const int xr = s.x == L.Lx - 1 ? 0 : s.x + 1;
    int acc = oth[rowbase + (xl >> 1)] + oth[rowbase + (xr >> 1)];
    const int j = s.x >> 1;
    if (L.ndim > 1) {
        const int yu = s.y == 0 ? L.Ly - 1 : s.y - 1;
        const int yd = s.y == L.Ly - 1 ? 0 : s.y + 1;
        acc += oth[((int64_t)s.z * L.Ly + yu) * L.half + j] + oth[((int64_t)s.z * L.Ly + yd) * L.half + j];
    }
    if (L.ndim > 2) {
        const int zu = s.z == 0 ? L.Lz - 1 : s.z - 1;
        const int zd = s.z == L.Lz - 1 ? 0 : s.z + 1;
        acc += oth[((int64_t)zu * L.Ly + s.y) * L.half + j] + oth[((int64_t)zd * L.Ly + s.y) * L.half + j];
    }
    return acc;
}

// ------------------------------------------------------------------ pack / unpack
__global__ void k_pack(LatView L, const int8_t *__restrict__ staging, int64_t N)
{
    const int chain = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int x = (int)(i % L.Lx);
    const int64_t row = i / L.Lx;
    const int y = (int)(row % L.Ly), z = (int)(row / L.Ly);
    const int colour = (x + y + z) & 1;
    const int v = staging[(int64_t)chain * N + i];
    const uint8_t enc = L.model == MCX_ISING ? (uint8_t)(v > 0) : (uint8_t)(v + 1);
    plane_ptr(L, chain, colour)[row * L.half + (x >> 1)] = enc;
}

__global__ void k_unpack(LatView L, int8_t *__restrict__ staging, int64_t N)
{
    const int chain = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int x = (int)(i % L.Lx);
    const int64_t row = i / L.Lx;
    const int y = (int)(row % L.Ly), z = (int)(row / L.Ly);
    const int colour = (x + y + z) & 1;
    const int enc = plane_ptr(L, chain, colour)[row * L.half + (x >> 1)];
    staging[(int64_t)chain * N + i] = L.model == MCX_ISING ? (int8_t)(2 * enc - 1) : (int8_t)(enc - 1);
}

// ------------------------------------------------------------------ init!(sys, mode; rng)
__global__ void k_init(LatView L, int64_t N, int mode, uint32_t seed_lo, uint32_t seed_hi, uint32_t first_chain)
{
    const int chain = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int x = (int)(i % L.Lx);
    const int64_t row = i / L.Lx;
    const int y = (int)(row % L.Ly), z = (int)(row / L.Ly);
    const int colour = (x + y + z) & 1;
    int s;   // physical spin
    if (mode == MCX_INIT_UP) s = 1;
    else if (mode == MCX_INIT_DOWN) s = -1;
    else if (mode == MCX_INIT_ZERO) s = 0;
    else if (L.model == MCX_ISING) {
        const Philox4 p = stream_block(seed_lo, seed_hi, first_chain + chain, TAG_INIT, 0, (uint32_t)(i >> 7), 0);
        const int w = (int)((i >> 5) & 3);
        const uint32_t word = w == 0 ? p.x : w == 1 ? p.y : w == 2 ? p.z : p.w;
        s = ((word >> (i & 31)) & 1u) ? 1 : -1;
    } else {
        const Philox4 p = stream_block(seed_lo, seed_hi, first_chain + chain, TAG_INIT, 0, (uint32_t)(i >> 3), 0);
        s = (int)((lane16(p, (int)(i & 7)) * 3u) >> 16) - 1;
    }
    const uint8_t enc = L.model == MCX_ISING ? (uint8_t)(s > 0) : (uint8_t)(s + 1);
    plane_ptr(L, chain, colour)[row * L.half + (x >> 1)] = enc;
}

// ------------------------------------------------------------------ _recompute_cached!
__global__ void k_zero_sums(long long *sums, int nchains, int keep_accepted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchains * SUM_FIELDS) return;
    if (keep_accepted && (i % SUM_FIELDS) == SUM_ACC) return;
    sums[i] = 0;
}

__global__ void k_recompute(LatView L, long long *__restrict__ sums)
{
    const int chain = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int pair = 0, spin = 0, spin2 = 0;
    if (q < L.halfN) {
        const uint8_t *p0 = plane_ptr(L, chain, 0), *p1 = plane_ptr(L, chain, 1);
        const Site s = site_of_slot(L, q, 0);
        const int raw = neighbour_raw_sum(L, p1, s);
        int s0, s1, nsum;
        if (L.model == MCX_ISING) {
            s0 = 2 * p0[q] - 1; s1 = 2 * p1[q] - 1; nsum = 2 * raw - L.nn;
        } else {
            s0 = p0[q] - 1; s1 = p1[q] - 1; nsum = raw - L.nn;
        }
        pair = s0 * nsum;            // every bond joins colour 0 to colour 1: counted once
        spin = s0 + s1;
        spin2 = s0 * s0 + s1 * s1;
    }
    pair = warp_sum(pair); spin = warp_sum(spin); spin2 = warp_sum(spin2);
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *o = (unsigned long long *)(sums + (int64_t)chain * SUM_FIELDS);
        atomicAdd(o + SUM_PAIR, (unsigned long long)(long long)pair);
        atomicAdd(o + SUM_SPIN, (unsigned long long)(long long)spin);
        atomicAdd(o + SUM_SPIN2, (unsigned long long)(long long)spin2);
    }
}

// ------------------------------------------------------------------ generic checkerboard sweep
template <int MODEL, int RULE>
__global__ void __launch_bounds__(128)
k_sweep_generic(LatView L, const uint32_t *__restrict__ thi, const uint32_t *__restrict__ tlo,
                const int32_t *__restrict__ labels, int table_len, long long *__restrict__ sums,
                uint32_t seed_lo, uint32_t seed_hi, uint64_t t, int colour, uint32_t first_chain)
{
    const int chain = blockIdx.y;
    const uint32_t chain_id = first_chain + (uint32_t)chain;
    const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nblk = (L.halfN + 7) >> 3;
    SiteAcc acc;
    if (blk < nblk) {
        uint8_t *tgt = plane_ptr(L, chain, colour);
        const uint8_t *oth = plane_ptr(L, chain, colour ^ 1);
        const int tb = labels[chain] * table_len;
        const int nn = L.nn;
        const Philox4 r0 = stream_block(seed_lo, seed_hi, chain_id, TAG_SWEEP, t, (uint32_t)blk, 0);
        Philox4 r2 = r0;
        if (MODEL == MCX_BLUME_CAPEL && RULE != MCX_HEATBATH)
            r2 = stream_block(seed_lo, seed_hi, chain_id, TAG_SWEEP, t, (uint32_t)blk, 2);
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            const int64_t q = blk * 8 + k;
            if (q >= L.halfN) break;
            const Site s = site_of_slot(L, q, colour);
            const int raw = neighbour_raw_sum(L, oth, s);
            const int so = tgt[q];
            // m < T with the low half fetched only on a tie of the high halves
            const int sn = site_update<MODEL, RULE>(so, raw, nn, r0, r2, k, thi + tb, tlo + tb, [&](uint32_t plane) {
                const Philox4 rl = stream_block(seed_lo, seed_hi, chain_id, TAG_SWEEP, t, (uint32_t)blk, plane);
                return lane16(rl, k);
            }, acc);
            if (sn != so) tgt[q] = (uint8_t)sn;
        }
    }
    flush_site_acc<MODEL>(acc, sums + (int64_t)chain * SUM_FIELDS);
}

// ------------------------------------------------------------------ launchers
static dim3 site_grid(const mcx_lattice *lat, int64_t n, int threads)
{
    return dim3((unsigned)((n + threads - 1) / threads), (unsigned)lat->nchains, 1);
}

void launch_pack(mcx_lattice *lat)
{
    if (lat->storage == MCX_STORAGE_BIT) { launch_pack_bits(lat); return; }
    if (launch_pack_ising2d(lat)) return;
    k_pack<<<site_grid(lat, lat->N, 256), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_staging, lat->N);
    lat->ctx->launches++;
}

void launch_unpack(mcx_lattice *lat)
{
    if (lat->storage == MCX_STORAGE_BIT) { launch_unpack_bits(lat); return; }
    if (launch_unpack_ising2d(lat)) return;
    k_unpack<<<site_grid(lat, lat->N, 256), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_staging, lat->N);
    lat->ctx->launches++;
}

void launch_init(mcx_lattice *lat, int mode, uint64_t seed)
{
    if (lat->storage == MCX_STORAGE_BIT) { launch_init_bits(lat, mode, seed); return; }
    if (launch_init_ising2d(lat, mode, seed)) return;
    k_init<<<site_grid(lat, lat->N, 256), 256, 0, lat->ctx->stream>>>(
        lat->view, lat->N, mode, (uint32_t)seed, (uint32_t)(seed >> 32), lat->first_chain);
    lat->ctx->launches++;
}

void launch_recompute(mcx_lattice *lat)
{
    const int n = lat->nchains * SUM_FIELDS;
    k_zero_sums<<<(n + 127) / 128, 128, 0, lat->ctx->stream>>>(lat->d_sums, lat->nchains, 1);
    lat->ctx->launches++;
    if (lat->storage == MCX_STORAGE_BIT) { launch_recompute_bits(lat); return; }
    if (launch_recompute_ising2d(lat)) return;
    k_recompute<<<site_grid(lat, lat->view.halfN, 256), 256, 0, lat->ctx->stream>>>(lat->view, lat->d_sums);
    lat->ctx->launches++;
}

template <int MODEL, int RULE>
static void launch_generic_t(mcx_lattice *lat, int colour, uint64_t t)
{
    const int64_t nblk = (lat->view.halfN + 7) >> 3;
    k_sweep_generic<MODEL, RULE><<<site_grid(lat, nblk, 128), 128, 0, lat->ctx->stream>>>(
        lat->view, lat->d_thi, lat->d_tlo, lat->d_labels, lat->table_len, lat->d_sums, (uint32_t)lat->seed,
        (uint32_t)(lat->seed >> 32), t, colour, lat->first_chain);
    lat->ctx->launches++;
}

void launch_sweep_generic(mcx_lattice *lat, int colour, uint64_t t)
{
    if (lat->model == MCX_ISING) {
        if (lat->rule == MCX_HEATBATH) launch_generic_t<MCX_ISING, MCX_HEATBATH>(lat, colour, t);
        else launch_generic_t<MCX_ISING, MCX_METROPOLIS>(lat, colour, t);   // Glauber differs only in its table
    } else {
        if (lat->rule == MCX_HEATBATH) launch_generic_t<MCX_BLUME_CAPEL, MCX_HEATBATH>(lat, colour, t);
        else launch_generic_t<MCX_BLUME_CAPEL, MCX_METROPOLIS>(lat, colour, t);
    }
}

}  // namespace mcx
