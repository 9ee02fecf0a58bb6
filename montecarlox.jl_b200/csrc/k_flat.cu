// k_flat.cu -- flat-histogram chains: multicanonical and Wang-Landau sweeps.
//
// The acceptance of these ensembles depends on the chain's GLOBAL observable
// (spin_flip!(sys, alg::AbstractImportanceSampling), SpinSystems/src/ising.jl:25-33 passes
// E_new, E_old), so a chain is serial by construction and the parallelism is across chains only
// (finding 4).  One thread owns one chain and visits sites 0..N-1 in order
// (FLAT stream); spins are stored chain-interleaved, spins[site][chain], so the 32 chains of a warp
// touch 32 consecutive bytes per site.  Per attempt (restating importance_sampling.jl:69-85,
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

constexpr int kFlatThreads = 32;          // one warp = 32 chains per block: spreads chains over SMs
constexpr int kSmemBins = 8192;           // 32 KB of uint32 counters

struct FlatParams {
    int8_t *spins;              // [N][nchains]
    long long *state;           // [nchains][4] pair, spin, spin2, accepted
    double *logweight;          // muca: [nbins]; WL: [nchains][nbins]
    unsigned long long *hist;   // [nbins]
    int *error;
    int64_t start, step, nbins, N;
    int Lx, Ly, Lz, ndim, nchains;
    double beta_pair, logf, J;
    uint32_t seed_lo, seed_hi, first_chain;
    uint64_t sweep0;
    int nsweeps, policy, step_shift, prefetch;
};

// rand < exp(log_ratio) for rand = m * 2^-32, m = hi << 16 | lo, decided exactly but cheaply:
//  1. a float estimate of p with a rigorous error margin settles all but ~2^-16 of the attempts from
//     the high half alone (no double-precision exp, no second Philox block);
//  2. otherwise the double-precision exp decides on the high half, and only if (hi, hi + 1) brackets
//     p * 2^16 is the low half fetched and the full 32-bit comparison made.
// The result equals (double)m * 2^-32 < exp(log_ratio) in every case.
template <class LoFn>
__device__ __forceinline__ bool draw_less_exp(uint32_t hi, double log_ratio, LoFn lo_fn)
{
    const float lf = (float)log_ratio;
    const float p16 = __expf(lf) * 65536.0f;
    const float eps = 3.0e-6f + fabsf(lf) * 1.0e-6f;       // > 4x the worst-case relative error of p16
    if ((float)(hi + 1) <= p16 * (1.0f - eps)) return true;
    if (hi > 0 && (float)hi >= p16 * (1.0f + eps)) return false;   // hi == 0: a denormal p could still win
    const double p = exp(log_ratio), pd16 = p * 65536.0;
    if ((double)(hi + 1) <= pd16) return true;
    if ((double)hi >= pd16) return false;
    const uint32_t lo = lo_fn();
    return (double)((hi << 16) | lo) * (1.0 / 4294967296.0) < p;
}

// div(x - start, step) truncating toward zero like Julia's div (binned_object.jl:22-24)
__device__ __forceinline__ int64_t bin_of(int64_t x, int64_t start, int64_t step, int shift)
{
    const int64_t d = x - start;
    if (shift >= 0) return (d + ((d >> 63) & (step - 1))) >> shift;
    return d / step;
}

template <int OBS, int KIND, bool SMEM_HIST>
__global__ void __launch_bounds__(kFlatThreads) k_flat_sweep(FlatParams P)
{
    extern __shared__ uint32_t s_hist[];
    if (SMEM_HIST) {
        for (int i = threadIdx.x; i < P.nbins; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < P.nchains) {
        const int nch = P.nchains;
        int8_t *sp = P.spins + c;
        long long *st = P.state + (int64_t)c * 4;
        long long pair = st[0], spin = st[1], spin2 = st[2], nacc = st[3];
        double *lw = P.logweight + (KIND == MCX_FLAT_WANG_LANDAU ? (int64_t)c * P.nbins : 0);
        const uint32_t chain_id = P.first_chain + (uint32_t)c;
        const int Lx = P.Lx, Ly = P.Ly, Lz = P.Lz;
        const int64_t sx = nch, sy = (int64_t)Lx * nch, sz = (int64_t)Lx * Ly * nch;
        const int shift = P.step_shift;
        // bin and log-weight of the current state are carried from attempt to attempt
        int64_t io = bin_of(OBS == MCX_OBS_ENERGY ? -pair : spin2, P.start, P.step, shift);
        bool dead = io < 0 || io >= P.nbins;
        if (dead) atomicExch(P.error, 1);
        double lw_old = dead ? 0.0 : lw[io];
        // run-length cache for the global histogram (a chain revisits its current bin many times)
        int64_t run_bin = io;
        unsigned long long run_cnt = 0;
        const int64_t pf = (int64_t)P.prefetch * nch;     // prefetch distance in bytes of the interleaved array
        const int64_t total = P.N * (int64_t)nch;
        for (int sw = 0; sw < P.nsweeps && !dead; ++sw) {
            const uint64_t t = P.sweep0 + (uint64_t)sw;
            Philox4 r0{}, r2{};
            int64_t i = 0;
            for (int z = 0; z < Lz && !dead; ++z)
                for (int y = 0; y < Ly && !dead; ++y)
                    for (int x = 0; x < Lx; ++x, ++i) {
                        const int lane = (int)(i & 7);
                        if (lane == 0) {
                            r0 = stream_block(P.seed_lo, P.seed_hi, chain_id, TAG_FLAT, t, (uint32_t)(i >> 3), 0);
                            if (OBS == MCX_OBS_SPIN2_WITH_PAIR_BOLTZMANN)
                                r2 = stream_block(P.seed_lo, P.seed_hi, chain_id, TAG_FLAT, t, (uint32_t)(i >> 3), 2);
                        }
                        int8_t *p = sp + i * nch;
                        if (pf) {
                            // lines first touched `prefetch` sites ahead: the site's own row of the next
                            // y (and, in 3-D, the z+1 and z-1 planes); addresses wrap like the lattice
                            int64_t a = i * nch + pf;
                            if (a >= total) a -= total;
                            int64_t ay = a + sy; if (ay >= total) ay -= total;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + ay));
                            if (P.ndim > 2) {
                                int64_t az = a + sz; if (az >= total) az -= total;
                                int64_t aw = a - sz; if (aw < 0) aw += total;
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + az));
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + aw));
                            }
                        }
                        const int s = *p;
                        int nb = p[x == 0 ? (Lx - 1) * sx : -sx] + p[x == Lx - 1 ? -(Lx - 1) * sx : sx];
                        if (P.ndim > 1) nb += p[y == 0 ? (Ly - 1) * sy : -sy] + p[y == Ly - 1 ? -(Ly - 1) * sy : sy];
                        if (P.ndim > 2) nb += p[z == 0 ? (Lz - 1) * sz : -sz] + p[z == Lz - 1 ? -(Lz - 1) * sz : sz];

                        int s_new, dpair, dspin, dspin2;
                        int64_t x_new;
                        uint32_t hi;
                        uint32_t lo_plane;
                        if (OBS == MCX_OBS_ENERGY) {
                            // flip_changes / delta_energy (ising.jl:187-198), integer path J = 1, h = 0
                            s_new = -s; dpair = -2 * s * nb; dspin = -2 * s; dspin2 = 0;
                            x_new = -pair - dpair;
                            hi = lane16(r0, lane); lo_plane = 1;
                        } else {
                            // _propose_state + propose_changes (blume_capel.jl:21-30,235-241),
                            // H = (J*sum_pair, sum_spins2) as in muca_BlumeCapel.jl:81-89
                            const int b = (int)(lane16(r0, lane) >> 15);
                            s_new = s == -1 ? (b ? 0 : 1) : s == 0 ? (b ? -1 : 1) : (b ? -1 : 0);
                            dspin = s_new - s; dspin2 = s_new * s_new - s * s; dpair = dspin * nb;
                            x_new = spin2 + dspin2;
                            hi = lane16(r2, lane); lo_plane = 3;
                        }
                        // _binindex for integer bins: div(x - start, step) + 1 (binned_object.jl:22-24); 0-based here
                        const int64_t in = bin_of(x_new, P.start, P.step, shift);
                        const bool inside = in >= 0 && in < P.nbins;
                        if (!inside && P.policy == 0) {                          // BoundsError
                            atomicExch(P.error, 1);
                            dead = true;
                            break;
                        }
                        bool accepted = false;
                        double lw_new = lw_old;
                        if (inside) {
                            lw_new = in == io ? lw_old : lw[in];
                            double log_ratio;
                            if (OBS == MCX_OBS_ENERGY) {
                                log_ratio = lw_new - lw_old;
                            } else {
                                const double Ho1 = P.J * (P.J * (double)pair);
                                const double Hn1 = Ho1 + P.J * (P.J * (double)dpair);
                                log_ratio = (-P.beta_pair * Hn1 + lw_new) - (-P.beta_pair * Ho1 + lw_old);
                            }
                            // _accept! (importance_sampling.jl:80-85)
                            if (log_ratio > 0) accepted = true;
                            else accepted = draw_less_exp(hi, log_ratio, [&]() {
                                const Philox4 rl = stream_block(P.seed_lo, P.seed_hi, chain_id, TAG_FLAT, t,
                                                                (uint32_t)(i >> 3), lo_plane);
                                return lane16(rl, lane);
                            });
                        }
                        if (accepted) {
                            *p = (int8_t)s_new;
                            pair += dpair; spin += dspin; spin2 += dspin2; nacc += 1;
                            io = in; lw_old = lw_new;
                        }
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
                            lw[io] = lw_old;
                        }
                    }
        }
        if (KIND == MCX_FLAT_MUCA && !SMEM_HIST && run_cnt) atomicAdd(P.hist + run_bin, run_cnt);
        st[0] = pair; st[1] = spin; st[2] = spin2; st[3] = nacc;
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < P.nbins; i += blockDim.x)
            if (s_hist[i]) atomicAdd(P.hist + i, (unsigned long long)s_hist[i]);
    }
}

// planes (colour-split, encoded) <-> chain-interleaved physical spins
__global__ void k_flat_load(LatView L, int8_t *__restrict__ il, int64_t N)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * L.nchains) return;
    const int c = (int)(idx % L.nchains);
    const int64_t i = idx / L.nchains;
    const int x = (int)(i % L.Lx);
    const int64_t row = i / L.Lx;
    const int y = (int)(row % L.Ly), z = (int)(row / L.Ly);
    const int enc = plane_ptr(L, c, (x + y + z) & 1)[row * L.half + (x >> 1)];
    il[idx] = L.model == MCX_ISING ? (int8_t)(2 * enc - 1) : (int8_t)(enc - 1);
}

__global__ void k_flat_store(LatView L, const int8_t *__restrict__ il, int64_t N)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * L.nchains) return;
    const int c = (int)(idx % L.nchains);
    const int64_t i = idx / L.nchains;
    const int x = (int)(i % L.Lx);
    const int64_t row = i / L.Lx;
    const int y = (int)(row % L.Ly), z = (int)(row / L.Ly);
    const int v = il[idx];
    plane_ptr(L, c, (x + y + z) & 1)[row * L.half + (x >> 1)] = L.model == MCX_ISING ? (uint8_t)(v > 0) : (uint8_t)(v + 1);
}

__global__ void k_flat_state_in(const long long *__restrict__ sums, long long *__restrict__ state, int nchains, int64_t N,
                                int model)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchains) return;
    state[c * 4 + 0] = sums[c * SUM_FIELDS + SUM_PAIR];
    state[c * 4 + 1] = sums[c * SUM_FIELDS + SUM_SPIN];
    state[c * 4 + 2] = model == MCX_ISING ? N : sums[c * SUM_FIELDS + SUM_SPIN2];
    state[c * 4 + 3] = 0;
}

__global__ void k_flat_state_out(long long *__restrict__ sums, const long long *__restrict__ state, int nchains)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchains) return;
    sums[c * SUM_FIELDS + SUM_PAIR] = state[c * 4 + 0];
    sums[c * SUM_FIELDS + SUM_SPIN] = state[c * 4 + 1];
    sums[c * SUM_FIELDS + SUM_SPIN2] = state[c * 4 + 2];
    sums[c * SUM_FIELDS + SUM_ACC] += state[c * 4 + 3];
}

// update!(ens::MulticanonicalEnsemble; mode=:simple): lw -= (h > 0 ? log(h) : 0) (multicanonical.jl:32-44)
__global__ void k_muca_update(double *__restrict__ lw, const unsigned long long *__restrict__ hist, int64_t nbins)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbins) return;
    const unsigned long long hh = hist[i];
    if (hh > 0) lw[i] -= log((double)hh);
}

template <int OBS, int KIND>
void launch_sweep_t(mcx_flat *f, const FlatParams &P)
{
    mcx_lattice *lat = f->lat;
    const int blocks = (lat->nchains + kFlatThreads - 1) / kFlatThreads;
    const bool smem = KIND == MCX_FLAT_MUCA && f->nbins <= kSmemBins;
    if (smem)
        k_flat_sweep<OBS, KIND, true><<<blocks, kFlatThreads, (size_t)f->nbins * sizeof(uint32_t), lat->ctx->stream>>>(P);
    else
        k_flat_sweep<OBS, KIND, false><<<blocks, kFlatThreads, 0, lat->ctx->stream>>>(P);
    lat->ctx->launches++;
}

}  // namespace

void launch_flat_load(mcx_flat *f)
{
    mcx_lattice *lat = f->lat;
    const int64_t n = lat->N * lat->nchains;
    k_flat_load<<<(unsigned)((n + 255) / 256), 256, 0, lat->ctx->stream>>>(lat->view, f->d_spins, lat->N);
    k_flat_state_in<<<(lat->nchains + 127) / 128, 128, 0, lat->ctx->stream>>>(lat->d_sums, f->d_state, lat->nchains, lat->N,
                                                                            lat->model);
    lat->ctx->launches += 2;
}

void launch_flat_store(mcx_flat *f)
{
    mcx_lattice *lat = f->lat;
    const int64_t n = lat->N * lat->nchains;
    k_flat_store<<<(unsigned)((n + 255) / 256), 256, 0, lat->ctx->stream>>>(lat->view, f->d_spins, lat->N);
    k_flat_state_out<<<(lat->nchains + 127) / 128, 128, 0, lat->ctx->stream>>>(lat->d_sums, f->d_state, lat->nchains);
    lat->ctx->launches += 2;
}

void launch_flat_sweep(mcx_flat *f, uint64_t sweep0, int nsweeps)
{
    mcx_lattice *lat = f->lat;
    FlatParams P;
    P.spins = f->d_spins; P.state = f->d_state; P.logweight = f->d_logweight;
    P.hist = (unsigned long long *)f->d_histogram; P.error = f->d_error;
    P.start = f->start; P.step = f->step; P.nbins = f->nbins; P.N = lat->N;
    P.Lx = lat->dims[0]; P.Ly = lat->dims[1]; P.Lz = lat->dims[2]; P.ndim = lat->ndim; P.nchains = lat->nchains;
    P.beta_pair = f->beta_pair; P.logf = f->logf; P.J = lat->J;
    P.seed_lo = (uint32_t)lat->seed; P.seed_hi = (uint32_t)(lat->seed >> 32); P.first_chain = lat->first_chain;
    P.sweep0 = sweep0; P.nsweeps = nsweeps; P.policy = f->policy;
    P.step_shift = -1;
    for (int sh = 0; sh < 62; ++sh)
        if (f->step == ((int64_t)1 << sh)) P.step_shift = sh;
    {
        // software prefetch only pays when the interleaved spins do not sit in L2 anyway
        const char *e = getenv("MCX_FLAT_PREFETCH");
        const double bytes = (double)lat->N * lat->nchains;
        P.prefetch = e ? atoi(e) : (bytes > 48.0e6 ? 24 : 0);
    }
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
        (e = cudaMalloc((void **)&f->d_spins, (size_t)lat->N * lat->nchains)) != cudaSuccess ||
        (e = cudaMalloc((void **)&f->d_state, sizeof(long long) * 4 * (size_t)lat->nchains)) != cudaSuccess ||
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
    cudaFree(f->d_logweight); cudaFree(f->d_histogram); cudaFree(f->d_spins); cudaFree(f->d_state); cudaFree(f->d_error);
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
    if (lat->sums_dirty) { launch_recompute(lat); lat->sums_dirty = false; }
    launch_flat_load(f);
    // shared-memory counters are 32-bit: bound the attempts one block can record per launch
    const double per_sweep = 32.0 * (double)lat->N;
    int64_t chunk = (int64_t)(4.0e9 / per_sweep);
    if (chunk < 1) chunk = 1;
    for (int64_t done = 0; done < nsweeps; done += chunk) {
        const int n = (int)((nsweeps - done) < chunk ? (nsweeps - done) : chunk);
        launch_flat_sweep(f, lat->sweep + (uint64_t)done, n);
    }
    launch_flat_store(f);
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
