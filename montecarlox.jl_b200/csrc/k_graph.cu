// k_graph.cu -- general topologies: IsingGraph (global J) and IsingMatrix (sparse J_ij) with no / uniform / per-site
// fields (SpinSystems/src/ising.jl:86-360), the part of the hot path SURVEY.md section 8f.4 lists after the grids.
//
// The checkerboard generalises to a greedy first-fit colouring in site order: sites of one colour share no edge, so a
// colour class is updated by one launch, one thread per site and chain.  Sites are stored sorted by (colour, index);
// the neighbour lists (CSR) keep the reference's adjacency order inside a row, so the Float64 sum
// sum_j s_i J_ij s_j (ising.jl:295-305) is formed in the reference's order and delta_energy is bit-identical.
// Couplings and fields are arbitrary doubles, so nothing can be tabulated: the thread evaluates the reference's own
// expression -- log_ratio = -beta * dE, exp / logistic in double (-fmad=false keeps multiplies and adds separate) -- and
// compares it with the positioned 32-bit draw u = m * 2^-32 (high half first, low half only when (hi, hi + 1) brackets
// p * 2^16).  RNG layout v1 with t = ncolours * sweep + colour and slot = rank of the site inside its colour class.
//
// The cached sums of the reference (sum_pair_interactions, sum_field_interactions) are running Float64 sums in flip
// order; a parallel update cannot reproduce that order, so sums are formed from the spins on read
// (_recompute_cached!, ising.jl:127-144) by a fixed-order reduction: exact, hence equal to the reference's, whenever
// the couplings and fields are exactly representable sums (integers, dyadic fractions); to rounding otherwise.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "mcx_internal.h"

struct mcx_graph {
    mcx_ctx *ctx;
    int64_t n, nnz;
    int nchains, ncolours;
    bool matrix;                 // per-entry couplings (IsingMatrix) or one global J (IsingGraph)
    double J;
    int hmode;                   // 0 none, 1 uniform, 2 per site
    double h;
    // device, sorted order (k = position after sorting by (colour, site index))
    int8_t *d_spins;             // [nchains][n], -1 / +1
    int64_t *d_rowptr;           // [n + 1]
    int32_t *d_col;              // [nnz] sorted index of the neighbour
    double *d_val;               // [nnz] or null
    double *d_h;                 // [n] or null
    int32_t *d_perm;             // [n] original site of sorted position k
    int8_t *d_staging;           // [nchains][n] original order
    long long *d_acc;            // [nchains] accepted moves
    double *d_obs;               // [nchains][3]: pair (unscaled by J in graph mode), field, spin
    std::vector<int64_t> class_off;   // [ncolours + 1]
    std::vector<int32_t> colour;      // [n] colour of original site
    int rule;
    double beta;
    uint64_t seed, sweep;
    uint32_t first_chain;
    int64_t steps;
};

namespace mcx {
namespace {

__device__ __forceinline__ double logistic(double x)            // src/infrastructure/utils.jl:82-88
{
    if (x >= 0) return 1.0 / (1.0 + exp(-x));
    const double ex = exp(x);
    return ex / (1.0 + ex);
}

// rand(rng) < p for the positioned draw of (chain, t, slot): decided on the high half unless (hi, hi + 1) brackets p * 2^16
__device__ __forceinline__ bool draw_less(double p, uint32_t seed_lo, uint32_t seed_hi, uint32_t chain, uint64_t t, uint32_t slot)
{
    const Philox4 r = stream_block(seed_lo, seed_hi, chain, TAG_SWEEP, t, slot >> 3, 0);
    const double hi = (double)lane16(r, (int)(slot & 7)), p16 = p * 65536.0;
    if (hi + 1.0 <= p16) return true;
    if (hi >= p16) return false;
    const Philox4 rl = stream_block(seed_lo, seed_hi, chain, TAG_SWEEP, t, slot >> 3, 1);
    const double m = hi * 65536.0 + (double)lane16(rl, (int)(slot & 7));
    return m * (1.0 / 4294967296.0) < p;
}

template <bool MATRIX>
__device__ __forceinline__ double local_pair(const int8_t *__restrict__ sp, const int64_t *__restrict__ rowptr,
                                             const int32_t *__restrict__ col, const double *__restrict__ val, int64_t k, int s)
{
    const int64_t p0 = rowptr[k], p1 = rowptr[k + 1];
    if (MATRIX) {
        double acc = 0.0;
        const double si = (double)s;
        for (int64_t p = p0; p < p1; ++p) {
            const int32_t j = col[p];
            if (j != k) acc += si * val[p] * (double)sp[j];
        }
        return acc;
    }
    int acc = 0;
    for (int64_t p = p0; p < p1; ++p) acc += s * sp[col[p]];
    return (double)acc;
}

template <bool MATRIX>
__global__ void __launch_bounds__(128)
k_graph_colour(int8_t *__restrict__ spins, int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
               const double *__restrict__ val, const double *__restrict__ hvec, int hmode, double h, double J, int rule, double beta,
               int64_t off, int64_t cnt, uint32_t seed_lo, uint32_t seed_hi, uint32_t first_chain, uint64_t t,
               long long *__restrict__ accepted)
{
    const int chain = blockIdx.y;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int acc = 0;
    if (slot < cnt) {
        int8_t *sp = spins + (int64_t)chain * n;
        const int64_t k = off + slot;
        const int s = sp[k];
        // flip_changes / delta_energy (ising.jl:187-198, 339-351)
        const double lp = local_pair<MATRIX>(sp, rowptr, col, val, k, s);
        const double dpair = MATRIX ? -2.0 * lp : (-2 * J) * lp;
        const int dspin = -2 * s;
        const double dfield = hmode == 0 ? 0.0 : hmode == 1 ? h * (double)dspin : hvec[k] * (double)dspin;
        const double dE = -dpair - dfield;
        const uint32_t chain_id = first_chain + (uint32_t)chain;
        bool flip;
        if (rule == MCX_HEATBATH) {
            // ising.jl:43-58: p(+1) = logistic(beta * s_old * dE)
            const double p_plus = logistic(beta * (double)s * dE);
            const int s_new = draw_less(p_plus, seed_lo, seed_hi, chain_id, t, (uint32_t)slot) ? 1 : -1;
            flip = s_new != s;
        } else {
            const double log_ratio = -beta * dE;
            if (rule == MCX_GLAUBER) flip = draw_less(logistic(log_ratio), seed_lo, seed_hi, chain_id, t, (uint32_t)slot);   // metropolis.jl:121-127
            else flip = log_ratio > 0 || draw_less(exp(log_ratio), seed_lo, seed_hi, chain_id, t, (uint32_t)slot);          // importance_sampling.jl:80-85
            acc = flip ? 1 : 0;
        }
        if (flip) sp[k] = (int8_t)(-s);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd((unsigned long long *)(accepted + chain), (unsigned long long)acc);
}

// pair / field / spin sums of one chain: one block, every thread a fixed stride of sites, fixed-order tree -> deterministic
template <bool MATRIX>
__global__ void __launch_bounds__(256)
k_graph_sums(const int8_t *__restrict__ spins, int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
             const double *__restrict__ val, const double *__restrict__ hvec, double *__restrict__ obs)
{
    __shared__ double s_pair[256], s_field[256], s_spin[256];
    const int chain = blockIdx.x;
    const int8_t *sp = spins + (int64_t)chain * n;
    double pair = 0.0, field = 0.0, spin = 0.0;
    for (int64_t k = threadIdx.x; k < n; k += 256) {
        const int s = sp[k];
        pair += local_pair<MATRIX>(sp, rowptr, col, val, k, s);
        if (hvec) field += hvec[k] * (double)s;
        spin += (double)s;
    }
    s_pair[threadIdx.x] = pair; s_field[threadIdx.x] = field; s_spin[threadIdx.x] = spin;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            s_pair[threadIdx.x] += s_pair[threadIdx.x + w];
            s_field[threadIdx.x] += s_field[threadIdx.x + w];
            s_spin[threadIdx.x] += s_spin[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        obs[3 * chain + 0] = s_pair[0] / 2;
        obs[3 * chain + 1] = s_field[0];
        obs[3 * chain + 2] = s_spin[0];
    }
}

__global__ void k_graph_permute_in(const int8_t *__restrict__ staging, int8_t *__restrict__ spins, const int32_t *__restrict__ perm, int64_t n)
{
    const int chain = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) spins[(int64_t)chain * n + k] = staging[(int64_t)chain * n + perm[k]] > 0 ? 1 : -1;
}
__global__ void k_graph_permute_out(const int8_t *__restrict__ spins, int8_t *__restrict__ staging, const int32_t *__restrict__ perm, int64_t n)
{
    const int chain = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) staging[(int64_t)chain * n + perm[k]] = spins[(int64_t)chain * n + k];
}
// init!(sys, mode; rng): the INIT stream is addressed by the ORIGINAL site index (bit i & 127 of block i >> 7)
__global__ void k_graph_init(int8_t *__restrict__ spins, const int32_t *__restrict__ perm, int64_t n, int mode, uint32_t seed_lo,
                             uint32_t seed_hi, uint32_t first_chain)
{
    const int chain = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int s = mode == MCX_INIT_DOWN ? -1 : 1;
    if (mode == MCX_INIT_RANDOM) {
        const int64_t i = perm[k];
        const Philox4 p = stream_block(seed_lo, seed_hi, first_chain + chain, TAG_INIT, 0, (uint32_t)(i >> 7), 0);
        const int w = (int)((i >> 5) & 3);
        const uint32_t word = w == 0 ? p.x : w == 1 ? p.y : w == 2 ? p.z : p.w;
        s = ((word >> (i & 31)) & 1u) ? 1 : -1;
    }
    spins[(int64_t)chain * n + k] = (int8_t)s;
}

}  // namespace
}  // namespace mcx

using namespace mcx;

int32_t mcx_set_error(int32_t code, const char *msg);   // mcx_api.cu

static int32_t gfail(int32_t code, const char *fmt, ...)
{
    char buf[400];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return mcx_set_error(code, buf);
}
#define GREQ(cond, code, ...) do { if (!(cond)) return gfail(code, __VA_ARGS__); } while (0)
#define GCUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return gfail(MCX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); } while (0)

static void graph_free(mcx_graph *g)
{
    if (!g) return;
    cudaFree(g->d_spins); cudaFree(g->d_rowptr); cudaFree(g->d_col); cudaFree(g->d_val); cudaFree(g->d_h); cudaFree(g->d_perm);
    cudaFree(g->d_staging); cudaFree(g->d_acc); cudaFree(g->d_obs);
    delete g;
}

static dim3 graph_grid(const mcx_graph *g, int64_t count, int threads)
{
    return dim3((unsigned)((count + threads - 1) / threads), (unsigned)g->nchains, 1);
}

extern "C" {

int32_t mcx_graph_create(mcx_ctx *ctx, int64_t n, const int64_t *rowptr, const int64_t *col, const double *J_ij, double J,
                         int32_t field_mode, double h, const double *h_i, int32_t nchains, mcx_graph **out)
{
    GREQ(ctx && rowptr && out, MCX_ERR_ARGUMENT, "NULL argument");
    GREQ(n >= 1 && n < ((int64_t)1 << 31), MCX_ERR_ARGUMENT, "need 1 <= n < 2^31 sites (got %lld)", (long long)n);
    GREQ(nchains >= 1 && nchains <= 65535, MCX_ERR_ARGUMENT, "nchains must be in [1, 65535] (got %d)", nchains);
    GREQ(field_mode >= 0 && field_mode <= 2, MCX_ERR_ARGUMENT, "field_mode must be 0 (none), 1 (uniform) or 2 (per site)");
    GREQ(field_mode != 2 || h_i, MCX_ERR_ARGUMENT, "Field vector length must match number of spins");
    GREQ(rowptr[0] == 0, MCX_ERR_ARGUMENT, "rowptr[0] must be 0");
    const int64_t nnz = rowptr[n];
    GREQ(nnz == 0 || col, MCX_ERR_ARGUMENT, "col is NULL");
    for (int64_t i = 0; i < n; ++i) {
        GREQ(rowptr[i + 1] >= rowptr[i], MCX_ERR_ARGUMENT, "rowptr must be non-decreasing");
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
            GREQ(col[p] >= 0 && col[p] < n, MCX_ERR_BOUNDS, "neighbour %lld of site %lld outside [0, %lld)", (long long)col[p], (long long)i, (long long)n);
    }
    // the neighbour relation must be symmetric (a SimpleGraph is; _check_symmetric, ising.jl:250-262, for a sparse J)
    {
        std::vector<std::pair<int64_t, int64_t>> e;
        std::vector<double> v;
        e.reserve((size_t)nnz);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
                if (col[p] != i) e.push_back({i, col[p]});
        std::vector<std::pair<int64_t, int64_t>> srt(e), rev(e.size());
        for (size_t k = 0; k < e.size(); ++k) rev[k] = {e[k].second, e[k].first};
        std::sort(srt.begin(), srt.end());
        std::sort(rev.begin(), rev.end());
        GREQ(srt == rev, MCX_ERR_STATE, "Sparse J must be symmetric");
        if (J_ij) {
            // J[row, col] == J[col, row]
            std::vector<std::pair<std::pair<int64_t, int64_t>, double>> a;
            a.reserve((size_t)nnz);
            for (int64_t i = 0; i < n; ++i)
                for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
                    if (col[p] != i) a.push_back({{std::min(i, col[p]), std::max(i, col[p])}, J_ij[p]});
            std::sort(a.begin(), a.end());
            for (size_t k = 0; k + 1 < a.size(); k += 2)
                GREQ(a[k].first == a[k + 1].first && a[k].second == a[k + 1].second, MCX_ERR_STATE, "Sparse J must be symmetric");
        }
    }
    GCUDA(cudaSetDevice(ctx->device));
    mcx_graph *g = new (std::nothrow) mcx_graph();
    GREQ(g, MCX_ERR_STATE, "out of host memory");
    g->ctx = ctx; g->n = n; g->nnz = nnz; g->nchains = nchains; g->matrix = J_ij != nullptr; g->J = J; g->hmode = field_mode; g->h = h;
    g->rule = -1; g->beta = 0.0; g->seed = 0; g->sweep = 0; g->first_chain = 0; g->steps = 0;
    g->d_spins = nullptr; g->d_rowptr = nullptr; g->d_col = nullptr; g->d_val = nullptr; g->d_h = nullptr; g->d_perm = nullptr;
    g->d_staging = nullptr; g->d_acc = nullptr; g->d_obs = nullptr;
    // greedy first-fit colouring in site order
    g->colour.assign((size_t)n, 0);
    int ncol = 0;
    {
        int64_t maxdeg = 0;
        for (int64_t i = 0; i < n; ++i) maxdeg = std::max(maxdeg, rowptr[i + 1] - rowptr[i]);
        std::vector<int64_t> mark((size_t)maxdeg + 2, -1);
        for (int64_t i = 0; i < n; ++i) {
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
                const int64_t j = col[p];
                if (j < i && g->colour[(size_t)j] <= maxdeg) mark[(size_t)g->colour[(size_t)j]] = i;
            }
            int c = 0;
            while (mark[(size_t)c] == i) ++c;
            g->colour[(size_t)i] = c;
            ncol = std::max(ncol, c + 1);
        }
    }
    g->ncolours = ncol;
    // sorted order: by (colour, site index)
    std::vector<int32_t> perm((size_t)n), inv((size_t)n);
    g->class_off.assign((size_t)ncol + 1, 0);
    for (int64_t i = 0; i < n; ++i) g->class_off[(size_t)g->colour[(size_t)i] + 1]++;
    for (int c = 0; c < ncol; ++c) g->class_off[(size_t)c + 1] += g->class_off[(size_t)c];
    {
        std::vector<int64_t> fill(g->class_off.begin(), g->class_off.end() - 1);
        for (int64_t i = 0; i < n; ++i) {
            const int64_t k = fill[(size_t)g->colour[(size_t)i]]++;
            perm[(size_t)k] = (int32_t)i;
            inv[(size_t)i] = (int32_t)k;
        }
    }
    std::vector<int64_t> rp((size_t)n + 1, 0);
    std::vector<int32_t> cs((size_t)std::max<int64_t>(nnz, 1));
    std::vector<double> vs(J_ij ? (size_t)std::max<int64_t>(nnz, 1) : 0), hs(field_mode == 2 ? (size_t)n : 0);
    for (int64_t k = 0; k < n; ++k) {
        const int64_t i = perm[(size_t)k];
        rp[(size_t)k + 1] = rp[(size_t)k] + (rowptr[i + 1] - rowptr[i]);
        for (int64_t p = rowptr[i], q = rp[(size_t)k]; p < rowptr[i + 1]; ++p, ++q) {   // the row keeps the reference's order
            cs[(size_t)q] = inv[(size_t)col[p]];
            if (J_ij) vs[(size_t)q] = J_ij[p];
        }
        if (field_mode == 2) hs[(size_t)k] = h_i[i];
    }
    cudaError_t e;
    const size_t sb = (size_t)n * (size_t)nchains;
    if ((e = cudaMalloc((void **)&g->d_spins, sb)) != cudaSuccess || (e = cudaMalloc((void **)&g->d_staging, sb)) != cudaSuccess ||
        (e = cudaMalloc((void **)&g->d_rowptr, sizeof(int64_t) * ((size_t)n + 1))) != cudaSuccess ||
        (e = cudaMalloc((void **)&g->d_col, sizeof(int32_t) * cs.size())) != cudaSuccess ||
        (e = cudaMalloc((void **)&g->d_perm, sizeof(int32_t) * (size_t)n)) != cudaSuccess ||
        (e = cudaMalloc((void **)&g->d_acc, sizeof(long long) * (size_t)nchains)) != cudaSuccess ||
        (e = cudaMalloc((void **)&g->d_obs, sizeof(double) * 3 * (size_t)nchains)) != cudaSuccess ||
        (J_ij && (e = cudaMalloc((void **)&g->d_val, sizeof(double) * vs.size())) != cudaSuccess) ||
        (field_mode == 2 && (e = cudaMalloc((void **)&g->d_h, sizeof(double) * (size_t)n)) != cudaSuccess)) {
        graph_free(g);
        return gfail(MCX_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    cudaMemcpy(g->d_rowptr, rp.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(g->d_col, cs.data(), sizeof(int32_t) * cs.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(g->d_perm, perm.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice);
    if (J_ij) cudaMemcpy(g->d_val, vs.data(), sizeof(double) * vs.size(), cudaMemcpyHostToDevice);
    if (field_mode == 2) cudaMemcpy(g->d_h, hs.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemsetAsync(g->d_acc, 0, sizeof(long long) * (size_t)nchains, ctx->stream);
    cudaMemsetAsync(g->d_spins, 1, sb, ctx->stream);                    // constructors start all-up (ising.jl:118, 271)
    if ((e = cudaGetLastError()) != cudaSuccess) { graph_free(g); return gfail(MCX_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)); }
    *out = g;
    return MCX_OK;
}

int32_t mcx_graph_destroy(mcx_graph *g)
{
    if (!g) return MCX_OK;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    graph_free(g);
    return MCX_OK;
}

int32_t mcx_graph_colours(mcx_graph *g, int32_t *ncolours, int32_t *colour_of_site)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    if (ncolours) *ncolours = g->ncolours;
    if (colour_of_site) memcpy(colour_of_site, g->colour.data(), sizeof(int32_t) * (size_t)g->n);
    return MCX_OK;
}

int32_t mcx_graph_upload(mcx_graph *g, const int8_t *host_spins)
{
    GREQ(g && host_spins, MCX_ERR_ARGUMENT, "NULL argument");
    GCUDA(cudaSetDevice(g->ctx->device));
    GCUDA(cudaMemcpyAsync(g->d_staging, host_spins, (size_t)g->n * (size_t)g->nchains, cudaMemcpyHostToDevice, g->ctx->stream));
    k_graph_permute_in<<<graph_grid(g, g->n, 256), 256, 0, g->ctx->stream>>>(g->d_staging, g->d_spins, g->d_perm, g->n);
    g->ctx->launches++;
    GCUDA(cudaGetLastError());
    return MCX_OK;
}

int32_t mcx_graph_download(mcx_graph *g, int8_t *host_spins)
{
    GREQ(g && host_spins, MCX_ERR_ARGUMENT, "NULL argument");
    GCUDA(cudaSetDevice(g->ctx->device));
    k_graph_permute_out<<<graph_grid(g, g->n, 256), 256, 0, g->ctx->stream>>>(g->d_spins, g->d_staging, g->d_perm, g->n);
    g->ctx->launches++;
    GCUDA(cudaMemcpyAsync(host_spins, g->d_staging, (size_t)g->n * (size_t)g->nchains, cudaMemcpyDeviceToHost, g->ctx->stream));
    GCUDA(cudaStreamSynchronize(g->ctx->stream));
    return MCX_OK;
}

int32_t mcx_graph_init(mcx_graph *g, int32_t mode, uint64_t seed)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    GREQ(mode == MCX_INIT_UP || mode == MCX_INIT_DOWN || mode == MCX_INIT_RANDOM, MCX_ERR_ARGUMENT, "Unknown initialization type: %d", mode);
    GCUDA(cudaSetDevice(g->ctx->device));
    k_graph_init<<<graph_grid(g, g->n, 256), 256, 0, g->ctx->stream>>>(g->d_spins, g->d_perm, g->n, mode, (uint32_t)seed,
                                                                       (uint32_t)(seed >> 32), g->first_chain);
    g->ctx->launches++;
    GCUDA(cudaGetLastError());
    return MCX_OK;
}

int32_t mcx_graph_set_rule(mcx_graph *g, int32_t rule, double beta)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    GREQ(rule >= MCX_METROPOLIS && rule <= MCX_HEATBATH, MCX_ERR_ARGUMENT, "unknown rule %d", rule);
    g->rule = rule; g->beta = beta;
    return MCX_OK;
}

int32_t mcx_graph_set_rng(mcx_graph *g, uint64_t seed, uint64_t next_sweep, uint32_t first_chain_id)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    GREQ(next_sweep < ((uint64_t)1 << 40), MCX_ERR_ARGUMENT, "sweep counter exceeds the time field");
    g->seed = seed; g->sweep = next_sweep; g->first_chain = first_chain_id;
    return MCX_OK;
}

int32_t mcx_graph_get_rng(mcx_graph *g, uint64_t *seed, uint64_t *next_sweep)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    if (seed) *seed = g->seed;
    if (next_sweep) *next_sweep = g->sweep;
    return MCX_OK;
}

int32_t mcx_graph_sweep(mcx_graph *g, int64_t nsweeps)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    GREQ(nsweeps >= 0, MCX_ERR_ARGUMENT, "nsweeps must be >= 0");
    GREQ(g->rule >= 0, MCX_ERR_STATE, "no update rule set: call mcx_graph_set_rule first");
    GCUDA(cudaSetDevice(g->ctx->device));
    for (int64_t s = 0; s < nsweeps; ++s) {
        for (int c = 0; c < g->ncolours; ++c) {
            const int64_t off = g->class_off[(size_t)c], cnt = g->class_off[(size_t)c + 1] - off;
            const uint64_t t = (uint64_t)g->ncolours * g->sweep + (uint64_t)c;
            if (g->matrix)
                k_graph_colour<true><<<graph_grid(g, cnt, 128), 128, 0, g->ctx->stream>>>(
                    g->d_spins, g->n, g->d_rowptr, g->d_col, g->d_val, g->d_h, g->hmode, g->h, g->J, g->rule, g->beta, off, cnt,
                    (uint32_t)g->seed, (uint32_t)(g->seed >> 32), g->first_chain, t, g->d_acc);
            else
                k_graph_colour<false><<<graph_grid(g, cnt, 128), 128, 0, g->ctx->stream>>>(
                    g->d_spins, g->n, g->d_rowptr, g->d_col, g->d_val, g->d_h, g->hmode, g->h, g->J, g->rule, g->beta, off, cnt,
                    (uint32_t)g->seed, (uint32_t)(g->seed >> 32), g->first_chain, t, g->d_acc);
            g->ctx->launches++;
        }
        g->sweep += 1;
    }
    g->steps += nsweeps * g->n;
    GCUDA(cudaGetLastError());
    return MCX_OK;
}

// sum_pair_interactions (with J), sum_spins, sum_field_interactions, alg.accepted, alg.steps per chain; synchronises
int32_t mcx_graph_observables(mcx_graph *g, double *pair_sum, int64_t *spin_sum, double *field_sum, int64_t *accepted, int64_t *steps)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    GCUDA(cudaSetDevice(g->ctx->device));
    if (g->matrix) k_graph_sums<true><<<g->nchains, 256, 0, g->ctx->stream>>>(g->d_spins, g->n, g->d_rowptr, g->d_col, g->d_val, g->d_h, g->d_obs);
    else k_graph_sums<false><<<g->nchains, 256, 0, g->ctx->stream>>>(g->d_spins, g->n, g->d_rowptr, g->d_col, g->d_val, g->d_h, g->d_obs);
    g->ctx->launches++;
    std::vector<double> obs((size_t)g->nchains * 3);
    std::vector<long long> acc((size_t)g->nchains);
    GCUDA(cudaMemcpyAsync(obs.data(), g->d_obs, sizeof(double) * obs.size(), cudaMemcpyDeviceToHost, g->ctx->stream));
    GCUDA(cudaMemcpyAsync(acc.data(), g->d_acc, sizeof(long long) * acc.size(), cudaMemcpyDeviceToHost, g->ctx->stream));
    GCUDA(cudaStreamSynchronize(g->ctx->stream));
    GCUDA(cudaGetLastError());
    for (int c = 0; c < g->nchains; ++c) {
        const double spin = obs[(size_t)c * 3 + 2];
        if (pair_sum) pair_sum[c] = g->matrix ? obs[(size_t)c * 3] : g->J * obs[(size_t)c * 3];      // _pair_sum (ising.jl:163-169, 307-313)
        if (spin_sum) spin_sum[c] = (int64_t)spin;
        if (field_sum) field_sum[c] = g->hmode == 0 ? 0.0 : g->hmode == 1 ? g->h * spin : obs[(size_t)c * 3 + 1];
        if (accepted) accepted[c] = acc[(size_t)c];
        if (steps) steps[c] = g->steps;
    }
    return MCX_OK;
}

int32_t mcx_graph_energies(mcx_graph *g, double *energy)
{
    GREQ(g && energy, MCX_ERR_ARGUMENT, "NULL argument");
    std::vector<double> pair((size_t)g->nchains), field((size_t)g->nchains);
    const int32_t st = mcx_graph_observables(g, pair.data(), nullptr, field.data(), nullptr, nullptr);
    if (st != MCX_OK) return st;
    for (int c = 0; c < g->nchains; ++c) energy[c] = -pair[(size_t)c] - field[(size_t)c];       // _full_energy (ising.jl:185, 337)
    return MCX_OK;
}

int32_t mcx_graph_reset_counters(mcx_graph *g)
{
    GREQ(g, MCX_ERR_ARGUMENT, "graph is NULL");
    GCUDA(cudaSetDevice(g->ctx->device));
    GCUDA(cudaMemsetAsync(g->d_acc, 0, sizeof(long long) * (size_t)g->nchains, g->ctx->stream));
    g->steps = 0;
    return MCX_OK;
}

}  // extern "C"
