// k_series.cu -- integrated autocorrelation time of an on-device measurement series
// (src/measurements/autocorrelations.jl:28-65), computed where mcx_sweep_series left its snapshots: the series of a
// whole batch of chains is reduced to one double per chain without crossing PCIe.
//
// One block per chain.  The reference's estimator: x centred on its mean, C(0) = <x, x> / n,
// C(lag) = <x[1:n-lag], x[1+lag:n]> / (n - lag) / C(0), tau = 1/2 + sum C(lag), stopping at the first C <= 0 or when
// lag > c * tau (self-consistent window), never below 1/2; zero variance gives 1/2.  The dot products are block
// reductions in a fixed order (deterministic; the last bits differ from a sequential BLAS dot).
#include <cstdarg>
#include <cstdio>

#include "mcx_internal.h"

int32_t mcx_set_error(int32_t code, const char *msg);   // mcx_api.cu

namespace mcx {
namespace {

constexpr int kTauThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *s_red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kTauThreads / 32; ++w) t += s_red[w];      // same order on every thread
    return t;
}

// observable: 0 energy = -J pair - h spin + D spin2, 1 magnetisation = sum s, 2 |sum s|
__global__ void __launch_bounds__(kTauThreads)
k_tau_int(const long long *__restrict__ series, int64_t n, int nchains, int observable, double J, double h, double D, int is_ising,
          int64_t N, int64_t lag_cap, double c, double *__restrict__ scratch, double *__restrict__ tau_out)
{
    __shared__ double s_red[kTauThreads / 32];
    const int chain = blockIdx.x;
    double *x = scratch + (int64_t)chain * n;
    double acc = 0.0;
    for (int64_t k = threadIdx.x; k < n; k += kTauThreads) {
        const long long *r = series + (k * nchains + chain) * SUM_FIELDS;
        double v;
        if (observable == 0) {
            v = -(J * (double)r[SUM_PAIR]);
            if (h != 0.0) v -= h * (double)r[SUM_SPIN];
            if (!is_ising) v += D * (double)r[SUM_SPIN2];
        } else {
            v = (double)r[SUM_SPIN];
            if (observable == 2) v = fabs(v);
        }
        x[k] = v;
        acc += v;
    }
    const double mu = block_sum(acc, s_red) / (double)n;
    acc = 0.0;
    for (int64_t k = threadIdx.x; k < n; k += kTauThreads) {
        const double d = x[k] - mu;
        x[k] = d;
        acc += d * d;
    }
    const double C0 = block_sum(acc, s_red) / (double)n;
    double tau = 0.5;
    if (C0 > 0) {
        for (int64_t lag = 1; lag <= lag_cap; ++lag) {
            acc = 0.0;
            for (int64_t k = threadIdx.x; k < n - lag; k += kTauThreads) acc += x[k] * x[k + lag];
            const double C = block_sum(acc, s_red) / (double)(n - lag) / C0;
            if (C <= 0) break;
            const double tau_next = tau + C;
            if ((double)lag > c * tau_next) break;
            tau = tau_next;
        }
    }
    if (threadIdx.x == 0) tau_out[chain] = tau > 0.5 ? tau : 0.5;
}

}  // namespace
}  // namespace mcx

using namespace mcx;

extern "C" int32_t mcx_series_tau_int(mcx_lattice *lat, int64_t nmeasure, int32_t observable, int64_t max_lag, double c, double *tau)
{
    if (!lat || !tau) return mcx_set_error(MCX_ERR_ARGUMENT, "NULL argument");
    if (nmeasure < 2) return mcx_set_error(MCX_ERR_ARGUMENT, "integrated_autocorrelation_time requires at least 2 samples");
    if (!(c > 0)) return mcx_set_error(MCX_ERR_ARGUMENT, "c must be positive");
    if (observable < 0 || observable > 2) return mcx_set_error(MCX_ERR_ARGUMENT, "observable must be 0 (energy), 1 (magnetisation) or 2 (|magnetisation|)");
    if (nmeasure > lat->series_n) return mcx_set_error(MCX_ERR_STATE, "no series of that length on the device: run mcx_sweep_series first");
    const int64_t cap_max = nmeasure / 2;
    const int64_t cap = max_lag == 0 ? cap_max : max_lag;
    if (cap < 1 || cap > cap_max) return mcx_set_error(MCX_ERR_ARGUMENT, "max_lag must satisfy 1 <= max_lag <= floor(length(samples)/2)");
    cudaSetDevice(lat->ctx->device);
    const size_t need = sizeof(double) * ((size_t)nmeasure + 1) * (size_t)lat->nchains;
    if (lat->tau_bytes < need) {
        cudaStreamSynchronize(lat->ctx->stream);
        cudaFree(lat->d_tau);
        lat->d_tau = nullptr; lat->tau_bytes = 0;
        if (cudaMalloc((void **)&lat->d_tau, need) != cudaSuccess) { cudaGetLastError(); return mcx_set_error(MCX_ERR_CUDA, "scratch allocation failed"); }
        lat->tau_bytes = need;
    }
    double *d_out = lat->d_tau + (size_t)nmeasure * (size_t)lat->nchains;
    k_tau_int<<<lat->nchains, kTauThreads, 0, lat->ctx->stream>>>(lat->d_series, nmeasure, lat->nchains, observable, lat->J, lat->h, lat->D,
                                                                 lat->model == MCX_ISING ? 1 : 0, lat->N, cap, c, lat->d_tau, d_out);
    lat->ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(tau, d_out, sizeof(double) * (size_t)lat->nchains, cudaMemcpyDeviceToHost, lat->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(lat->ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return mcx_set_error(MCX_ERR_CUDA, cudaGetErrorString(e));
    return MCX_OK;
}
