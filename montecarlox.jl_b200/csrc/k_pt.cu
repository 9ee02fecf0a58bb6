// k_pt.cu -- replica exchange on the device.
//
// Restates update!(rx::ReplicaExchange{ThreadsBackend}, xs) (src/algorithms/replica_exchange.jl:158-178):
// the pairs of one stage are disjoint, so they are decided in parallel; ensembles (labels) move
// between slots, lattices stay (:133).  Every rank holds the full ladder state and runs this
// kernel on the same all-gathered energies, so all ranks reach identical decisions without the
// pairwise send/recv + Allgather(new_index) of the MPI backend (:193-244).
#include "mcx_internal.h"

namespace mcx {

// energy(sys) per local replica from the integer sums (ising.jl:175-178, blume_capel.jl:222-224)
__global__ void k_pt_publish(const long long *__restrict__ sums, double *__restrict__ x, int nlocal, int first_slot,
                             double J, double h, double D, int model)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nlocal) return;
    const long long *s = sums + (int64_t)c * SUM_FIELDS;
    double e = -(J * (double)s[SUM_PAIR]);
    if (h != 0.0) e -= h * (double)s[SUM_SPIN];
    if (model == MCX_BLUME_CAPEL) e += D * (double)s[SUM_SPIN2];
    x[first_slot + c] = e;
}

// thread 0 of a block: wait until every rank has published round `value`; 20 s give up and raise *err
__device__ __forceinline__ void pt_wait_all(const unsigned long long *arrived, int nranks, unsigned long long value, int *err,
                                            int *ctx_err)
{
    const volatile unsigned long long *a = arrived;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int r = 0; r < nranks; ++r)
        while (a[r] < value) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 20000000000ull) {
                *err = 1;
                if (ctx_err) { *(volatile int *)ctx_err = ASYNC_ERR_PT_PEERS; __threadfence_system(); }   // the host's next call fails
                return;
            }
        }
    __threadfence_system();
}

// clock != nullptr (a round replayed from a CUDA graph): the round comes from the device clock; `stage` then holds the
// parity of stage - round, which no exchange changes, and `x` the base of the two round-parity buffers (xstride apart).
__global__ void k_pt_exchange(int n, int stage, uint64_t round, const double *__restrict__ betas,
                              const double *x, int32_t *__restrict__ index,
                              int32_t *__restrict__ slot_of, long long *__restrict__ steps,
                              long long *__restrict__ accepted, int32_t *__restrict__ labels, int nlocal,
                              int first_slot, uint32_t seed_lo, uint32_t seed_hi, const unsigned long long *arrived,
                              int nranks, int *err, int *ctx_err, const PtClock *clock, int xstride)
{
    if (clock) {
        round = *(const volatile unsigned long long *)&clock->round;
        stage = (int)((round + (uint64_t)stage) & 1);
        x += (size_t)(round & 1) * (size_t)xstride;
    }
    if (arrived) {      // energies arrive by peer stores: wait for every rank's publish of this round
        if (threadIdx.x == 0) pt_wait_all(arrived, nranks, round + 1, err, ctx_err);
        __syncthreads();
    }
    // 0-based pair k joins ladder indices k and k+1; stage 0 takes k = 0,2,4,.. (reference first=1)
    for (int k = (stage & 1) + 2 * (blockIdx.x * blockDim.x + threadIdx.x); k < n - 1;
         k += 2 * blockDim.x * gridDim.x) {
        const int ri = slot_of[k], rj = slot_of[k + 1];
        steps[k] += 1;
        // u = rand(algorithm(rx, ri).rng): EXCHANGE stream of slot ri at this round, 53 bits
        const Philox4 p = stream_block(seed_lo, seed_hi, (uint32_t)ri, TAG_EXCHANGE, round, 0, 0);
        const uint64_t w = ((uint64_t)p.y << 32) | p.x;
        const double u = (double)(w >> 11) * (1.0 / 9007199254740992.0);
        const double bi = betas[k], bj = betas[k + 1];
        const double xi = ((const volatile double *)x)[ri], xj = ((const volatile double *)x)[rj];
        // exchange_log_ratio (:110-113) with logweight(E) = -beta*E (ensembles/boltzmann.jl:28)
        const double lr = ((-bi * xj) - (-bi * xi)) + ((-bj * xi) - (-bj * xj));
        const bool acc = (lr > 0) || (u < exp(lr));
        if (acc) {
            accepted[k] += 1;
            index[ri] = k + 1; index[rj] = k;
            slot_of[k] = rj; slot_of[k + 1] = ri;
            if (ri >= first_slot && ri < first_slot + nlocal) labels[ri - first_slot] = k + 1;
            if (rj >= first_slot && rj < first_slot + nlocal) labels[rj - first_slot] = k;
        }
    }
}

// The same energies stored straight into every rank's buffer (peer memory over NVLink), then this rank's
// arrival counter bumped on every rank: the all-gather of replica_exchange.jl:239 fused into the publish.
__global__ void k_pt_publish_peers(const long long *__restrict__ sums, double *const *__restrict__ peer_x,
                                   unsigned long long *const *__restrict__ peer_arrived, int nranks, int rank, int nlocal,
                                   int first_slot, int offset, unsigned long long value, double J, double h, double D, int model,
                                   const PtClock *clock, int n)
{
    if (clock) {                                               // graph replay: the round parity buffer and the arrival value of this round
        const unsigned long long round = *(const volatile unsigned long long *)&clock->round;
        offset = (int)(round & 1) * n;
        value = round + 1;
    }
    for (int c = threadIdx.x; c < nlocal; c += blockDim.x) {
        const long long *s = sums + (int64_t)c * SUM_FIELDS;
        double e = -(J * (double)s[SUM_PAIR]);
        if (h != 0.0) e -= h * (double)s[SUM_SPIN];
        if (model == MCX_BLUME_CAPEL) e += D * (double)s[SUM_SPIN2];
        for (int r = 0; r < nranks; ++r) peer_x[r][offset + first_slot + c] = e;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < nranks) {
        *(volatile unsigned long long *)(peer_arrived[threadIdx.x] + rank) = value;
        __threadfence_system();
    }
}

// The tail of a graph-replayed round in ONE launch of one block: publish (k_pt_publish / k_pt_publish_peers), the wait for
// the other ranks, the exchange (k_pt_exchange) and the clock advance -- three dependent launches fewer per round, which at
// an exchange after every sweep is a sixth of the round.  Same expressions, same order per pair as the separate kernels.
__global__ void __launch_bounds__(256)
k_pt_round_tail(const long long *__restrict__ sums, double *const *__restrict__ peer_x, unsigned long long *const *__restrict__ peer_arrived,
                const unsigned long long *arrived, int nranks, int rank, int nlocal, int first_slot, double J, double h, double D,
                int model, PtClock *clock, int n, int stage_par, const double *__restrict__ betas, double *x_base,
                int32_t *__restrict__ index, int32_t *__restrict__ slot_of, long long *__restrict__ steps,
                long long *__restrict__ accepted, int32_t *__restrict__ labels, uint32_t seed_lo, uint32_t seed_hi, int *err,
                int *ctx_err, unsigned long long dt)
{
    const unsigned long long round = *(const volatile unsigned long long *)&clock->round;
    const bool peers = peer_x != nullptr;
    const int offset = peers ? (int)(round & 1) * n : 0;
    for (int c = threadIdx.x; c < nlocal; c += blockDim.x) {
        const long long *s = sums + (int64_t)c * SUM_FIELDS;
        double e = -(J * (double)s[SUM_PAIR]);
        if (h != 0.0) e -= h * (double)s[SUM_SPIN];
        if (model == MCX_BLUME_CAPEL) e += D * (double)s[SUM_SPIN2];
        if (peers) { for (int r = 0; r < nranks; ++r) peer_x[r][offset + first_slot + c] = e; }
        else x_base[first_slot + c] = e;
    }
    if (peers) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < nranks) {
            *(volatile unsigned long long *)(peer_arrived[threadIdx.x] + rank) = round + 1;
            __threadfence_system();
        }
        if (threadIdx.x == 0) pt_wait_all(arrived, nranks, round + 1, err, ctx_err);
    } else {
        __threadfence();
    }
    __syncthreads();
    const int stage = (int)((round + (unsigned long long)stage_par) & 1);
    const double *x = x_base + offset;
    for (int k = (stage & 1) + 2 * (int)threadIdx.x; k < n - 1; k += 2 * (int)blockDim.x) {
        const int ri = slot_of[k], rj = slot_of[k + 1];
        steps[k] += 1;
        const Philox4 p = stream_block(seed_lo, seed_hi, (uint32_t)ri, TAG_EXCHANGE, round, 0, 0);
        const uint64_t w = ((uint64_t)p.y << 32) | p.x;
        const double u = (double)(w >> 11) * (1.0 / 9007199254740992.0);
        const double bi = betas[k], bj = betas[k + 1];
        const double xi = ((const volatile double *)x)[ri], xj = ((const volatile double *)x)[rj];
        const double lr = ((-bi * xj) - (-bi * xi)) + ((-bj * xi) - (-bj * xj));
        const bool acc = (lr > 0) || (u < exp(lr));
        if (acc) {
            accepted[k] += 1;
            index[ri] = k + 1; index[rj] = k;
            slot_of[k] = rj; slot_of[k + 1] = ri;
            if (ri >= first_slot && ri < first_slot + nlocal) labels[ri - first_slot] = k + 1;
            if (rj >= first_slot && rj < first_slot + nlocal) labels[rj - first_slot] = k;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        clock->t_base += dt;
        clock->round = round + 1;
    }
}

void launch_pt_round_tail(mcx_pt *pt, int64_t sweeps)
{
    mcx_lattice *lat = pt->lat;
    const int stage_par = (int)(((uint64_t)pt->stage + pt->round) & 1);
    k_pt_round_tail<<<1, 256, 0, lat->ctx->stream>>>(lat->d_sums, pt->peers ? pt->d_peer_x : nullptr, pt->d_peer_arrived, pt->d_arrived,
                                                   pt->nranks, pt->rank, lat->nchains, pt->first_slot, lat->J, lat->h, lat->D, lat->model,
                                                   pt->d_clock, pt->n, stage_par, pt->d_betas, pt->d_x, pt->d_index, pt->d_slot_of,
                                                   pt->d_steps, pt->d_accepted, lat->d_labels, (uint32_t)lat->seed,
                                                   (uint32_t)(lat->seed >> 32), pt->d_err, lat->ctx->d_err, 2 * (unsigned long long)sweeps);
    lat->ctx->launches++;
}

__global__ void k_pt_clock_set(PtClock *clock, unsigned long long t_base, unsigned long long round)
{
    clock->t_base = t_base; clock->round = round;
}
__global__ void k_pt_clock_advance(PtClock *clock, unsigned long long dt)
{
    clock->t_base += dt; clock->round += 1;
}

void launch_pt_clock_set(mcx_pt *pt)
{
    k_pt_clock_set<<<1, 1, 0, pt->lat->ctx->stream>>>(pt->d_clock, 2 * pt->lat->sweep, pt->round);
    pt->lat->ctx->launches++;
}
void launch_pt_clock_advance(mcx_pt *pt, int64_t sweeps)
{
    k_pt_clock_advance<<<1, 1, 0, pt->lat->ctx->stream>>>(pt->d_clock, 2 * (unsigned long long)sweeps);
    pt->lat->ctx->launches++;
}

void launch_pt_publish(mcx_pt *pt, const PtClock *clock)
{
    if (pt->peers) {
        mcx_lattice *lat = pt->lat;
        k_pt_publish_peers<<<1, 256, 0, lat->ctx->stream>>>(lat->d_sums, pt->d_peer_x, pt->d_peer_arrived, pt->nranks, pt->rank,
                                                          lat->nchains, pt->first_slot, (int)(pt->round & 1) * pt->n,
                                                          pt->round + 1, lat->J, lat->h, lat->D, lat->model, clock, pt->n);
        lat->ctx->launches++;
        return;
    }
    {
    mcx_lattice *lat = pt->lat;
    k_pt_publish<<<(lat->nchains + 127) / 128, 128, 0, lat->ctx->stream>>>(
        lat->d_sums, pt->d_x, lat->nchains, pt->first_slot, lat->J, lat->h, lat->D, lat->model);
    lat->ctx->launches++;
    }
}

void launch_pt_exchange(mcx_pt *pt, const PtClock *clock)
{
    mcx_lattice *lat = pt->lat;
    const int npairs = (pt->n - 1 + 1) / 2;
    const int blocks = npairs > 0 ? (npairs + 127) / 128 : 1;
    // graph replay: stage - round keeps its parity, the kernel rebuilds the stage and the parity buffer from the clock's round
    const int stage = clock ? (int)(((uint64_t)pt->stage + pt->round) & 1) : pt->stage;
    const double *x = clock ? pt->d_x : pt->d_x + (pt->peers ? (pt->round & 1) * pt->n : 0);
    k_pt_exchange<<<blocks, 128, 0, lat->ctx->stream>>>(
        pt->n, stage, pt->round, pt->d_betas, x, pt->d_index, pt->d_slot_of,
        pt->d_steps, pt->d_accepted, lat->d_labels, lat->nchains, pt->first_slot, (uint32_t)lat->seed,
        (uint32_t)(lat->seed >> 32), pt->peers ? pt->d_arrived : nullptr, pt->nranks, pt->d_err, lat->ctx->d_err, clock,
        pt->peers ? pt->n : 0);
    lat->ctx->launches++;
}

}  // namespace mcx
