/*
 * mcx_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See mcx_oracle.h for scope and parity status.  Citations are file:line
 * relative to /root/reference (MonteCarloX.jl + SpinSystems, Julia).
 *
 * The reference updates one random site per call (SpinSystems/src/ising.jl:25-58).
 * The "Philox-driven reference" the CUDA kernels must match is defined as: the
 * reference's per-site primitives (flip_changes, delta_energy, accept!, modify!)
 * called in checkerboard order, with an injected counter-based AbstractRNG that
 * is positioned at (chain, sweep, colour, site) before every attempt.  That is
 * mode 1 below.  Mode 2 is the reference's own random-site loop.
 */
#include "mcx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

/* ===================================================================== */
/* Philox4x32-10 (Salmon et al., SC'11; constants as in                   */
/* /usr/local/cuda/include/curand_philox4x32_x.h:88-91)                   */
/* ===================================================================== */
void mcxo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream addressing (DESIGN.md "RNG layout v1"):
 *   key = (seed lo, seed hi)
 *   ctr = ( q >> 3,  t lo32,  (t>>32 & 0xffff) | plane<<16 | tag<<24,  chain )
 *   the 16-bit lane of slot q is  (out[(q&7)>>1] >> 16*((q&7)&1)) & 0xffff
 * Draw slot n of a position uses plane 2n (high half) and plane 2n+1 (low half). */
static void stream_block(uint64_t seed, uint32_t chain, uint32_t tag, uint64_t t, uint64_t blk,
                         uint32_t plane, uint32_t out[4])
{
    uint32_t ctr[4], key[2];
    ctr[0] = (uint32_t)blk;
    ctr[1] = (uint32_t)t;
    ctr[2] = (uint32_t)((t >> 32) & 0xffffu) | (plane << 16) | (tag << 24);
    ctr[3] = chain;
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    mcxo_philox4x32_10(ctr, key, out);
}

void mcxo_rng_position(mcxo_rng *r, uint32_t tag, uint64_t t, uint64_t q)
{
    r->tag = tag; r->t = t; r->q = q; r->draw = 0;
}

uint32_t mcxo_rng_lane16(const mcxo_rng *r, uint32_t plane)
{
    uint32_t out[4];
    stream_block(r->seed, r->chain, r->tag, r->t, r->q >> 3, plane, out);
    uint32_t lane = (uint32_t)(r->q & 7);
    return (out[lane >> 1] >> (16 * (lane & 1))) & 0xffffu;
}

/* rand(rng)::Float64 for PhiloxRNG: 32 random bits, u = m * 2^-32 (exact in Float64).
 * Sites the reference draws at: importance_sampling.jl:82, metropolis.jl:124,
 * ising.jl:50, blume_capel.jl:75. */
double mcxo_rand_f64(mcxo_rng *r)
{
    uint32_t hi = mcxo_rng_lane16(r, 2 * r->draw);
    uint32_t lo = mcxo_rng_lane16(r, 2 * r->draw + 1);
    r->draw += 1;
    return (double)((hi << 16) | lo) * (1.0 / 4294967296.0);
}

/* rand(rng, Bool) for PhiloxRNG: top bit of the high half (blume_capel.jl:22). */
int mcxo_rand_bool(mcxo_rng *r)
{
    uint32_t hi = mcxo_rng_lane16(r, 2 * r->draw);
    r->draw += 1;
    return (int)(hi >> 15);
}

/* The u of replica_exchange.jl:168: one 53-bit Float64 from slot `chain`'s
 * EXCHANGE stream at exchange round `round`. */
double mcxo_exchange_u(uint64_t seed, uint32_t chain, uint64_t round)
{
    uint32_t out[4];
    stream_block(seed, chain, MCXO_TAG_EXCHANGE, round, 0, 0, out);
    uint64_t w = ((uint64_t)out[1] << 32) | out[0];
    return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}

/* ===================================================================== */
/* xoshiro256++ (what Julia's Xoshiro is), splitmix64-seeded.  Only used  */
/* for mode 2 (CPU baseline and statistical cross-checks).                */
/* ===================================================================== */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
void mcxo_xoshiro_seed(mcxo_xoshiro *x, uint64_t seed)
{
    for (int i = 0; i < 4; ++i) {
        uint64_t z = (seed += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        x->s[i] = z ^ (z >> 31);
    }
}
uint64_t mcxo_xoshiro_next(mcxo_xoshiro *x)
{
    uint64_t *s = x->s;
    uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
static inline double xoshiro_f64(mcxo_xoshiro *x)
{
    return (double)(mcxo_xoshiro_next(x) >> 11) * (1.0 / 9007199254740992.0);
}

/* ===================================================================== */
/* systems                                                                */
/* ===================================================================== */
/* Neighbour table for a periodic hyper-cubic lattice, x fastest.  Restates the
 * explicit construction of IsingLatticeOptim (ising.jl:444-456: neighbour order
 * left, right, up(y-1), down(y+1)), extended to 3-D as i = x + Lx*(y + Ly*z). */
mcxo_system *mcxo_system_create(int model, int ndim, const int64_t *dims, double J, double h, double D)
{
    mcxo_system *s = (mcxo_system *)calloc(1, sizeof(*s));
    s->model = model; s->ndim = ndim; s->nn = 2 * ndim;
    s->dims[0] = dims[0]; s->dims[1] = ndim > 1 ? dims[1] : 1; s->dims[2] = ndim > 2 ? dims[2] : 1;
    s->N = s->dims[0] * s->dims[1] * s->dims[2];
    s->J = J; s->h = h; s->D = D;
    s->spins = (int8_t *)malloc((size_t)s->N);
    s->nbr = (int64_t *)malloc(sizeof(int64_t) * (size_t)s->N * (size_t)s->nn);
    int64_t Lx = s->dims[0], Ly = s->dims[1], Lz = s->dims[2];
    for (int64_t z = 0; z < Lz; ++z)
        for (int64_t y = 0; y < Ly; ++y)
            for (int64_t x = 0; x < Lx; ++x) {
                int64_t i = x + Lx * (y + Ly * z);
                int64_t *nb = s->nbr + i * s->nn;
                int64_t xl = (x == 0) ? Lx - 1 : x - 1, xr = (x == Lx - 1) ? 0 : x + 1;
                nb[0] = xl + Lx * (y + Ly * z);
                nb[1] = xr + Lx * (y + Ly * z);
                if (ndim > 1) {
                    int64_t yu = (y == 0) ? Ly - 1 : y - 1, yd = (y == Ly - 1) ? 0 : y + 1;
                    nb[2] = x + Lx * (yu + Ly * z);
                    nb[3] = x + Lx * (yd + Ly * z);
                }
                if (ndim > 2) {
                    int64_t zu = (z == 0) ? Lz - 1 : z - 1, zd = (z == Lz - 1) ? 0 : z + 1;
                    nb[4] = x + Lx * (y + Ly * zu);
                    nb[5] = x + Lx * (y + Ly * zd);
                }
            }
    /* constructors start all-up (ising.jl:118, blume_capel.jl:156) */
    memset(s->spins, 1, (size_t)s->N);
    mcxo_recompute(s);
    return s;
}

void mcxo_system_destroy(mcxo_system *s)
{
    if (!s) return;
    free(s->spins); free(s->nbr); free(s);
}

void mcxo_system_set_spins(mcxo_system *s, const int8_t *spins)
{
    memcpy(s->spins, spins, (size_t)s->N);
    mcxo_recompute(s);
}

void mcxo_system_get_spins(const mcxo_system *s, int8_t *spins) { memcpy(spins, s->spins, (size_t)s->N); }

/* init!(sys, :random; rng) (ising.jl:60-78, blume_capel.jl:87-110) with the PhiloxRNG
 * INIT stream: Ising site i takes bit (i & 127) of block i>>7 (set -> +1);
 * Blume-Capel site i takes state (3 * lane16(i)) >> 16 from {-1,0,1}. */
void mcxo_system_init_random(mcxo_system *s, uint64_t seed, uint32_t chain)
{
    uint32_t out[4];
    if (s->model == MCXO_ISING) {
        for (int64_t i = 0; i < s->N; ++i) {
            if ((i & 127) == 0 || i == 0) stream_block(seed, chain, MCXO_TAG_INIT, 0, (uint64_t)i >> 7, 0, out);
            uint32_t bit = (out[(i >> 5) & 3] >> (i & 31)) & 1u;
            s->spins[i] = bit ? 1 : -1;
        }
    } else {
        for (int64_t i = 0; i < s->N; ++i) {
            if ((i & 7) == 0) stream_block(seed, chain, MCXO_TAG_INIT, 0, (uint64_t)i >> 3, 0, out);
            uint32_t lane = (uint32_t)(i & 7);
            uint32_t v = (out[lane >> 1] >> (16 * (lane & 1))) & 0xffffu;
            s->spins[i] = (int8_t)((int)((v * 3u) >> 16) - 1);
        }
    }
    mcxo_recompute(s);
}

/* local_pair_interactions: s_i * sum_j s_j (abstractions.jl:41-48, ising.jl:463-469) */
int64_t mcxo_local_pair_interactions(const mcxo_system *s, int64_t i)
{
    int64_t si = s->spins[i], acc = 0;
    const int64_t *nb = s->nbr + i * s->nn;
    for (int k = 0; k < s->nn; ++k) acc += si * s->spins[nb[k]];
    return acc;
}

int64_t mcxo_pair_count(const mcxo_system *s)
{
    int64_t acc = 0;
    for (int64_t i = 0; i < s->N; ++i) acc += mcxo_local_pair_interactions(s, i);
    return acc / 2;   /* div(pair_unweighted, 2): ising.jl:164, :476 */
}

int64_t mcxo_spin2_sum(const mcxo_system *s)
{
    int64_t acc = 0;
    for (int64_t i = 0; i < s->N; ++i) acc += (int64_t)s->spins[i] * s->spins[i];
    return acc;
}

/* _recompute_cached! (ising.jl:141-145,500-504; blume_capel.jl:207-212) */
void mcxo_recompute(mcxo_system *s)
{
    s->sum_pair = s->J * (double)mcxo_pair_count(s);
    int64_t m = 0;
    for (int64_t i = 0; i < s->N; ++i) m += s->spins[i];
    s->sum_spins = m;
    s->sum_spins2 = mcxo_spin2_sum(s);
}

/* energy(sys; full) (ising.jl:17, :175-178; blume_capel.jl:18, :222-226) */
double mcxo_energy(const mcxo_system *s, int full)
{
    double pair = full ? s->J * (double)mcxo_pair_count(s) : s->sum_pair;
    int64_t m = full ? mcxo_magnetization(s, 1) : s->sum_spins;
    double e = -pair - s->h * (double)m;
    if (s->model == MCXO_BLUME_CAPEL) e += s->D * (double)(full ? mcxo_spin2_sum(s) : s->sum_spins2);
    return e;
}

int64_t mcxo_magnetization(const mcxo_system *s, int full)
{
    if (!full) return s->sum_spins;
    int64_t m = 0;
    for (int64_t i = 0; i < s->N; ++i) m += s->spins[i];
    return m;
}

/* flip_changes + delta_energy for Ising (ising.jl:187-198, :484-491):
 *   dpair = -2*J*lpi ; dspin = -2*s ; dE = -dpair - h*dspin */
static inline void ising_flip_changes(const mcxo_system *s, int64_t i, double *dpair, int64_t *dspin)
{
    *dpair = (-2.0 * s->J) * (double)mcxo_local_pair_interactions(s, i);
    *dspin = -2 * (int64_t)s->spins[i];
}
double mcxo_delta_energy_flip(const mcxo_system *s, int64_t i)
{
    double dpair; int64_t dspin;
    ising_flip_changes(s, i, &dpair, &dspin);
    return -dpair - s->h * (double)dspin;
}

/* Blume-Capel: local_coupling, propose_changes, delta_energy (blume_capel.jl:227-248) */
static inline double bc_local_coupling(const mcxo_system *s, int64_t i)
{
    int64_t acc = 0;
    const int64_t *nb = s->nbr + i * s->nn;
    for (int k = 0; k < s->nn; ++k) acc += s->spins[nb[k]];
    return s->J * (double)acc;
}
static inline void bc_propose_changes(const mcxo_system *s, int64_t i, int s_new, double *dpair,
                                      int64_t *dspin, int64_t *dspin2)
{
    int s_old = s->spins[i];
    *dspin = s_new - s_old;
    *dspin2 = (int64_t)s_new * s_new - (int64_t)s_old * s_old;   /* _sqdiff :32-34 */
    *dpair = (double)(*dspin) * bc_local_coupling(s, i);
}
double mcxo_delta_energy_bc(const mcxo_system *s, int64_t i, int s_new)
{
    double dpair; int64_t dspin, dspin2;
    bc_propose_changes(s, i, s_new, &dpair, &dspin, &dspin2);
    return -dpair - s->h * (double)dspin + s->D * (double)dspin2;
}

/* logistic (src/infrastructure/utils.jl:82-88) */
double mcxo_logistic(double x)
{
    if (x >= 0) return 1.0 / (1.0 + exp(-x));
    double ex = exp(x);
    return ex / (1.0 + ex);
}

/* _propose_state (blume_capel.jl:21-30) */
int mcxo_propose_state(int u, int s_old)
{
    if (s_old == -1) return u ? 0 : 1;
    if (s_old == 0) return u ? -1 : 1;
    return u ? -1 : 0;
}

/* modify! (ising.jl:200-205,493-498; blume_capel.jl:250-256) */
static inline void ising_modify(mcxo_system *s, int64_t i, double dpair, int64_t dspin)
{
    s->spins[i] = (int8_t)(-s->spins[i]);
    s->sum_pair += dpair;
    s->sum_spins += dspin;
}
static inline void bc_modify(mcxo_system *s, int64_t i, int s_new, double dpair, int64_t dspin, int64_t dspin2)
{
    s->spins[i] = (int8_t)s_new;
    s->sum_pair += dpair;
    s->sum_spins += dspin;
    s->sum_spins2 += dspin2;
}

/* accept!(alg::AbstractMetropolis, dE) -> logweight(Boltzmann, dE) = -beta*dE (boltzmann.jl:28,
 * metropolis.jl:14-17) -> _accept! (importance_sampling.jl:80-85); Glauber (metropolis.jl:121-127). */
static inline int accept_delta(mcxo_alg *a, double dE, mcxo_rng *r)
{
    double log_ratio = -a->beta * dE;
    int accepted;
    a->steps += 1;
    if (a->rule == MCXO_GLAUBER) {
        accepted = mcxo_rand_f64(r) < mcxo_logistic(log_ratio);
    } else {
        accepted = (log_ratio > 0) || (mcxo_rand_f64(r) < exp(log_ratio));
    }
    a->accepted += accepted;
    return accepted;
}

/* One attempt at site i: spin_flip! without pick_site.
 * Ising:  Metropolis/Glauber ising.jl:35-41, heat bath ising.jl:43-58.
 * Blume-Capel: Metropolis/Glauber blume_capel.jl:52-59, heat bath :61-85. */
void mcxo_attempt_at(mcxo_system *s, mcxo_alg *a, int64_t i, mcxo_rng *r)
{
    if (s->model == MCXO_ISING) {
        double dpair; int64_t dspin;
        ising_flip_changes(s, i, &dpair, &dspin);
        double dE = -dpair - s->h * (double)dspin;
        if (a->rule == MCXO_HEATBATH) {
            int s_old = s->spins[i];
            double p_plus = mcxo_logistic(a->beta * (double)s_old * dE);
            int s_new = mcxo_rand_f64(r) < p_plus ? 1 : -1;
            if (s_new != s_old) ising_modify(s, i, dpair, dspin);
            a->steps += 1;
        } else {
            if (accept_delta(a, dE, r)) ising_modify(s, i, dpair, dspin);
        }
    } else {
        if (a->rule == MCXO_HEATBATH) {
            double coupling = bc_local_coupling(s, i);
            double h_i = s->h;
            double e1 = -(-1) * coupling - h_i * (-1) + s->D;
            double e2 = 0.0;
            double e3 = -(1) * coupling - h_i * (1) + s->D;
            double w1 = exp(-a->beta * e1), w2 = exp(-a->beta * e2), w3 = exp(-a->beta * e3);
            double z = w1 + w2 + w3;
            double rr = mcxo_rand_f64(r) * z;
            int s_new = rr < w1 ? -1 : (rr < (w1 + w2) ? 0 : 1);
            if (s_new != s->spins[i]) {
                double dpair; int64_t dspin, dspin2;
                bc_propose_changes(s, i, s_new, &dpair, &dspin, &dspin2);
                bc_modify(s, i, s_new, dpair, dspin, dspin2);
            }
            a->steps += 1;
        } else {
            int s_new = mcxo_propose_state(mcxo_rand_bool(r), s->spins[i]);
            double dpair; int64_t dspin, dspin2;
            bc_propose_changes(s, i, s_new, &dpair, &dspin, &dspin2);
            double dE = -dpair - s->h * (double)dspin + s->D * (double)dspin2;
            if (accept_delta(a, dE, r)) bc_modify(s, i, s_new, dpair, dspin, dspin2);
        }
    }
}

/* ---- mode 1: checkerboard order, stream positioned per (chain, sweep, colour, site) ---- */
void mcxo_sweep_checkerboard(mcxo_system *s, mcxo_alg *a, uint64_t seed, uint32_t chain,
                             uint64_t sweep0, int64_t nsweeps)
{
    int64_t Lx = s->dims[0], Ly = s->dims[1];
    int64_t half = Lx / 2;
    mcxo_rng r; r.seed = seed; r.chain = chain;
    for (int64_t sw = 0; sw < nsweeps; ++sw)
        for (int colour = 0; colour < 2; ++colour) {
            uint64_t t = 2 * (sweep0 + (uint64_t)sw) + (uint64_t)colour;
            for (int64_t i = 0; i < s->N; ++i) {
                int64_t x = i % Lx, row = i / Lx;          /* row = y + Ly*z */
                int64_t y = row % Ly, z = row / Ly;
                if (((x + y + z) & 1) != colour) continue;
                mcxo_rng_position(&r, MCXO_TAG_SWEEP, t, (uint64_t)(row * half + (x >> 1)));
                mcxo_attempt_at(s, a, i, &r);
            }
        }
}

/* ---- mode 2: the reference's own loop ---- */
/* pick_site: rand(rng, UInt) % N + 1 (abstractions.jl:19); accept with exp per attempt, or the
 * TableMetropolis of docs/src/examples/spin_systems/importance_Ising2D.jl:74-92 (use_table). */
void mcxo_sweep_random_site(mcxo_system *s, mcxo_alg *a, mcxo_xoshiro *x, int64_t nattempts, int use_table)
{
    double p4 = exp(-4 * a->beta), p8 = exp(-8 * a->beta);
    for (int64_t n = 0; n < nattempts; ++n) {
        int64_t i = (int64_t)(mcxo_xoshiro_next(x) % (uint64_t)s->N);
        if (s->model == MCXO_ISING && a->rule == MCXO_METROPOLIS) {
            double dpair; int64_t dspin;
            ising_flip_changes(s, i, &dpair, &dspin);
            double dE = -dpair - s->h * (double)dspin;
            int accepted;
            a->steps += 1;
            if (use_table) {
                int idE = (int)dE;
                if (idE <= 0) accepted = 1;
                else accepted = xoshiro_f64(x) < (idE == 4 ? p4 : p8);
            } else {
                double log_ratio = -a->beta * dE;
                accepted = (log_ratio > 0) || (xoshiro_f64(x) < exp(log_ratio));
            }
            a->accepted += accepted;
            if (accepted) ising_modify(s, i, dpair, dspin);
        } else {
            /* generic path: serve the positioned-stream interface from xoshiro by
             * re-keying a Philox position with fresh sequential randomness */
            mcxo_rng r; r.seed = mcxo_xoshiro_next(x); r.chain = 0;
            mcxo_rng_position(&r, MCXO_TAG_FLAT, (uint64_t)n, (uint64_t)i);
            mcxo_attempt_at(s, a, i, &r);
        }
    }
}

/* One chain per thread, like ThreadsBackend (src/infrastructure/parallel_chains.jl:99-105).
 * Returns seconds of wall time for the sweeps (excludes setup). */
typedef struct {
    mcxo_system **sys; mcxo_xoshiro *rng; double *am, *ae;
    int c0, c1, use_table; int64_t sweeps; double beta;
} baseline_job;

static void *baseline_worker(void *arg)
{
    baseline_job *j = (baseline_job *)arg;
    for (int c = j->c0; c < j->c1; ++c) {
        mcxo_alg a = { MCXO_METROPOLIS, j->beta, 0, 0 };
        int64_t N = j->sys[c]->N;
        for (int64_t sw = 0; sw < j->sweeps; ++sw) {
            mcxo_sweep_random_site(j->sys[c], &a, &j->rng[c], N, j->use_table);
            j->am[c] += fabs((double)j->sys[c]->sum_spins) / (double)N;
            j->ae[c] += -j->sys[c]->sum_pair / (double)N;
        }
    }
    return 0;
}

double mcxo_baseline_random_site(int L, double beta, int nchains, int64_t sweeps, int nthreads,
                                 int use_table, uint64_t seed, double *mean_abs_m, double *mean_e)
{
    int64_t dims[2] = { L, L };
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nchains) nthreads = nchains;
    mcxo_system **sys = (mcxo_system **)malloc(sizeof(*sys) * (size_t)nchains);
    mcxo_xoshiro *rng = (mcxo_xoshiro *)malloc(sizeof(*rng) * (size_t)nchains);
    double *am = (double *)calloc((size_t)nchains, sizeof(double));
    double *ae = (double *)calloc((size_t)nchains, sizeof(double));
    for (int c = 0; c < nchains; ++c) {
        sys[c] = mcxo_system_create(MCXO_ISING, 2, dims, 1.0, 0.0, 0.0);
        mcxo_system_init_random(sys[c], seed, (uint32_t)c);
        mcxo_xoshiro_seed(&rng[c], seed + 1000u + (uint64_t)c);
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    baseline_job *jobs = (baseline_job *)malloc(sizeof(baseline_job) * (size_t)nthreads);
    struct timespec ts0, ts1;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    for (int t = 0; t < nthreads; ++t) {
        baseline_job jb = { sys, rng, am, ae, (int)((int64_t)nchains * t / nthreads),
                            (int)((int64_t)nchains * (t + 1) / nthreads), use_table, sweeps, beta };
        jobs[t] = jb;
        pthread_create(&th[t], 0, baseline_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    double m = 0, e = 0;
    for (int c = 0; c < nchains; ++c) {
        m += am[c] / (double)sweeps; e += ae[c] / (double)sweeps;
        mcxo_system_destroy(sys[c]);
    }
    if (mean_abs_m) *mean_abs_m = m / nchains;
    if (mean_e) *mean_e = e / nchains;
    free(sys); free(rng); free(am); free(ae); free(th); free(jobs);
    return (double)(ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double)(ts1.tv_nsec - ts0.tv_nsec);
}

/* Per-chain time averages of e = E/N, |m|, m^2, m^4 from the reference's own loop (random site,
 * xoshiro256++, exp per uphill attempt), measured every `interval` sweeps after `therm` sweeps.
 * This stands in for "the reference's Xoshiro runs" in the statistical parity tests.
 * out is [nchains][4]. */
typedef struct { int L; double beta; int64_t therm, sweeps, interval; uint64_t seed; int c0, c1; double *out; } stats_job;

static void *stats_worker(void *arg)
{
    stats_job *j = (stats_job *)arg;
    int64_t dims[2] = { j->L, j->L };
    for (int c = j->c0; c < j->c1; ++c) {
        mcxo_system *s = mcxo_system_create(MCXO_ISING, 2, dims, 1.0, 0.0, 0.0);
        mcxo_xoshiro x;
        mcxo_system_init_random(s, j->seed, (uint32_t)c);
        mcxo_xoshiro_seed(&x, j->seed * 7919u + (uint64_t)c);
        mcxo_alg a = { MCXO_METROPOLIS, j->beta, 0, 0 };
        const int64_t N = s->N;
        mcxo_sweep_random_site(s, &a, &x, N * j->therm, 0);
        double se = 0, sm = 0, sm2 = 0, sm4 = 0; int64_t n = 0;
        for (int64_t sw = 0; sw < j->sweeps; sw += j->interval) {
            mcxo_sweep_random_site(s, &a, &x, N * j->interval, 0);
            double m = (double)s->sum_spins / (double)N, e = -s->sum_pair / (double)N;
            se += e; sm += fabs(m); sm2 += m * m; sm4 += m * m * m * m; n++;
        }
        j->out[4 * c + 0] = se / n; j->out[4 * c + 1] = sm / n; j->out[4 * c + 2] = sm2 / n; j->out[4 * c + 3] = sm4 / n;
        mcxo_system_destroy(s);
    }
    return 0;
}

void mcxo_stats_random_site(int L, double beta, int nchains, int64_t therm, int64_t sweeps, int64_t interval,
                            int nthreads, uint64_t seed, double *out)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nchains) nthreads = nchains;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    stats_job *jobs = (stats_job *)malloc(sizeof(stats_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        stats_job jb = { L, beta, therm, sweeps, interval, seed, (int)((int64_t)nchains * t / nthreads),
                         (int)((int64_t)nchains * (t + 1) / nthreads), out };
        jobs[t] = jb;
        pthread_create(&th[t], 0, stats_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    free(th); free(jobs);
}

/* Lean variant of the same loop for lattices whose 32 B/site neighbour table (ising.jl:436,
 * nbr4::Vector{NTuple{4,Int}}) would not fit comfortably in host memory (L = 16384: 8.6 GB):
 * identical algorithm and draws, neighbours computed arithmetically instead of looked up.
 * One chain per thread; returns seconds for `nattempts` attempts per chain. */
typedef struct {
    int8_t *spins; int L; double beta; int64_t nattempts; mcxo_xoshiro rng; int64_t accepted; int use_table;
    pthread_barrier_t *bar; double t0, t1;
} lean_job;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *lean_worker(void *arg)
{
    lean_job *j = (lean_job *)arg;
    const int L = j->L;
    const uint64_t N = (uint64_t)L * (uint64_t)L;
    int8_t *sp = j->spins;
    for (uint64_t i = 0; i < N; i += 64) {                      /* init!(sys, :random), untimed */
        uint64_t w = mcxo_xoshiro_next(&j->rng);
        for (uint64_t b = 0; b < 64 && i + b < N; ++b) sp[i + b] = ((w >> b) & 1) ? 1 : -1;
    }
    pthread_barrier_wait(j->bar);
    j->t0 = now_s();
    const double p4 = exp(-4 * j->beta), p8 = exp(-8 * j->beta);
    int64_t acc = 0;
    for (int64_t n = 0; n < j->nattempts; ++n) {
        uint64_t i = mcxo_xoshiro_next(&j->rng) % N;           /* pick_site, abstractions.jl:19 */
        int x = (int)(i % (uint64_t)L), y = (int)(i / (uint64_t)L);
        int xl = x == 0 ? L - 1 : x - 1, xr = x == L - 1 ? 0 : x + 1;
        int yu = y == 0 ? L - 1 : y - 1, yd = y == L - 1 ? 0 : y + 1;
        int s = sp[i];
        int lpi = s * (sp[(uint64_t)y * L + xl] + sp[(uint64_t)y * L + xr] + sp[(uint64_t)yu * L + x] + sp[(uint64_t)yd * L + x]);
        int dE = 2 * lpi;                                       /* -dpair, ising.jl:484-491 */
        int accepted;
        if (j->use_table) {                                     /* TableMetropolis, importance_Ising2D.jl:74-92 */
            if (dE <= 0) accepted = 1;
            else accepted = xoshiro_f64(&j->rng) < (dE == 4 ? p4 : p8);
        } else {                                                /* _accept!, importance_sampling.jl:80-85 */
            double log_ratio = -j->beta * (double)dE;
            accepted = (log_ratio > 0) || (xoshiro_f64(&j->rng) < exp(log_ratio));
        }
        if (accepted) { sp[i] = (int8_t)-s; acc++; }
    }
    j->accepted = acc;
    j->t1 = now_s();
    return 0;
}

double mcxo_baseline_lean(int L, double beta, int nchains, int64_t nattempts, int use_table, uint64_t seed,
                          double *accept_rate)
{
    const uint64_t N = (uint64_t)L * (uint64_t)L;
    lean_job *jobs = (lean_job *)calloc((size_t)nchains, sizeof(lean_job));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nchains);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, 0, (unsigned)nchains);
    for (int c = 0; c < nchains; ++c) {
        jobs[c].spins = (int8_t *)malloc(N);
        mcxo_xoshiro_seed(&jobs[c].rng, seed + (uint64_t)c);
        jobs[c].L = L; jobs[c].beta = beta; jobs[c].nattempts = nattempts; jobs[c].use_table = use_table;
        jobs[c].bar = &bar;
    }
    for (int c = 0; c < nchains; ++c) pthread_create(&th[c], 0, lean_worker, &jobs[c]);
    for (int c = 0; c < nchains; ++c) pthread_join(th[c], 0);
    pthread_barrier_destroy(&bar);
    int64_t acc = 0;
    double t0 = jobs[0].t0, t1 = jobs[0].t1;
    for (int c = 0; c < nchains; ++c) {
        acc += jobs[c].accepted; free(jobs[c].spins);
        if (jobs[c].t0 < t0) t0 = jobs[c].t0;
        if (jobs[c].t1 > t1) t1 = jobs[c].t1;
    }
    if (accept_rate) *accept_rate = (double)acc / ((double)nattempts * nchains);
    free(jobs); free(th);
    return t1 - t0;
}

/* ===================================================================== */
/* Integer threshold tables.  For a draw u = m*2^-32 and a Float64 p:     */
/*   u < p  <=>  m < ceil(p * 2^32)   (p*2^32 is exact).                  */
/* The host evaluates the reference's own float expression for every      */
/* local configuration and hands the device only the integers.            */
/* Table index conventions (also in include/mcx_b200.h):                  */
/*  Ising (nn = 2*ndim):      idx = s*(nn+1) + nup, s in {0:down,1:up},   */
/*     Metropolis/Glauber: flip iff m < T;  HeatBath: new spin up iff m<T */
/*  Blume-Capel Metropolis/Glauber: idx = (so*2+b)*(2nn+1) + (sum+nn),    */
/*     so in {0,1,2} = {-1,0,+1}, b the Bool draw; accept iff m < T       */
/*  Blume-Capel HeatBath: idx = k*(2nn+1) + (sum+nn), k in {0,1}:         */
/*     new = m<T0 ? -1 : (m<T1 ? 0 : +1)                                  */
/* ===================================================================== */
static uint64_t thr_from_p(double p)
{
    if (!(p > 0)) return 0;
    double x = ceil(p * 4294967296.0);
    if (x >= 4294967296.0) return 4294967296ull;
    return (uint64_t)x;
}
static uint64_t thr_accept(int rule, double log_ratio)
{
    if (rule == MCXO_GLAUBER) return thr_from_p(mcxo_logistic(log_ratio));
    if (log_ratio > 0) return 4294967296ull;
    return thr_from_p(exp(log_ratio));
}
/* smallest m in [0,2^32] with !(m*2^-32*z < w): threshold of the monotone predicate */
static uint64_t thr_scaled(double z, double w)
{
    uint64_t lo = 0, hi = 4294967296ull;   /* predicate true on [0,lo), find first false */
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        double rr = ((double)mid * (1.0 / 4294967296.0)) * z;
        if (rr < w) lo = mid + 1; else hi = mid;
    }
    return lo;
}

int mcxo_table_len(int model, int rule, int ndim)
{
    int nn = 2 * ndim;
    if (model == MCXO_ISING) return 2 * (nn + 1);
    if (rule == MCXO_HEATBATH) return 2 * (2 * nn + 1);
    return 6 * (2 * nn + 1);
}

void mcxo_build_table(int model, int rule, int ndim, double beta, double J, double h, double D,
                      uint64_t *T)
{
    int nn = 2 * ndim;
    if (model == MCXO_ISING) {
        for (int sb = 0; sb < 2; ++sb)
            for (int nup = 0; nup <= nn; ++nup) {
                int s = sb ? 1 : -1;
                int64_t lpi = (int64_t)s * (2 * nup - nn);
                double dpair = (-2.0 * J) * (double)lpi;
                int64_t dspin = -2 * s;
                double dE = -dpair - h * (double)dspin;
                uint64_t t;
                if (rule == MCXO_HEATBATH) t = thr_from_p(mcxo_logistic(beta * (double)s * dE));
                else t = thr_accept(rule, -beta * dE);
                T[sb * (nn + 1) + nup] = t;
            }
    } else if (rule == MCXO_HEATBATH) {
        for (int sum = -nn; sum <= nn; ++sum) {
            double coupling = J * (double)sum;
            double e1 = -(-1) * coupling - h * (-1) + D;
            double e2 = 0.0;
            double e3 = -(1) * coupling - h * (1) + D;
            double w1 = exp(-beta * e1), w2 = exp(-beta * e2), w3 = exp(-beta * e3);
            double z = w1 + w2 + w3;
            T[0 * (2 * nn + 1) + (sum + nn)] = thr_scaled(z, w1);
            T[1 * (2 * nn + 1) + (sum + nn)] = thr_scaled(z, w1 + w2);
        }
    } else {
        for (int so = 0; so < 3; ++so)
            for (int b = 0; b < 2; ++b)
                for (int sum = -nn; sum <= nn; ++sum) {
                    int s_old = so - 1;
                    int s_new = mcxo_propose_state(b, s_old);
                    int64_t dspin = s_new - s_old;
                    int64_t dspin2 = (int64_t)s_new * s_new - (int64_t)s_old * s_old;
                    double dpair = (double)dspin * (J * (double)sum);
                    double dE = -dpair - h * (double)dspin + D * (double)dspin2;
                    T[(so * 2 + b) * (2 * nn + 1) + (sum + nn)] = thr_accept(rule, -beta * dE);
                }
    }
}

/* ===================================================================== */
/* BinnedObject (discrete, 1-D) + multicanonical / Wang-Landau            */
/* ===================================================================== */
/* _binindex for Integer bins: div(x - start, step) + 1 (binned_object.jl:22-24).
 * Julia's div truncates toward zero, as C's / does. */
int64_t mcxo_binindex(int64_t start, int64_t step, int64_t x) { return (x - start) / step + 1; }
/* Float bins: Int(round((x - start)/step)) + 1 (binned_object.jl:18-20); round = ties-to-even */
int64_t mcxo_binindex_f(double start, double step, double x) { return (int64_t)rint((x - start) / step) + 1; }

/* update!(::MulticanonicalEnsemble; mode=:simple) (ensembles/multicanonical.jl:32-44) */
void mcxo_muca_update(double *logweight, const double *histogram, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        double hh = histogram[i];
        double logh = hh > 0 ? log(hh) : 0.0;
        logweight[i] -= logh;
    }
}

/* accept!(alg, x_new, x_old): generic (importance_sampling.jl:69-78) with
 * MulticanonicalEnsemble.record_visit! (ensembles/multicanonical.jl:25-30) or the
 * Wang-Landau variant (algorithms/wang_landau.jl:29-37).  Returns accepted (0/1) or -1 for
 * an out-of-range lookup (BoundsError; steps unchanged, test_multicanonical.jl:39-43). */
int mcxo_flat_accept(mcxo_alg *a, mcxo_flat *f, int kind, int64_t x_new, int64_t x_old, double u)
{
    int64_t in = mcxo_binindex(f->start, f->step, x_new), io = mcxo_binindex(f->start, f->step, x_old);
    if (in < 1 || in > f->num || io < 1 || io > f->num) return -1;
    double log_ratio = f->logweight[in - 1] - f->logweight[io - 1];
    a->steps += 1;
    int accepted = (log_ratio > 0) || (u < exp(log_ratio));
    a->accepted += accepted;
    int64_t iv = accepted ? in : io;
    if (kind == 0) f->histogram[iv - 1] += 1;
    else f->logweight[iv - 1] -= f->logf;
    return accepted;
}

/* Flat-histogram sweeps: the reference's spin_flip!(sys, alg::ImportanceSampling) (ising.jl:25-33;
 * muca_BlumeCapel.jl:81-89 for the (pair, spin^2) tuple observable) applied once per site per sweep,
 * one site after the other (the acceptance depends on the chain's global observable), in
 * checkerboard order: all colour-0 sites in slot order, then all colour-1 sites.  The stream is
 * positioned at (chain, FLAT, 2*sweep + colour, slot) with slot = row*(Lx/2) + (x>>1) as for SWEEP. */
/* policy 0: an out-of-range bin is the reference's BoundsError (test/test_multicanonical.jl:39-43; returns -1).
 * policy 1: energy windows (BASELINE.json configs[4]).  The reference has no windows; this restates the
 * contract include/mcx_b200.h gives mcx_flat_create: a proposal that leaves the binned range is a rejected
 * attempt -- steps += 1, no draw, the visit (record_visit! / lw -= logf) goes to the current bin.  A chain
 * whose CURRENT state is outside the range is an error under both policies. */
int mcxo_flat_sweep_policy(mcxo_system *s, mcxo_alg *a, mcxo_flat *f, int kind, int observable,
                           double beta_pair, uint64_t seed, uint32_t chain, uint64_t sweep0, int64_t nsweeps,
                           int policy)
{
    mcxo_rng r; r.seed = seed; r.chain = chain;
    const int64_t Lx = s->dims[0], Ly = s->dims[1], half = Lx / 2;
    for (int64_t sw = 0; sw < nsweeps; ++sw)
        for (int colour = 0; colour < 2; ++colour)
        for (int64_t i = 0; i < s->N; ++i) {
            const int64_t xx = i % Lx, row = i / Lx, yy = row % Ly, zz = row / Ly;
            if (((xx + yy + zz) & 1) != colour) continue;
            mcxo_rng_position(&r, MCXO_TAG_FLAT, 2 * (sweep0 + (uint64_t)sw) + (uint64_t)colour,
                              (uint64_t)(row * half + (xx >> 1)));
            if (observable == 0) {
                double dpair; int64_t dspin;
                ising_flip_changes(s, i, &dpair, &dspin);
                double dE = -dpair - s->h * (double)dspin;
                int64_t E_old = (int64_t)mcxo_energy(s, 0);
                int64_t E_new = E_old + (int64_t)dE;
                int64_t in = mcxo_binindex(f->start, f->step, E_new), io = mcxo_binindex(f->start, f->step, E_old);
                if (io < 1 || io > f->num) return -1;
                const int outside = in < 1 || in > f->num;
                if (outside && policy == 0) return -1;
                a->steps += 1;
                int accepted = 0;
                if (!outside) {
                    double log_ratio = f->logweight[in - 1] - f->logweight[io - 1];
                    accepted = (log_ratio > 0) || (mcxo_rand_f64(&r) < exp(log_ratio));
                }
                a->accepted += accepted;
                int64_t iv = accepted ? in : io;
                if (kind == 0) f->histogram[iv - 1] += 1; else f->logweight[iv - 1] -= f->logf;
                if (accepted) ising_modify(s, i, dpair, dspin);
            } else {
                int s_new = mcxo_propose_state(mcxo_rand_bool(&r), s->spins[i]);
                double dpair; int64_t dspin, dspin2;
                bc_propose_changes(s, i, s_new, &dpair, &dspin, &dspin2);
                double Ho1 = s->J * s->sum_pair; int64_t Ho2 = s->sum_spins2;
                double Hn1 = Ho1 + s->J * dpair; int64_t Hn2 = Ho2 + dspin2;
                int64_t in = mcxo_binindex(f->start, f->step, Hn2), io = mcxo_binindex(f->start, f->step, Ho2);
                if (io < 1 || io > f->num) return -1;
                const int outside = in < 1 || in > f->num;
                if (outside && policy == 0) return -1;
                a->steps += 1;
                int accepted = 0;
                if (!outside) {
                    /* logweight(CustomEnsemble, H) = -beta*H[1] + lw2(H[2]) (muca_BlumeCapel.jl:49-51) */
                    double log_ratio = (-beta_pair * Hn1 + f->logweight[in - 1]) - (-beta_pair * Ho1 + f->logweight[io - 1]);
                    accepted = (log_ratio > 0) || (mcxo_rand_f64(&r) < exp(log_ratio));
                }
                a->accepted += accepted;
                int64_t iv = accepted ? in : io;
                if (kind == 0) f->histogram[iv - 1] += 1; else f->logweight[iv - 1] -= f->logf;
                if (accepted) bc_modify(s, i, s_new, dpair, dspin, dspin2);
            }
        }
    return 0;
}

int mcxo_flat_sweep(mcxo_system *s, mcxo_alg *a, mcxo_flat *f, int kind, int observable,
                    double beta_pair, uint64_t seed, uint32_t chain, uint64_t sweep0, int64_t nsweeps)
{
    return mcxo_flat_sweep_policy(s, a, f, kind, observable, beta_pair, seed, chain, sweep0, nsweeps, 0);
}

/* ===================================================================== */
/* Replica exchange (src/algorithms/replica_exchange.jl)                  */
/* ===================================================================== */
/* exchange_log_ratio :110-113 with BoltzmannEnsemble logweight(E) = -beta*E (boltzmann.jl:28) */
double mcxo_exchange_log_ratio(double bi, double bj, double xi, double xj)
{
    return ((-bi * xj) - (-bi * xi)) + ((-bj * xi) - (-bj * xj));
}
/* _accept_exchange :115 */
int mcxo_accept_exchange(double log_ratio, double u) { return (log_ratio > 0) || (u < exp(log_ratio)); }

/* _resolve_pair :138-149 -> out = {active, pair_id, partner_index} */
void mcxo_resolve_pair(int64_t my_index, int64_t stage, int64_t nranks, int64_t out[3])
{
    int64_t first = (stage % 2 == 0) ? 1 : 2;
    int64_t offset = my_index - first;
    if (offset >= 0 && offset % 2 == 0 && my_index < nranks) {
        out[0] = 1; out[1] = my_index; out[2] = my_index + 1;
    } else if (offset > 0 && (offset % 2 != 0) && my_index - 1 >= first) {
        out[0] = 1; out[1] = my_index - 1; out[2] = my_index - 1;
    } else {
        out[0] = 0; out[1] = 0; out[2] = 0;
    }
}

/* set_betas (parallel_tempering.jl:146-162): range(bmax, bmin, length=n), optionally in log space */
void mcxo_set_betas(int64_t n, double bmin, double bmax, int geometric, double *out)
{
    long double a = geometric ? (long double)log(bmax) : (long double)bmax;
    long double b = geometric ? (long double)log(bmin) : (long double)bmin;
    for (int64_t i = 0; i < n; ++i) {
        long double v = (a * (long double)(n - 1 - i) + b * (long double)i) / (long double)(n - 1);
        out[i] = geometric ? exp((double)v) : (double)v;
    }
}

/* update!(rx::ReplicaExchange{ThreadsBackend}, xs) (:158-178).  indices[r] is the 1-based ladder
 * position of slot r; the ensembles (betas) move between slots on accept (:133), the lattices
 * stay.  u_of_slot[r] is what rand(algorithm(rx, r).rng) returns for this round. */
void mcxo_rx_update(int64_t n, int64_t *stage, int64_t *indices, int64_t *steps, int64_t *accepted,
                    double *beta_of_slot, const double *xs, const double *u_of_slot)
{
    int64_t first = (*stage % 2 == 0) ? 1 : 2;
    for (int64_t pair_id = first; pair_id <= n - 1; pair_id += 2) {
        int64_t ri = -1, rj = -1;
        for (int64_t r = 0; r < n; ++r) {
            if (indices[r] == pair_id && ri < 0) ri = r;
            if (indices[r] == pair_id + 1 && rj < 0) rj = r;
        }
        steps[pair_id - 1] += 1;
        double u = u_of_slot[ri];
        double lr = mcxo_exchange_log_ratio(beta_of_slot[ri], beta_of_slot[rj], xs[ri], xs[rj]);
        if (mcxo_accept_exchange(lr, u)) {
            double tb = beta_of_slot[ri]; beta_of_slot[ri] = beta_of_slot[rj]; beta_of_slot[rj] = tb;
            accepted[pair_id - 1] += 1;
            int64_t ti = indices[ri]; indices[ri] = indices[rj]; indices[rj] = ti;
        }
    }
    *stage = 1 - *stage;
}

/* =====================================================================================================
 * General topologies: IsingGraph (global J, SpinSystems/src/ising.jl:86-205) and IsingMatrix (sparse J_ij,
 * ising.jl:207-360) with no / uniform / per-site fields.  Neighbour lists are CSR in the reference's adjacency
 * order (ascending neighbour index: Graphs.jl adjacency lists and SparseMatrixCSC columns are sorted).
 * Sweeps visit the colour classes of a greedy first-fit colouring in site order (generalising the checkerboard);
 * the random stream of an attempt is positioned at (chain, t = ncolours * sweep + colour, slot = rank of the site
 * inside its colour class).
 * ===================================================================================================== */
mcxo_graph *mcxo_graph_create(int64_t n, const int64_t *rowptr, const int64_t *col, const double *val, double J,
                              int hmode, double h, const double *hvec)
{
    mcxo_graph *g = (mcxo_graph *)calloc(1, sizeof(*g));
    int64_t nnz = rowptr[n];
    g->n = n;
    g->rowptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    g->col = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    memcpy(g->rowptr, rowptr, sizeof(int64_t) * (size_t)(n + 1));
    memcpy(g->col, col, sizeof(int64_t) * (size_t)nnz);
    g->matrix = val != 0;
    if (val) {
        g->val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
        memcpy(g->val, val, sizeof(double) * (size_t)nnz);
    }
    g->J = J; g->hmode = hmode; g->h = h;
    if (hmode == 2) {
        g->hvec = (double *)malloc(sizeof(double) * (size_t)n);
        memcpy(g->hvec, hvec, sizeof(double) * (size_t)n);
    }
    g->spins = (int8_t *)malloc((size_t)n);
    for (int64_t i = 0; i < n; ++i) g->spins[i] = 1;            /* constructors start all-up (ising.jl:118, 271) */
    mcxo_graph_recompute(g);
    return g;
}

void mcxo_graph_destroy(mcxo_graph *g)
{
    if (!g) return;
    free(g->rowptr); free(g->col); free(g->val); free(g->hvec); free(g->spins); free(g);
}

void mcxo_graph_set_spins(mcxo_graph *g, const int8_t *spins) { memcpy(g->spins, spins, (size_t)g->n); mcxo_graph_recompute(g); }
void mcxo_graph_get_spins(const mcxo_graph *g, int8_t *spins) { memcpy(spins, g->spins, (size_t)g->n); }

void mcxo_graph_init_random(mcxo_graph *g, uint64_t seed, uint32_t chain)
{
    uint32_t out[4];
    for (int64_t i = 0; i < g->n; ++i) {
        if ((i & 127) == 0) stream_block(seed, chain, MCXO_TAG_INIT, 0, (uint64_t)i >> 7, 0, out);
        g->spins[i] = ((out[(i >> 5) & 3] >> (i & 31)) & 1u) ? 1 : -1;
    }
    mcxo_graph_recompute(g);
}

/* local_pair_interactions: graph s_i * sum_j s_j (abstractions.jl:41-48); matrix sum_j s_i J_ij s_j, j != i (ising.jl:295-305) */
double mcxo_graph_local_pair(const mcxo_graph *g, int64_t i)
{
    if (!g->matrix) {
        int64_t si = g->spins[i], acc = 0;
        for (int64_t p = g->rowptr[i]; p < g->rowptr[i + 1]; ++p) acc += si * g->spins[g->col[p]];
        return (double)acc;
    }
    double acc = 0.0, si = (double)g->spins[i];
    for (int64_t p = g->rowptr[i]; p < g->rowptr[i + 1]; ++p) {
        int64_t j = g->col[p];
        if (j != i) acc += si * g->val[p] * (double)g->spins[j];
    }
    return acc;
}

/* _pair_sum (ising.jl:163-169, 307-313), _field_sum (:171-179, 315-323) */
static double graph_pair_sum(const mcxo_graph *g)
{
    double acc = 0.0;
    for (int64_t i = 0; i < g->n; ++i) acc += mcxo_graph_local_pair(g, i);
    return g->matrix ? acc / 2 : g->J * (acc / 2);
}
static double graph_field_sum(const mcxo_graph *g)
{
    if (g->hmode == 0) return 0.0;
    if (g->hmode == 1) {
        int64_t s = 0;
        for (int64_t i = 0; i < g->n; ++i) s += g->spins[i];
        return g->h * (double)s;
    }
    double acc = 0.0;
    for (int64_t i = 0; i < g->n; ++i) acc += g->hvec[i] * (double)g->spins[i];
    return acc;
}

void mcxo_graph_recompute(mcxo_graph *g)
{
    g->sum_pair = graph_pair_sum(g);
    g->sum_spins = 0;
    for (int64_t i = 0; i < g->n; ++i) g->sum_spins += g->spins[i];
    g->sum_field = graph_field_sum(g);
}

double mcxo_graph_energy(const mcxo_graph *g, int full)
{
    if (full) return -graph_pair_sum(g) - graph_field_sum(g);     /* _full_energy */
    return g->hmode == 0 ? -g->sum_pair : -g->sum_pair - g->sum_field;
}
int64_t mcxo_graph_magnetization(const mcxo_graph *g) { return g->sum_spins; }

/* flip_changes + delta_energy (ising.jl:187-198, 339-351) */
static inline void graph_flip_changes(const mcxo_graph *g, int64_t i, double *dpair, int64_t *dspin)
{
    double lp = mcxo_graph_local_pair(g, i);
    *dpair = g->matrix ? -2.0 * lp : (-2 * g->J) * lp;
    *dspin = -2 * (int64_t)g->spins[i];
}
static inline double graph_dfield(const mcxo_graph *g, int64_t dspin, int64_t i)
{
    return g->hmode == 0 ? 0.0 : g->hmode == 1 ? g->h * (double)dspin : g->hvec[i] * (double)dspin;
}
double mcxo_graph_delta_energy(const mcxo_graph *g, int64_t i)
{
    double dpair; int64_t dspin;
    graph_flip_changes(g, i, &dpair, &dspin);
    return -dpair - graph_dfield(g, dspin, i);
}
/* modify! (ising.jl:200-214, 353-367) */
void mcxo_graph_flip(mcxo_graph *g, int64_t i)
{
    double dpair; int64_t dspin;
    graph_flip_changes(g, i, &dpair, &dspin);
    g->spins[i] = (int8_t)(-g->spins[i]);
    g->sum_pair += dpair;
    g->sum_spins += dspin;
    g->sum_field += graph_dfield(g, dspin, i);
}

/* spin_flip! without pick_site (ising.jl:35-58) */
void mcxo_graph_attempt_at(mcxo_graph *g, mcxo_alg *a, int64_t i, mcxo_rng *r)
{
    double dE = mcxo_graph_delta_energy(g, i);
    if (a->rule == MCXO_HEATBATH) {
        int s_old = g->spins[i];
        double p_plus = mcxo_logistic(a->beta * (double)s_old * dE);
        int s_new = mcxo_rand_f64(r) < p_plus ? 1 : -1;
        if (s_new != s_old) mcxo_graph_flip(g, i);
        a->steps += 1;
    } else if (accept_delta(a, dE, r)) {
        mcxo_graph_flip(g, i);
    }
}

/* greedy first-fit colouring in site order; returns the number of colours */
int mcxo_graph_colour(const mcxo_graph *g, int32_t *colour)
{
    int ncol = 0;
    int64_t maxdeg = 0;
    for (int64_t i = 0; i < g->n; ++i) {
        int64_t d = g->rowptr[i + 1] - g->rowptr[i];
        if (d > maxdeg) maxdeg = d;
    }
    int64_t *mark = (int64_t *)malloc(sizeof(int64_t) * (size_t)(maxdeg + 2));
    for (int64_t k = 0; k < maxdeg + 2; ++k) mark[k] = -1;
    for (int64_t i = 0; i < g->n; ++i) {
        for (int64_t p = g->rowptr[i]; p < g->rowptr[i + 1]; ++p) {
            int64_t j = g->col[p];
            if (j < i && j != i && colour[j] <= maxdeg) mark[colour[j]] = i;
        }
        int c = 0;
        while (mark[c] == i) ++c;
        colour[i] = c;
        if (c + 1 > ncol) ncol = c + 1;
    }
    free(mark);
    return ncol;
}

void mcxo_graph_sweep_coloured(mcxo_graph *g, mcxo_alg *a, uint64_t seed, uint32_t chain, uint64_t sweep0, int64_t nsweeps,
                               const int32_t *colour, int ncolours)
{
    mcxo_rng r; r.seed = seed; r.chain = chain;
    for (int64_t sw = 0; sw < nsweeps; ++sw)
        for (int c = 0; c < ncolours; ++c) {
            uint64_t t = (uint64_t)ncolours * (sweep0 + (uint64_t)sw) + (uint64_t)c;
            uint64_t slot = 0;
            for (int64_t i = 0; i < g->n; ++i) {
                if (colour[i] != c) continue;
                mcxo_rng_position(&r, MCXO_TAG_SWEEP, t, slot++);
                mcxo_graph_attempt_at(g, a, i, &r);
            }
        }
}
