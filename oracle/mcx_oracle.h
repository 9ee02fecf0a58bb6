/*
 * mcx_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the lattice sampling hot path of MonteCarloX.jl /
 * SpinSystems (reference tree /root/reference, Julia).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (montecarlox.jl_b200/csrc) never does.
 *
 * Parity status: the reference is Julia and cannot run in this image (no julia
 * binary).  The restatement is pinned against every known-answer value the
 * reference's own tests hold for this path (see
 * tests/test_oracle_known_answers.py) and against the reference's only golden
 * file (SpinSystems/data/exact_solutions/ising2D_8x8.csv) through an
 * independent transfer-matrix enumeration (tests/golden/gen_exact_dos.py).
 * No reference test pins a spin *trajectory* (SURVEY.md section 4), so the
 * trajectory-level definition ("the reference's per-site primitives called in
 * checkerboard order with an injected Philox AbstractRNG") is pinned only by
 * construction: parity of trajectories is "unpinned by the reference".
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef MCX_ORACLE_H
#define MCX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- Philox4x32-10 + positioned stream ("PhiloxRNG") -------- */
void mcxo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

enum { MCXO_TAG_SWEEP = 0, MCXO_TAG_EXCHANGE = 1, MCXO_TAG_INIT = 2, MCXO_TAG_FLAT = 3 };

typedef struct {
    uint64_t seed;   /* Philox key */
    uint32_t chain;  /* counter word 3 */
    uint32_t tag;    /* stream family */
    uint64_t t;      /* time index (2*sweep+colour for SWEEP) */
    uint64_t q;      /* site slot within the colour plane (or linear site) */
    uint32_t draw;   /* next draw slot at this position */
} mcxo_rng;

void     mcxo_rng_position(mcxo_rng *r, uint32_t tag, uint64_t t, uint64_t q);
uint32_t mcxo_rng_lane16(const mcxo_rng *r, uint32_t plane);
double   mcxo_rand_f64(mcxo_rng *r);   /* stand-in for rand(rng)::Float64 */
int      mcxo_rand_bool(mcxo_rng *r);  /* stand-in for rand(rng, Bool)   */
double   mcxo_exchange_u(uint64_t seed, uint32_t chain, uint64_t round);

/* ---------------- sequential RNG for the reference's own loop ------------ */
typedef struct { uint64_t s[4]; } mcxo_xoshiro;
void     mcxo_xoshiro_seed(mcxo_xoshiro *x, uint64_t seed);
uint64_t mcxo_xoshiro_next(mcxo_xoshiro *x);

/* ---------------- models -------------------------------------------------- */
enum { MCXO_ISING = 0, MCXO_BLUME_CAPEL = 1 };
enum { MCXO_METROPOLIS = 0, MCXO_GLAUBER = 1, MCXO_HEATBATH = 2 };

typedef struct {
    int      model, ndim, nn;
    int64_t  dims[3];
    int64_t  N;
    int8_t  *spins;          /* reference order: x fastest */
    int64_t *nbr;            /* [N][nn] neighbour table (ising.jl:430-461 generalised) */
    double   J, h, D;        /* couplings (Ising J=1,h=0 is the integer path) */
    /* cached sums, as the reference caches them */
    double   sum_pair;       /* J * sum_<ij> s_i s_j   (ising.jl:90, blume_capel.jl:123) */
    int64_t  sum_spins;
    int64_t  sum_spins2;
} mcxo_system;

typedef struct {
    int      rule;
    double   beta;
    int64_t  steps, accepted;
} mcxo_alg;

mcxo_system *mcxo_system_create(int model, int ndim, const int64_t *dims, double J, double h, double D);
void    mcxo_system_destroy(mcxo_system *s);
void    mcxo_system_set_spins(mcxo_system *s, const int8_t *spins);
void    mcxo_system_get_spins(const mcxo_system *s, int8_t *spins);
void    mcxo_system_init_random(mcxo_system *s, uint64_t seed, uint32_t chain);
void    mcxo_recompute(mcxo_system *s);
double  mcxo_energy(const mcxo_system *s, int full);
int64_t mcxo_magnetization(const mcxo_system *s, int full);
int64_t mcxo_pair_count(const mcxo_system *s);   /* unweighted sum_<ij> s_i s_j, full recompute */
int64_t mcxo_spin2_sum(const mcxo_system *s);

/* per-site primitives */
int64_t mcxo_local_pair_interactions(const mcxo_system *s, int64_t i);
double  mcxo_delta_energy_flip(const mcxo_system *s, int64_t i);             /* Ising */
double  mcxo_delta_energy_bc(const mcxo_system *s, int64_t i, int s_new);    /* Blume-Capel */
double  mcxo_logistic(double x);
int     mcxo_propose_state(int u_bool, int s_old);

/* one attempt at a fixed site with a positioned stream (spin_flip! minus pick_site) */
void mcxo_attempt_at(mcxo_system *s, mcxo_alg *a, int64_t i, mcxo_rng *r);

/* mode 1: checkerboard + Philox sweeps (parity target for the CUDA kernels) */
void mcxo_sweep_checkerboard(mcxo_system *s, mcxo_alg *a, uint64_t seed, uint32_t chain,
                             uint64_t sweep0, int64_t nsweeps);

/* mode 2: the reference's actual loop: random site, sequential RNG (xoshiro256++) */
void mcxo_sweep_random_site(mcxo_system *s, mcxo_alg *a, mcxo_xoshiro *x, int64_t nattempts,
                            int use_table);
/* multi-chain variant of mode 2, one chain per thread (ThreadsBackend) */
double mcxo_baseline_random_site(int L, double beta, int nchains, int64_t sweeps, int nthreads,
                                 int use_table, uint64_t seed, double *mean_abs_m, double *mean_e);

/* per-chain averages of e, |m|, m^2, m^4 from the reference's random-site loop (out: [nchains][4]) */
void mcxo_stats_random_site(int L, double beta, int nchains, int64_t therm, int64_t sweeps, int64_t interval,
                            int nthreads, uint64_t seed, double *out);

/* lean multi-chain baseline for very large L (neighbours computed, not tabulated) */
double mcxo_baseline_lean(int L, double beta, int nchains, int64_t nattempts, int use_table, uint64_t seed,
                          double *accept_rate);

/* ---------------- integer threshold tables (what the host hands the GPU) -- */
int mcxo_table_len(int model, int rule, int ndim);
void mcxo_build_table(int model, int rule, int ndim, double beta, double J, double h, double D,
                      uint64_t *thresholds);

/* ---------------- BinnedObject (discrete 1-D), muca / WL ------------------ */
typedef struct {
    int64_t start, step, num;
    double *logweight;
    double *histogram;
    double  logf;
} mcxo_flat;

int64_t mcxo_binindex(int64_t start, int64_t step, int64_t x);   /* 1-based like the reference */
int64_t mcxo_binindex_f(double start, double step, double x);
void    mcxo_muca_update(double *logweight, const double *histogram, int64_t n);
/* sequential-site flat-histogram sweeps with Philox (tag FLAT). kind 0 = muca, 1 = WL.
 * observable: 0 = energy (Ising, integer E), 1 = sum s^2 with Boltzmann(beta_pair) on the
 * pair term (muca_BlumeCapel.jl). Returns 0, or -1 on an out-of-range bin (BoundsError). */
int mcxo_flat_sweep(mcxo_system *s, mcxo_alg *a, mcxo_flat *f, int kind, int observable,
                    double beta_pair, uint64_t seed, uint32_t chain, uint64_t sweep0, int64_t nsweeps);
/* the same with the out-of-range policy of include/mcx_b200.h mcx_flat_create: 0 = BoundsError (the
 * reference), 1 = energy window (a proposal leaving the range is a rejected attempt; not in the reference) */
int mcxo_flat_sweep_policy(mcxo_system *s, mcxo_alg *a, mcxo_flat *f, int kind, int observable,
                           double beta_pair, uint64_t seed, uint32_t chain, uint64_t sweep0, int64_t nsweeps,
                           int policy);
/* single generic accept!(alg, x_new, x_old) on a table, with explicit u (test helper) */
int mcxo_flat_accept(mcxo_alg *a, mcxo_flat *f, int kind, int64_t x_new, int64_t x_old, double u);

/* ---------------- replica exchange --------------------------------------- */
double mcxo_exchange_log_ratio(double beta_i, double beta_j, double x_i, double x_j);
int    mcxo_accept_exchange(double log_ratio, double u);
void   mcxo_resolve_pair(int64_t my_index, int64_t stage, int64_t nranks, int64_t out[3]);
void   mcxo_set_betas(int64_t n, double bmin, double bmax, int geometric, double *out);
/* ReplicaExchange{ThreadsBackend} update!: indices 1-based ladder positions per slot,
 * betas_of_slot are the ensembles currently held by each slot (swapped on accept). */
void   mcxo_rx_update(int64_t n, int64_t *stage, int64_t *indices, int64_t *steps, int64_t *accepted,
                      double *beta_of_slot, const double *xs, const double *u_of_slot);


/* ---------------- general topologies (IsingGraph / IsingMatrix, ising.jl:86-360) ---------------- */
typedef struct {
    int64_t  n;
    int64_t *rowptr, *col;   /* CSR neighbour lists, 0-based, the reference's adjacency order (ascending) */
    double  *val;            /* J_ij per entry (IsingMatrix) or NULL (IsingGraph: global J) */
    int      matrix;
    double   J;
    int      hmode;          /* 0: no field, 1: uniform h, 2: per-site h_i */
    double   h, *hvec;
    int8_t  *spins;
    double   sum_pair;       /* sum_pair_interactions */
    int64_t  sum_spins;
    double   sum_field;      /* sum_field_interactions */
} mcxo_graph;

mcxo_graph *mcxo_graph_create(int64_t n, const int64_t *rowptr, const int64_t *col, const double *val, double J,
                              int hmode, double h, const double *hvec);
void    mcxo_graph_destroy(mcxo_graph *g);
void    mcxo_graph_set_spins(mcxo_graph *g, const int8_t *spins);
void    mcxo_graph_get_spins(const mcxo_graph *g, int8_t *spins);
void    mcxo_graph_init_random(mcxo_graph *g, uint64_t seed, uint32_t chain);
void    mcxo_graph_recompute(mcxo_graph *g);
double  mcxo_graph_local_pair(const mcxo_graph *g, int64_t i);
double  mcxo_graph_energy(const mcxo_graph *g, int full);
int64_t mcxo_graph_magnetization(const mcxo_graph *g);
double  mcxo_graph_delta_energy(const mcxo_graph *g, int64_t i);
void    mcxo_graph_flip(mcxo_graph *g, int64_t i);                       /* modify!(sys, i, flip_changes(sys, i)...) */
void    mcxo_graph_attempt_at(mcxo_graph *g, mcxo_alg *a, int64_t i, mcxo_rng *r);
int     mcxo_graph_colour(const mcxo_graph *g, int32_t *colour);         /* greedy first-fit in site order */
void    mcxo_graph_sweep_coloured(mcxo_graph *g, mcxo_alg *a, uint64_t seed, uint32_t chain, uint64_t sweep0,
                                  int64_t nsweeps, const int32_t *colour, int ncolours);

#ifdef __cplusplus
}
#endif
#endif
