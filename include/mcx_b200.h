/*
 * mcx_b200.h -- C ABI of the B200-native lattice sampling library (libmcx_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of MonteCarloX.jl / SpinSystems: checkerboard
 * Metropolis / Glauber / heat-bath sweeps over Ising and Blume-Capel lattices, batched over
 * chains and parallel-tempering replicas, with multicanonical / Wang-Landau table updates.
 *
 * The reference (Julia, /root/reference) has no FFI: its extension mechanism is multiple
 * dispatch.  Each entry point below names the reference method(s) a Julia
 * shim would forward to it (julia/MonteCarloXB200.jl shows the ccall side; INTEGRATION.md the
 * wiring).  Plain pointers and sizes only; no torch / CUDA types in signatures (a CUDA stream is
 * passed as void*).
 *
 * Conventions
 *  - every function returns an int32 status (MCX_OK == 0); mcx_last_error() gives the message
 *    (thread-local).  The Julia shim maps MCX_ERR_ARGUMENT -> ArgumentError, MCX_ERR_BOUNDS ->
 *    BoundsError, MCX_ERR_STATE -> AssertionError, the rest -> ErrorException.
 *  - the library owns device memory behind opaque handles; the caller owns every host pointer and
 *    keeps it alive for the duration of the call.
 *  - handles are not thread-safe; distinct handles are.  One host thread/process per GPU.
 *  - work is enqueued on the context's stream; calls that return host data synchronise it.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    MCX_ERR_CUDA.
 *
 * Site order of host spin buffers is the reference's: i = x + Lx*(y + Ly*z), x fastest
 * (SpinSystems/src/ising.jl:444-456), values -1/+1 (Ising) or -1/0/+1 (Blume-Capel), int8.
 *
 * RNG layout v1 (bit-exactness contract): Philox4x32-10,
 *   key = (seed lo32, seed hi32)
 *   ctr = (slot >> 3, t lo32, (t >> 32 & 0xffff) | plane << 16 | tag << 24, chain id)
 *   16-bit lane of a slot = (out[(slot & 7) >> 1] >> 16 * (slot & 1)) & 0xffff
 *   draw n of a position: high half from plane 2n, low half from plane 2n+1;
 *   rand(Float64) = (hi << 16 | lo) * 2^-32, rand(Bool) = hi >> 15.
 *   tag SWEEP(0): t = 2*sweep + colour, slot = row*(Lx/2) + (x >> 1), row = y + Ly*z,
 *   colour = (x + y + z) & 1.  tag EXCHANGE(1), INIT(2), FLAT(3): see DESIGN.md.
 */
#ifndef MCX_B200_H
#define MCX_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCX_ABI_VERSION 1

enum mcx_status {
    MCX_OK = 0,
    MCX_ERR_ARGUMENT = 1,
    MCX_ERR_BOUNDS = 2,
    MCX_ERR_CUDA = 3,
    MCX_ERR_STATE = 4,
    MCX_ERR_UNSUPPORTED = 5
};
enum mcx_model { MCX_ISING = 0, MCX_BLUME_CAPEL = 1 };
enum mcx_rule { MCX_METROPOLIS = 0, MCX_GLAUBER = 1, MCX_HEATBATH = 2 };
/* device storage of the spins (the reference's `spins::Vector{Int8}`, ising.jl:433): one byte per spin, or -- Ising in 2 / 3
 * dimensions with Lx % 32 == 0 -- one bit per spin.  Same RNG positions, same decisions: trajectories do not depend on it. */
enum mcx_storage { MCX_STORAGE_INT8 = 0, MCX_STORAGE_BIT = 1 };
enum mcx_init_mode { MCX_INIT_UP = 0, MCX_INIT_DOWN = 1, MCX_INIT_ZERO = 2, MCX_INIT_RANDOM = 3 };
enum mcx_flat_kind { MCX_FLAT_MUCA = 0, MCX_FLAT_WANG_LANDAU = 1 };
enum mcx_flat_observable { MCX_OBS_ENERGY = 0, MCX_OBS_SPIN2_WITH_PAIR_BOLTZMANN = 1 };

typedef struct mcx_ctx mcx_ctx;
typedef struct mcx_lattice mcx_lattice;
typedef struct mcx_pt mcx_pt;
typedef struct mcx_flat mcx_flat;
typedef struct mcx_graph mcx_graph;

/* ---- library / context ------------------------------------------------------------------ */
int32_t     mcx_abi_version(void);
const char *mcx_last_error(void);

/* Binds a context to a CUDA device.  `stream` is a cudaStream_t to enqueue on (NULL: the library
 * creates its own non-blocking stream).  Replaces nothing in the reference; plays the role of
 * parallel_backends.jl:86 `init(:mode)` for a GPU backend. */
int32_t mcx_ctx_create(int32_t device, void *stream, mcx_ctx **out);
int32_t mcx_ctx_destroy(mcx_ctx *ctx);                   /* parallel_backends.jl:104 finalize! */
int32_t mcx_ctx_set_stream(mcx_ctx *ctx, void *stream);
int32_t mcx_ctx_sync(mcx_ctx *ctx);
int32_t mcx_ctx_info(mcx_ctx *ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor,
                     uint64_t *total_mem_bytes);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int32_t mcx_ctx_launch_count(mcx_ctx *ctx, uint64_t *count);
/* Device-side waits (ticket-queue dependencies, slab neighbours, replica-exchange peers) give up after 10-25 s
 * instead of hanging the GPU and store a code into a zero-copy error word of the context.  Every later call on
 * the context (mcx_sweep, mcx_ctx_sync, mcx_observables, mcx_pt_run, mcx_pt_state, ...) then fails at once with
 * MCX_ERR_CUDA -- no synchronisation is needed to see it; the reference has no counterpart (its MPI waits block).
 * mcx_ctx_async_error reads the code (0: none) without failing; mcx_ctx_clear_error synchronises and resets it. */
int32_t mcx_ctx_async_error(mcx_ctx *ctx, int32_t *code);
int32_t mcx_ctx_clear_error(mcx_ctx *ctx);

/* ---- lattice: SpinSystems Ising / BlumeCapel constructors ---------------------------------
 * mcx_lattice_create  <- Ising(dims) ising.jl:413, IsingLatticeOptim(Lx,Ly) ising.jl:441,
 *                        BlumeCapel(dims) blume_capel.jl:500; `nchains` independent lattices
 *                        (ParallelChains, parallel_chains.jl:14).  Periodic, every dim even, >= 4.
 *                        All chains start all-up like the reference constructors. */
int32_t mcx_lattice_create(mcx_ctx *ctx, int32_t model, int32_t ndim, const int32_t *dims,
                           int32_t nchains, int32_t storage, mcx_lattice **out);
int32_t mcx_lattice_destroy(mcx_lattice *lat);
/* couplings used to form energies from the integer sums: E = -J*pair - h*spin + D*spin2
 * (ising.jl:175-178, blume_capel.jl:222-224).  The update rule itself sees only the tables. */
int32_t mcx_lattice_set_couplings(mcx_lattice *lat, double J, double h, double D);
/* RNG chain id of local chain c is first_chain_id + c (global slot id when sharded over GPUs) */
int32_t mcx_lattice_set_first_chain_id(mcx_lattice *lat, uint32_t first_chain_id);
/* sys.spins .= host (any caller-side init!), then _recompute_cached! (ising.jl:500-504).
 * host_spins is [nchains][N]. */
int32_t mcx_lattice_upload(mcx_lattice *lat, const int8_t *host_spins);
int32_t mcx_lattice_download(mcx_lattice *lat, int8_t *host_spins);      /* read sys.spins */
/* The same assignment split in two, for callers that stream configurations through the device (the reference
 * has no counterpart: `sys.spins .= host` is a host-memory copy).  _begin copies host_spins into the handle's
 * staging buffer on an internal copy stream and returns at once; the lattice itself is not touched, so sweeps
 * already queued keep running while the bytes cross PCIe.  _commit orders the context's stream after that copy,
 * converts the staging buffer into the lattice (+ _recompute_cached!) and is otherwise mcx_lattice_upload.
 * host_spins must stay valid and unchanged until _commit has been called and the stream has passed it; pinned
 * memory is what makes the copy overlap.  One upload may be pending per handle; mcx_lattice_download while one
 * is pending is MCX_ERR_STATE. */
int32_t mcx_lattice_upload_begin(mcx_lattice *lat, const int8_t *host_spins);
int32_t mcx_lattice_upload_commit(mcx_lattice *lat);
/* The same assignments / read with host buffers at one bit per spin: site i of a chain is bit (i & 7) of byte (i >> 3) of
 * that chain's N / 8 bytes, 1 = up (+1), 0 = down (-1); [nchains][N / 8].  An eighth of the bytes over PCIe; either device
 * storage takes them.  Ising only, N % 32 == 0.  _bits_begin pairs with mcx_lattice_upload_commit like _upload_begin. */
int32_t mcx_lattice_upload_bits(mcx_lattice *lat, const uint8_t *host_bits);
int32_t mcx_lattice_upload_bits_begin(mcx_lattice *lat, const uint8_t *host_bits);
int32_t mcx_lattice_download_bits(mcx_lattice *lat, uint8_t *host_bits);
/* init!(sys, :up/:down/:zero/:random; rng) ising.jl:74, blume_capel.jl:106 (INIT stream) */
int32_t mcx_lattice_init(mcx_lattice *lat, int32_t mode, uint64_t seed);

/* ---- canonical update rule: Metropolis(rng; beta) metropolis.jl:95, Glauber :118,
 * HeatBath heat_bath.jl:17, and the accept! methods metropolis.jl:14-17,121-127,
 * importance_sampling.jl:80-85, the heat-bath bodies ising.jl:43-58, blume_capel.jl:61-85.
 * The host evaluates the reference's float expression for each local configuration and passes
 * integer thresholds T in [0, 2^32] (u < p <=> m < ceil(p*2^32) for u = m*2^-32), n_labels
 * tables of table_len entries, label-major.  Index conventions (nn = 2*ndim):
 *   Ising:  idx = s*(nn+1) + nup, s in {0:down, 1:up}, nup = #up neighbours.
 *           Metropolis/Glauber: flip iff m < T.  HeatBath: new spin is up iff m < T.
 *   Blume-Capel Metropolis/Glauber: idx = (so*2 + b)*(2nn+1) + (sum+nn), so in {0,1,2} for
 *           {-1,0,+1}, b the Bool draw of _propose_state (blume_capel.jl:21-30), sum the
 *           neighbour spin sum; accept iff m < T.
 *   Blume-Capel HeatBath: idx = k*(2nn+1) + (sum+nn), k in {0,1}: new = m<T0 ? -1 : m<T1 ? 0 : +1.
 */
int32_t mcx_set_rule(mcx_lattice *lat, int32_t rule, const uint64_t *thresholds, int32_t n_labels,
                     int32_t table_len);
/* which table (ensemble) each chain currently holds: the ensemble swap of
 * replica_exchange.jl:133 moves labels, not lattices */
int32_t mcx_set_labels(mcx_lattice *lat, const int32_t *label_of_chain);
int32_t mcx_get_labels(mcx_lattice *lat, int32_t *label_of_chain);
/* PhiloxRNG(seed) and the position of the next sweep (checkpoint/restore is exact) */
int32_t mcx_set_rng(mcx_lattice *lat, uint64_t seed, uint64_t next_sweep);
int32_t mcx_get_rng(mcx_lattice *lat, uint64_t *seed, uint64_t *next_sweep);

/* sweep!(sys, alg, nsweeps): nsweeps x (colour 0 half-sweep, colour 1 half-sweep) on every chain.
 * One attempt per site per sweep = the reference's `for _ in 1:N; spin_flip!(sys, alg); end`
 * (docs/src/examples/spin_systems/pt_Ising2D.jl:52-57) in checkerboard order.  Asynchronous.
 * How the sweeps are launched (row bands, chain groups, one resident / ticket-queue launch for a whole series, a CUDA graph
 * of 32 sweeps replayed for series of >= 65 sweeps of one big lattice) never changes the trajectory; the first long series
 * of a handle pays the one-off capture of its graph (MCX_SWEEP_GRAPH=0 switches the replay off). */
int32_t mcx_sweep(mcx_lattice *lat, int64_t nsweeps);

/* measure!(measurements, sys, i) with an interval schedule (src/measurements/measurements.jl:192-200)
 * without a host round trip per measurement: nmeasure x (interval sweeps, snapshot of the sums).
 * out is int64 [nmeasure][nchains][4] = {pair_sum, spin_sum, spin2_sum, accepted}; synchronises. */
int32_t mcx_sweep_series(mcx_lattice *lat, int64_t nmeasure, int64_t interval, int64_t *out);

/* integrated_autocorrelation_time(samples; max_lag, c) (src/measurements/autocorrelations.jl:28-65) of the series the last
 * mcx_sweep_series left on the device, one value per chain, without moving the series: observable 0 = energy(sys),
 * 1 = magnetization(sys), 2 = |magnetization|; max_lag 0 = floor(n / 2); first `nmeasure` snapshots; synchronises.
 * Same estimator, same errors (MCX_ERR_ARGUMENT for the reference's ArgumentErrors); sums are deterministic block
 * reductions, so the value agrees with the sequential formula to rounding (1e-12 relative), not bit for bit. */
int32_t mcx_series_tau_int(mcx_lattice *lat, int64_t nmeasure, int32_t observable, int64_t max_lag, double c, double *tau);

/* cached sums per chain (any pointer may be NULL), synchronises:
 *   pair_sum  = sum_<ij> s_i s_j (unweighted; sys.sum_pair_interactions / J)   ising.jl:90
 *   spin_sum  = sys.sum_spins, spin2_sum = sys.sum_spins2                       blume_capel.jl:124-125
 *   accepted, steps = alg.accepted, alg.steps                                   importance_sampling.jl:26-27 */
int32_t mcx_observables(mcx_lattice *lat, int64_t *pair_sum, int64_t *spin_sum, int64_t *spin2_sum,
                        int64_t *accepted, int64_t *steps);
int32_t mcx_energies(mcx_lattice *lat, double *energy);                  /* energy(sys) per chain */
int32_t mcx_reset_counters(mcx_lattice *lat);                            /* reset!(alg) importance_sampling.jl:106 */
/* restore alg.accepted (per chain) and alg.steps from a checkpoint (checkpointing.jl:95-101) */
int32_t mcx_set_counters(mcx_lattice *lat, const int64_t *accepted, int64_t steps);
int32_t mcx_recompute(mcx_lattice *lat);                                 /* energy(sys; full=true) */
/* on (default): sweeps keep pair/spin sums current per flip, like modify! (ising.jl:200-205).
 * off: sweeps only count accepted moves; the sums are recomputed from the spins on the next
 * read (same values, fewer instructions per attempt). */
int32_t mcx_set_tracking(mcx_lattice *lat, int32_t on);
/* ---- slab decomposition: one lattice over several GPUs (not in the reference, whose
 * IsingLatticeOptim (ising.jl:430-461) is a single Vector{Int8}) -----------------------------------------
 * A handle created for dims = [Lx, Ly_local] becomes rows [row_offset, row_offset + Ly_local) of a lattice
 * with global_Ly rows.  The half-sweep kernel reads the row above / below the slab directly from the
 * neighbour handle's device memory (same process: mcx_slab_attach_local; other process / GPU: the 128-byte
 * token of mcx_slab_export, opened with CUDA IPC and read over NVLink).  Randomness is positioned by global
 * row, so trajectories equal those of the unsplit lattice.  Remote slabs are advanced by mcx_sweep (device
 * flags order the half-sweeps between GPUs, no host synchronisation); local slabs by calling
 * mcx_slab_half_sweep on every slab of the lattice in turn.  Observables, upload, download and init act on
 * the slab's own rows; sums are the slab's share (add them over slabs). */
int32_t mcx_slab_configure(mcx_lattice *lat, int32_t global_Ly, int32_t row_offset);
int32_t mcx_slab_export(mcx_lattice *lat, void *handle128);
int32_t mcx_slab_attach_ipc(mcx_lattice *lat, const void *up_handle128, const void *dn_handle128);
int32_t mcx_slab_attach_local(mcx_lattice *lat, mcx_lattice *up, mcx_lattice *dn);
int32_t mcx_slab_half_sweep(mcx_lattice *lat);
/* synchronises the stream; timed_out != 0 if a wait for a neighbour gave up (20 s) */
int32_t mcx_slab_status(mcx_lattice *lat, int32_t *timed_out, uint64_t *half_sweeps_done);

/* device pointer of the int64 [nchains][4] accumulator block {pair, spin, spin2, accepted}
 * for zero-copy plumbing (collectives) by the host runtime */
int32_t mcx_lattice_device_sums(mcx_lattice *lat, void **device_ptr);

/* ---- replica exchange / parallel tempering -------------------------------------------------
 * ReplicaExchange(backend, algs) replica_exchange.jl:52-66, ParallelTempering(betas; seed)
 * parallel_tempering.jl:24-53, update!(rx, xs) :158-178 / :227-244, index(rx) :48-50,
 * acceptance_rates :76-87, reset! :68-74.
 * n_global replicas over all ranks; this rank's lattice holds slots
 * [first_slot, first_slot + nchains).  betas[k] is the inverse temperature of ladder index k
 * (table label k).  Slot r starts at ladder index r. */
int32_t mcx_pt_create(mcx_lattice *lat, int32_t n_global, int32_t first_slot, const double *betas,
                      mcx_pt **out);
int32_t mcx_pt_destroy(mcx_pt *pt);
/* device buffer double[n_global] of per-slot energies x; mcx_pt_publish fills this rank's slice
 * from the lattice sums (energy(sys) per replica, pt_Ising2D.jl:96).  With more than one rank the
 * host runtime all-gathers the buffer in place (NCCL) between publish and exchange. */
int32_t mcx_pt_energy_buffer(mcx_pt *pt, void **device_ptr);
int32_t mcx_pt_publish(mcx_pt *pt);
/* All-gather by peer stores instead of a collective call (replaces the Allgather of replica_exchange.jl:239):
 * every rank exports a 128-byte token (CUDA IPC handles of its energy buffer and arrival counters), the host
 * runtime hands all tokens, in rank order, to mcx_pt_attach_peers.  From then on mcx_pt_publish stores this
 * rank's energies straight into every rank's buffer over NVLink and bumps its arrival counter there, and
 * mcx_pt_exchange waits on the device until every rank's energies of the round have arrived (two buffers by
 * round parity; a wait that lasts 20 s gives up and is reported by mcx_pt_peer_status). */
int32_t mcx_pt_export(mcx_pt *pt, void *handle128);
int32_t mcx_pt_attach_peers(mcx_pt *pt, int32_t nranks, int32_t rank, const void *handles /* [nranks][128] */);
int32_t mcx_pt_peer_status(mcx_pt *pt, int32_t *timed_out);
/* update!(rx, xs): all pairs of the current stage decided on the device with
 * u = EXCHANGE stream of the lower slot (replica_exchange.jl:168), labels swapped, stage toggled */
int32_t mcx_pt_exchange(mcx_pt *pt);
/* the user loop of pt_Ising2D.jl:52-57 (sweeps, update!(pt) every `interval` sweeps) in one call:
 * nrounds x (sweeps_per_round sweeps, publish, exchange); one rank or attached peers, else MCX_ERR_UNSUPPORTED.
 * For 2-D Ising int8 lattices with Lx % 32 == 0 all rounds run in ONE persistent kernel launch: the warp that
 * finishes a round's last work item stores the energies into every rank's buffer (NVLink), waits for the other
 * ranks' energies, decides the exchanges and releases the next round -- same decisions, same trajectories as
 * mcx_sweep + mcx_pt_publish + mcx_pt_exchange per round (MCX_PT_PERSIST=0 forces that path). */
int32_t mcx_pt_run(mcx_pt *pt, int64_t nrounds, int64_t sweeps_per_round);
/* how the last mcx_pt_run was executed: *path = 1 one persistent launch (*strip_rows = rows per work item), 2 = rounds
 * replayed from a CUDA graph of eight rounds whose kernels read the half-sweep index and the round from a device clock
 * (intervals of one or two sweeps, where queuing the ~25 stream operations of a round bounds the rate; MCX_PT_GRAPH=0
 * disables it), 0 = rounds queued from the host launch by launch.  Diagnostics for bench.py and the tests. */
int32_t mcx_pt_run_info(mcx_pt *pt, int32_t *path, int32_t *strip_rows);
int32_t mcx_pt_state(mcx_pt *pt, int64_t *indices /*[n] 1-based*/, int64_t *steps /*[n-1]*/,
                     int64_t *accepted /*[n-1]*/, int64_t *stage, int64_t *round);
int32_t mcx_pt_reset(mcx_pt *pt);
/* restore a checkpointed ladder: rx.indices (1-based), rx.steps, rx.accepted, rx.stage and the
 * exchange round of the EXCHANGE stream (checkpointing.jl:95-101 restoring replica_exchange.jl:13-19) */
int32_t mcx_pt_set_state(mcx_pt *pt, const int64_t *indices, const int64_t *steps, const int64_t *accepted,
                         int64_t stage, int64_t round);

/* ---- flat-histogram ensembles ---------------------------------------------------------------
 * Multicanonical(rng, bins) algorithms/multicanonical.jl:9, WangLandau(rng, bins; logf)
 * algorithms/wang_landau.jl:10, accept!(alg, x_new, x_old) importance_sampling.jl:69-78 and
 * wang_landau.jl:29-37, record_visit! ensembles/multicanonical.jl:25-30, update! :32-44 and
 * ensembles/wang_landau.jl:23, reset! algorithms/multicanonical.jl:27-33,
 * merge_histograms!/distribute_logweight! parallel_multicanonical.jl:38-73.
 * Chains are serial in the global observable, so parallelism is across chains only; the
 * sweep visits every site once, serially, in checkerboard order (all colour-0 slots ascending,
 * then all colour-1 slots; FLAT stream), which lets a warp prepare the proposals of a batch of
 * same-colour sites in parallel.  Integer bins start:step:start+step*(n-1)
 * (BinnedObject, binned_object.jl:13-24); an out-of-range lookup makes the next synchronising
 * call return MCX_ERR_BOUNDS (out_of_range_policy 0, the reference's behaviour,
 * test/test_multicanonical.jl:39-43) or is rejected as a move (policy 1: energy windows, which the
 * reference does not have).  Multicanonical chains share one log-weight table and one histogram
 * (the state after merge_histograms!/distribute_logweight!); Wang-Landau chains own one table
 * each ([nchains][nbins]), like one WangLandauEnsemble per algorithm object. */
int32_t mcx_flat_create(mcx_lattice *lat, int32_t kind, int32_t observable, int64_t bin_start,
                        int64_t bin_step, int64_t nbins, double beta_pair, int32_t out_of_range_policy,
                        mcx_flat **out);
int32_t mcx_flat_destroy(mcx_flat *flat);
int32_t mcx_flat_set_logweight(mcx_flat *flat, const double *logweight);
int32_t mcx_flat_get_logweight(mcx_flat *flat, double *logweight);
int32_t mcx_flat_get_histogram(mcx_flat *flat, double *histogram);
int32_t mcx_flat_reset_histogram(mcx_flat *flat);
int32_t mcx_flat_set_logf(mcx_flat *flat, double logf);
int32_t mcx_flat_get_logf(mcx_flat *flat, double *logf);
int32_t mcx_flat_sweep(mcx_flat *flat, int64_t nsweeps);
int32_t mcx_flat_update(mcx_flat *flat);                 /* update!(ens): muca :simple, WL halves logf */
int32_t mcx_flat_device_histogram(mcx_flat *flat, void **device_ptr, int64_t *nbins);
int32_t mcx_flat_device_logweight(mcx_flat *flat, void **device_ptr, int64_t *nbins);

/* ---- general topologies: IsingGraph / IsingMatrix with site fields -----------------------------
 * mcx_graph_create <- Ising(graph::SimpleGraph, J::Real; h) / IsingGraph ising.jl:117-139 (J_ij == NULL: one global J),
 *                     Ising(J::SparseMatrixCSC; h) / IsingMatrix ising.jl:264-293 and Ising(graph, J::Vector) :383-404
 *                     (J_ij: one coupling per CSR entry), Ising(dims; periodic=false) :406-417 (a grid graph).
 * rowptr[n + 1] / col[nnz]: 0-based neighbour lists in the reference's adjacency order (ascending neighbour index);
 * field_mode 0: h = 0, 1: uniform h, 2: per-site h_i[n] (ising.jl:95-115).  An asymmetric neighbour relation or J_ij
 * is MCX_ERR_STATE (the reference's AssertionError "Sparse J must be symmetric", :250-262).  `nchains` independent copies.
 * Sweeps visit the colour classes of a greedy first-fit colouring in site order (the checkerboard generalised); the
 * stream of an attempt is (chain, t = ncolours * sweep + colour, slot = rank of the site inside its colour class) of
 * RNG layout v1.  Couplings are doubles, so the acceptance is the reference's Float64 expression evaluated on the
 * device (accept! metropolis.jl:14-17,121-127; _accept! importance_sampling.jl:80-85; heat bath ising.jl:43-58). */
int32_t mcx_graph_create(mcx_ctx *ctx, int64_t n, const int64_t *rowptr, const int64_t *col, const double *J_ij, double J,
                         int32_t field_mode, double h, const double *h_i, int32_t nchains, mcx_graph **out);
int32_t mcx_graph_destroy(mcx_graph *g);
/* number of colours and (optional, int32[n]) the colour of every site */
int32_t mcx_graph_colours(mcx_graph *g, int32_t *ncolours, int32_t *colour_of_site);
int32_t mcx_graph_upload(mcx_graph *g, const int8_t *host_spins);      /* sys.spins .= host, [nchains][n], -1 / +1 */
int32_t mcx_graph_download(mcx_graph *g, int8_t *host_spins);
int32_t mcx_graph_init(mcx_graph *g, int32_t mode, uint64_t seed);     /* init!(sys, :up / :down / :random; rng) ising.jl:74 */
int32_t mcx_graph_set_rule(mcx_graph *g, int32_t rule, double beta);   /* Metropolis / Glauber / HeatBath(rng; beta) */
int32_t mcx_graph_set_rng(mcx_graph *g, uint64_t seed, uint64_t next_sweep, uint32_t first_chain_id);
int32_t mcx_graph_get_rng(mcx_graph *g, uint64_t *seed, uint64_t *next_sweep);
int32_t mcx_graph_sweep(mcx_graph *g, int64_t nsweeps);                /* nsweeps x every colour class once; asynchronous */
/* per chain (any pointer may be NULL), formed from the spins like _recompute_cached! (ising.jl:127-144); synchronises:
 * pair_sum = sum_pair_interactions, spin_sum = sum_spins, field_sum = sum_field_interactions, alg.accepted, alg.steps */
int32_t mcx_graph_observables(mcx_graph *g, double *pair_sum, int64_t *spin_sum, double *field_sum, int64_t *accepted,
                              int64_t *steps);
int32_t mcx_graph_energies(mcx_graph *g, double *energy);              /* energy(sys; full=true) per chain */
int32_t mcx_graph_reset_counters(mcx_graph *g);

#ifdef __cplusplus
}
#endif
#endif
