// mcx_internal.h -- host-side handle layouts and kernel launcher prototypes (not part of the ABI).
#pragma once
#include "mcx_common.cuh"

struct mcx_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int sm_count, cc_major, cc_minor;
    size_t total_mem;
    uint64_t launches;
    // auxiliary streams for overlapping the half-sweeps of chain groups (created on first use)
    cudaStream_t aux[16];
    cudaEvent_t aux_fork, aux_join[16];
    bool aux_ready;
    // sticky error word of the device-side spin waits (ticket-queue dependencies, slab neighbours, replica-exchange
    // peers): zero-copy host memory, written by the kernel that gives up a wait, read by the host on every later
    // call of this context without a synchronisation (mcx::async_error)
    int *h_err, *d_err;
};

// slab decomposition state of a lattice handle (k_slab.cu)
struct mcx_slab {
    int global_Ly;
    bool attached, remote;             // remote: neighbours live in other processes (IPC), ordered by device flags
    int colour;                        // colour of the next half-sweep
    unsigned long long epoch;          // half-sweeps completed since attach
    unsigned long long *d_ctl;         // control block, layout in mcx_common.cuh (LatView::slab_ctl)
    void *ipc_opened[4];
};

struct mcx_lattice {
    mcx_ctx *ctx;
    mcx::LatView view;
    int model, ndim, nn, storage, nchains;
    int dims[3];
    int64_t N;
    long long *d_sums;          // [nchains][SUM_FIELDS]
    int rule, n_labels, table_len;
    uint32_t *d_thi, *d_tlo;    // [n_labels][table_len]: T >> 16 (0..65536) and T & 0xffff
    uint64_t *h_table;          // host copy of the thresholds
    int32_t *d_labels;          // [nchains]
    uint64_t seed, sweep;
    uint32_t first_chain;
    int64_t steps;              // attempts per chain since the last reset (same for every chain)
    double J, h, D;
    int8_t *d_staging;          // [nchains][N] reference-order spins for upload/download
    void *d_hostbits;           // [nchains][N / 8] reference-order spins at one bit each (mcx_lattice_upload_bits / _download_bits)
    bool pending_bits;          // the pending split upload is a bit buffer (in d_hostbits), not an int8 one
    // split upload (mcx_lattice_upload_begin / _commit): H2D into d_staging on a copy stream of the handle
    cudaStream_t copy_stream;
    cudaEvent_t ev_copied, ev_packed;   // copy finished; last conversion out of d_staging finished
    bool upload_pending, packed_recorded;
    // series of sweeps of one big lattice replayed from a CUDA graph (launch_sweeps_ising2d_banded_graph): the device clock
    // its kernels read, the instantiated graph of sweep_graph_K sweeps and what it was captured for
    unsigned long long *d_tclock;
    cudaGraphExec_t sweep_graph;
    int sweep_graph_K;
    uint64_t sweep_graph_launches;
    struct SweepGraphKey { uint64_t seed; uint32_t first_chain; int rule, track, bands, band_rows; const void *planes, *thi, *labels, *sums; } sweep_graph_key;
    void *d_queue;              // k_queue.cu / k_persist.cu: control words (ticket counter, ...) and per-item progress words
    size_t queue_bytes;
    long long *d_series;        // mcx_sweep_series: snapshots of the sums, grown on demand
    size_t series_bytes;
    int64_t series_n;           // snapshots the last mcx_sweep_series left in d_series
    double *d_tau;              // mcx_series_tau_int: centred series per chain + the results
    size_t tau_bytes;
    bool fast2d;                // Lx % 32 == 0 && ndim == 2: row-aligned 128-bit kernels apply
    bool track_sums;            // fast kernels accumulate pair/spin sums per flip (else recompute lazily)
    bool sums_dirty;            // pair/spin sums are stale (untracked sweeps ran)
    mcx_slab *slab;             // non-null: this handle is a slab of a taller lattice
};

struct mcx_pt {
    mcx_lattice *lat;
    int n, first_slot;
    void *d_dev;                // k_persist.cu: device copy of the ladder view the persistent rounds read
    int persist_R;              // tallest strip the in-kernel rounds last ran with (reported by mcx_pt_run_info)
    int last_path;              // 0: host-queued rounds, 1: one persistent launch (k_persist.cu), 2: rounds replayed from a CUDA graph
    double *d_betas;            // [n] ladder
    double *d_x;                // [n] per-slot energies
    int32_t *d_index;           // [n] 0-based ladder index held by slot
    int32_t *d_slot_of;         // [n] inverse permutation
    long long *d_steps, *d_accepted;   // [n-1]
    int stage;
    uint64_t round;
    // all-gather by peer stores (mcx_pt_attach_peers): d_x holds two buffers of n energies (round parity);
    // the publish kernel stores this rank's slice into every rank's buffer over NVLink and bumps its arrival
    // counter there, the exchange kernel waits for all counters -- no collective call, no host sync
    bool peers;
    int nranks, rank;
    unsigned long long *d_arrived;     // [kMaxPtRanks] rounds published by each rank (they write it)
    double **d_peer_x;                 // [nranks] every rank's d_x (own included)
    unsigned long long **d_peer_arrived;   // [nranks] every rank's d_arrived
    int *d_err;
    void *ipc_opened[2 * 64];
    // rounds replayed from a CUDA graph (mcx_pt_run with short intervals): the clock the captured kernels read, the
    // instantiated graph of graph_rounds rounds and what it was captured for
    mcx::PtClock *d_clock;
    cudaGraphExec_t graph_exec;
    int graph_rounds;
    uint64_t graph_launches;    // kernel launches one replay stands for
    struct GraphKey { uint64_t seed; uint32_t first_chain; int rule, S, n, nchains, peers; const void *thi, *labels, *sums, *planes, *x; } graph_key;
};
constexpr int kMaxPtRanks = 64;

struct mcx_flat {
    mcx_lattice *lat;
    int kind, observable, policy, ntables;
    int64_t start, step, nbins;
    double beta_pair, logf;
    double *d_logweight;        // [nbins] shared by every chain of this rank
    void *d_histogram;          // [nbins] unsigned long long visit counters
    int8_t *d_spins;            // [N][nchains] chain-interleaved copy used by the serial sweeps
    long long *d_state;         // [nchains][4]: pair, spin, spin2, accepted
    int *d_error;               // out-of-range flag
    uint64_t sweep;
    bool spins_valid;
};

namespace mcx {

// Tuning and test hooks (MCX_* environment variables), read once per public sweep call -- not per launch;
// -1: variable not set.
struct Knobs {
    int rows_per_strip, ctas_per_sm, variant, full, groups, bands, bc2d, ising3d;
    int resident, resident_cluster, resident_rows, resident_threads, force_generic;
    int queue_rows;   // MCX_QUEUE_ROWS: strip height of the ticket-queue kernel (tuning hook)
    int queue;        // MCX_QUEUE: 1 = series of sweeps through the ticket-queue kernel (k_queue.cu)
    int queue_grid;   // MCX_QUEUE_GRID: CTAs of the persistent rounds kernel (tuning hook; default: every resident slot)
    int sweep_graph;  // MCX_SWEEP_GRAPH: 0 = series of row-band sweeps are never replayed from a CUDA graph; n > 1: n sweeps per replay (default 32)
    int pt_graph;     // MCX_PT_GRAPH: 0 = mcx_pt_run never replays its rounds from a CUDA graph
    int pt_persist;   // MCX_PT_PERSIST: 1 = mcx_pt_run as one persistent launch whenever the shape allows, 0 = never
    int flat_window;  // MCX_FLAT_WINDOW: 0 = flat-histogram chains read the log-weight table from global memory (no shared-memory window)
    int band_rows;    // MCX_BAND_ROWS: strip height of the row-band launches (tuning hook; default 16)
    int wl_spec;      // MCX_WL_SPEC: Wang-Landau attempts decided at once (0 = serial loop, 8, 32; unset = adaptive)
};
const Knobs &knobs();
void knobs_refresh();

// chain sub-range, row band and stream of the launch being issued by the series launchers (k_ising2d.cu);
// default: the whole batch on the context's stream.  Read by the half-sweep launchers of k_ising2d.cu / k_bc2d.cu.
struct LaunchRange {
    int chain0 = 0, nchains = -1;
    int row0 = 0, nrows = -1;              // row band [row0, row0 + nrows) of the lattice (whole lattice: nrows < 0)
    int R = 0;                             // strip height of a band launch
    cudaStream_t stream = nullptr;
    bool use_stream = false;
};
extern thread_local LaunchRange g_launch_range;

// k_generic.cu
void launch_pack(mcx_lattice *lat);      // staging (reference order, -1/0/+1) -> colour planes
void launch_unpack(mcx_lattice *lat);    // colour planes -> staging
void launch_init(mcx_lattice *lat, int mode, uint64_t seed);
void launch_recompute(mcx_lattice *lat);
void launch_sweep_generic(mcx_lattice *lat, int colour, uint64_t t);

// k_ising2d.cu
bool launch_sweep_ising2d(mcx_lattice *lat, int colour, uint64_t t);   // false: shape not supported
// nsweeps sweeps of a small batch with the chains dealt into groups on auxiliary streams, so that one
// group's launch fills the SMs that the previous launch's tail leaves idle; false: not applicable
bool launch_sweeps_ising2d_grouped(mcx_lattice *lat, int64_t nsweeps);
// nsweeps sweeps of one big lattice with its rows dealt into bands on auxiliary streams; a band's half-sweep
// waits (events) only for its own and its two neighbour bands' previous half-sweep; false: not applicable
bool launch_sweeps_ising2d_banded(mcx_lattice *lat, int64_t nsweeps);
int64_t launch_sweeps_ising2d_banded_graph(mcx_lattice *lat, int64_t nsweeps);   // sweeps done by graph replays (0: not applicable)
// k_bc2d.cu: vectorised 2-D Blume-Capel half-sweep (Metropolis / Glauber, Lx % 32 == 0)
bool launch_sweep_bc2d(mcx_lattice *lat, int colour, uint64_t t);       // false: not applicable, nothing launched

// k_bits.cu: MCX_STORAGE_BIT (one bit per spin; Ising, Lx % 32 == 0, 2-D / 3-D): layout conversion, init, recompute and the
// 2-D half-sweep (whole batch, chain group or row band as set in g_launch_range); host bit buffers <-> staging for both storages
bool bits_shape_ok(int model, int ndim, const int32_t *dims);
void launch_pack_bits(mcx_lattice *lat);
void launch_unpack_bits(mcx_lattice *lat);
void launch_init_bits(mcx_lattice *lat, int mode, uint64_t seed);
void launch_recompute_bits(mcx_lattice *lat);
void launch_half_sweep_bits2d(mcx_lattice *lat, int colour, uint64_t t);
void launch_hostbits_to_staging(mcx_lattice *lat, const void *d_bits, cudaStream_t stream);
void launch_staging_to_hostbits(mcx_lattice *lat, void *d_bits, cudaStream_t stream);

// k_ising3d.cu: vectorised 3-D Ising half-sweep (Lx % 32 == 0), int8 or bit planes
bool launch_sweep_ising3d(mcx_lattice *lat, int colour, uint64_t t);    // false: not applicable, nothing launched

// k_bc3d.cu: vectorised 3-D Blume-Capel half-sweep, all three rules (Lx % 32 == 0)
bool launch_sweep_bc3d(mcx_lattice *lat, int colour, uint64_t t);      // false: not applicable, nothing launched

// k_slab.cu
int32_t slab_half_sweep(mcx_lattice *lat);
void slab_free(mcx_lattice *lat);

// k_resident.cu: nsweeps whole sweeps of lat->sweep .. in one launch, lattice resident in cluster shared memory
bool launch_sweeps_resident(mcx_lattice *lat, int64_t nsweeps);         // false: not applicable, nothing launched
bool launch_recompute_ising2d(mcx_lattice *lat);
bool launch_pack_ising2d(mcx_lattice *lat);
bool launch_unpack_ising2d(mcx_lattice *lat);
bool launch_init_ising2d(mcx_lattice *lat, int mode, uint64_t seed);                        // false: shape not supported

// k_queue.cu: nsweeps whole sweeps in one launch, work items of all half-sweeps taken from one ticket counter
bool launch_sweeps_ising2d_queue(mcx_lattice *lat, int64_t nsweeps);   // false: not applicable, nothing launched
// k_persist.cu: nrounds x (sweeps_per_round sweeps, energies to all ranks, exchange) in ONE launch; false: not applicable
bool launch_pt_rounds_persistent(mcx_pt *pt, int64_t nrounds, int64_t sweeps_per_round);

// codes a kernel stores into mcx_ctx::d_err when a device-side wait gives up
enum { ASYNC_ERR_QUEUE_DEP = 1, ASYNC_ERR_SLAB = 2, ASYNC_ERR_PT_PEERS = 3, ASYNC_ERR_PT_ROUND = 4 };
const char *async_error_text(int code);

// k_rows8.cu
bool launch_sweep_rows8(mcx_lattice *lat, int colour, uint64_t t);     // false: shape not supported

// k_pt.cu
void launch_pt_publish(mcx_pt *pt, const mcx::PtClock *clock = nullptr);     // clock: round read on the device (graph replay)
void launch_pt_exchange(mcx_pt *pt, const mcx::PtClock *clock = nullptr);
void launch_pt_clock_set(mcx_pt *pt);                      // clock <- the host's sweep index and round
void launch_pt_clock_advance(mcx_pt *pt, int64_t sweeps);  // last node of a captured round
void launch_pt_round_tail(mcx_pt *pt, int64_t sweeps);     // publish + wait + exchange + clock advance of a captured round in one launch
extern thread_local const unsigned long long *g_t_clock;   // k_ising2d launches add *g_t_clock to their half-sweep index

// k_flat.cu
void launch_flat_load(mcx_flat *f);      // planes -> interleaved + state
void launch_flat_store(mcx_flat *f);     // interleaved -> planes, sums
void launch_flat_sweep(mcx_flat *f, uint64_t sweep0, int nsweeps);
void launch_flat_update(mcx_flat *f);

}  // namespace mcx
