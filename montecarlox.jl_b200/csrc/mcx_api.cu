// mcx_api.cu -- the C ABI of libmcx_b200.so (include/mcx_b200.h).  Host-side handle management only;
// all compute is in the k_*.cu kernels.  There is no CPU fallback: every path below either
// launches CUDA work or fails with MCX_ERR_CUDA.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "mcx_internal.h"

using namespace mcx;

static thread_local char g_err[512] = "";

static int32_t fail(int32_t code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// shared with k_flat.cu
int32_t mcx_set_error(int32_t code, const char *msg) { return fail(code, "%s", msg); }

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(MCX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define REQUIRE(cond, code, ...)                    \
    do {                                            \
        if (!(cond)) return fail(code, __VA_ARGS__); \
    } while (0)

static int32_t check_launch(mcx_ctx *ctx)
{
    (void)ctx;
    CUDA_TRY(cudaGetLastError());
    return MCX_OK;
}

// ------------------------------------------------------------------------------------ knobs
namespace mcx {
static thread_local Knobs g_knobs;
static thread_local bool g_knobs_ready = false;
static int knob(const char *name)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : -1;
}
void knobs_refresh()
{
    Knobs &k = g_knobs;
    k.rows_per_strip = knob("MCX_ROWS_PER_STRIP"); k.ctas_per_sm = knob("MCX_CTAS_PER_SM"); k.variant = knob("MCX_VARIANT");
    k.full = knob("MCX_FULL"); k.groups = knob("MCX_GROUPS"); k.bands = knob("MCX_BANDS"); k.bc2d = knob("MCX_BC2D");
    k.ising3d = knob("MCX_ISING3D"); k.resident = knob("MCX_RESIDENT"); k.resident_cluster = knob("MCX_RESIDENT_CLUSTER");
    k.resident_rows = knob("MCX_RESIDENT_ROWS"); k.resident_threads = knob("MCX_RESIDENT_THREADS");
    k.force_generic = knob("MCX_FORCE_GENERIC"); k.wl_spec = knob("MCX_WL_SPEC"); k.band_rows = knob("MCX_BAND_ROWS"); k.queue = knob("MCX_QUEUE"); k.queue_rows = knob("MCX_QUEUE_ROWS");
    k.queue_grid = knob("MCX_QUEUE_GRID"); k.pt_persist = knob("MCX_PT_PERSIST"); k.pt_graph = knob("MCX_PT_GRAPH"); k.sweep_graph = knob("MCX_SWEEP_GRAPH"); k.flat_window = knob("MCX_FLAT_WINDOW");
    g_knobs_ready = true;
}
const Knobs &knobs()
{
    if (!g_knobs_ready) knobs_refresh();
    return g_knobs;
}
const char *async_error_text(int code)
{
    switch (code) {
    case ASYNC_ERR_QUEUE_DEP: return "sweep series (ticket queue): a dependency wait timed out";
    case ASYNC_ERR_SLAB: return "slab half-sweep: the wait for a neighbour GPU timed out";
    case ASYNC_ERR_PT_PEERS: return "replica exchange: the wait for another rank's energies timed out";
    case ASYNC_ERR_PT_ROUND: return "parallel-tempering rounds: the wait for the previous round's exchange timed out";
    default: return "a device-side wait timed out";
    }
}
}  // namespace mcx

// A kernel that gave up a device-side wait has stored its code into the context's zero-copy error word: every later
// call on the context fails at once (no synchronisation needed to see it) until mcx_ctx_clear_error.
static int32_t async_error(mcx_ctx *ctx)
{
    const int code = ctx->h_err ? *(volatile int *)ctx->h_err : 0;
    if (code == 0) return MCX_OK;
    return fail(MCX_ERR_CUDA, "%s; lattices of this context are not valid (mcx_ctx_clear_error to go on)", async_error_text(code));
}
#define ASYNC_CHECK(ctx)                         \
    do {                                         \
        const int32_t a__ = async_error(ctx);    \
        if (a__ != MCX_OK) return a__;           \
    } while (0)


extern "C" {

int32_t mcx_abi_version(void) { return MCX_ABI_VERSION; }
const char *mcx_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------ context
int32_t mcx_ctx_create(int32_t device, void *stream, mcx_ctx **out)
{
    REQUIRE(out, MCX_ERR_ARGUMENT, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(MCX_ERR_CUDA, "no CUDA device available (%s); libmcx_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    REQUIRE(device >= 0 && device < count, MCX_ERR_ARGUMENT, "device %d out of range [0,%d)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major >= 10, MCX_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
            device, prop.major, prop.minor);
    mcx_ctx *c = new (std::nothrow) mcx_ctx();
    REQUIRE(c, MCX_ERR_STATE, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->total_mem = prop.totalGlobalMem;
    c->launches = 0;
    c->h_err = c->d_err = nullptr;
    if (cudaHostAlloc((void **)&c->h_err, sizeof(int) * 4, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&c->d_err, c->h_err, 0) != cudaSuccess) {
        const cudaError_t he = cudaGetLastError();
        if (c->h_err) cudaFreeHost(c->h_err);
        delete c;
        return fail(MCX_ERR_CUDA, "zero-copy error word: %s", cudaGetErrorString(he));
    }
    memset(c->h_err, 0, sizeof(int) * 4);
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) {
            cudaFreeHost(c->h_err);
            delete c;
            return fail(MCX_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(se));
        }
        c->own_stream = true;
    }
    *out = c;
    return MCX_OK;
}

int32_t mcx_ctx_destroy(mcx_ctx *ctx)
{
    if (!ctx) return MCX_OK;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->aux_ready) {
        for (int i = 0; i < 16; ++i) { cudaStreamDestroy(ctx->aux[i]); cudaEventDestroy(ctx->aux_join[i]); }
        cudaEventDestroy(ctx->aux_fork);
    }
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    delete ctx;
    return MCX_OK;
}

int32_t mcx_ctx_set_stream(mcx_ctx *ctx, void *stream)
{
    REQUIRE(ctx, MCX_ERR_ARGUMENT, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) {
        if (!stream) return MCX_OK;
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return MCX_OK;
}

int32_t mcx_ctx_sync(mcx_ctx *ctx)
{
    REQUIRE(ctx, MCX_ERR_ARGUMENT, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ASYNC_CHECK(ctx);
    return MCX_OK;
}

int32_t mcx_ctx_async_error(mcx_ctx *ctx, int32_t *code)
{
    REQUIRE(ctx && code, MCX_ERR_ARGUMENT, "NULL argument");
    *code = ctx->h_err ? *(volatile int *)ctx->h_err : 0;
    return MCX_OK;
}

int32_t mcx_ctx_clear_error(mcx_ctx *ctx)
{
    REQUIRE(ctx, MCX_ERR_ARGUMENT, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_err) *(volatile int *)ctx->h_err = 0;
    return MCX_OK;
}

int32_t mcx_ctx_info(mcx_ctx *ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, uint64_t *total_mem_bytes)
{
    REQUIRE(ctx, MCX_ERR_ARGUMENT, "ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (total_mem_bytes) *total_mem_bytes = ctx->total_mem;
    return MCX_OK;
}

int32_t mcx_ctx_launch_count(mcx_ctx *ctx, uint64_t *count)
{
    REQUIRE(ctx && count, MCX_ERR_ARGUMENT, "NULL argument");
    *count = ctx->launches;
    return MCX_OK;
}

// ------------------------------------------------------------------------------------ lattice
static void lattice_free(mcx_lattice *lat)
{
    if (!lat) return;
    cudaSetDevice(lat->ctx->device);
    slab_free(lat);
    cudaFree(lat->view.planes);
    cudaFree(lat->d_sums);
    cudaFree(lat->d_thi);
    cudaFree(lat->d_tlo);
    cudaFree(lat->d_labels);
    cudaFree(lat->d_staging);
    cudaFree(lat->d_hostbits);
    cudaFree(lat->d_queue);
    if (lat->sweep_graph) cudaGraphExecDestroy(lat->sweep_graph);
    cudaFree(lat->d_tclock);
    cudaFree(lat->d_series);
    cudaFree(lat->d_tau);
    if (lat->copy_stream) {
        cudaStreamSynchronize(lat->copy_stream);
        cudaStreamDestroy(lat->copy_stream);
        cudaEventDestroy(lat->ev_copied);
        cudaEventDestroy(lat->ev_packed);
    }
    free(lat->h_table);
    delete lat;
}

int32_t mcx_lattice_create(mcx_ctx *ctx, int32_t model, int32_t ndim, const int32_t *dims, int32_t nchains,
                           int32_t storage, mcx_lattice **out)
{
    REQUIRE(ctx && dims && out, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(model == MCX_ISING || model == MCX_BLUME_CAPEL, MCX_ERR_ARGUMENT, "unknown model %d", model);
    REQUIRE(ndim >= 1 && ndim <= 3, MCX_ERR_ARGUMENT, "ndim must be 1, 2 or 3 (got %d)", ndim);
    REQUIRE(nchains >= 1 && nchains <= 65535, MCX_ERR_ARGUMENT, "nchains must be in [1, 65535] (got %d)", nchains);
    REQUIRE(storage == MCX_STORAGE_INT8 || storage == MCX_STORAGE_BIT, MCX_ERR_ARGUMENT, "unknown storage %d", storage);
    REQUIRE(storage == MCX_STORAGE_INT8 || bits_shape_ok(model, ndim, dims), MCX_ERR_UNSUPPORTED,
            "one-bit storage holds Ising lattices in 2 or 3 dimensions with Lx %% 32 == 0 (Blume-Capel has three states; other shapes: int8)");
    int64_t N = 1;
    for (int d = 0; d < ndim; ++d) {
        REQUIRE(dims[d] >= 4 && dims[d] % 2 == 0, MCX_ERR_ARGUMENT,
                "periodic checkerboard needs every dimension even and >= 4 (dims[%d] = %d)", d, dims[d]);
        N *= dims[d];
    }
    REQUIRE(N / 16 < ((int64_t)1 << 32), MCX_ERR_ARGUMENT, "lattice too large for the 32-bit block counter");
    CUDA_TRY(cudaSetDevice(ctx->device));
    mcx_lattice *lat = new (std::nothrow) mcx_lattice();
    REQUIRE(lat, MCX_ERR_STATE, "out of host memory");
    memset(lat, 0, sizeof(*lat));
    lat->ctx = ctx;
    lat->model = model; lat->ndim = ndim; lat->nn = 2 * ndim; lat->storage = storage; lat->nchains = nchains;
    lat->dims[0] = dims[0]; lat->dims[1] = ndim > 1 ? dims[1] : 1; lat->dims[2] = ndim > 2 ? dims[2] : 1;
    lat->N = N;
    lat->J = 1.0; lat->h = 0.0; lat->D = 0.0;
    lat->rule = -1;
    LatView &v = lat->view;
    v.Lx = lat->dims[0]; v.Ly = lat->dims[1]; v.Lz = lat->dims[2]; v.half = v.Lx / 2;
    v.ndim = ndim; v.nn = 2 * ndim; v.model = model; v.nchains = nchains;
    v.halfN = N / 2;
    v.plane_stride = ((storage == MCX_STORAGE_BIT ? v.halfN / 8 : v.halfN) + 255) / 256 * 256;     // bytes
    lat->fast2d = (ndim == 2) && (v.Lx % 32 == 0);
    lat->track_sums = true;
    const size_t plane_bytes = (size_t)v.plane_stride * 2 * (size_t)nchains;
    // slack behind the last plane: the strip loop's L2 prefetch (k_strip.cuh) names a few rows past a strip
    const size_t plane_slack = ((size_t)8 * (size_t)v.half + 255) / 256 * 256;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&v.planes, plane_bytes + plane_slack)) != cudaSuccess ||
        (e = cudaMalloc((void **)&lat->d_sums, sizeof(long long) * SUM_FIELDS * (size_t)nchains)) != cudaSuccess ||
        (e = cudaMalloc((void **)&lat->d_labels, sizeof(int32_t) * (size_t)nchains)) != cudaSuccess) {
        lattice_free(lat);
        return fail(MCX_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    v.up_planes = v.dn_planes = v.planes; v.row_offset = 0; v.slab_sides = 3; v.slab_ctl = nullptr;
    v.err = ctx->d_err;
    cudaMemsetAsync(lat->d_labels, 0, sizeof(int32_t) * (size_t)nchains, ctx->stream);
    cudaMemsetAsync(lat->d_sums, 0, sizeof(long long) * SUM_FIELDS * (size_t)nchains, ctx->stream);
    // constructors start all-up (ising.jl:118, blume_capel.jl:156)
    launch_init(lat, MCX_INIT_UP, 0);
    launch_recompute(lat);
    int32_t st = check_launch(ctx);
    if (st != MCX_OK) { lattice_free(lat); return st; }
    *out = lat;
    return MCX_OK;
}

int32_t mcx_lattice_destroy(mcx_lattice *lat)
{
    if (!lat) return MCX_OK;
    cudaSetDevice(lat->ctx->device);
    cudaStreamSynchronize(lat->ctx->stream);
    lattice_free(lat);
    return MCX_OK;
}

int32_t mcx_lattice_set_couplings(mcx_lattice *lat, double J, double h, double D)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    lat->J = J; lat->h = h; lat->D = D;
    return MCX_OK;
}

int32_t mcx_lattice_set_first_chain_id(mcx_lattice *lat, uint32_t first_chain_id)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    lat->first_chain = first_chain_id;
    return MCX_OK;
}

static int32_t ensure_staging(mcx_lattice *lat)
{
    if (lat->d_staging) return MCX_OK;
    CUDA_TRY(cudaMalloc((void **)&lat->d_staging, (size_t)lat->N * (size_t)lat->nchains));
    return MCX_OK;
}

int32_t mcx_lattice_upload(mcx_lattice *lat, const int8_t *host_spins)
{
    REQUIRE(lat && host_spins, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(!lat->upload_pending, MCX_ERR_STATE, "an upload is pending on this handle (the staging buffer is in use): commit it first");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    int32_t st = ensure_staging(lat);
    if (st != MCX_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(lat->d_staging, host_spins, (size_t)lat->N * (size_t)lat->nchains,
                             cudaMemcpyHostToDevice, lat->ctx->stream));
    launch_pack(lat);
    if (lat->copy_stream) {      // a later upload_begin may overwrite d_staging only after this conversion
        CUDA_TRY(cudaEventRecord(lat->ev_packed, lat->ctx->stream));
        lat->packed_recorded = true;
    }
    launch_recompute(lat);
    lat->sums_dirty = false;
    return check_launch(lat->ctx);
}

int32_t mcx_lattice_upload_begin(mcx_lattice *lat, const int8_t *host_spins)
{
    REQUIRE(lat && host_spins, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(!lat->upload_pending, MCX_ERR_STATE, "an upload is already pending on this handle: commit it first");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    int32_t st = ensure_staging(lat);
    if (st != MCX_OK) return st;
    if (!lat->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&lat->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&lat->ev_copied, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&lat->ev_packed, cudaEventDisableTiming));
        // whatever the context's stream has queued so far may still read or write d_staging
        CUDA_TRY(cudaEventRecord(lat->ev_packed, lat->ctx->stream));
        lat->packed_recorded = true;
    }
    if (lat->packed_recorded) CUDA_TRY(cudaStreamWaitEvent(lat->copy_stream, lat->ev_packed, 0));
    CUDA_TRY(cudaMemcpyAsync(lat->d_staging, host_spins, (size_t)lat->N * (size_t)lat->nchains,
                             cudaMemcpyHostToDevice, lat->copy_stream));
    CUDA_TRY(cudaEventRecord(lat->ev_copied, lat->copy_stream));
    lat->upload_pending = true;
    return MCX_OK;
}

int32_t mcx_lattice_upload_commit(mcx_lattice *lat)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    REQUIRE(lat->upload_pending, MCX_ERR_STATE, "no upload pending on this handle: call mcx_lattice_upload_begin first");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    CUDA_TRY(cudaStreamWaitEvent(lat->ctx->stream, lat->ev_copied, 0));
    if (lat->pending_bits) launch_hostbits_to_staging(lat, lat->d_hostbits, lat->ctx->stream);
    lat->pending_bits = false;
    launch_pack(lat);
    CUDA_TRY(cudaEventRecord(lat->ev_packed, lat->ctx->stream));
    lat->packed_recorded = true;
    launch_recompute(lat);
    lat->sums_dirty = false;
    lat->upload_pending = false;
    return check_launch(lat->ctx);
}

int32_t mcx_lattice_download(mcx_lattice *lat, int8_t *host_spins)
{
    REQUIRE(lat && host_spins, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(!lat->upload_pending, MCX_ERR_STATE, "an upload is pending on this handle (the staging buffer is in use): commit it first");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    int32_t st = ensure_staging(lat);
    if (st != MCX_OK) return st;
    launch_unpack(lat);
    CUDA_TRY(cudaMemcpyAsync(host_spins, lat->d_staging, (size_t)lat->N * (size_t)lat->nchains,
                             cudaMemcpyDeviceToHost, lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    return check_launch(lat->ctx);
}

// ---- the same transfers with host buffers at one bit per spin (bit i & 7 of byte i >> 3 of a chain = site i, 1 = up):
// an eighth of the bytes over PCIe; converted on the device through the staging buffer, so either storage takes them
static int32_t ensure_hostbits(mcx_lattice *lat)
{
    REQUIRE(lat->model == MCX_ISING, MCX_ERR_UNSUPPORTED, "bit buffers hold two-state (Ising) spins only");
    REQUIRE(lat->N % 32 == 0, MCX_ERR_UNSUPPORTED, "bit buffers need a site count divisible by 32 (N = %lld)", (long long)lat->N);
    if (lat->d_hostbits) return MCX_OK;
    CUDA_TRY(cudaMalloc(&lat->d_hostbits, (size_t)lat->N / 8 * (size_t)lat->nchains));
    return MCX_OK;
}

int32_t mcx_lattice_upload_bits(mcx_lattice *lat, const uint8_t *host_bits)
{
    REQUIRE(lat && host_bits, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(!lat->upload_pending, MCX_ERR_STATE, "an upload is pending on this handle (the staging buffer is in use): commit it first");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    int32_t st = ensure_staging(lat);
    if (st == MCX_OK) st = ensure_hostbits(lat);
    if (st != MCX_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(lat->d_hostbits, host_bits, (size_t)lat->N / 8 * (size_t)lat->nchains, cudaMemcpyHostToDevice,
                             lat->ctx->stream));
    launch_hostbits_to_staging(lat, lat->d_hostbits, lat->ctx->stream);
    launch_pack(lat);
    if (lat->copy_stream) {
        CUDA_TRY(cudaEventRecord(lat->ev_packed, lat->ctx->stream));
        lat->packed_recorded = true;
    }
    launch_recompute(lat);
    lat->sums_dirty = false;
    return check_launch(lat->ctx);
}

int32_t mcx_lattice_upload_bits_begin(mcx_lattice *lat, const uint8_t *host_bits)
{
    REQUIRE(lat && host_bits, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(!lat->upload_pending, MCX_ERR_STATE, "an upload is already pending on this handle: commit it first");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    int32_t st = ensure_staging(lat);
    if (st == MCX_OK) st = ensure_hostbits(lat);
    if (st != MCX_OK) return st;
    if (!lat->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&lat->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&lat->ev_copied, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&lat->ev_packed, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(lat->ev_packed, lat->ctx->stream));
        lat->packed_recorded = true;
    }
    // d_hostbits is read by the conversion that mcx_lattice_upload_commit queues before it records ev_packed
    if (lat->packed_recorded) CUDA_TRY(cudaStreamWaitEvent(lat->copy_stream, lat->ev_packed, 0));
    CUDA_TRY(cudaMemcpyAsync(lat->d_hostbits, host_bits, (size_t)lat->N / 8 * (size_t)lat->nchains, cudaMemcpyHostToDevice,
                             lat->copy_stream));
    CUDA_TRY(cudaEventRecord(lat->ev_copied, lat->copy_stream));
    lat->upload_pending = true;
    lat->pending_bits = true;
    return MCX_OK;
}

int32_t mcx_lattice_download_bits(mcx_lattice *lat, uint8_t *host_bits)
{
    REQUIRE(lat && host_bits, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(!lat->upload_pending, MCX_ERR_STATE, "an upload is pending on this handle (the staging buffer is in use): commit it first");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    int32_t st = ensure_staging(lat);
    if (st == MCX_OK) st = ensure_hostbits(lat);
    if (st != MCX_OK) return st;
    launch_unpack(lat);
    launch_staging_to_hostbits(lat, lat->d_hostbits, lat->ctx->stream);
    CUDA_TRY(cudaMemcpyAsync(host_bits, lat->d_hostbits, (size_t)lat->N / 8 * (size_t)lat->nchains, cudaMemcpyDeviceToHost,
                             lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    return check_launch(lat->ctx);
}

int32_t mcx_lattice_init(mcx_lattice *lat, int32_t mode, uint64_t seed)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    REQUIRE(mode >= MCX_INIT_UP && mode <= MCX_INIT_RANDOM, MCX_ERR_ARGUMENT, "Unknown initialization type: %d", mode);
    REQUIRE(!(mode == MCX_INIT_ZERO && lat->model == MCX_ISING), MCX_ERR_ARGUMENT,
            "Unknown initialization type: zero is only defined for Blume-Capel");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    launch_init(lat, mode, seed);
    launch_recompute(lat);
    lat->sums_dirty = false;
    return check_launch(lat->ctx);
}

// ------------------------------------------------------------------------------------ rule
static int expected_table_len(const mcx_lattice *lat, int rule)
{
    const int nn = lat->nn;
    if (lat->model == MCX_ISING) return 2 * (nn + 1);
    return rule == MCX_HEATBATH ? 2 * (2 * nn + 1) : 6 * (2 * nn + 1);
}

int32_t mcx_set_rule(mcx_lattice *lat, int32_t rule, const uint64_t *thresholds, int32_t n_labels, int32_t table_len)
{
    REQUIRE(lat && thresholds, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(rule >= MCX_METROPOLIS && rule <= MCX_HEATBATH, MCX_ERR_ARGUMENT, "unknown rule %d", rule);
    REQUIRE(n_labels >= 1, MCX_ERR_ARGUMENT, "n_labels must be >= 1");
    REQUIRE(table_len == expected_table_len(lat, rule), MCX_ERR_ARGUMENT, "table_len %d != %d expected for this model/rule",
            table_len, expected_table_len(lat, rule));
    const size_t n = (size_t)n_labels * (size_t)table_len;
    std::vector<uint32_t> hi(n), lo(n);
    for (size_t i = 0; i < n; ++i) {
        REQUIRE(thresholds[i] <= ((uint64_t)1 << 32), MCX_ERR_ARGUMENT, "threshold %zu out of range [0, 2^32]", i);
        hi[i] = (uint32_t)(thresholds[i] >> 16);
        lo[i] = (uint32_t)(thresholds[i] & 0xffffu);
    }
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    if (n_labels != lat->n_labels || table_len != lat->table_len || !lat->d_thi) {
        cudaFree(lat->d_thi); cudaFree(lat->d_tlo);
        lat->d_thi = lat->d_tlo = nullptr;
        CUDA_TRY(cudaMalloc((void **)&lat->d_thi, n * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc((void **)&lat->d_tlo, n * sizeof(uint32_t)));
        free(lat->h_table);
        lat->h_table = (uint64_t *)malloc(n * sizeof(uint64_t));
    }
    memcpy(lat->h_table, thresholds, n * sizeof(uint64_t));
    CUDA_TRY(cudaMemcpy(lat->d_thi, hi.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(lat->d_tlo, lo.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    lat->rule = rule; lat->n_labels = n_labels; lat->table_len = table_len;
    // labels outside the new range would index past the tables
    std::vector<int32_t> lab(lat->nchains);
    CUDA_TRY(cudaMemcpy(lab.data(), lat->d_labels, sizeof(int32_t) * lat->nchains, cudaMemcpyDeviceToHost));
    bool fix = false;
    for (auto &l : lab) if (l < 0 || l >= n_labels) { l = 0; fix = true; }
    if (fix) CUDA_TRY(cudaMemcpy(lat->d_labels, lab.data(), sizeof(int32_t) * lat->nchains, cudaMemcpyHostToDevice));
    return MCX_OK;
}

int32_t mcx_set_labels(mcx_lattice *lat, const int32_t *label_of_chain)
{
    REQUIRE(lat && label_of_chain, MCX_ERR_ARGUMENT, "NULL argument");
    for (int c = 0; c < lat->nchains; ++c)
        REQUIRE(label_of_chain[c] >= 0 && label_of_chain[c] < (lat->n_labels > 0 ? lat->n_labels : 1), MCX_ERR_BOUNDS,
                "label %d of chain %d outside [0, %d)", label_of_chain[c], c, lat->n_labels);
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(lat->d_labels, label_of_chain, sizeof(int32_t) * lat->nchains, cudaMemcpyHostToDevice,
                             lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    return MCX_OK;
}

int32_t mcx_get_labels(mcx_lattice *lat, int32_t *label_of_chain)
{
    REQUIRE(lat && label_of_chain, MCX_ERR_ARGUMENT, "NULL argument");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(label_of_chain, lat->d_labels, sizeof(int32_t) * lat->nchains, cudaMemcpyDeviceToHost,
                             lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    return MCX_OK;
}

int32_t mcx_set_rng(mcx_lattice *lat, uint64_t seed, uint64_t next_sweep)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    REQUIRE(next_sweep < ((uint64_t)1 << 47), MCX_ERR_ARGUMENT, "sweep counter exceeds the 48-bit time field");
    lat->seed = seed; lat->sweep = next_sweep;
    return MCX_OK;
}

int32_t mcx_get_rng(mcx_lattice *lat, uint64_t *seed, uint64_t *next_sweep)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    if (seed) *seed = lat->seed;
    if (next_sweep) *next_sweep = lat->sweep;
    return MCX_OK;
}

int32_t mcx_set_tracking(mcx_lattice *lat, int32_t on)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    lat->track_sums = on != 0;
    return MCX_OK;
}

// ------------------------------------------------------------------------------------ sweep
int32_t mcx_sweep(mcx_lattice *lat, int64_t nsweeps)
{
    knobs_refresh();
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    REQUIRE(nsweeps >= 0, MCX_ERR_ARGUMENT, "nsweeps must be >= 0");
    REQUIRE(lat->rule >= 0, MCX_ERR_STATE, "no update rule set: call mcx_set_rule first");
    ASYNC_CHECK(lat->ctx);
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    if (lat->track_sums && lat->sums_dirty) { launch_recompute(lat); lat->sums_dirty = false; }
    if (lat->slab) {
        // a slab of a taller lattice: half-sweeps ordered against the neighbour GPUs by device flags
        REQUIRE(lat->slab->attached && lat->slab->remote, MCX_ERR_STATE,
                "slabs attached inside one process advance in lockstep through mcx_slab_half_sweep");
        REQUIRE(lat->slab->colour == 0, MCX_ERR_STATE, "slab is in the middle of a sweep");
        // a tall slab: row bands on auxiliary streams, the first / last band ordered against the neighbour GPUs
        if (nsweeps > 0 && launch_sweeps_ising2d_banded(lat, nsweeps)) {
            if (!lat->track_sums) lat->sums_dirty = true;
            lat->sweep += (uint64_t)nsweeps;
            lat->steps += nsweeps * lat->N;
            lat->slab->epoch += 2 * (uint64_t)nsweeps;
            return check_launch(lat->ctx);
        }
        for (int64_t s = 0; s < 2 * nsweeps; ++s) {
            const int32_t st = slab_half_sweep(lat);
            if (st != MCX_OK) return fail(st, "slab half-sweep could not be launched");
        }
        return check_launch(lat->ctx);
    }
    // test hooks: MCX_FORCE_GENERIC=1 -> shape-generic kernel, =2 -> rows-of-8 kernel
    const int force = knobs().force_generic > 0 && lat->storage == MCX_STORAGE_INT8 ? knobs().force_generic : 0;   // byte-plane kernels
    bool try_series = force == 0;        // whole-series launchers (resident kernel, chain groups) still worth asking
    for (int64_t s = 0; s < nsweeps;) {
        if (try_series) {
            // MCX_QUEUE=1: the whole series in one launch, work items of all half-sweeps from one ticket counter
            if (knobs().queue > 0 && launch_sweeps_ising2d_queue(lat, nsweeps - s)) {      // advances lat->sweep itself
                if (!lat->track_sums) lat->sums_dirty = true;
                s = nsweeps;
                continue;
            }
            // series of sweeps over small lattices: one launch with the lattice resident in shared memory
            const int64_t chunk = nsweeps - s < 16384 ? nsweeps - s : 16384;
            if (launch_sweeps_resident(lat, chunk)) {
                if (!lat->track_sums) lat->sums_dirty = true;
                lat->sweep += (uint64_t)chunk;
                s += chunk;
                continue;
            }
            // small batches of lattices too big for shared memory: the ticket-queue series (its own policy, k_queue.cu)
            if (knobs().queue < 0 && launch_sweeps_ising2d_queue(lat, nsweeps - s)) {
                if (!lat->track_sums) lat->sums_dirty = true;
                s = nsweeps;
                continue;
            }
            // one big lattice, a long series: the row-band launches of four sweeps replayed from a CUDA graph (device clock)
            if (lat->nchains == 1) {
                const int64_t gdone = launch_sweeps_ising2d_banded_graph(lat, nsweeps - s);
                if (gdone > 0) {
                    if (!lat->track_sums) lat->sums_dirty = true;
                    lat->sweep += (uint64_t)gdone;
                    s += gdone;
                    continue;
                }
            }
            // chain groups (batches) or row bands (one big lattice) on auxiliary streams overlap each other's launch tails
            if (lat->nchains > 1 ? launch_sweeps_ising2d_grouped(lat, nsweeps - s) : launch_sweeps_ising2d_banded(lat, nsweeps - s)) {
                if (!lat->track_sums) lat->sums_dirty = true;
                lat->sweep += (uint64_t)(nsweeps - s);
                s = nsweeps;
                continue;
            }
            try_series = false;
        }
        for (int colour = 0; colour < 2; ++colour) {
            const uint64_t t = 2 * lat->sweep + (uint64_t)colour;
            if (force == 1) launch_sweep_generic(lat, colour, t);
            else if (force == 2 && launch_sweep_rows8(lat, colour, t)) {}
            else if (launch_sweep_ising2d(lat, colour, t) || launch_sweep_bc2d(lat, colour, t) || launch_sweep_ising3d(lat, colour, t) ||
                     launch_sweep_bc3d(lat, colour, t)) {
                if (!lat->track_sums) lat->sums_dirty = true;
            }
            else if (!launch_sweep_rows8(lat, colour, t)) launch_sweep_generic(lat, colour, t);
        }
        lat->sweep += 1;
        s += 1;
    }
    lat->steps += nsweeps * lat->N;
    return check_launch(lat->ctx);
}

// measure!(measurements, sys, i) with an interval schedule (src/measurements/measurements.jl:192-200),
// kept on the device: nmeasure x (interval sweeps, then a snapshot of the per-chain sums), one
// device-to-host copy at the end.  out is [nmeasure][nchains][4] = {pair, spin, spin2, accepted}.
int32_t mcx_sweep_series(mcx_lattice *lat, int64_t nmeasure, int64_t interval, int64_t *out)
{
    REQUIRE(lat && out, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(nmeasure >= 0 && interval >= 1, MCX_ERR_ARGUMENT, "need nmeasure >= 0 and interval >= 1");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    const size_t stride = (size_t)lat->nchains * SUM_FIELDS;
    if (nmeasure == 0) return MCX_OK;
    ASYNC_CHECK(lat->ctx);
    // the snapshot buffer belongs to the handle and only ever grows: no allocation (an implicit device
    // synchronisation) inside a series once a handle has run one of this length
    const size_t series_bytes = sizeof(long long) * stride * (size_t)nmeasure;
    if (lat->series_bytes < series_bytes) {
        CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
        cudaFree(lat->d_series);
        lat->d_series = nullptr; lat->series_bytes = 0;
        CUDA_TRY(cudaMalloc((void **)&lat->d_series, series_bytes));
        lat->series_bytes = series_bytes;
    }
    long long *d_series = lat->d_series;
    const bool was_tracking = lat->track_sums;
    lat->track_sums = true;                       // the snapshots need current sums after every interval
    int32_t st = MCX_OK;
    if (lat->sums_dirty) { launch_recompute(lat); lat->sums_dirty = false; }
    for (int64_t k = 0; k < nmeasure && st == MCX_OK; ++k) {
        st = mcx_sweep(lat, interval);
        if (st == MCX_OK && cudaMemcpyAsync(d_series + k * stride, lat->d_sums, sizeof(long long) * stride,
                                            cudaMemcpyDeviceToDevice, lat->ctx->stream) != cudaSuccess)
            st = fail(MCX_ERR_CUDA, "snapshot copy failed");
    }
    lat->track_sums = was_tracking;
    lat->series_n = st == MCX_OK ? nmeasure : 0;
    if (st == MCX_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_series, sizeof(long long) * stride * (size_t)nmeasure, cudaMemcpyDeviceToHost,
                                        lat->ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(lat->ctx->stream);
        if (e != cudaSuccess) st = fail(MCX_ERR_CUDA, "series copy failed: %s", cudaGetErrorString(e));
    }
    if (st == MCX_OK) st = async_error(lat->ctx);
    if (lat->model == MCX_ISING && st == MCX_OK)
        for (int64_t i = 0; i < nmeasure * lat->nchains; ++i) out[i * SUM_FIELDS + SUM_SPIN2] = lat->N;
    return st;
}

static int32_t refresh_sums(mcx_lattice *lat)
{
    if (lat->sums_dirty) {
        launch_recompute(lat);
        lat->sums_dirty = false;
    }
    return MCX_OK;
}

int32_t mcx_observables(mcx_lattice *lat, int64_t *pair_sum, int64_t *spin_sum, int64_t *spin2_sum, int64_t *accepted,
                        int64_t *steps)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    refresh_sums(lat);
    std::vector<long long> h((size_t)lat->nchains * SUM_FIELDS);
    CUDA_TRY(cudaMemcpyAsync(h.data(), lat->d_sums, h.size() * sizeof(long long), cudaMemcpyDeviceToHost,
                             lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    CUDA_TRY(cudaGetLastError());
    ASYNC_CHECK(lat->ctx);                // a device-side wait of an earlier sweep gave up
    for (int c = 0; c < lat->nchains; ++c) {
        if (pair_sum) pair_sum[c] = h[(size_t)c * SUM_FIELDS + SUM_PAIR];
        if (spin_sum) spin_sum[c] = h[(size_t)c * SUM_FIELDS + SUM_SPIN];
        if (spin2_sum) spin2_sum[c] = lat->model == MCX_ISING ? lat->N : h[(size_t)c * SUM_FIELDS + SUM_SPIN2];
        if (accepted) accepted[c] = h[(size_t)c * SUM_FIELDS + SUM_ACC];
        if (steps) steps[c] = lat->steps;
    }
    return MCX_OK;
}

int32_t mcx_energies(mcx_lattice *lat, double *energy)
{
    REQUIRE(lat && energy, MCX_ERR_ARGUMENT, "NULL argument");
    std::vector<int64_t> pair(lat->nchains), spin(lat->nchains), spin2(lat->nchains);
    int32_t st = mcx_observables(lat, pair.data(), spin.data(), spin2.data(), nullptr, nullptr);
    if (st != MCX_OK) return st;
    for (int c = 0; c < lat->nchains; ++c) {
        double e = -(lat->J * (double)pair[c]);
        if (lat->h != 0.0) e -= lat->h * (double)spin[c];
        if (lat->model == MCX_BLUME_CAPEL) e += lat->D * (double)spin2[c];
        energy[c] = e;
    }
    return MCX_OK;
}

int32_t mcx_reset_counters(mcx_lattice *lat)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    std::vector<long long> h((size_t)lat->nchains * SUM_FIELDS);
    CUDA_TRY(cudaMemcpyAsync(h.data(), lat->d_sums, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    for (int c = 0; c < lat->nchains; ++c) h[(size_t)c * SUM_FIELDS + SUM_ACC] = 0;
    CUDA_TRY(cudaMemcpyAsync(lat->d_sums, h.data(), h.size() * sizeof(long long), cudaMemcpyHostToDevice, lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    lat->steps = 0;
    return MCX_OK;
}

// restore checkpointed counters: alg.accepted per chain and alg.steps (importance_sampling.jl:26-27 through
// checkpointing.jl:95-101)
int32_t mcx_set_counters(mcx_lattice *lat, const int64_t *accepted, int64_t steps)
{
    REQUIRE(lat && accepted, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(steps >= 0, MCX_ERR_ARGUMENT, "steps must be >= 0");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    std::vector<long long> h((size_t)lat->nchains * SUM_FIELDS);
    CUDA_TRY(cudaMemcpyAsync(h.data(), lat->d_sums, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    for (int c = 0; c < lat->nchains; ++c) h[(size_t)c * SUM_FIELDS + SUM_ACC] = accepted[c];
    CUDA_TRY(cudaMemcpyAsync(lat->d_sums, h.data(), h.size() * sizeof(long long), cudaMemcpyHostToDevice, lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    lat->steps = steps;
    return MCX_OK;
}

int32_t mcx_recompute(mcx_lattice *lat)
{
    REQUIRE(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    launch_recompute(lat);
    lat->sums_dirty = false;
    return check_launch(lat->ctx);
}

int32_t mcx_lattice_device_sums(mcx_lattice *lat, void **device_ptr)
{
    REQUIRE(lat && device_ptr, MCX_ERR_ARGUMENT, "NULL argument");
    *device_ptr = lat->d_sums;
    return MCX_OK;
}

// ------------------------------------------------------------------------------------ parallel tempering
int32_t mcx_pt_create(mcx_lattice *lat, int32_t n_global, int32_t first_slot, const double *betas, mcx_pt **out)
{
    REQUIRE(lat && betas && out, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(n_global >= 2, MCX_ERR_ARGUMENT, "need at least 2 replicas");
    REQUIRE(first_slot >= 0 && first_slot + lat->nchains <= n_global, MCX_ERR_ARGUMENT,
            "slots [%d, %d) do not fit in %d replicas", first_slot, first_slot + lat->nchains, n_global);
    REQUIRE(lat->n_labels == n_global, MCX_ERR_STATE,
            "the lattice must hold one rule table per ladder index (n_labels = %d, replicas = %d)", lat->n_labels, n_global);
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    mcx_pt *pt = new (std::nothrow) mcx_pt();
    REQUIRE(pt, MCX_ERR_STATE, "out of host memory");
    memset(pt, 0, sizeof(*pt));
    pt->lat = lat; pt->n = n_global; pt->first_slot = first_slot;
    const size_t n = (size_t)n_global;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&pt->d_betas, n * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc((void **)&pt->d_x, 2 * n * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc((void **)&pt->d_arrived, kMaxPtRanks * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc((void **)&pt->d_err, sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc((void **)&pt->d_index, n * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMalloc((void **)&pt->d_slot_of, n * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMalloc((void **)&pt->d_steps, n * sizeof(long long))) != cudaSuccess ||
        (e = cudaMalloc((void **)&pt->d_accepted, n * sizeof(long long))) != cudaSuccess) {
        mcx_pt_destroy(pt);
        return fail(MCX_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    CUDA_TRY(cudaMemcpy(pt->d_betas, betas, n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(pt->d_x, 0, 2 * n * sizeof(double)));
    CUDA_TRY(cudaMemset(pt->d_arrived, 0, kMaxPtRanks * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(pt->d_err, 0, sizeof(int)));
    lat->first_chain = (uint32_t)first_slot;
    lat->track_sums = true;
    *out = pt;
    return mcx_pt_reset(pt);
}

int32_t mcx_pt_destroy(mcx_pt *pt)
{
    if (!pt) return MCX_OK;
    cudaSetDevice(pt->lat->ctx->device);
    cudaStreamSynchronize(pt->lat->ctx->stream);
    for (void *p : pt->ipc_opened)
        if (p) cudaIpcCloseMemHandle(p);
    cudaFree(pt->d_betas); cudaFree(pt->d_x); cudaFree(pt->d_index); cudaFree(pt->d_slot_of);
    cudaFree(pt->d_steps); cudaFree(pt->d_accepted); cudaFree(pt->d_arrived); cudaFree(pt->d_err);
    cudaFree(pt->d_peer_x); cudaFree(pt->d_peer_arrived); cudaFree(pt->d_dev);
    if (pt->graph_exec) cudaGraphExecDestroy(pt->graph_exec);
    cudaFree(pt->d_clock);
    delete pt;
    return MCX_OK;
}

int32_t mcx_pt_reset(mcx_pt *pt)
{
    REQUIRE(pt, MCX_ERR_ARGUMENT, "pt is NULL");
    mcx_lattice *lat = pt->lat;
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    std::vector<int32_t> id(pt->n);
    for (int i = 0; i < pt->n; ++i) id[i] = i;
    CUDA_TRY(cudaMemcpy(pt->d_index, id.data(), sizeof(int32_t) * pt->n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pt->d_slot_of, id.data(), sizeof(int32_t) * pt->n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(pt->d_steps, 0, sizeof(long long) * pt->n));
    CUDA_TRY(cudaMemset(pt->d_accepted, 0, sizeof(long long) * pt->n));
    CUDA_TRY(cudaMemcpy(lat->d_labels, id.data() + pt->first_slot, sizeof(int32_t) * lat->nchains, cudaMemcpyHostToDevice));
    pt->stage = 0;
    pt->round = 0;
    return MCX_OK;
}

// restore a checkpointed ladder (checkpointing.jl:95-101 + replica_exchange.jl:13-19 fields)
int32_t mcx_pt_set_state(mcx_pt *pt, const int64_t *indices, const int64_t *steps, const int64_t *accepted, int64_t stage,
                         int64_t round)
{
    REQUIRE(pt && indices, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(stage == 0 || stage == 1, MCX_ERR_ARGUMENT, "stage must be 0 or 1");
    mcx_lattice *lat = pt->lat;
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    std::vector<int32_t> idx(pt->n), slot(pt->n, -1);
    for (int r = 0; r < pt->n; ++r) {
        REQUIRE(indices[r] >= 1 && indices[r] <= pt->n, MCX_ERR_BOUNDS, "ladder index %lld of slot %d outside 1..%d",
                (long long)indices[r], r, pt->n);
        idx[r] = (int32_t)indices[r] - 1;
        REQUIRE(slot[idx[r]] < 0, MCX_ERR_ARGUMENT, "Replica-exchange local index permutation is inconsistent");
        slot[idx[r]] = r;
    }
    CUDA_TRY(cudaMemcpy(pt->d_index, idx.data(), sizeof(int32_t) * pt->n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pt->d_slot_of, slot.data(), sizeof(int32_t) * pt->n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(lat->d_labels, idx.data() + pt->first_slot, sizeof(int32_t) * lat->nchains, cudaMemcpyHostToDevice));
    if (steps) CUDA_TRY(cudaMemcpy(pt->d_steps, steps, sizeof(long long) * (pt->n - 1), cudaMemcpyHostToDevice));
    if (accepted) CUDA_TRY(cudaMemcpy(pt->d_accepted, accepted, sizeof(long long) * (pt->n - 1), cudaMemcpyHostToDevice));
    // arrival counters of the peer-store all-gather count rounds: start over (all ranks restore, then barrier)
    CUDA_TRY(cudaMemset(pt->d_arrived, 0, kMaxPtRanks * sizeof(unsigned long long)));
    pt->stage = (int)stage;
    pt->round = (uint64_t)round;
    return MCX_OK;
}

int32_t mcx_pt_energy_buffer(mcx_pt *pt, void **device_ptr)
{
    REQUIRE(pt && device_ptr, MCX_ERR_ARGUMENT, "NULL argument");
    *device_ptr = pt->d_x;
    return MCX_OK;
}

// all-gather by peer stores: a 128-byte token (CUDA IPC handles of the energy buffer and the arrival counters) ...
int32_t mcx_pt_export(mcx_pt *pt, void *handle128)
{
    REQUIRE(pt && handle128, MCX_ERR_ARGUMENT, "NULL argument");
    CUDA_TRY(cudaSetDevice(pt->lat->ctx->device));
    cudaIpcMemHandle_t h[2];
    CUDA_TRY(cudaIpcGetMemHandle(&h[0], pt->d_x));
    CUDA_TRY(cudaIpcGetMemHandle(&h[1], pt->d_arrived));
    memcpy(handle128, h, 128);
    return MCX_OK;
}

// ... and the tokens of all ranks, in rank order (this rank's own entry is not opened)
int32_t mcx_pt_attach_peers(mcx_pt *pt, int32_t nranks, int32_t rank, const void *handles)
{
    REQUIRE(pt && handles, MCX_ERR_ARGUMENT, "NULL argument");
    REQUIRE(nranks >= 1 && nranks <= kMaxPtRanks && rank >= 0 && rank < nranks, MCX_ERR_ARGUMENT, "bad rank %d of %d", rank, nranks);
    REQUIRE(!pt->peers, MCX_ERR_STATE, "peers already attached");
    CUDA_TRY(cudaSetDevice(pt->lat->ctx->device));
    std::vector<double *> px(nranks);
    std::vector<unsigned long long *> pa(nranks);
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { px[r] = pt->d_x; pa[r] = pt->d_arrived; continue; }
        cudaIpcMemHandle_t h[2];
        memcpy(h, (const char *)handles + (size_t)r * 128, 128);
        CUDA_TRY(cudaIpcOpenMemHandle(&pt->ipc_opened[2 * r], h[0], cudaIpcMemLazyEnablePeerAccess));
        CUDA_TRY(cudaIpcOpenMemHandle(&pt->ipc_opened[2 * r + 1], h[1], cudaIpcMemLazyEnablePeerAccess));
        px[r] = (double *)pt->ipc_opened[2 * r];
        pa[r] = (unsigned long long *)pt->ipc_opened[2 * r + 1];
    }
    CUDA_TRY(cudaMalloc((void **)&pt->d_peer_x, sizeof(double *) * nranks));
    CUDA_TRY(cudaMalloc((void **)&pt->d_peer_arrived, sizeof(unsigned long long *) * nranks));
    CUDA_TRY(cudaMemcpy(pt->d_peer_x, px.data(), sizeof(double *) * nranks, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(pt->d_peer_arrived, pa.data(), sizeof(unsigned long long *) * nranks, cudaMemcpyHostToDevice));
    pt->nranks = nranks; pt->rank = rank; pt->peers = true;
    return MCX_OK;
}

int32_t mcx_pt_peer_status(mcx_pt *pt, int32_t *timed_out)
{
    REQUIRE(pt && timed_out, MCX_ERR_ARGUMENT, "NULL argument");
    CUDA_TRY(cudaSetDevice(pt->lat->ctx->device));
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, pt->d_err, sizeof(int), cudaMemcpyDeviceToHost, pt->lat->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(pt->lat->ctx->stream));
    *timed_out = err;
    return MCX_OK;
}

int32_t mcx_pt_publish(mcx_pt *pt)
{
    REQUIRE(pt, MCX_ERR_ARGUMENT, "pt is NULL");
    ASYNC_CHECK(pt->lat->ctx);
    CUDA_TRY(cudaSetDevice(pt->lat->ctx->device));
    refresh_sums(pt->lat);
    launch_pt_publish(pt);
    return check_launch(pt->lat->ctx);
}

int32_t mcx_pt_exchange(mcx_pt *pt)
{
    REQUIRE(pt, MCX_ERR_ARGUMENT, "pt is NULL");
    ASYNC_CHECK(pt->lat->ctx);
    CUDA_TRY(cudaSetDevice(pt->lat->ctx->device));
    launch_pt_exchange(pt);
    pt->stage = 1 - pt->stage;
    pt->round += 1;
    return check_launch(pt->lat->ctx);
}

}  // extern "C"

// One round of mcx_pt_run queued with the kernels reading the device clock: sweeps (chain groups or plain launches of
// k_ising2d), publish, exchange, clock advance.  false: some launch is not available for this lattice.
static bool pt_enqueue_clocked_round(mcx_pt *pt, int64_t S)
{
    mcx_lattice *lat = pt->lat;
    const uint64_t sweep0 = lat->sweep;
    lat->sweep = 0;                                             // launch arguments relative to the clock
    mcx::g_t_clock = &pt->d_clock->t_base;
    bool ok = mcx::launch_sweeps_ising2d_grouped(lat, S);
    if (!ok) {
        ok = true;
        for (int64_t s = 0; s < S && ok; ++s)
            for (int colour = 0; colour < 2 && ok; ++colour) ok = mcx::launch_sweep_ising2d(lat, colour, 2 * (uint64_t)s + (uint64_t)colour);
    }
    mcx::g_t_clock = nullptr;
    lat->sweep = sweep0;
    if (!ok) return false;
    if (mcx::knobs().pt_graph == 2) {                         // MCX_PT_GRAPH=2: the tail as three launches (A/B hook)
        mcx::launch_pt_publish(pt, pt->d_clock);
        mcx::launch_pt_exchange(pt, pt->d_clock);
        mcx::launch_pt_clock_advance(pt, S);
    } else {
        mcx::launch_pt_round_tail(pt, S);
    }
    return true;
}

// Rounds of mcx_pt_run replayed from a CUDA graph.  Returns the number of rounds done (0: not applicable, the caller
// queues them launch by launch), -1 on a CUDA error.
static int64_t pt_run_graph(mcx_pt *pt, int64_t nrounds, int64_t S)
{
    using namespace mcx;
    mcx_lattice *lat = pt->lat;
    mcx_ctx *ctx = lat->ctx;
    const Knobs &k = knobs();
    const int kGraphRounds = k.pt_graph > 2 ? (k.pt_graph > 256 ? 256 : k.pt_graph) : 8;   // rounds per replay (MCX_PT_GRAPH=n > 2 sets it)
    if (k.pt_graph == 0 || S >= 3 || nrounds < 2 * kGraphRounds || pt->graph_rounds < 0) return 0;   // < 0: the stream cannot be captured
    if (!lat->fast2d || lat->model != MCX_ISING || lat->storage != MCX_STORAGE_INT8 || lat->slab || lat->N < (1 << 20)) return 0;
    if (k.variant >= 0 || k.rows_per_strip >= 0 || k.force_generic > 0 || k.queue > 0 || k.resident > 0) return 0;
    if (!pt->d_clock && cudaMalloc((void **)&pt->d_clock, sizeof(PtClock)) != cudaSuccess) { cudaGetLastError(); pt->d_clock = nullptr; return 0; }
    if (lat->sums_dirty) { launch_recompute(lat); lat->sums_dirty = false; }
    mcx_pt::GraphKey key;                                       // what the captured launch arguments depend on
    memset(&key, 0, sizeof(key));
    key.seed = lat->seed; key.first_chain = lat->first_chain; key.rule = lat->rule; key.S = (int)S + 1024 * kGraphRounds + (k.groups == 0 ? 1 << 20 : 0); key.n = pt->n;
    key.nchains = lat->nchains;
    // the captured tail holds the parity of stage - round, which exchanges keep but mcx_pt_set_state / mcx_pt_reset may change
    key.peers = (pt->peers ? 1 : 0) | (int)((((uint64_t)pt->stage + pt->round) & 1) << 1);
    key.thi = lat->d_thi; key.labels = lat->d_labels; key.sums = lat->d_sums; key.planes = lat->view.planes; key.x = pt->d_x;
    if (pt->graph_exec && memcmp(&key, &pt->graph_key, sizeof(key)) != 0) {
        cudaGraphExecDestroy(pt->graph_exec);
        pt->graph_exec = nullptr;
    }
    if (!pt->graph_exec) {
        // one round launch by launch first: whatever the launchers create lazily (streams, events) exists before the capture
        const int32_t st0 = mcx_sweep(lat, S);
        if (st0 != MCX_OK || mcx_pt_publish(pt) != MCX_OK || mcx_pt_exchange(pt) != MCX_OK) return -1;
        nrounds -= 1;
        const uint64_t launches0 = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            pt->graph_rounds = -1;                             // do not try again on this handle
            return 1;
        }
        bool ok = true;
        for (int r = 0; r < kGraphRounds && ok; ++r) ok = pt_enqueue_clocked_round(pt, S);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        pt->graph_launches = ctx->launches - launches0;
        ctx->launches = launches0;
        if (!ok || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return 1;                                          // the warm-up round is done; the caller queues the rest
        }
        const cudaError_t ei = cudaGraphInstantiate(&pt->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { cudaGetLastError(); pt->graph_exec = nullptr; return 1; }
        pt->graph_key = key;
        pt->graph_rounds = kGraphRounds;
        // the warm-up round counts as done
        const int64_t rest = pt_run_graph(pt, nrounds, S);
        return rest < 0 ? -1 : 1 + rest;
    }
    const int64_t replays = nrounds / pt->graph_rounds;
    if (replays < 1) return 0;
    launch_pt_clock_set(pt);
    for (int64_t g = 0; g < replays; ++g) {
        if (cudaGraphLaunch(pt->graph_exec, ctx->stream) != cudaSuccess) return -1;
        ctx->launches += pt->graph_launches;
    }
    const int64_t done = replays * pt->graph_rounds;
    lat->sweep += (uint64_t)(done * S);
    lat->steps += done * S * lat->N;
    pt->round += (uint64_t)done;
    pt->stage = (int)((pt->stage + done) & 1);
    return done;
}

extern "C" {

// The user loop of pt_Ising2D.jl:52-57 -- `for i in 1:n; sweep; i % interval == 0 && update!(pt); end` -- queued
// in one call: nrounds x (sweeps_per_round sweeps, publish, exchange).  Needs the energies to reach all ranks
// without the host (one rank, or peer stores attached); otherwise MCX_ERR_UNSUPPORTED and the caller loops.
int32_t mcx_pt_run(mcx_pt *pt, int64_t nrounds, int64_t sweeps_per_round)
{
    knobs_refresh();
    REQUIRE(pt, MCX_ERR_ARGUMENT, "pt is NULL");
    REQUIRE(nrounds >= 0 && sweeps_per_round >= 1, MCX_ERR_ARGUMENT, "need nrounds >= 0 and sweeps_per_round >= 1");
    mcx_lattice *lat = pt->lat;
    REQUIRE(pt->peers || lat->nchains == pt->n, MCX_ERR_UNSUPPORTED,
            "replicas on other ranks: attach peers (mcx_pt_attach_peers) or drive publish / all-gather / exchange from the host");
    REQUIRE(lat->rule >= 0, MCX_ERR_STATE, "no update rule set: call mcx_set_rule first");
    ASYNC_CHECK(lat->ctx);
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    if (nrounds == 0) return MCX_OK;
    pt->last_path = 0;
    // Intervals of one or two sweeps on lattices too big for shared memory: queued launch by launch the rounds are bound by
    // the host (some 25 stream operations per round), so they are replayed from a CUDA graph of eight rounds whose kernels
    // read the half-sweep index and the round from a device clock.  On one B200 this is as fast as the persistent launch
    // below at 32 replicas of 1024 x 1024 (23749 against 23650 PT sweeps/s) and faster at 64 (14817 against 12732), and it
    // works across ranks (profiles/r02_call34_pt_graph.log); MCX_PT_PERSIST=1 still prefers the persistent launch.
    if (knobs().pt_persist <= 0) {
        const bool was_tracking = lat->track_sums;
        lat->track_sums = sweeps_per_round < 3;
        const int64_t done = pt_run_graph(pt, nrounds, sweeps_per_round);
        if (done < 0) return fail(MCX_ERR_CUDA, "replaying the rounds from a CUDA graph failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (done > 0) pt->last_path = 2; else lat->track_sums = was_tracking;
        nrounds -= done;
        if (nrounds == 0) return check_launch(lat->ctx);
    }
    // all rounds in ONE persistent launch (k_persist.cu): sweeps, energies to every rank, exchange, labels -- the
    // host queues nothing between rounds
    if (launch_pt_rounds_persistent(pt, nrounds, sweeps_per_round)) {
        if (pt->last_path == 0) pt->last_path = 1;
        return check_launch(lat->ctx);
    }
    // exchanging every sweep or two: keep the energy sums current per flip; longer intervals: recompute on publish
    lat->track_sums = sweeps_per_round < 3;
    for (int64_t r = 0; r < nrounds; ++r) {
        int32_t st = mcx_sweep(lat, sweeps_per_round);
        if (st == MCX_OK) st = mcx_pt_publish(pt);
        if (st == MCX_OK) st = mcx_pt_exchange(pt);
        if (st != MCX_OK) return st;
    }
    return MCX_OK;
}

/* how the last mcx_pt_run was executed: path 1 = one persistent launch for all rounds (strip_rows = its strip
 * height), 0 = rounds queued from the host */
int32_t mcx_pt_run_info(mcx_pt *pt, int32_t *path, int32_t *strip_rows)
{
    REQUIRE(pt, MCX_ERR_ARGUMENT, "pt is NULL");
    if (path) *path = pt->last_path;
    if (strip_rows) *strip_rows = pt->persist_R;
    return MCX_OK;
}

int32_t mcx_pt_state(mcx_pt *pt, int64_t *indices, int64_t *steps, int64_t *accepted, int64_t *stage, int64_t *round)
{
    REQUIRE(pt, MCX_ERR_ARGUMENT, "pt is NULL");
    mcx_lattice *lat = pt->lat;
    CUDA_TRY(cudaSetDevice(lat->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(lat->ctx->stream));
    ASYNC_CHECK(lat->ctx);                // exchanges decided on energies that never arrived are not a state
    if (indices) {
        std::vector<int32_t> id(pt->n);
        CUDA_TRY(cudaMemcpy(id.data(), pt->d_index, sizeof(int32_t) * pt->n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < pt->n; ++i) indices[i] = id[i] + 1;   // 1-based like rx.indices
    }
    if (steps) CUDA_TRY(cudaMemcpy(steps, pt->d_steps, sizeof(long long) * (pt->n - 1), cudaMemcpyDeviceToHost));
    if (accepted) CUDA_TRY(cudaMemcpy(accepted, pt->d_accepted, sizeof(long long) * (pt->n - 1), cudaMemcpyDeviceToHost));
    if (stage) *stage = pt->stage;
    if (round) *round = (int64_t)pt->round;
    return MCX_OK;
}

}  // extern "C"
