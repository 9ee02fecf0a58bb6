// k_slab.cu -- one 2-D Ising lattice split by rows over several GPUs.
//
// A slab handle holds rows [row_offset, row_offset + Ly) of a lattice with global_Ly rows.  The checkerboard
// half-sweep needs, of the other colour plane, the row above the slab's first row and the row below its last
// row.  Those rows are not copied: the half-sweep kernel (k_ising2d<.., SLAB = true>) loads them straight from
// the neighbour slab's plane array -- another lattice of the same process (attach_local) or the device memory
// of the neighbour GPU's process mapped through CUDA IPC and read over NVLink (attach_ipc).  16 bytes per
// thread of a strip's first / last row are all that ever cross the link.
//
// Ordering between GPUs.  The boundary strips of half-sweep e of a slab may start once both neighbours have
// FINISHED the boundary strips of half-sweep e - 1: they have then written the rows it reads and are done
// reading the rows it overwrites.  Every slab owns a control block in its own memory (LatView::slab_ctl)
// with one 64-bit progress counter per neighbour.  Inside the half-sweep kernel the strip order is rotated so
// that the two boundary strips are the first items; only the CTAs that hold them poll the two local counters
// (the interior never waits), and the last of them to finish bumps "its" counter in both neighbours' blocks
// (system-scope fence, one 8-byte store over NVLink) while the interior is still being swept -- so by the
// time a neighbour's next half-sweep starts, its wait is already satisfied.  No extra launches, no host
// synchronisation.  A wait that lasts 20 s gives up and raises an error flag instead of hanging the GPU.
//
// Randomness is positioned by GLOBAL row (LatView::row_offset), so the trajectory of the split lattice is
// bit-identical to the same lattice on one GPU (tests/test_gpu_slab.py).
#include <cstring>
#include <new>

#include "mcx_internal.h"

namespace mcx {

// one colour of the current sweep of a slab, with the cross-GPU ordering when the neighbours are remote
int32_t slab_half_sweep(mcx_lattice *lat)
{
    mcx_slab *s = lat->slab;
    const uint64_t t = 2 * lat->sweep + (uint64_t)s->colour;
    // remote neighbours: the kernel's boundary CTAs wait for / signal the neighbour GPUs (LatView::slab_ctl)
    if (!launch_sweep_ising2d(lat, s->colour, t)) return MCX_ERR_UNSUPPORTED;
    if (!lat->track_sums) lat->sums_dirty = true;
    s->epoch += 1;
    if (s->colour == 1) { lat->sweep += 1; lat->steps += lat->N; }
    s->colour ^= 1;
    return MCX_OK;
}

void slab_free(mcx_lattice *lat)
{
    mcx_slab *s = lat->slab;
    if (!s) return;
    for (void *p : s->ipc_opened)
        if (p) cudaIpcCloseMemHandle(p);
    cudaFree(s->d_ctl);
    delete s;
    lat->slab = nullptr;
}

}  // namespace mcx

using namespace mcx;

int32_t mcx_set_error(int32_t code, const char *msg);   // mcx_api.cu

#define SREQ(cond, code, msg) do { if (!(cond)) return mcx_set_error(code, msg); } while (0)
#define SCUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return mcx_set_error(MCX_ERR_CUDA, cudaGetErrorString(e__)); } while (0)

extern "C" {

int32_t mcx_slab_configure(mcx_lattice *lat, int32_t global_Ly, int32_t row_offset)
{
    SREQ(lat, MCX_ERR_ARGUMENT, "lat is NULL");
    SREQ(lat->fast2d && lat->model == MCX_ISING && lat->storage == MCX_STORAGE_INT8, MCX_ERR_UNSUPPORTED,
         "slabs need a 2-D Ising int8 lattice with Lx % 32 == 0");
    SREQ(!lat->slab, MCX_ERR_STATE, "lattice is already a slab");
    SREQ(lat->nchains == 1, MCX_ERR_UNSUPPORTED, "a slab handle holds one chain");
    SREQ(row_offset >= 0 && row_offset % 2 == 0 && lat->view.Ly % 2 == 0 && row_offset + lat->view.Ly <= global_Ly,
         MCX_ERR_ARGUMENT, "a slab is an even number of rows at an even offset inside the global lattice");
    SCUDA(cudaSetDevice(lat->ctx->device));
    mcx_slab *s = new (std::nothrow) mcx_slab();
    SREQ(s, MCX_ERR_STATE, "out of host memory");
    memset(s, 0, sizeof(*s));
    s->global_Ly = global_Ly;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&s->d_ctl, SLAB_CTL_WORDS * sizeof(unsigned long long))) != cudaSuccess) {
        delete s;
        return mcx_set_error(MCX_ERR_CUDA, cudaGetErrorString(e));
    }
    SCUDA(cudaMemset(s->d_ctl, 0, SLAB_CTL_WORDS * sizeof(unsigned long long)));
    lat->slab = s;
    lat->view.row_offset = row_offset;
    return MCX_OK;
}

int32_t mcx_slab_export(mcx_lattice *lat, void *handle128)
{
    SREQ(lat && lat->slab && handle128, MCX_ERR_ARGUMENT, "needs a configured slab and a 128-byte buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle layout");
    SCUDA(cudaSetDevice(lat->ctx->device));
    cudaIpcMemHandle_t h[2];
    SCUDA(cudaIpcGetMemHandle(&h[0], lat->view.planes));
    SCUDA(cudaIpcGetMemHandle(&h[1], lat->slab->d_ctl));
    memcpy(handle128, h, 128);
    return MCX_OK;
}

int32_t mcx_slab_attach_ipc(mcx_lattice *lat, const void *up_handle128, const void *dn_handle128)
{
    SREQ(lat && lat->slab && up_handle128 && dn_handle128, MCX_ERR_ARGUMENT, "needs a configured slab and two handles");
    SREQ(!lat->slab->attached, MCX_ERR_STATE, "slab is already attached");
    SCUDA(cudaSetDevice(lat->ctx->device));
    mcx_slab *s = lat->slab;
    const bool same = memcmp(up_handle128, dn_handle128, 128) == 0;      // two slabs: one neighbour on both sides
    cudaIpcMemHandle_t h[2];
    memcpy(h, up_handle128, 128);
    SCUDA(cudaIpcOpenMemHandle(&s->ipc_opened[0], h[0], cudaIpcMemLazyEnablePeerAccess));
    SCUDA(cudaIpcOpenMemHandle(&s->ipc_opened[1], h[1], cudaIpcMemLazyEnablePeerAccess));
    if (same) {
        s->ipc_opened[2] = s->ipc_opened[3] = nullptr;
    } else {
        memcpy(h, dn_handle128, 128);
        SCUDA(cudaIpcOpenMemHandle(&s->ipc_opened[2], h[0], cudaIpcMemLazyEnablePeerAccess));
        SCUDA(cudaIpcOpenMemHandle(&s->ipc_opened[3], h[1], cudaIpcMemLazyEnablePeerAccess));
    }
    lat->view.up_planes = (uint8_t *)s->ipc_opened[0];
    lat->view.dn_planes = (uint8_t *)(same ? s->ipc_opened[0] : s->ipc_opened[2]);
    // I am the up neighbour's DOWN neighbour and the down neighbour's UP neighbour
    unsigned long long *up_ctl = (unsigned long long *)s->ipc_opened[1];
    unsigned long long *dn_ctl = (unsigned long long *)(same ? s->ipc_opened[1] : s->ipc_opened[3]);
    unsigned long long ctl[SLAB_CTL_WORDS] = {0, 0, 2 * lat->sweep, (unsigned long long)(uintptr_t)(up_ctl + SLAB_FLAG_DN),
                                              (unsigned long long)(uintptr_t)(dn_ctl + SLAB_FLAG_UP), 0, 0, 0};
    SCUDA(cudaMemcpy(s->d_ctl, ctl, sizeof(ctl), cudaMemcpyHostToDevice));
    lat->view.slab_ctl = s->d_ctl;
    s->remote = true;
    s->attached = true;
    return MCX_OK;
}

int32_t mcx_slab_attach_local(mcx_lattice *lat, mcx_lattice *up, mcx_lattice *dn)
{
    SREQ(lat && lat->slab && up && dn, MCX_ERR_ARGUMENT, "needs a configured slab and two neighbours");
    SREQ(up->ctx == lat->ctx && dn->ctx == lat->ctx, MCX_ERR_ARGUMENT, "local neighbours share the slab's context (one stream)");
    SREQ(up->view.half == lat->view.half && dn->view.half == lat->view.half && up->view.Ly == lat->view.Ly &&
         dn->view.Ly == lat->view.Ly && up->nchains == lat->nchains && dn->nchains == lat->nchains,
         MCX_ERR_ARGUMENT, "slabs of one lattice have the same shape");
    lat->view.up_planes = up->view.planes;
    lat->view.dn_planes = dn->view.planes;
    lat->slab->remote = false;
    lat->slab->attached = true;
    return MCX_OK;
}

int32_t mcx_slab_half_sweep(mcx_lattice *lat)
{
    SREQ(lat && lat->slab && lat->slab->attached, MCX_ERR_STATE, "needs an attached slab");
    SREQ(lat->rule >= 0, MCX_ERR_STATE, "no update rule set: call mcx_set_rule first");
    SCUDA(cudaSetDevice(lat->ctx->device));
    knobs_refresh();
    const int32_t st = slab_half_sweep(lat);
    if (st != MCX_OK) return mcx_set_error(st, "slab half-sweep could not be launched");
    SCUDA(cudaGetLastError());
    return MCX_OK;
}

int32_t mcx_slab_status(mcx_lattice *lat, int32_t *timed_out, uint64_t *half_sweeps_done)
{
    SREQ(lat && lat->slab, MCX_ERR_ARGUMENT, "needs a configured slab");
    SCUDA(cudaSetDevice(lat->ctx->device));
    unsigned long long err = 0;
    SCUDA(cudaMemcpyAsync(&err, lat->slab->d_ctl + SLAB_ERR, sizeof(err), cudaMemcpyDeviceToHost, lat->ctx->stream));
    SCUDA(cudaStreamSynchronize(lat->ctx->stream));
    if (timed_out) *timed_out = (int32_t)err;
    if (half_sweeps_done) *half_sweeps_done = lat->slab->epoch;
    return MCX_OK;
}

}  // extern "C"
