// k_flat.cu -- flat-histogram (multicanonical / Wang-Landau) chains.  Placeholder entry points:
// the serial-chain kernels land in the next milestone; until then these fail loudly.
#include "mcx_internal.h"

static int32_t unsupported()
{
    return MCX_ERR_UNSUPPORTED;
}

extern "C" {
int32_t mcx_flat_create(mcx_lattice *, int32_t, int32_t, int64_t, int64_t, int64_t, double, mcx_flat **) { return unsupported(); }
int32_t mcx_flat_destroy(mcx_flat *) { return unsupported(); }
int32_t mcx_flat_set_logweight(mcx_flat *, const double *) { return unsupported(); }
int32_t mcx_flat_get_logweight(mcx_flat *, double *) { return unsupported(); }
int32_t mcx_flat_get_histogram(mcx_flat *, double *) { return unsupported(); }
int32_t mcx_flat_reset_histogram(mcx_flat *) { return unsupported(); }
int32_t mcx_flat_set_logf(mcx_flat *, double) { return unsupported(); }
int32_t mcx_flat_sweep(mcx_flat *, int64_t) { return unsupported(); }
int32_t mcx_flat_update(mcx_flat *) { return unsupported(); }
int32_t mcx_flat_device_histogram(mcx_flat *, void **, int64_t *) { return unsupported(); }
int32_t mcx_flat_device_logweight(mcx_flat *, void **, int64_t *) { return unsupported(); }
}
