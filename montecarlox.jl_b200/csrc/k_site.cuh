// k_site.cuh -- the table-driven per-site rule shared by the generic and the rows-of-8 kernels.
// Restates (through the host-built integer tables, include/mcx_b200.h):
//   Ising   spin_flip! ising.jl:35-58            Blume-Capel spin_flip! blume_capel.jl:52-85
//   accept! metropolis.jl:14-17,121-127          _accept! importance_sampling.jl:80-85
#pragma once
#include "mcx_common.cuh"

namespace mcx {

struct SiteAcc {
    int dpair = 0, dspin = 0, dspin2 = 0, nacc = 0;
};

// so: current encoding (Ising 0/1, Blume-Capel 0/1/2); raw: sum of the neighbours' encodings.
// r0 / r2: Philox blocks of planes 0 and 2 for this group of eight slots; k: lane in the group.
// fetch_lo(plane) returns the lane's low half from plane `plane` (called only on a 16-bit tie).
template <int MODEL, int RULE, class LoFn>
__device__ __forceinline__ int site_update(int so, int raw, int nn, const Philox4 &r0, const Philox4 &r2, int k,
                                           const uint32_t *thi, const uint32_t *tlo, LoFn fetch_lo, SiteAcc &acc)
{
    auto less_than = [&](uint32_t hi, int idx, uint32_t lo_plane) -> bool {
        const uint32_t th = thi[idx], tl = tlo[idx];
        if (hi < th) return true;
        if (hi > th || tl == 0) return false;
        return fetch_lo(lo_plane) < tl;
    };
    int sn = so;
    if (MODEL == MCX_ISING) {
        const bool lt = less_than(lane16(r0, k), so * (nn + 1) + raw, 1);
        if (RULE == MCX_HEATBATH) sn = lt ? 1 : 0;
        else sn = lt ? (so ^ 1) : so;
        if (sn != so) {
            const int sgn = 2 * so - 1, nsum = 2 * raw - nn;
            acc.dpair += -2 * sgn * nsum;
            acc.dspin += -2 * sgn;
            acc.nacc += 1;
        }
    } else {
        const int nsum = raw - nn;
        if (RULE == MCX_HEATBATH) {
            const uint32_t hi = lane16(r0, k);
            const bool lt0 = less_than(hi, raw, 1);
            const bool lt1 = lt0 ? true : less_than(hi, (2 * nn + 1) + raw, 1);
            sn = lt0 ? 0 : (lt1 ? 1 : 2);
        } else {
            const int b = (int)(lane16(r0, k) >> 15);
            const int prop = so == 0 ? (b ? 1 : 2) : so == 1 ? (b ? 0 : 2) : (b ? 0 : 1);
            const bool lt = less_than(lane16(r2, k), (so * 2 + b) * (2 * nn + 1) + raw, 3);
            sn = lt ? prop : so;
        }
        if (sn != so) {
            const int d = sn - so;
            acc.dpair += d * nsum;
            acc.dspin += d;
            acc.dspin2 += (sn - 1) * (sn - 1) - (so - 1) * (so - 1);
            acc.nacc += 1;
        }
    }
    return sn;
}

template <int MODEL>
__device__ __forceinline__ void flush_site_acc(SiteAcc acc, long long *sums_chain)
{
    const int dpair = warp_sum(acc.dpair), dspin = warp_sum(acc.dspin), nacc = warp_sum(acc.nacc);
    const int dspin2 = MODEL == MCX_BLUME_CAPEL ? warp_sum(acc.dspin2) : 0;
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *o = (unsigned long long *)sums_chain;
        if (dpair) atomicAdd(o + SUM_PAIR, (unsigned long long)(long long)dpair);
        if (dspin) atomicAdd(o + SUM_SPIN, (unsigned long long)(long long)dspin);
        if (MODEL == MCX_BLUME_CAPEL && dspin2) atomicAdd(o + SUM_SPIN2, (unsigned long long)(long long)dspin2);
        if (nacc) atomicAdd(o + SUM_ACC, (unsigned long long)(long long)nacc);
    }
}

}  // namespace mcx
