// k_bits.cuh -- MCX_STORAGE_BIT: Ising colour planes at one bit per spin (SpinSystems/src/ising.jl:430-461 keeps
// `spins::Vector{Int8}`; this is the multispin storage the hot path asks for).
//
// Layout.  A colour plane of a lattice with `rows` = Ly * Lz rows of `half` = Lx / 2 sites (Lx % 32 == 0, Ly even) is
// a [rows / 2][half / 16] matrix of 32-bit words.  Word (m, g) holds the 32 sites (row 2m + rho, column 16 g + j),
// rho in {0, 1}, j in 0..15, with site (rho, j = 4 w + b) at bit
//
//         8 b + 4 rho + w                                       (b = j & 3, w = j >> 2)
//
// i.e. transposed inside the word so that the kernels' working form -- four 32-bit words of 0/1 bytes per 16-site
// thread-row, byte b of word w = site 4 w + b, exactly what the int8 planes hold -- is one shift and one mask away:
//
//         word_w(rho) = (bits >> (4 rho + w)) & 0x01010101
//
// and the way back is one shift-add per word.  The two rows of a word are the two rows a thread updates in one loop
// trip of the half-sweep kernels, so a trip loads ONE word of the target plane and ONE new word of the other plane
// (a warp: 128 contiguous bytes each) where the int8 kernel loads four 128-bit vectors.  Everything after the
// expansion is the int8 kernels' own code (update_row / update_row3), so trajectories are identical by construction.
#pragma once
#include "mcx_internal.h"

namespace mcx {

__host__ __device__ __forceinline__ int bit_pos(int rho, int j) { return 8 * (j & 3) + 4 * rho + (j >> 2); }

// the 16 sites of row `rho` of a word as four words of 0/1 bytes
__device__ __forceinline__ uint4 bits_expand(uint32_t bits, int rho)
{
    const uint32_t v = bits >> (4 * rho);
    return make_uint4(v & 0x01010101u, (v >> 1) & 0x01010101u, (v >> 2) & 0x01010101u, (v >> 3) & 0x01010101u);
}

// four words of 0/1 bytes -> the 16 bits of one row, placed for rho = 0 (shift the result by 4 for rho = 1)
__device__ __forceinline__ uint32_t bits_compress(uint4 q) { return q.x + 2u * q.y + 4u * q.z + 8u * q.w; }

// word index of (row, segment) in a colour plane and its base pointer
__device__ __forceinline__ const uint32_t *bits_word(const uint8_t *plane, int64_t row, int nseg, int seg)
{
    return reinterpret_cast<const uint32_t *>(plane) + (row >> 1) * nseg + seg;
}

// one site out of a plane: used for the out-of-segment in-row neighbour of edge lanes
__device__ __forceinline__ uint32_t bits_site(const uint8_t *plane, int64_t row, int nseg, int col)
{
    const uint32_t w = *bits_word(plane, row, nseg, col >> 4);
    return (w >> bit_pos((int)(row & 1), col & 15)) & 1u;
}

}  // namespace mcx
