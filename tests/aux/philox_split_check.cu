// Host-side check (built by tests/test_philox_split.py with nvcc, no GPU needed): the split form of a Philox4x32-10 block used by
// the strip loop -- philox_tail(philox_head(c1, c2, c3, k0, k1), c0) -- equals philox4x32_10(c0, c1, c2, c3, k0, k1), and the
// plain block reproduces the Random123 known-answer vectors.
#include "../../montecarlox.jl_b200/csrc/mcx_common.cuh"
#include <cstdio>
#include <cstdint>

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint32_t next32()
{
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 16);
}

int main()
{
    using namespace mcx;
    int bad = 0;
    for (int i = 0; i < 200000; ++i) {
        uint32_t c0 = next32(), c1 = next32(), c2 = next32(), c3 = next32(), k0 = next32(), k1 = next32();
        if (i < 64) { c0 = (uint32_t)i; c1 = c2 = c3 = 0; }                 // small counters as the kernels use them
        if (i >= 64 && i < 128) { c0 = 0xffffffffu - (uint32_t)i; k0 = k1 = 0xffffffffu; }
        const Philox4 a = philox4x32_10(c0, c1, c2, c3, k0, k1);
        const Philox4 b = philox_tail(philox_head(c1, c2, c3, k0, k1), c0);
        if (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) ++bad;
    }
    // Random123 kat_vectors: philox4x32 10 rounds
    const Philox4 z = philox4x32_10(0, 0, 0, 0, 0, 0);
    const bool kat0 = z.x == 0x6627e8d5u && z.y == 0xe169c58du && z.z == 0xbc57ac4cu && z.w == 0x9b00dbd8u;
    const Philox4 f = philox4x32_10(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    const bool kat1 = f.x == 0x408f276du && f.y == 0x41c83b0eu && f.z == 0xa20bc7c6u && f.w == 0x6d5451fdu;
    printf("mismatches=%d kat0=%d kat1=%d\n", bad, (int)kat0, (int)kat1);
    return (bad == 0 && kat0 && kat1) ? 0 : 1;
}
