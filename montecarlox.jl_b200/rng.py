"""PhiloxRNG: the counter-based AbstractRNG injected into the reference's algorithms
(`alg.rng`, metropolis.jl:79,95; heat_bath.jl:17; parallel_tempering.jl:24-28 calls `rng(seed+i)`).

RNG layout v1 (include/mcx_b200.h).  The device kernels generate exactly these numbers; this host
class exists so that reference-style host code (`rand(alg.rng)`, replica_exchange.jl:168) and the
Julia shim have a bit-identical definition.  Precedent for RNG injection in the reference:
MutableRandomNumbers (src/infrastructure/rng.jl:55)."""

TAG_SWEEP, TAG_EXCHANGE, TAG_INIT, TAG_FLAT = 0, 1, 2, 3
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def stream_block(seed, chain, tag, t, blk, plane):
    ctr = (blk & _MASK, t & _MASK, ((t >> 32) & 0xFFFF) | (plane << 16) | (tag << 24), chain & _MASK)
    return philox4x32_10(ctr, (seed & _MASK, (seed >> 32) & _MASK))


class PhiloxRNG:
    """Positioned Philox stream.  `PhiloxRNG(seed)` mirrors `Xoshiro(seed)` as a constructor;
    `chain` is the replica / chain slot (counter word 3)."""

    def __init__(self, seed=0, chain=0):
        self.seed = int(seed)
        self.chain = int(chain)
        self.tag, self.t, self.q, self.draw = TAG_SWEEP, 0, 0, 0

    def position(self, tag, t, q=0):
        self.tag, self.t, self.q, self.draw = tag, int(t), int(q), 0
        return self

    def _next_draw(self):
        """planes 2n, 2n+1 of draw n; the plane field is 8 bits, so after 128 draws at one position the stream moves
        on to the same lane of the next block (q += 8) instead of aliasing the tag field"""
        if self.draw >= 128:
            self.q += 8
            self.draw = 0
        n = self.draw
        self.draw += 1
        return n

    def _lane16(self, plane):
        out = stream_block(self.seed, self.chain, self.tag, self.t, (self.q >> 3) & _MASK, plane & 0xFF)
        lane = self.q & 7
        return (out[lane >> 1] >> (16 * (lane & 1))) & 0xFFFF

    def rand(self):
        """rand(rng)::Float64 -- 32 bits, exact in Float64."""
        if self.tag == TAG_EXCHANGE:
            return exchange_u(self.seed, self.chain, self.t)
        n = self._next_draw()
        hi = self._lane16(2 * n)
        lo = self._lane16(2 * n + 1)
        return ((hi << 16) | lo) / 4294967296.0

    def rand_bool(self):
        """rand(rng, Bool)."""
        hi = self._lane16(2 * self._next_draw())
        return bool(hi >> 15)


def exchange_u(seed, chain, round_):
    """u of replica_exchange.jl:168 for slot `chain` at exchange round `round_` (53 bits)."""
    out = stream_block(seed, chain, TAG_EXCHANGE, round_, 0, 0)
    return float(((out[1] << 32) | out[0]) >> 11) / 9007199254740992.0
