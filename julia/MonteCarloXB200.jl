# MonteCarloXB200.jl -- Julia shim over libmcx_b200.so (include/mcx_b200.h).
#
# NOT RUNNABLE IN THIS IMAGE (no julia binary, SURVEY.md section 0 finding 2).  It is the
# reference-side binding a maintainer would add: it shows, entry point by entry point, what the
# C ABI replaces and how the reference's own types keep working.  The Python package
# montecarlox.jl_b200/ mirrors exactly this file and is what the tests drive.
#
# Three pieces:
#   1. PhiloxRNG <: AbstractRNG        -- the counter-based RNG injected into `alg.rng`
#   2. sweep!(sys, alg, n) on CPU      -- the executable definition of "the Philox-driven
#                                         reference": the UNMODIFIED reference spin_flip! visiting
#                                         sites in checkerboard order
#   3. DeviceIsing / DeviceBlumeCapel  -- the same verbs forwarded to the GPU through ccall
module MonteCarloXB200

using Random
using MonteCarloX
using SpinSystems
import MonteCarloX: update!, reset!, acceptance_rate
import SpinSystems: energy, magnetization, init!, spin_flip!

const libmcx = get(ENV, "MCX_B200_LIB", "libmcx_b200.so")

# ----------------------------------------------------------------------------------------------
# status codes -> the exceptions the reference throws (SURVEY.md 8b)
# ----------------------------------------------------------------------------------------------
function check(status::Int32)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:mcx_last_error, libmcx), Cstring, ()))
    status == 1 && throw(ArgumentError(msg))
    status == 2 && throw(BoundsError(msg))
    status == 4 && throw(AssertionError(msg))
    error("libmcx_b200 status $status: $msg")
end

# ----------------------------------------------------------------------------------------------
# 1. PhiloxRNG (RNG layout v1, include/mcx_b200.h).  Precedent for RNG injection in the
#    reference: MutableRandomNumbers <: AbstractRNG (src/infrastructure/rng.jl:55).
# ----------------------------------------------------------------------------------------------
const TAG_SWEEP, TAG_EXCHANGE, TAG_INIT, TAG_FLAT = UInt32(0), UInt32(1), UInt32(2), UInt32(3)

mutable struct PhiloxRNG <: AbstractRNG
    seed::UInt64
    chain::UInt32
    tag::UInt32
    t::UInt64
    q::UInt64      # slot: row*(Lx/2) + (x>>1) for SWEEP, linear site for FLAT
    site::UInt64   # 0-based linear site served by the next rand(rng, UInt)
    draw::UInt32
end
PhiloxRNG(seed::Integer; chain::Integer=0) = PhiloxRNG(UInt64(seed), UInt32(chain), TAG_SWEEP, 0, 0, 0, 0)

function philox4x32_10(c0::UInt32, c1::UInt32, c2::UInt32, c3::UInt32, k0::UInt32, k1::UInt32)
    for _ in 1:10
        p0 = UInt64(0xD2511F53) * c0
        p1 = UInt64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = (UInt32(p1 >> 32) ⊻ c1 ⊻ k0), UInt32(p1 & 0xffffffff),
                         (UInt32(p0 >> 32) ⊻ c3 ⊻ k1), UInt32(p0 & 0xffffffff)
        k0 += 0x9E3779B9
        k1 += 0xBB67AE85
    end
    return (c0, c1, c2, c3)
end

function lane16(r::PhiloxRNG, plane::UInt32)
    c2 = UInt32((r.t >> 32) & 0xffff) | (plane << 16) | (r.tag << 24)
    out = philox4x32_10(UInt32((r.q >> 3) & 0xffffffff), UInt32(r.t & 0xffffffff), c2, r.chain,
                        UInt32(r.seed & 0xffffffff), UInt32(r.seed >> 32))
    lane = r.q & 7
    return (out[(lane >> 1) + 1] >> (16 * (lane & 1))) & 0xffff
end

"position the stream before one attempt: (tag, time, slot) and the site pick_site will return"
function position!(r::PhiloxRNG, tag::UInt32, t::Integer, q::Integer, site::Integer)
    r.tag, r.t, r.q, r.site, r.draw = tag, UInt64(t), UInt64(q), UInt64(site), 0
    return r
end

# Every rand flavour the hot path uses is overridden, so nothing depends on Julia-version defaults
# (SURVEY.md 7.3).  rand(rng)::Float64 = m * 2^-32, exact:
function Random.rand(r::PhiloxRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}})
    if r.tag == TAG_EXCHANGE          # replica_exchange.jl:168: 53 bits of block 0
        c2 = UInt32((r.t >> 32) & 0xffff) | (TAG_EXCHANGE << 24)
        o = philox4x32_10(UInt32(0), UInt32(r.t & 0xffffffff), c2, r.chain,
                          UInt32(r.seed & 0xffffffff), UInt32(r.seed >> 32))
        return Float64(((UInt64(o[2]) << 32) | o[1]) >> 11) * 2.0^-53
    end
    hi = lane16(r, 2 * r.draw); lo = lane16(r, 2 * r.draw + UInt32(1)); r.draw += 1
    return Float64((UInt32(hi) << 16) | UInt32(lo)) * 2.0^-32
end
# rand(rng, Bool) (blume_capel.jl:22): top bit of the high half of the next draw slot
function Random.rand(r::PhiloxRNG, ::Random.SamplerType{Bool})
    hi = lane16(r, 2 * r.draw); r.draw += 1
    return (hi >> 15) == 1
end
# rand(rng, UInt) is only drawn by pick_site (abstractions.jl:19: rand(rng, UInt) % N + 1): it
# returns the positioned site, so the UNMODIFIED reference spin_flip! visits exactly that site.
Random.rand(r::PhiloxRNG, ::Random.SamplerType{UInt64}) = r.site

# ----------------------------------------------------------------------------------------------
# 2. The Philox-driven reference: checkerboard order, unmodified spin_flip! per site.
#    Works on every CPU system of SpinSystems built on a periodic grid (ising.jl:413, blume_capel.jl:500);
#    this is what oracle/mcx_oracle.c (mcxo_sweep_checkerboard) restates in C.
# ----------------------------------------------------------------------------------------------
function sweep!(sys::SpinSystems.AbstractSpinSystem, alg, dims::Vector{Int}, sweep0::Integer, nsweeps::Integer)
    rng = alg.rng::PhiloxRNG
    # canonical rules (local acceptance) use the SWEEP stream; flat-histogram ensembles
    # (ImportanceSampling with a Multicanonical / WangLandau ensemble: ising.jl:25-33) the FLAT stream
    tag = (alg isa MonteCarloX.AbstractMetropolis || alg isa MonteCarloX.HeatBath) ? TAG_SWEEP : TAG_FLAT
    Lx = dims[1]; Ly = length(dims) > 1 ? dims[2] : 1
    half = Lx ÷ 2
    for s in 0:nsweeps-1, colour in 0:1
        t = 2 * (sweep0 + s) + colour
        for i in 0:length(sys.spins)-1
            x = i % Lx; row = i ÷ Lx; y = row % Ly; z = row ÷ Ly
            ((x + y + z) & 1) == colour || continue
            position!(rng, tag, t, row * half + (x >> 1), i)
            spin_flip!(sys, alg)           # reference code, untouched (ising.jl:35-58, blume_capel.jl:52-85)
        end
    end
    return nothing
end

# ----------------------------------------------------------------------------------------------
# 3. Device-backed systems.  Host-built integer tables: the reference's own float expression is
#    evaluated IN JULIA (its own exp) for every local configuration; the device only compares
#    integers (generalises TableMetropolis, docs/src/examples/spin_systems/importance_Ising2D.jl:74-92).
# ----------------------------------------------------------------------------------------------
const TWO32 = UInt64(1) << 32
thr(p::Float64) = p > 0 ? min(UInt64(ceil(p * 4294967296.0)), TWO32) : UInt64(0)
thr_accept(::MonteCarloX.Glauber, lr) = thr(MonteCarloX.logistic(lr))                 # metropolis.jl:124
thr_accept(::MonteCarloX.AbstractMetropolis, lr) = lr > 0 ? TWO32 : thr(exp(lr))      # importance_sampling.jl:82

mutable struct DeviceCtx
    h::Ptr{Cvoid}
end
function DeviceCtx(device::Integer=0; stream::Ptr{Cvoid}=C_NULL)
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mcx_ctx_create, libmcx), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, stream, out))
    return DeviceCtx(out[])
end

abstract type AbstractDeviceSystem end
mutable struct DeviceIsing <: SpinSystems.AbstractIsing     # dispatches like any reference Ising
    h::Ptr{Cvoid}
    dims::Vector{Int}
    nchains::Int
    J::Float64
    hfield::Float64
    rule_key::Any
end

# storage = :int8 (one byte per spin, like `spins::Vector{Int8}` of ising.jl:433) or :bit (one bit per spin on the device;
# 2-D / 3-D with Lx % 32 == 0).  Trajectories and observables do not depend on it.
function DeviceIsing(ctx::DeviceCtx, dims::Vector{Int}; J=1, h=0, nchains::Integer=1, storage::Symbol=:int8)
    storage in (:int8, :bit) || throw(ArgumentError("storage must be :int8 or :bit"))
    out = Ref{Ptr{Cvoid}}()
    d = Int32.(dims)
    check(ccall((:mcx_lattice_create, libmcx), Int32,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Int32, Int32, Ref{Ptr{Cvoid}}),
                ctx.h, 0, length(d), d, nchains, storage === :bit ? 1 : 0, out))
    check(ccall((:mcx_lattice_set_couplings, libmcx), Int32, (Ptr{Cvoid}, Float64, Float64, Float64), out[], J, h, 0.0))
    sys = DeviceIsing(out[], dims, nchains, J, h, nothing)
    finalizer(s -> ccall((:mcx_lattice_destroy, libmcx), Int32, (Ptr{Cvoid},), s.h), sys)
    return sys
end

# sys.spins: examples read length(sys.spins) (muca_Ising2D.jl:81) and copy it
function Base.getproperty(sys::DeviceIsing, name::Symbol)
    if name === :spins
        buf = Vector{Int8}(undef, prod(getfield(sys, :dims)) * getfield(sys, :nchains))
        GC.@preserve buf check(ccall((:mcx_lattice_download, libmcx), Int32, (Ptr{Cvoid}, Ptr{Int8}), getfield(sys, :h), buf))
        return buf
    end
    return getfield(sys, name)
end

# Split assignment of sys.spins for callers that stream configurations through the device: the H2D copy runs on the
# handle's copy stream beside sweeps that are already queued; `host` must stay alive and unchanged until the commit.
upload_begin!(sys::DeviceIsing, host::Vector{Int8}) =
    check(ccall((:mcx_lattice_upload_begin, libmcx), Int32, (Ptr{Cvoid}, Ptr{Int8}), getfield(sys, :h), host))
upload_commit!(sys::DeviceIsing) =
    check(ccall((:mcx_lattice_upload_commit, libmcx), Int32, (Ptr{Cvoid},), getfield(sys, :h)))

# The same transfers with a BitVector-like host buffer: site i is bit (i-1) & 7 of byte ((i-1) >> 3) + 1, 1 = up
# (`reinterpret(UInt8, (spins .> 0).chunks)` is this format); an eighth of the bytes over PCIe.
upload_bits!(sys::DeviceIsing, host::Vector{UInt8}) =
    check(ccall((:mcx_lattice_upload_bits, libmcx), Int32, (Ptr{Cvoid}, Ptr{UInt8}), getfield(sys, :h), host))
upload_bits_begin!(sys::DeviceIsing, host::Vector{UInt8}) =
    check(ccall((:mcx_lattice_upload_bits_begin, libmcx), Int32, (Ptr{Cvoid}, Ptr{UInt8}), getfield(sys, :h), host))
function download_bits(sys::DeviceIsing)
    buf = Vector{UInt8}(undef, (prod(getfield(sys, :dims)) * getfield(sys, :nchains)) >> 3)
    GC.@preserve buf check(ccall((:mcx_lattice_download_bits, libmcx), Int32, (Ptr{Cvoid}, Ptr{UInt8}), getfield(sys, :h), buf))
    return buf
end

function sums(sys::DeviceIsing)
    n = sys.nchains
    pair, spin, spin2, acc, steps = (Vector{Int64}(undef, n) for _ in 1:5)
    check(ccall((:mcx_observables, libmcx), Int32,
                (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                sys.h, pair, spin, spin2, acc, steps))
    return pair, spin, spin2, acc, steps
end

function energy(sys::DeviceIsing; full=false)            # ising.jl:17
    full && check(ccall((:mcx_recompute, libmcx), Int32, (Ptr{Cvoid},), sys.h))
    pair, spin, = sums(sys)
    e = -(sys.J .* pair) .- sys.hfield .* spin
    return sys.nchains == 1 ? e[1] : e
end
function magnetization(sys::DeviceIsing; full=false)     # ising.jl:18
    full && check(ccall((:mcx_recompute, libmcx), Int32, (Ptr{Cvoid},), sys.h))
    m = sums(sys)[2]
    return sys.nchains == 1 ? m[1] : m
end

function init!(sys::DeviceIsing, type::Symbol; rng=nothing)     # ising.jl:74-78
    mode = type == :up ? 0 : type == :down ? 1 : type == :random ? 3 : error("Unknown initialization type: $type")
    seed = UInt64(0)
    if type == :random
        @assert rng !== nothing "Random initialization requires rng"
        seed = (rng::PhiloxRNG).seed
        check(ccall((:mcx_lattice_set_first_chain_id, libmcx), Int32, (Ptr{Cvoid}, UInt32), sys.h, rng.chain))
    end
    check(ccall((:mcx_lattice_init, libmcx), Int32, (Ptr{Cvoid}, Int32, UInt64), sys.h, mode, seed))
    return sys
end

"Ising rule table: idx = s*(nn+1) + nup (include/mcx_b200.h), every entry from the reference's expression"
function ising_table(alg, ndim::Int, J, h)
    nn = 2 * ndim
    β = alg isa MonteCarloX.HeatBath ? alg.β : MonteCarloX.ensemble(alg).beta
    T = Vector{UInt64}(undef, 2 * (nn + 1))
    for sb in 0:1, nup in 0:nn
        s = sb == 1 ? 1 : -1
        lpi = s * (2nup - nn)                          # local_pair_interactions, abstractions.jl:41-48
        Δpair = -2 * J * lpi; Δspin = -2 * s           # flip_changes, ising.jl:187-192
        ΔE = -Δpair - h * Δspin                        # delta_energy, ising.jl:198
        T[sb * (nn + 1) + nup + 1] = alg isa MonteCarloX.HeatBath ?
            thr(MonteCarloX.logistic(β * float(s) * ΔE)) :                     # ising.jl:49
            thr_accept(alg, MonteCarloX.logweight(MonteCarloX.ensemble(alg), ΔE))   # metropolis.jl:14-17
    end
    return T
end

rule_code(::MonteCarloX.Glauber) = Int32(1)
rule_code(::MonteCarloX.AbstractMetropolis) = Int32(0)
rule_code(::MonteCarloX.HeatBath) = Int32(2)

"sweep!(sys, alg, n): n*N attempts = `for _ in 1:n*N; spin_flip!(sys, alg); end` in checkerboard order"
function sweep!(sys::DeviceIsing, alg, nsweeps::Integer=1)
    rng = alg.rng::PhiloxRNG
    key = (typeof(alg), alg isa MonteCarloX.HeatBath ? alg.β : MonteCarloX.ensemble(alg).beta, rng.seed, rng.chain)
    if key != sys.rule_key
        T = ising_table(alg, length(sys.dims), sys.J, sys.hfield)
        GC.@preserve T check(ccall((:mcx_set_rule, libmcx), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt64}, Int32, Int32),
                                   sys.h, rule_code(alg), T, 1, length(T)))
        check(ccall((:mcx_lattice_set_first_chain_id, libmcx), Int32, (Ptr{Cvoid}, UInt32), sys.h, rng.chain))
        s = Ref{UInt64}(); n = Ref{UInt64}()
        check(ccall((:mcx_get_rng, libmcx), Int32, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}), sys.h, s, n))
        check(ccall((:mcx_set_rng, libmcx), Int32, (Ptr{Cvoid}, UInt64, UInt64), sys.h, rng.seed, n[]))
        sys.rule_key = key
    end
    acc0 = sum(sums(sys)[4])
    check(ccall((:mcx_sweep, libmcx), Int32, (Ptr{Cvoid}, Int64), sys.h, nsweeps))
    alg.steps += nsweeps * prod(sys.dims)                       # importance_sampling.jl:81
    hasproperty(alg, :accepted) && (alg.accepted += sum(sums(sys)[4]) - acc0)
    return nothing
end

# spin_flip!(sys::DeviceIsing, alg) itself is deliberately NOT defined: a single random-site attempt
# per ccall would defeat the device; callers replace their `for _ in 1:N; spin_flip!(...)` loop by sweep!.

# ----------------------------------------------------------------------------------------------
# GPU backend for ParallelChains / ReplicaExchange (parallel_backends.jl:25-70, replica_exchange.jl)
# ----------------------------------------------------------------------------------------------
struct GPUBackend
    rank::Int
    nranks::Int
    allgather!::Function     # (device_ptr::Ptr{Float64}, n) -> nothing ; NCCL.jl / MPI.jl (CUDA-aware) plumbing
end
MonteCarloX.rank(b::GPUBackend) = b.rank
Base.size(b::GPUBackend) = b.nranks
MonteCarloX.is_root(b::GPUBackend) = b.rank == 0

mutable struct DeviceReplicaExchange
    h::Ptr{Cvoid}
    sys::DeviceIsing
    backend::GPUBackend
    n::Int
end

"position of the lattice's next sweep in the SWEEP stream (kept when an algorithm is bound to a lattice that has already run)"
function next_sweep(sys::DeviceIsing)
    seed = Ref{UInt64}(); nxt = Ref{UInt64}()
    check(ccall((:mcx_get_rng, libmcx), Int32, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}), sys.h, seed, nxt))
    return nxt[]
end

"ParallelTempering(betas; seed, rng = s -> PhiloxRNG(seed; chain = s - seed - 1)) bound to a batched lattice"
function attach(backend::GPUBackend, sys::DeviceIsing, algs::Vector)
    n = length(algs)
    per = n ÷ backend.nranks
    first = backend.rank * per
    T = reduce(vcat, [ising_table(a, length(sys.dims), sys.J, sys.hfield) for a in algs])
    GC.@preserve T check(ccall((:mcx_set_rule, libmcx), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt64}, Int32, Int32),
                               sys.h, rule_code(algs[1]), T, n, length(T) ÷ n))
    betas = Float64[MonteCarloX.ensemble(a).beta for a in algs]
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mcx_set_rng, libmcx), Int32, (Ptr{Cvoid}, UInt64, UInt64), sys.h, algs[1].rng.seed, next_sweep(sys)))
    GC.@preserve betas check(ccall((:mcx_pt_create, libmcx), Int32,
                                   (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ref{Ptr{Cvoid}}), sys.h, n, first, betas, out))
    rx = DeviceReplicaExchange(out[], sys, backend, n)
    finalizer(r -> (r.h == C_NULL || ccall((:mcx_pt_destroy, libmcx), Int32, (Ptr{Cvoid},), r.h); r.h = C_NULL), rx)
    return rx
end

"update!(rx): replica_exchange.jl:158-178 on the device; only the energies cross NVLink"
function update!(rx::DeviceReplicaExchange)
    check(ccall((:mcx_pt_publish, libmcx), Int32, (Ptr{Cvoid},), rx.h))
    if rx.backend.nranks > 1
        p = Ref{Ptr{Cvoid}}()
        check(ccall((:mcx_pt_energy_buffer, libmcx), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), rx.h, p))
        rx.backend.allgather!(Ptr{Float64}(p[]), rx.n)       # in-place all-gather of n doubles
    end
    check(ccall((:mcx_pt_exchange, libmcx), Int32, (Ptr{Cvoid},), rx.h))
    return nothing
end

function index(rx::DeviceReplicaExchange)                    # replica_exchange.jl:48-50
    idx = Vector{Int64}(undef, rx.n)
    check(ccall((:mcx_pt_state, libmcx), Int32,
                (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                rx.h, idx, C_NULL, C_NULL, C_NULL, C_NULL))
    return idx
end

# ----------------------------------------------------------------------------------------------
# 5. One lattice over several GPUs (beyond the reference: IsingLatticeOptim, ising.jl:430-461, is one
#    Vector{Int8}).  Each rank holds a slab of rows; the half-sweep kernel reads the rows above / below
#    the slab straight from the neighbour GPU's memory (CUDA IPC, NVLink), device flags order the
#    half-sweeps, and `sweep!` is the ordinary one.  `allgather_bytes` is whatever the host runtime
#    offers (MPI.Allgather on a UInt8 buffer).
# ----------------------------------------------------------------------------------------------
struct SlabIsing
    part::DeviceIsing          # this rank's rows
    dims::Vector{Int}          # global [Lx, Ly]
    backend::GPUBackend
end

function SlabIsing(ctx::DeviceCtx, dims::Vector{Int}, backend::GPUBackend, allgather_bytes)
    Lx, Ly = dims
    n, r = backend.nranks, backend.rank
    Ly % (2n) == 0 || throw(ArgumentError("Ly = $Ly does not split into $n slabs of an even number of rows"))
    rows = Ly ÷ n
    part = DeviceIsing(ctx, [Lx, rows])
    check(ccall((:mcx_slab_configure, libmcx), Int32, (Ptr{Cvoid}, Int32, Int32), part.h, Ly, r * rows))
    token = Vector{UInt8}(undef, 128)
    GC.@preserve token check(ccall((:mcx_slab_export, libmcx), Int32, (Ptr{Cvoid}, Ptr{UInt8}), part.h, token))
    all = allgather_bytes(token)                       # 128 bytes per rank, rank order
    up, dn = all[mod(r - 1, n) + 1], all[mod(r + 1, n) + 1]
    GC.@preserve up dn check(ccall((:mcx_slab_attach_ipc, libmcx), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}), part.h, up, dn))
    return SlabIsing(part, dims, backend)
end

sweep!(sys::SlabIsing, alg, nsweeps::Integer=1) = sweep!(sys.part, alg, nsweeps)
# energy / magnetization of the whole lattice: sum of the slabs' shares (the host reduces, e.g. MPI.Allreduce)
local_pair_sum(sys::SlabIsing) = sums(sys.part)[1]
local_spin_sum(sys::SlabIsing) = sums(sys.part)[2]

# ----------------------------------------------------------------------------------------------
# 5b. Multicanonical chains on the device and the whole parallel-tempering loop in one call.
#     The reference's host objects stay the source of truth: `ens.logweight.values` / `ens.histogram.values`
#     are ordinary Float64 arrays that the examples read and modify (muca_Ising2D.jl:36-39,89-90); the device
#     mirror is refreshed from / into them around every sweep! call.
# ----------------------------------------------------------------------------------------------
mutable struct DeviceMulticanonical
    h::Ptr{Cvoid}              # mcx_flat, kind MUCA: ONE log-weight table and ONE histogram shared by all chains of `sys`
    sys::DeviceIsing
    alg                        # ImportanceSampling{<:MulticanonicalEnsemble} (algorithms/multicanonical.jl:9-25)
end

function DeviceMulticanonical(sys::DeviceIsing, alg)
    ens = MonteCarloX.ensemble(alg)::MonteCarloX.MulticanonicalEnsemble
    b = ens.logweight.bins[1]                                   # DiscreteBinning(start, step, num), binned_object.jl:13-17
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mcx_flat_create, libmcx), Int32,
                (Ptr{Cvoid}, Int32, Int32, Int64, Int64, Int64, Float64, Int32, Ref{Ptr{Cvoid}}),
                sys.h, 0, 0, b.start, b.step, b.num, 0.0, 0, out))           # policy 0: out of range = BoundsError
    rng = alg.rng::PhiloxRNG
    check(ccall((:mcx_lattice_set_first_chain_id, libmcx), Int32, (Ptr{Cvoid}, UInt32), sys.h, rng.chain))
    # the FLAT stream is keyed by the algorithm's seed, exactly like the canonical bind (runs with different
    # PhiloxRNG(seed) must differ); the lattice keeps its sweep position
    check(ccall((:mcx_set_rng, libmcx), Int32, (Ptr{Cvoid}, UInt64, UInt64), sys.h, rng.seed, next_sweep(sys)))
    mc = DeviceMulticanonical(out[], sys, alg)
    # the flat handle points into the lattice: destroy it first (mc.sys keeps the lattice alive until then)
    finalizer(x -> (x.h == C_NULL || ccall((:mcx_flat_destroy, libmcx), Int32, (Ptr{Cvoid},), x.h); x.h = C_NULL), mc)
    return mc
end

"n*N attempts per chain: spin_flip!(sys, alg::ImportanceSampling) (ising.jl:25-33) with record_visit! (multicanonical.jl:25-30)"
function sweep!(mc::DeviceMulticanonical, nsweeps::Integer=1)
    ens = MonteCarloX.ensemble(mc.alg)
    lw, hist = ens.logweight.values, ens.histogram.values
    acc0 = sum(sums(mc.sys)[4])
    GC.@preserve lw check(ccall((:mcx_flat_set_logweight, libmcx), Int32, (Ptr{Cvoid}, Ptr{Float64}), mc.h, lw))
    all(iszero, hist) && check(ccall((:mcx_flat_reset_histogram, libmcx), Int32, (Ptr{Cvoid},), mc.h))   # after reset!(alg)
    check(ccall((:mcx_flat_sweep, libmcx), Int32, (Ptr{Cvoid}, Int64), mc.h, nsweeps))
    GC.@preserve hist check(ccall((:mcx_flat_get_histogram, libmcx), Int32, (Ptr{Cvoid}, Ptr{Float64}), mc.h, hist))
    mc.alg.steps += nsweeps * prod(mc.sys.dims) * mc.sys.nchains               # importance_sampling.jl:81
    mc.alg.accepted += sum(sums(mc.sys)[4]) - acc0
    return nothing
end
# update!(ensemble(alg)) and reset!(alg) are the reference's own methods on the host arrays
# (ensembles/multicanonical.jl:32-44, algorithms/multicanonical.jl:27-33); nothing to override.

"merge_histograms! over ranks (parallel_multicanonical.jl:38-52): one all-reduce on the device histogram; every rank then
 applies the same update!, which replaces the Bcast of distribute_logweight! (:59-73)"
function merge_histograms!(mc::DeviceMulticanonical, allreduce_sum!::Function)
    p = Ref{Ptr{Cvoid}}(); n = Ref{Int64}()
    check(ccall((:mcx_flat_device_histogram, libmcx), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Int64}), mc.h, p, n))
    allreduce_sum!(Ptr{UInt64}(p[]), n[])                        # NCCL.jl / CUDA-aware MPI.jl on the device pointer
    hist = MonteCarloX.ensemble(mc.alg).histogram.values
    GC.@preserve hist check(ccall((:mcx_flat_get_histogram, libmcx), Int32, (Ptr{Cvoid}, Ptr{Float64}), mc.h, hist))
    return nothing
end

"the user loop `for i in 1:n; sweeps; i % interval == 0 && update!(pt); end` (pt_Ising2D.jl:52-57) queued in one call"
function run!(rx::DeviceReplicaExchange, nrounds::Integer, sweeps_per_round::Integer)
    # 2-D Ising replicas: all rounds in ONE persistent kernel launch (energies, exchange decisions and labels inside the
    # kernel; k_persist.cu); otherwise the rounds are queued by the library.  A device-side wait that gave up (a rank that
    # never arrived) makes this and every later call fail with MCX_ERR_CUDA until mcx_ctx_clear_error.
    check(ccall((:mcx_pt_run, libmcx), Int32, (Ptr{Cvoid}, Int64, Int64), rx.h, nrounds, sweeps_per_round))
    return nothing
end

"code of a device-side wait that timed out on this context (0: none); clear_error! synchronises and resets it"
function async_error(ctx_handle::Ptr{Cvoid})
    code = Ref{Int32}()
    check(ccall((:mcx_ctx_async_error, libmcx), Int32, (Ptr{Cvoid}, Ref{Int32}), ctx_handle, code))
    return code[]
end
clear_error!(ctx_handle::Ptr{Cvoid}) = check(ccall((:mcx_ctx_clear_error, libmcx), Int32, (Ptr{Cvoid},), ctx_handle))

"restore alg.accepted (per chain) and alg.steps of a checkpointed lattice (checkpointing.jl:95-101)"
function set_counters!(sys::DeviceIsing, accepted::Vector{Int64}, steps::Integer)
    GC.@preserve accepted check(ccall((:mcx_set_counters, libmcx), Int32, (Ptr{Cvoid}, Ptr{Int64}, Int64), sys.h, accepted, steps))
end

# ----------------------------------------------------------------------------------------------
# 6. Wang-Landau in energy windows (BASELINE.json configs[4]; beyond the reference, whose lookups outside the
#    binned range throw BoundsError, test/test_multicanonical.jl:39-43).  One DeviceIsing batch per window
#    (one chain per walker, own DeviceCtx = own stream), `mcx_flat_create(..., out_of_range_policy = 1)`:
#    a proposal leaving the window is a rejected attempt that visits the current bin.  Windows are dealt to
#    the ranks in contiguous blocks; nothing is exchanged while sampling.  The Python mirror
#    (montecarlox.jl_b200/windows.py) adds the drive of the walkers into their windows and is what the tests run.
# ----------------------------------------------------------------------------------------------
mutable struct DeviceWangLandau
    h::Ptr{Cvoid}              # mcx_flat: one log-weight table per chain, like one WangLandauEnsemble per algorithm
    sys::DeviceIsing
    bins::StepRange{Int,Int}
    logf::Float64
end

"WangLandau(rng, bins; logf) (algorithms/wang_landau.jl:10-18) for every chain of `sys`, restricted to `bins`; walker c
 draws from PhiloxRNG(rng.seed; chain = rng.chain + c)"
function DeviceWangLandau(sys::DeviceIsing, rng::PhiloxRNG, bins::StepRange{Int,Int}; logf::Float64=1.0, window::Bool=true)
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mcx_flat_create, libmcx), Int32,
                (Ptr{Cvoid}, Int32, Int32, Int64, Int64, Int64, Float64, Int32, Ref{Ptr{Cvoid}}),
                sys.h, 1, 0, first(bins), step(bins), length(bins), 0.0, window ? 1 : 0, out))
    check(ccall((:mcx_lattice_set_first_chain_id, libmcx), Int32, (Ptr{Cvoid}, UInt32), sys.h, rng.chain))
    check(ccall((:mcx_set_rng, libmcx), Int32, (Ptr{Cvoid}, UInt64, UInt64), sys.h, rng.seed, next_sweep(sys)))
    wl = DeviceWangLandau(out[], sys, bins, logf)
    finalizer(x -> (x.h == C_NULL || ccall((:mcx_flat_destroy, libmcx), Int32, (Ptr{Cvoid},), x.h); x.h = C_NULL), wl)
    return wl
end

"n*N attempts per walker: spin_flip!(sys, alg::ImportanceSampling) (ising.jl:25-33) + accept! (wang_landau.jl:29-37)"
function sweep!(wl::DeviceWangLandau, nsweeps::Integer=1)
    check(ccall((:mcx_flat_set_logf, libmcx), Int32, (Ptr{Cvoid}, Float64), wl.h, wl.logf))
    check(ccall((:mcx_flat_sweep, libmcx), Int32, (Ptr{Cvoid}, Int64), wl.h, nsweeps))
    return nothing
end

"log-weight tables [nbins, nchains] as host Float64 (examples read `ens.logweight.values`, muca_Ising2D.jl:89-90)"
function logweights(wl::DeviceWangLandau)
    lw = Matrix{Float64}(undef, length(wl.bins), wl.sys.nchains)
    GC.@preserve lw check(ccall((:mcx_flat_get_logweight, libmcx), Int32, (Ptr{Cvoid}, Ptr{Float64}), wl.h, lw))
    return lw
end

MonteCarloX.update!(wl::DeviceWangLandau; power::Real=0.5) = (wl.logf *= power; nothing)   # ensembles/wang_landau.jl:23

"equal-width windows over bin indices 1:nbins, neighbours sharing `overlap` of a window -> Vector of UnitRange"
function partition_windows(nbins::Int, nwindows::Int; overlap::Float64=0.5)
    nwindows == 1 && return [1:nbins]
    width = min(max(ceil(Int, nbins / (1 + (nwindows - 1) * (1 - overlap))), 2), nbins)
    stride = (nbins - width) / (nwindows - 1)
    firsts = [round(Int, k * stride) for k in 0:nwindows-1]
    firsts[end] = nbins - width
    return [f+1:f+width for f in firsts]
end

"join window pieces of log g (0.0 = never visited) at their overlaps: shift by the mean difference, cross-fade"
function join_logdos(pieces::Vector{Vector{Float64}}, windows::Vector{UnitRange{Int}}, nbins::Int)
    out = fill(NaN, nbins)
    prev_end = 0
    for (k, (p, w)) in enumerate(zip(pieces, windows))
        vis = p .!= 0.0
        if k == 1
            out[w[vis]] = p[vis]; prev_end = last(w); continue
        end
        both = isfinite.(out[w]) .& vis
        any(both) || throw(ArgumentError("windows $(k-1) and $k share no visited bin"))
        shift = sum(out[w][both] .- p[both]) / count(both)
        span = max(min(prev_end, last(w)) - first(w) + 1, 1)
        for (i, b) in enumerate(w)
            vis[i] || continue
            t = clamp((i - 0.5) / span, 0.0, 1.0)
            out[b] = both[i] ? (1 - t) * out[b] + t * (p[i] + shift) : p[i] + shift
        end
        prev_end = max(prev_end, last(w))
    end
    return out
end

# ----------------------------------------------------------------------------------------------
# 7. General topologies: IsingGraph / IsingMatrix with fields (SpinSystems/src/ising.jl:86-417) on the device.
#    Built from the reference's own objects: `DeviceIsingGraph(ctx, sys_cpu)` takes the neighbour lists
#    (`sys.nbrs`, ascending) or the sparse matrix (`J.colptr / rowval / nzval`; a symmetric CSC is its own CSR).
#    sweep! visits the colour classes of a greedy first-fit colouring in site order; couplings are Float64, so the
#    acceptance is evaluated on the device from (rule, beta) with the reference's own expression.
# ----------------------------------------------------------------------------------------------
mutable struct DeviceIsingGraph <: SpinSystems.AbstractIsing
    h::Ptr{Cvoid}
    n::Int
    nchains::Int
    rule_key::Any
end

function _graph_create(ctx::DeviceCtx, rowptr::Vector{Int64}, col::Vector{Int64}, val, J, hmode::Integer, hs, hv, nchains)
    out = Ref{Ptr{Cvoid}}()
    n = length(rowptr) - 1
    GC.@preserve rowptr col val hv check(ccall((:mcx_graph_create, libmcx), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Float64, Int32, Float64, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, n, rowptr, col, val === nothing ? C_NULL : pointer(val), Float64(J), hmode, Float64(hs),
        hv === nothing ? C_NULL : pointer(hv), nchains, out))
    sys = DeviceIsingGraph(out[], n, nchains, nothing)
    finalizer(s -> ccall((:mcx_graph_destroy, libmcx), Int32, (Ptr{Cvoid},), s.h), sys)
    return sys
end

_field(sys) = hasproperty(sys, :h) ? (sys.h isa AbstractVector ? (2, 0.0, Float64.(sys.h)) : (1, Float64(sys.h), nothing)) : (0, 0.0, nothing)

function DeviceIsingGraph(ctx::DeviceCtx, sys::SpinSystems.IsingGraph; nchains::Integer=1)
    rowptr = Int64[0]; col = Int64[]
    for nb in sys.nbrs                                   # ising.jl:120: collect(Graphs.neighbors(graph, i)), ascending
        append!(col, Int64.(nb) .- 1); push!(rowptr, length(col))
    end
    hmode, hs, hv = _field(sys)
    return _graph_create(ctx, rowptr, col, nothing, sys.J, hmode, hs, hv, nchains)
end

function DeviceIsingGraph(ctx::DeviceCtx, sys::SpinSystems.IsingMatrix; nchains::Integer=1)
    J = sys.J                                            # symmetric SparseMatrixCSC (ising.jl:250-262): column i = row i
    hmode, hs, hv = _field(sys)
    return _graph_create(ctx, Int64.(J.colptr) .- 1, Int64.(J.rowval) .- 1, Float64.(J.nzval), 0.0, hmode, hs, hv, nchains)
end

function sweep!(sys::DeviceIsingGraph, alg, nsweeps::Integer=1)
    rng = alg.rng::PhiloxRNG
    β = alg isa MonteCarloX.HeatBath ? alg.β : MonteCarloX.ensemble(alg).beta
    key = (typeof(alg), β, rng.seed, rng.chain)
    if key != sys.rule_key
        s = Ref{UInt64}(); n = Ref{UInt64}()
        check(ccall((:mcx_graph_get_rng, libmcx), Int32, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}), sys.h, s, n))
        check(ccall((:mcx_graph_set_rule, libmcx), Int32, (Ptr{Cvoid}, Int32, Float64), sys.h, rule_code(alg), β))
        check(ccall((:mcx_graph_set_rng, libmcx), Int32, (Ptr{Cvoid}, UInt64, UInt64, UInt32), sys.h, rng.seed, n[], rng.chain))
        sys.rule_key = key
    end
    check(ccall((:mcx_graph_sweep, libmcx), Int32, (Ptr{Cvoid}, Int64), sys.h, nsweeps))
    alg.steps += nsweeps * sys.n * sys.nchains
    return nothing
end

function SpinSystems.energy(sys::DeviceIsingGraph; full::Bool=false)
    e = Vector{Float64}(undef, sys.nchains)
    check(ccall((:mcx_graph_energies, libmcx), Int32, (Ptr{Cvoid}, Ptr{Float64}), sys.h, e))
    return sys.nchains == 1 ? e[1] : e
end

# ----------------------------------------------------------------------------------------------
# 8. integrated_autocorrelation_time (src/measurements/autocorrelations.jl:28-65) of a series measured on the
#    device: `sweep_series!` leaves the snapshots there, `tau_int` reduces them to one Float64 per chain in place.
# ----------------------------------------------------------------------------------------------
function sweep_series!(sys::DeviceIsing, nmeasure::Integer, interval::Integer=1)
    out = Array{Int64}(undef, 4, sys.nchains, nmeasure)           # {pair, spin, spin2, accepted} per chain and snapshot
    check(ccall((:mcx_sweep_series, libmcx), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}), sys.h, nmeasure, interval, out))
    return out
end

function tau_int(sys::DeviceIsing, nmeasure::Integer; observable::Symbol=:energy, max_lag::Integer=0, c::Real=5.0)
    code = observable === :energy ? 0 : observable === :magnetization ? 1 : 2
    tau = Vector{Float64}(undef, sys.nchains)
    check(ccall((:mcx_series_tau_int, libmcx), Int32, (Ptr{Cvoid}, Int64, Int32, Int64, Float64, Ptr{Float64}),
                sys.h, nmeasure, code, max_lag, c, tau))
    return tau
end

end # module
