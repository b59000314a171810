# PowerSystemAdequacyB200.jl -- drop-in for GeneratingAdequacy/PowerSystemAdequacy.jl.
#
# Same module surface as the reference (PowerSystemAdequacy.jl:8-10): Generator, LoadModel,
# ReliabilityResult, run_analytical, run_non_sequential_mc, run_sequential_mc, compare_results.
# Every engine body is one `ccall` into libpsra_b200.so (include/psra_b200.h); nothing is computed
# on the CPU.  A driver such as run_full_comparison.jl switches by replacing
#
#     include("PowerSystemAdequacy.jl");      using .PowerSystemAdequacy
# with
#     include("PowerSystemAdequacyB200.jl");  using .PowerSystemAdequacyB200
#
# NOTE: Julia is not installed in the build image, so this file has been written against the header
# and cross-checked by hand (argument order / element types / struct layouts mirror the ctypes
# binding in powersystemsreliabilityassessment_b200/_lib.py, which IS exercised by the tests).
# Only whole arrays cross the boundary, so 1-based indexing never leaks.
module PowerSystemAdequacyB200

using Printf

export Generator, LoadModel, ReliabilityResult, unit_importance,
       run_analytical, run_non_sequential_mc, run_sequential_mc, compare_results,
       SequentialIndices, run_sequential_indices, tail_risk, PSRA_INIT_ALL_UP, PSRA_INIT_STATIONARY,
       DetailedGenerator, run_detailed_mc

const LIB = get(ENV, "PSRA_B200_LIB", joinpath(@__DIR__, "..", "powersystemsreliabilityassessment_b200", "libpsra_b200.so"))
const PSRA_INIT_ALL_UP = Int32(0)
const PSRA_INIT_STATIONARY = Int32(1)

# ---------------------------------------------------------------- data model (PSA.jl:20-53)
struct Generator
    id::Int
    capacity::Float64
    mttf::Float64
    mttr::Float64
    lambda::Float64
    mu::Float64
    for_rate::Float64
end
function Generator(id::Int, capacity::Float64, mttf::Float64, mttr::Float64)
    λ = 1.0 / mttf; μ = 1.0 / mttr
    return Generator(id, capacity, mttf, mttr, λ, μ, λ / (λ + μ))
end

struct LoadModel
    hourly_load::Vector{Float64}
    peak_load::Float64
    LoadModel(hourly_load::Vector{Float64}) = new(hourly_load, maximum(hourly_load))
end

struct ReliabilityResult
    method::String
    lole_hours_yr::Float64
    eue_mwh_yr::Float64
    computation_time::Float64
    convergence_history::Vector{Float64}
end

# ---------------------------------------------------------------- C structs (psra_b200.h)
struct PsraConfig
    device::Int32; warps_per_block::Int32; seg_hours::Int32; blocks_per_sm::Int32
    reserved::NTuple{4,Int32}
    ngpus::Int32; ev_cap::Int32; tail_bins::Int32
    reserved2::NTuple{5,Int32}
end
mutable struct PsraSeqSummary
    years::Int64; sum_lol_hours::Int64; sum_ens_fp::Int64; sum_entries::Int64; years_with_loss::Int64
    sum_lol_sq::UInt64; sum_ens_sq_lo::UInt64; sum_ens_sq_hi::UInt64; events::UInt64
    kernel_ms::Float32; redone::Int32
    PsraSeqSummary() = new(0, 0, 0, 0, 0, 0, 0, 0, 0, 0f0, 0)
end
struct PsraSeqOutputs
    lol_hours::Ptr{UInt32}; ens_fp::Ptr{Int64}; entries::Ptr{UInt32}; fail_count::Ptr{UInt32}
    group_lol::Ptr{Int64}; group::Int32; keep_on_device::Int32; history::Ptr{Float64}
    tail_hist::Int32; reserved::Int32
end
mutable struct PsraNonseqSummary
    samples::Int64; sum_lol_hours::Int64; sum_ens_fp::Int64; samples_with_loss::Int64
    sum_lol_sq::UInt64; sum_ens_sq_lo::UInt64; sum_ens_sq_hi::UInt64
    kernel_ms::Float32; reserved::Int32
    PsraNonseqSummary() = new(0, 0, 0, 0, 0, 0, 0, 0f0, 0)
end
struct PsraNonseqOutputs
    lol_hours::Ptr{UInt32}; ens_fp::Ptr{Int64}; cap_avail::Ptr{Int32}; states::Ptr{UInt32}
    group_lol::Ptr{Int64}; group::Int32; reserved::Int32; history::Ptr{Float64}
end
struct PsraTailOut
    var::Float64; cvar::Float64; n_tail::Int64; x_lo::Int64; x_hi::Int64
end

# ---------------------------------------------------------------- handle
mutable struct Engine
    h::Ptr{Cvoid}
    fp_scale::Float64
end

function check(e::Engine, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:psra_last_error, LIB), Cstring, (Ptr{Cvoid},), e.h))
    error("libpsra_b200 error $rc: $msg")
end

"""
    Engine(; device=0, ngpus=1)

One libpsra_b200 handle.  `ngpus = G > 1` spans the devices `device .. device+G-1` of this process: the Monte Carlo
calls shard their years / samples over them and combine the integer accumulators, per-hour failure counts and the ENS
histogram with `ncclAllReduce` inside the library -- `run_sequential_mc(gens, load, years; engine=Engine(ngpus=8))`.
"""
function Engine(; device::Integer=0, ngpus::Integer=1)
    z4 = (Int32(0), Int32(0), Int32(0), Int32(0)); z5 = (Int32(0), Int32(0), Int32(0), Int32(0), Int32(0))
    cfg = Ref(PsraConfig(Int32(device), 0, 0, 0, z4, Int32(ngpus), Int32(0), Int32(0), z5))
    hp = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:psra_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{PsraConfig}), hp, cfg)
    e = Engine(hp[], 1.0)
    if rc != 0
        msg = hp[] == C_NULL ? "psra_create failed" :
              unsafe_string(ccall((:psra_last_error, LIB), Cstring, (Ptr{Cvoid},), hp[]))
        hp[] != C_NULL && ccall((:psra_destroy, LIB), Cvoid, (Ptr{Cvoid},), hp[])
        error("libpsra_b200 error $rc: $msg (there is no CPU fallback)")
    end
    finalizer(x -> (x.h != C_NULL && ccall((:psra_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), e)
    return e
end

const _engine = Ref{Union{Nothing,Engine}}(nothing)
default_engine() = (_engine[] === nothing && (_engine[] = Engine()); _engine[]::Engine)

# Float64 MW -> int32 fixed point.  Capacities must lie on the grid (error otherwise: choose a finer fp_scale).
function fixed(x::AbstractVector{Float64}, scale::Float64)
    v = x .* scale; r = round.(v)
    all(abs.(v .- r) .<= 1e-6) || error("capacity is not representable at fp_scale=$scale; choose a finer scale")
    return Int32.(r)
end
# Loads: the reference tests `cap_avail < load` in Float64 (PowerSystemAdequacy.jl:192,253); with capacities on the
# grid, c < L  <=>  c < ceil(L), so ceil keeps every loss-of-load hour exactly (LOLE / LOLF unbiased; the deficit is
# over-stated by < 1 grid unit per loss hour -- a finer fp_scale shrinks that).  Loads on the grid are unchanged.
function fixed_load(x::AbstractVector{Float64}, scale::Float64)
    v = x .* scale; r = round.(v)
    return Int32.(ifelse.(abs.(v .- r) .<= 1e-6, r, ceil.(v)))
end

function set_system!(e::Engine, gens::Vector{Generator}, load::LoadModel; fp_scale::Float64=1.0)
    cap = fixed([g.capacity for g in gens], fp_scale)
    mttf = [g.mttf for g in gens]; mttr = [g.mttr for g in gens]
    ld = fixed_load(load.hourly_load, fp_scale)
    GC.@preserve cap mttf mttr ld begin
        check(e, ccall((:psra_set_system, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int32),
                       e.h, cap, mttf, mttr, Int32(length(cap))))
        check(e, ccall((:psra_set_load, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int32), e.h, ld, Int32(length(ld))))
    end
    e.fp_scale = fp_scale
    return nothing
end

# ---------------------------------------------------------------- engines
"run_analytical, PowerSystemAdequacy.jl:113-163"
function run_analytical(gens::Vector{Generator}, load::LoadModel; step_size::Float64=10.0,
                        engine::Engine=default_engine())
    t_start = time()
    cap = [g.capacity for g in gens]; q = [g.for_rate for g in gens]
    max_len = Int32(ceil(Int, sum(cap) / step_size) + 2 * length(cap) + 8)
    probs = zeros(Float64, max_len); n = Ref{Int32}(0)
    lole = Ref{Float64}(0.0); eue = Ref{Float64}(0.0)
    total_installed = sum(g.capacity for g in gens)
    GC.@preserve cap q probs begin
        check(engine, ccall((:psra_copt, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Int32, Ref{Int32}),
                            engine.h, cap, q, Int32(length(cap)), step_size, probs, max_len, n))
        check(engine, ccall((:psra_copt_indices, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Float64, Ptr{Float64}, Int32, Ref{Float64}, Ref{Float64}),
                            engine.h, probs, n[], step_size, total_installed, load.hourly_load,
                            Int32(length(load.hourly_load)), lole, eue))
    end
    return ReliabilityResult("Analytical", lole[], eue[], time() - t_start, Float64[])
end

"run_non_sequential_mc, PowerSystemAdequacy.jl:169-208 (history: running mean every 100 iterations)"
function run_non_sequential_mc(gens::Vector{Generator}, load::LoadModel, iterations::Int;
                               seed::Integer=42, fp_scale::Float64=1.0, engine::Engine=default_engine())
    t_start = time()
    set_system!(engine, gens, load; fp_scale=fp_scale)
    history = zeros(Float64, div(iterations, 100))
    s = PsraNonseqSummary()
    GC.@preserve history begin
        out = Ref(PsraNonseqOutputs(C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, Int32(100), Int32(0), pointer(history)))
        check(engine, ccall((:psra_nonseq_mc, LIB), Cint,
                            (Ptr{Cvoid}, Int64, Int64, UInt64, Ref{PsraNonseqOutputs}, Ref{PsraNonseqSummary}),
                            engine.h, 0, iterations, UInt64(seed), out, s))
    end
    return ReliabilityResult("Non-Sequential MC", s.sum_lol_hours / iterations,
                             s.sum_ens_fp / iterations / fp_scale, time() - t_start, history)
end

struct SequentialIndices
    years::Int64
    lole::Float64; eens::Float64; lolf::Float64; lold::Float64
    p_loss_year::Float64; events::UInt64; kernel_ms::Float32
    history::Vector{Float64}
end

"Sequential MC with the full index set (LOLE, EENS, LOLF = mean NLC, LOLD; Montecarlo_seq/seqMain.m:160-213)"
function run_sequential_indices(gens::Vector{Generator}, load::LoadModel, years::Int; seed::Integer=42,
                                year0::Integer=0, fp_scale::Float64=1.0, init_mode::Int32=PSRA_INIT_STATIONARY,
                                years_per_chain::Integer=1, keep_on_device::Bool=false, tail_hist::Bool=false,
                                engine::Engine=default_engine())
    set_system!(engine, gens, load; fp_scale=fp_scale)
    history = zeros(Float64, div(years, 10))          # running mean every 10 years, computed on the device
    s = PsraSeqSummary()
    GC.@preserve history begin
        out = Ref(PsraSeqOutputs(C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, Int32(10), Int32(keep_on_device), pointer(history),
                                 Int32(tail_hist), Int32(0)))
        check(engine, ccall((:psra_seq_mc, LIB), Cint,
                            (Ptr{Cvoid}, Int64, Int64, UInt64, Int32, Int32, Ref{PsraSeqOutputs}, Ref{PsraSeqSummary}),
                            engine.h, year0, years, UInt64(seed), init_mode, Int32(years_per_chain), out, s))
    end
    n = max(s.years, 1)
    return SequentialIndices(s.years, s.sum_lol_hours / n, s.sum_ens_fp / n / fp_scale, s.sum_entries / n,
                             s.sum_entries > 0 ? s.sum_lol_hours / s.sum_entries : 0.0,
                             s.years_with_loss / n, s.events, s.kernel_ms, history)
end

"""
run_sequential_mc, PowerSystemAdequacy.jl:214-269 (history: running mean every 10 years).

Default semantics differ from the reference in one documented way (INTEGRATION.md section 4): the reference runs ONE
chain that starts all-up and carries the unit states across all years (:223-224); the default here is independent
years from the stationary law (`init_mode=PSRA_INIT_STATIONARY, years_per_chain=1`): the same expectation without the
all-up start bias, and what lets the years shard over GPUs.  `init_mode=PSRA_INIT_ALL_UP, years_per_chain=years`
reproduces the reference's chain.  `engine=Engine(ngpus=8)` uses eight GPUs from this one call.
"""
function run_sequential_mc(gens::Vector{Generator}, load::LoadModel, years::Int; kwargs...)
    t_start = time()
    r = run_sequential_indices(gens, load, years; kwargs...)
    return ReliabilityResult("Sequential MC", r.lole, r.eens, time() - t_start, r.history)
end

"Weak-point detection of Montecarlo_seq/seqMain.m:140-150,225-231 at HL1: (comp_importance, down_in_loss, indices)"
function unit_importance(gens::Vector{Generator}, load::LoadModel, years::Int; seed::Integer=42, year0::Integer=0,
                         fp_scale::Float64=1.0, init_mode::Int32=PSRA_INIT_STATIONARY, years_per_chain::Integer=1,
                         engine::Engine=default_engine())
    set_system!(engine, gens, load; fp_scale=fp_scale)
    cnt = zeros(UInt64, length(gens))
    s = PsraSeqSummary()
    GC.@preserve cnt begin
        check(engine, ccall((:psra_seq_unit_importance, LIB), Cint,
                            (Ptr{Cvoid}, Int64, Int64, UInt64, Int32, Int32, Ptr{UInt64}, Ptr{Cvoid}, Ref{PsraSeqSummary}),
                            engine.h, year0, years, UInt64(seed), init_mode, Int32(years_per_chain), cnt, C_NULL, s))
    end
    n = max(s.years, 1)
    idx = SequentialIndices(s.years, s.sum_lol_hours / n, s.sum_ens_fp / n / fp_scale, s.sum_entries / n,
                            s.sum_entries > 0 ? s.sum_lol_hours / s.sum_entries : 0.0,
                            s.years_with_loss / n, s.events, s.kernel_ms, Float64[])
    imp = s.sum_lol_hours > 0 ? Float64.(cnt) ./ s.sum_lol_hours : zeros(Float64, length(gens))
    return imp, cnt, idx
end

"VaR / CVaR of the per-year ENS of the last run_sequential_indices(...; tail_hist=true) (exact, from the device's ENS histogram: no per-year vector) or (...; keep_on_device=true) (radix select over the kept vector)"
function tail_risk(alphas::Vector{Float64}=[0.95, 0.99]; engine::Engine=default_engine())
    outs = Vector{PsraTailOut}(undef, length(alphas))
    GC.@preserve alphas outs begin
        check(engine, ccall((:psra_tail, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Int32, Ptr{PsraTailOut}, Ptr{Int64}, Int32, Int64),
                            engine.h, C_NULL, 0, alphas, Int32(length(alphas)), outs, C_NULL, Int32(0), 1))
    end
    return [(alpha=a, var=o.var / engine.fp_scale, cvar=o.cvar / engine.fp_scale, n_tail=o.n_tail)
            for (a, o) in zip(alphas, outs)]
end

# ---------------------------------------------------------------- tail_risk.jl / MCvsMarkovProcess.jl engine
"mutable struct Generator of generating_adequacy_comprehensive.jl:11-24"
mutable struct DetailedGenerator
    name::String
    capacity::Float64
    for_rate::Float64
    maintenance_weeks::Int
    energy_limit::Float64
    effective_q::Float64
    scheduled_outage_start::Int
end
DetailedGenerator(name, cap, q, maint, elim) = DetailedGenerator(name, cap, q, maint, elim, q, 0)

struct PsraDetailedSystem
    capacity_mw::Ptr{Float64}; for_rate::Ptr{Float64}; maint_start_week::Ptr{Int32}; maint_weeks::Ptr{Int32}
    energy_limit_mwh::Ptr{Float64}; n_units::Int32; reserved::Int32
end

"run_detailed_mc, tail_risk.jl:12-91 -> (yearly_lole_distribution, hourly_failure_prob)"
function run_detailed_mc(gens::Vector{DetailedGenerator}, base_load::Vector{Float64}, lfu_sigma_percent::Float64,
                         n_years::Int; seed::Integer=42, engine::Engine=default_engine())
    cap = [g.capacity for g in gens]; q = [g.for_rate for g in gens]
    ms = Int32[g.scheduled_outage_start for g in gens]; mw = Int32[g.maintenance_weeks for g in gens]
    el = [g.energy_limit for g in gens]
    lfu_std_dev = maximum(base_load) * (lfu_sigma_percent / 100.0)
    yl = zeros(UInt32, n_years); hf = zeros(UInt32, length(base_load)); ms_k = Ref{Float32}(0f0)
    GC.@preserve cap q ms mw el base_load yl hf begin
        sys = Ref(PsraDetailedSystem(pointer(cap), pointer(q), pointer(ms), pointer(mw), pointer(el),
                                     Int32(length(gens)), Int32(0)))
        check(engine, ccall((:psra_detailed_mc, LIB), Cint,
                            (Ptr{Cvoid}, Ref{PsraDetailedSystem}, Ptr{Float64}, Int32, Float64, Int64, Int64, UInt64,
                             Ptr{UInt32}, Ptr{UInt32}, Ref{Float32}),
                            engine.h, sys, base_load, Int32(length(base_load)), lfu_std_dev, 0, n_years, UInt64(seed),
                            yl, hf, ms_k))
    end
    return Float64.(yl), Float64.(hf) ./ n_years
end

"compare_results, PowerSystemAdequacy.jl:275-285 (table only; plotting stays with the caller)"
function compare_results(results::Vector{ReliabilityResult})
    println("\n==========================================")
    println("       METHOD COMPARISON SUMMARY")
    println("==========================================")
    @printf "%-20s | %-10s | %-10s | %-10s\n" "Method" "LOLE(h/yr)" "EUE(MWh)" "Time(s)"
    println("-"^60)
    for r in results
        @printf "%-20s | %-10.4f | %-10.2f | %-10.4f\n" r.method r.lole_hours_yr r.eue_mwh_yr r.computation_time
    end
    println("-"^60)
end

end # module
