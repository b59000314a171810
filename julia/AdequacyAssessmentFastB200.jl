# AdequacyAssessmentFastB200.jl -- drop-in for GeneratingAdequacy/AdequacyAssessmentII.jl
# (module AdequacyAssessmentFast): same exports -- Generator, TieLine, Area, System,
# run_fast_sequential_simulation, SupportPolicy, ISOLATED, INTERCONNECTED -- the simulation body is one
# `ccall` of psra_multi_area_mc (include/psra_b200.h) into libpsra_b200.so.
#
# NOTE: Julia is not installed in the build image; the file is written against the header and mirrors the
# ctypes binding (powersystemsreliabilityassessment_b200/_lib.py: AreaSystem / AreaOutputs / AreaSummary),
# which the GPU tests exercise (tests/test_gpu_multi_area.py).
module AdequacyAssessmentFastB200

using Printf

export Generator, TieLine, Area, System, run_fast_sequential_simulation, SupportPolicy, ISOLATED, INTERCONNECTED

const LIB = get(ENV, "PSRA_B200_LIB", joinpath(@__DIR__, "..", "powersystemsreliabilityassessment_b200", "libpsra_b200.so"))
const PSRA_MAX_AREAS = 8

# ---------------------------------------------------------------- data model (AdequacyAssessmentII.jl:15-63)
struct Generator                 # the reference's mutable state (current_state, time_to_transition) lives on the device
    id::String
    capacity::Float64
    mttf::Float64
    mttr::Float64
end
struct TieLine
    from_area::Int
    to_area::Int
    capacity::Float64
end
struct Area
    id::Int
    name::String
    generators::Vector{Generator}
    hourly_load::Vector{Float64}
end
struct System
    areas::Vector{Area}
    tie_lines::Vector{TieLine}
    topology_matrix::Matrix{Float64}
end
function System(areas, lines)
    n = length(areas)
    mat = zeros(Float64, n, n)
    for line in lines                                   # bidirectional, :56-60
        mat[line.from_area, line.to_area] += line.capacity
        mat[line.to_area, line.from_area] += line.capacity
    end
    return System(areas, lines, mat)
end
@enum SupportPolicy ISOLATED INTERCONNECTED

# ---------------------------------------------------------------- C structs (psra_b200.h)
struct PsraConfig
    device::Int32; warps_per_block::Int32; seg_hours::Int32; blocks_per_sm::Int32
    reserved::NTuple{4,Int32}
    ngpus::Int32; ev_cap::Int32; tail_bins::Int32
    reserved2::NTuple{5,Int32}
end
struct PsraAreaSystem
    n_areas::Int32; n_units::Int32; n_hours::Int32; reserved::Int32
    unit_area::Ptr{Int32}; cap_fp::Ptr{Int32}; mttf_h::Ptr{Float64}; mttr_h::Ptr{Float64}
    load_fp::Ptr{Int32}; topology_fp::Ptr{Int32}
end
struct PsraAreaOutputs
    lol_hours::Ptr{UInt32}; ens_fp::Ptr{Int64}
end
mutable struct PsraAreaSummary
    years::Int64
    sum_lol_hours::NTuple{8,Int64}
    sum_ens_fp::NTuple{8,Int64}
    events::UInt64
    kernel_ms::Float32
    n_areas::Int32
    PsraAreaSummary() = new(0, ntuple(_ -> Int64(0), 8), ntuple(_ -> Int64(0), 8), 0, 0f0, 0)
end

# capacities and tie capacities must lie on the fixed-point grid (error otherwise); loads are ceiled onto it, which
# keeps the reference's Float64 test `margin < 0` exact for whole-grid capacities (c < L <=> c < ceil(L))
function fixed(x::AbstractVector{Float64}, scale::Float64)
    v = x .* scale; r = round.(v)
    all(abs.(v .- r) .<= 1e-6) || error("value not representable at fp_scale=$scale; choose a finer scale")
    return Int32.(r)
end
function fixed_load(x::AbstractVector{Float64}, scale::Float64)
    v = x .* scale; r = round.(v)
    return Int32.(ifelse.(abs.(v .- r) .<= 1e-6, r, ceil.(v)))
end

"""
    run_fast_sequential_simulation(sys, policy, n_years; seed=42, fp_scale=1.0)

Same call and return value as AdequacyAssessmentII.jl:185-250: a vector of `(area, lole, eue)` named tuples.
Years are independent (own Philox streams keyed (seed; year, unit)); MW values are converted to fixed point
(capacities must lie on the grid; loads are ceiled onto it).
"""
function run_fast_sequential_simulation(sys::System, policy::SupportPolicy, n_years::Int; seed::Integer=42,
                                        fp_scale::Float64=1.0, device::Integer=0)
    t_start = time()
    println("--- Running FAST Adequacy Assessment ---")
    println("Policy: $policy | Years: $n_years")
    n_areas = length(sys.areas)
    n_areas <= PSRA_MAX_AREAS || error("at most $PSRA_MAX_AREAS areas")
    H = length(sys.areas[1].hourly_load)
    unit_area = Int32[]; cap = Float64[]; mttf = Float64[]; mttr = Float64[]
    for (i, area) in enumerate(sys.areas), g in area.generators
        push!(unit_area, Int32(i - 1)); push!(cap, g.capacity); push!(mttf, g.mttf); push!(mttr, g.mttr)
    end
    capi = fixed(cap, fp_scale)
    loads = Int32[]                                     # [n_areas][H], row-major for C
    for area in sys.areas
        length(area.hourly_load) == H || error("all areas need load curves of the same length")
        append!(loads, fixed_load(area.hourly_load, fp_scale))
    end
    topo = fixed(vec(permutedims(sys.topology_matrix)), fp_scale)   # symmetric; row-major for C

    h = Ref{Ptr{Cvoid}}(C_NULL)
    cfg = Ref(PsraConfig(Int32(device), 0, 0, 0, (Int32(0), Int32(0), Int32(0), Int32(0)), Int32(1), Int32(0), Int32(0),
                         (Int32(0), Int32(0), Int32(0), Int32(0), Int32(0))))
    rc = ccall((:psra_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{PsraConfig}), h, cfg)
    rc == 0 || error("psra_create failed ($rc): no CUDA device / library? (there is no CPU fallback)")
    s = PsraAreaSummary()
    try
        GC.@preserve unit_area capi mttf mttr loads topo begin
            a = Ref(PsraAreaSystem(Int32(n_areas), Int32(length(capi)), Int32(H), Int32(0), pointer(unit_area), pointer(capi),
                                   pointer(mttf), pointer(mttr), pointer(loads), pointer(topo)))
            o = Ref(PsraAreaOutputs(C_NULL, C_NULL))
            rc = ccall((:psra_multi_area_mc, LIB), Cint,          # replaces the year/hour/area loops, :199-236
                       (Ptr{Cvoid}, Ref{PsraAreaSystem}, Int32, Int64, Int64, UInt64, Int32, Ref{PsraAreaOutputs}, Ref{PsraAreaSummary}),
                       h[], a, Int32(Int(policy)), 0, n_years, UInt64(seed), Int32(1), o, s)
        end
        rc == 0 || error(unsafe_string(ccall((:psra_last_error, LIB), Cstring, (Ptr{Cvoid},), h[])))
    finally
        ccall((:psra_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
    elapsed = time() - t_start
    println("Simulation completed in $(round(elapsed, digits=2)) seconds.")
    results = []
    for i in 1:n_areas
        push!(results, (area = sys.areas[i].name,
                        lole = s.sum_lol_hours[i] / n_years,
                        eue = s.sum_ens_fp[i] / n_years / fp_scale))
    end
    return results
end

# The reference's demo driver (AdequacyAssessmentII.jl:256-291, run_adequacy_assessmentII.jl) stays with the reference:
# include this module instead of AdequacyAssessmentII.jl and its `run_demo` body works unchanged.

end # module
