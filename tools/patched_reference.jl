# tools/patched_reference.jl -- runs the REFERENCE's own run_sequential_mc on injected duration lists.
#
# The reference draws every time to failure / repair from Julia's unseeded global RNG
# (GeneratingAdequacy/PowerSystemAdequacy.jl:224,243,246), so its random stream cannot be reproduced and the library's
# bit-exact parity is established against a CPU restatement (oracle/psra_oracle.c).  Whoever has Julia can close that
# gap with this script: it reads the reference source AS TEXT from the checkout given on the command line, replaces the
# three `-log(rand())/rate` draws by reads from per-unit duration lists (and records the per-year sums), evaluates the
# patched module, runs it on the fixture written by tools/export_injected_fixture.py and compares every year with the
# values the oracle and the CUDA path produce for the same lists.  Nothing of the reference is stored in this repository.
#
#   python tools/export_injected_fixture.py /tmp/fixture
#   julia tools/patched_reference.jl /path/to/PowerSystemsReliabilityAssessment /tmp/fixture
#
# NOT RUN in the build image (no Julia there); the substitutions are plain string replacements that fail loudly
# when the reference text changes.

module PSRA_INJ
mutable struct Store
    D::Matrix{Float64}          # D[unit, k]
    used::Vector{Int}
    year_lole::Vector{Float64}
    year_eue::Vector{Float64}
end
const S = Store(zeros(0, 0), Int[], Float64[], Float64[])
function next(i::Int)
    S.used[i] += 1
    S.used[i] <= size(S.D, 2) || error("injected durations exhausted for unit $i")
    return S.D[i, S.used[i]]
end
function record(lole::Float64, eue::Float64)
    push!(S.year_lole, lole)
    push!(S.year_eue, eue)
    return nothing
end
end # module

function read_f64(path::String)
    n = div(filesize(path), 8)
    v = Vector{Float64}(undef, n)
    open(path, "r") do io
        read!(io, v)
    end
    return ltoh.(v)
end

function must_replace(src::String, old::String, new::String)
    occursin(old, src) || error("reference text changed: cannot find `$old`")
    return replace(src, old => new)
end

function main()
    length(ARGS) == 2 || error("usage: julia tools/patched_reference.jl <reference checkout> <fixture directory>")
    ref_file = joinpath(ARGS[1], "GeneratingAdequacy", "PowerSystemAdequacy.jl")
    fx = ARGS[2]
    meta = parse.(Int, split(strip(read(joinpath(fx, "meta.txt"), String))))
    U, K, years, H = meta
    cap = read_f64(joinpath(fx, "cap.f64")); mttf = read_f64(joinpath(fx, "mttf.f64")); mttr = read_f64(joinpath(fx, "mttr.f64"))
    load = read_f64(joinpath(fx, "load.f64")); dur = read_f64(joinpath(fx, "dur.f64"))
    lol_ref = read_f64(joinpath(fx, "lol.f64")); eue_ref = read_f64(joinpath(fx, "eue.f64"))
    length(dur) == U * K && length(load) == H && length(lol_ref) == years || error("fixture sizes do not match meta.txt")
    # numpy wrote dur[u][k] row-major: element (u, k) sits at k + K * u
    PSRA_INJ.S.D = permutedims(reshape(dur, K, U))
    PSRA_INJ.S.used = zeros(Int, U)

    src = read(ref_file, String)
    src = must_replace(src, "ttf = [-log(rand())/g.lambda for g in gens]",
                       "ttf = [Main.PSRA_INJ.next(i) for (i, g) in enumerate(gens)]")
    src = must_replace(src, "ttf[i] += -log(rand())/g.mu", "ttf[i] += Main.PSRA_INJ.next(i)")
    src = must_replace(src, "ttf[i] += -log(rand())/g.lambda", "ttf[i] += Main.PSRA_INJ.next(i)")
    src = must_replace(src, "cum_eue += year_eue",
                       "cum_eue += year_eue; Main.PSRA_INJ.record(year_lole, year_eue)")
    # the plotting package is only needed by compare_results
    src = replace(src, r"^\s*using\s+Plots.*$"m => "")
    include_string(Main, src, ref_file)

    PSA = getfield(Main, :PowerSystemAdequacy)
    # the patched module was defined after main() started: call it in the latest world
    gens = [Base.invokelatest(PSA.Generator, i, cap[i], mttf[i], mttr[i]) for i in 1:U]
    gens = convert(Vector{PSA.Generator}, gens)
    lm = Base.invokelatest(PSA.LoadModel, load)
    res = Base.invokelatest(PSA.run_sequential_mc, gens, lm, years)

    ok = PSRA_INJ.S.year_lole == lol_ref && PSRA_INJ.S.year_eue == eue_ref
    println("reference LOLE = ", res.lole_hours_yr, " h/yr, EUE = ", res.eue_mwh_yr, " MWh/yr over ", years, " years")
    println("per-year LOL hours  reference: ", PSRA_INJ.S.year_lole, "  fixture: ", lol_ref)
    println("per-year ENS        reference: ", PSRA_INJ.S.year_eue, "  fixture: ", eue_ref)
    println(ok ? "PATCHED_REFERENCE_PARITY_OK (bit-exact per year)" : "PATCHED_REFERENCE_PARITY_FAILED")
    exit(ok ? 0 : 1)
end

main()
