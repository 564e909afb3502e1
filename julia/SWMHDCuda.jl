# SWMHDCuda.jl — Julia-side binding of libswmhd_cuda.so for the reference scripts.
#
# UNTESTED IN THE BUILD CONTAINER (no Julia there).  It is the glue a maintainer adds so that
# jacobian_formulation/SWMHD_example.jl:21-97 and divergence_formulation/divergence_sw_mhd.jl:19-96
# run unchanged: the model object stays a real Oceananigans.ShallowWaterModel (so callbacks,
# AbstractOperations diagnostics and OutputWriters keep working), only `time_step!` is redirected.
#
#   include("SWMHDCuda.jl"); using .SWMHDCuda
#   model = ShallowWaterModel(...)            # exactly as in the reference script
#   set!(model, u = uᵢ, v = vᵢ, h = hᵢ, A = Aᵢ)
#   SWMHDCuda.attach!(model)                  # uploads the parent arrays, builds the GPU context
#   run!(simulation)                          # time_step! now runs on the B200
module SWMHDCuda

using Oceananigans
using Oceananigans.Models.ShallowWaterModels: ShallowWaterModel, VectorInvariantFormulation, ConservativeFormulation
using Oceananigans.Grids: topology, Bounded
using Oceananigans.TimeSteppers: tick!
import Oceananigans.TimeSteppers: time_step!

const LIB = get(ENV, "SWMHD_LIB", joinpath(@__DIR__, "..", "swmhd_b200", "libswmhd_cuda.so"))

# struct swmhd_config (include/swmhd.h) — field order and types must match exactly
struct Config
    abi_version::Int32
    Nx::Int32; Ny::Int32; Hx::Int32; Hy::Int32
    topo_x::Int32; topo_y::Int32
    formulation::Int32; arith::Int32; flags::Int32
    dx::Float64; dy::Float64; g::Float64; f::Float64
    weno_eps::Float64; h_ref::Float64
    A_gradient_bc::Int32; device::Int32
    A_grad_south::Float64; A_grad_north::Float64
    slab_j0::Int32; slab_ny::Int32; rank::Int32; world::Int32
end

struct Diag
    ke::Float64; me::Float64; pe::Float64; total::Float64
    max_abs_u::Float64; max_abs_A::Float64; min_h::Float64
    max_abs_div_hB::Float64; sum_h::Float64
    all_finite::Int32; reserved::Int32
end

const CTX = IdDict{Any,Ptr{Cvoid}}()

check(rc, ctx = C_NULL) = rc == 0 ? nothing :
    error("libswmhd_cuda error $rc: " * unsafe_string(ccall((:swmhd_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

fields_of(model) = (model.solution[1], model.solution[2], model.solution.h, model.tracers.A)

function attach!(model::ShallowWaterModel; arith = 0, device = 0, A_gradient = nothing)
    grid = model.grid
    form = model.formulation isa VectorInvariantFormulation ? 0 : 1
    by = topology(grid, 2) == Bounded ? 1 : 0
    gs, gn = A_gradient === nothing ? (0.0, 0.0) : A_gradient
    cfg = Config(1, grid.Nx, grid.Ny, grid.Hx, grid.Hy, 0, by, form, arith, 0,
                 grid.Δxᶜᵃᵃ, grid.Δyᵃᶜᵃ, model.gravitational_acceleration, model.coriolis.f,
                 1e-6, 1.0, A_gradient === nothing ? 0 : 1, device, gs, gn, 0, grid.Ny, 0, 1)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:swmhd_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, ctx))
    CTX[model] = ctx[]
    upload!(model)
    return model
end

"set!(model, ...) happened on the Julia side: push the parent arrays and refill halos."
function upload!(model)
    ctx = CTX[model]
    for (k, f) in enumerate(fields_of(model))
        p = parent(f)                                   # (Nx+6) x (Ny_f+6) x 1, column-major: the ABI layout
        check(ccall((:swmhd_set_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Csize_t), ctx, k - 1, p, length(p)), ctx)
    end
    check(ccall((:swmhd_fill_halos, LIB), Cint, (Ptr{Cvoid},), ctx), ctx)
end

"Refresh the Julia-side fields (with halos) before callbacks / output writers read them."
function download!(model)
    ctx = CTX[model]
    for (k, f) in enumerate(fields_of(model))
        p = parent(f)
        check(ccall((:swmhd_get_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Csize_t), ctx, k - 1, p, length(p)), ctx)
    end
end

# time_step!(model, Δt) — SWMHD_example.jl:97 via run!: one RK3 step on the GPU
function time_step!(model::ShallowWaterModel, Δt; callbacks = nothing, euler = false)
    haskey(CTX, model) || return invoke(time_step!, Tuple{Any,Any}, model, Δt)
    ctx = CTX[model]
    check(ccall((:swmhd_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cint), ctx, Δt, 1), ctx)
    tick!(model.clock, Δt)          # upstream ticks stage by stage; the sum is the same to round-off
    download!(model)                # callbacks and writers of the reference scripts read every iteration
    return nothing
end

function diagnostics(model)
    d = Ref{Diag}()
    check(ccall((:swmhd_diagnostics, LIB), Cint, (Ptr{Cvoid}, Ref{Diag}), CTX[model], d), CTX[model])
    return d[]
end

detach!(model) = (ccall((:swmhd_destroy, LIB), Cvoid, (Ptr{Cvoid},), pop!(CTX, model)); nothing)

end # module
