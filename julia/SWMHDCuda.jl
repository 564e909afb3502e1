# SWMHDCuda.jl — Julia-side binding of libswmhd_cuda.so for the reference scripts.
#
# UNTESTED IN THE BUILD CONTAINER (no Julia there).  It is the glue a maintainer adds so that
# jacobian_formulation/SWMHD_example.jl:21-97 and divergence_formulation/divergence_sw_mhd.jl:19-96
# run unchanged: the model object stays a real Oceananigans.ShallowWaterModel (so callbacks,
# AbstractOperations diagnostics and OutputWriters keep working), only `time_step!` is redirected
# for models that were attached.
#
#   include("SWMHDCuda.jl"); using .SWMHDCuda
#   model = ShallowWaterModel(...)            # exactly as in the reference script
#   set!(model, u = uᵢ, v = vᵢ, h = hᵢ, A = Aᵢ)
#   SWMHDCuda.attach!(model)                  # uploads the parent arrays, builds the GPU context
#   SWMHDCuda.attach!(model; n_gpus = 8)      # ... or one context driving 8 B200s (y-slabs, NCCL inside the library)
#   run!(simulation)                          # time_step! now runs on the B200(s)
#
# Two download policies (attach!(...; download = ...)):
#   :every_step  (default) the four parents are refreshed after every step, so the reference's progress callback
#                (SWMHD_example.jl:47-65 reads model.solution.u, .h, model.tracers.A every iteration) and its writers
#                work with no change to the script.  538 MB per step at 4096^2: PCIe-bound.
#   :on_demand   nothing is copied per step.  The step also evaluates the diagnostics on the device
#                (swmhd_step_diag, fused into the stage-1 kernel): use `SWMHDCuda.progress` as the progress callback
#                (same message, max|u|, max|A|, min h from the device) and add
#                    simulation.callbacks[:sync] = Callback(SWMHDCuda.sync_fields!, TimeInterval(0.1))
#                with the schedule of the field writer (SWMHD_example.jl:81-84): callbacks run before writers.
module SWMHDCuda

using Oceananigans
using Oceananigans: AbstractModel
using Oceananigans.Models.ShallowWaterModels: ShallowWaterModel, VectorInvariantFormulation, ConservativeFormulation
using Oceananigans.Grids: topology, Bounded
using Oceananigans.TimeSteppers: tick!, RungeKutta3TimeStepper
using Oceananigans.Utils: prettytime
using Printf
import Oceananigans.TimeSteppers: time_step!

const LIB = get(ENV, "SWMHD_LIB", joinpath(@__DIR__, "..", "swmhd_b200", "libswmhd_cuda.so"))
const ABI_VERSION = 2

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
    n_gpus::Int32; device_ids::NTuple{8,Int32}; reserved0::Int32
end

struct Diag
    ke::Float64; me::Float64; pe::Float64; total::Float64
    max_abs_u::Float64; max_abs_A::Float64; min_h::Float64
    max_abs_div_hB::Float64; sum_h::Float64
    all_finite::Int32; reserved::Int32
end

mutable struct Attached
    ctx::Ptr{Cvoid}
    download::Symbol
    stale::Bool             # the Julia-side parents are older than the device state
    last::Diag              # diagnostics of the state at the start of the last step (:on_demand)
end

const CTX = IdDict{Any,Attached}()

check(rc, ctx = C_NULL) = rc == 0 ? nothing :
    error("libswmhd_cuda error $rc: " * unsafe_string(ccall((:swmhd_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

fields_of(model) = (model.solution[1], model.solution[2], model.solution.h, model.tracers.A)

"""
    attach!(model; n_gpus = 1, device_ids = 0:n_gpus-1, arith = 0, A_gradient = nothing, download = :every_step)

Build the GPU context for `model` and upload its fields.  `n_gpus > 1`: ONE context splits the grid into y-slabs, one per
device, and exchanges the halo rows itself (ncclSend/ncclRecv inside libswmhd_cuda.so); the parents passed to
`swmhd_set_field` / `swmhd_get_field` stay the global arrays, so nothing else changes on the Julia side.
"""
function attach!(model::ShallowWaterModel; n_gpus = 1, device_ids = 0:n_gpus-1, arith = 0, A_gradient = nothing, download = :every_step)
    grid = model.grid
    form = model.formulation isa VectorInvariantFormulation ? 0 : 1
    by = topology(grid, 2) == Bounded ? 1 : 0
    gs, gn = A_gradient === nothing ? (0.0, 0.0) : A_gradient
    ids = ntuple(k -> k <= n_gpus ? Int32(device_ids[k]) : Int32(0), 8)
    cfg = Config(ABI_VERSION, grid.Nx, grid.Ny, grid.Hx, grid.Hy, 0, by, form, arith, 0,
                 grid.Δxᶜᵃᵃ, grid.Δyᵃᶜᵃ, model.gravitational_acceleration, model.coriolis.f,
                 1e-6, 1.0, A_gradient === nothing ? 0 : 1, ids[1], gs, gn, 0, grid.Ny, 0, 1,
                 n_gpus > 1 ? n_gpus : 0, ids, 0)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:swmhd_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, ctx))
    CTX[model] = Attached(ctx[], download, false, Diag(0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0))
    upload!(model)
    return model
end

"set!(model, ...) happened on the Julia side: push the parent arrays and refill halos."
function upload!(model)
    a = CTX[model]
    for (k, f) in enumerate(fields_of(model))
        p = parent(f)                                   # (Nx+6) x (Ny_f+6) x 1, column-major: the ABI layout
        check(ccall((:swmhd_set_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Csize_t), a.ctx, k - 1, p, length(p)), a.ctx)
    end
    check(ccall((:swmhd_fill_halos, LIB), Cint, (Ptr{Cvoid},), a.ctx), a.ctx)
    a.stale = false
end

"Refresh the Julia-side fields (with halos) before callbacks / output writers read them."
function download!(model)
    a = CTX[model]
    for (k, f) in enumerate(fields_of(model))
        p = parent(f)
        check(ccall((:swmhd_get_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Csize_t), a.ctx, k - 1, p, length(p)), a.ctx)
    end
    a.stale = false
end

"Callback for the :on_demand policy: `Callback(SWMHDCuda.sync_fields!, TimeInterval(0.1))` (the field writer's schedule)."
sync_fields!(sim) = (haskey(CTX, sim.model) && CTX[sim.model].stale && download!(sim.model); nothing)

"Progress callback with the message of SWMHD_example.jl:47-63; the three reductions come from the fused device diagnostics."
function progress(sim)
    d = CTX[sim.model].last
    @info @sprintf("Iter: %d, time: %s, Δt: %s, max|u|: %.3e, max|A|: %.3e, min h: %.3e",
                   sim.model.clock.iteration, prettytime(sim.model.clock.time), prettytime(sim.Δt), d.max_abs_u, d.max_abs_A, d.min_h)
end

# time_step!(model, Δt) — SWMHD_example.jl:97 via run!: one RK3 step on the GPU.
# Upstream's method is typed on AbstractModel{<:RungeKutta3TimeStepper}; this one on the intersection with
# ShallowWaterModel, which is more specific than either (no ambiguity).  Models that were not attached take upstream's path.
const Steppable = typeintersect(ShallowWaterModel, AbstractModel{<:RungeKutta3TimeStepper})

function time_step!(model::Steppable, Δt; kw...)
    haskey(CTX, model) || return invoke(time_step!, Tuple{AbstractModel{<:RungeKutta3TimeStepper},Any}, model, Δt; kw...)
    a = CTX[model]
    if a.download === :on_demand
        d = Ref{Diag}()
        check(ccall((:swmhd_step_diag, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cint, Ref{Diag}), a.ctx, Δt, 1, d), a.ctx)
        a.last = d[]
        a.stale = true
    else
        check(ccall((:swmhd_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cint), a.ctx, Δt, 1), a.ctx)
        download!(model)            # the reference's callbacks and writers read the fields every iteration
    end
    tick!(model.clock, Δt)          # upstream ticks stage by stage; the sum is the same to round-off
    return nothing
end

function diagnostics(model)
    d = Ref{Diag}()
    a = CTX[model]
    check(ccall((:swmhd_diagnostics, LIB), Cint, (Ptr{Cvoid}, Ref{Diag}), a.ctx, d), a.ctx)
    return d[]
end

detach!(model) = (ccall((:swmhd_destroy, LIB), Cvoid, (Ptr{Cvoid},), pop!(CTX, model).ctx); nothing)

end # module
