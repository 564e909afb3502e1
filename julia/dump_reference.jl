# dump_reference.jl — run on any machine with Julia + Oceananigans v0.76.x to produce golden
# binaries from the REAL reference (closes the "parity unpinned" gap of SURVEY 8c).
# Writes raw Float64 parent arrays (with halos) after 0, 1, 10, 100 and 1000 RK3 steps:
#     golden_<form>_<N>_step<k>_<field>.f64     (same layout as swmhd_set_field / swmhd_get_field)
# Compare with:  python tools/compare_reference_dump.py <dir>
using Oceananigans
using Oceananigans.Models.ShallowWaterModels: VectorInvariantFormulation, ConservativeFormulation
using Oceananigans.Advection: VelocityStencil
using Oceananigans.Operators
using Oceananigans.Grids: topology

refroot = get(ENV, "SWMHD_REFERENCE", joinpath(@__DIR__, "..", "..", "reference"))

function build(form, N)
    grid = RectilinearGrid(size = (N, N), x = (-5, 5), y = (-5, 5), topology = (Periodic, Periodic, Flat))
    if form == :jacobian
        include(joinpath(refroot, "jacobian_formulation", "sw_mhd_jacobian_functions.jl"))
        model = ShallowWaterModel(grid = grid, timestepper = :RungeKutta3,
            momentum_advection = WENO5(vector_invariant = VelocityStencil()), mass_advection = WENO5(), tracer_advection = WENO5(),
            gravitational_acceleration = 9.81, coriolis = FPlane(f = 1), tracers = (:A),
            forcing = (u = Forcing(lorentz_force_func_x, discrete_form = true), v = Forcing(lorentz_force_func_y, discrete_form = true)),
            formulation = VectorInvariantFormulation())
        set!(model, u = (x, y, z) -> 5y * exp(-(x^2 + y^2)), v = (x, y, z) -> -5x * exp(-(x^2 + y^2)), h = 1, A = (x, y, z) -> 0.5abs(y))
    else
        include(joinpath(refroot, "divergence_formulation", "sw_mhd_divergence_functions.jl"))
        model = ShallowWaterModel(grid = grid, timestepper = :RungeKutta3,
            momentum_advection = WENO5(), mass_advection = WENO5(), tracer_advection = WENO5(),
            gravitational_acceleration = 9.81, coriolis = FPlane(f = 1), tracers = (:A),
            forcing = (uh = Forcing(div_lorentz_x, discrete_form = true), vh = Forcing(div_lorentz_y, discrete_form = true)),
            formulation = ConservativeFormulation())
        set!(model, h = 1, A = (x, y, z) -> 0.5exp(-((x - 0.5)^2 + y^2)) - 0.5exp(-((x + 0.5)^2 + y^2)))
    end
    return model
end

# Bounded-y with activity AT the walls (pins Appendix-C item C10: halo depth of the no-flux / gradient BCs, the WENO wall
# fallback, v beyond the wall — the published Bounded-y figures cannot, their wall region is quiescent, profiles/r02_c10_probe.md).
# A = -0.05 y + 0.02 exp(-(x^2 + (y-4)^2)) with GradientBoundaryCondition(-0.05) (divergence_sw_mhd.jl:17), vortex centred at (0, 3.5).
function build_wall(form, N)
    grid = RectilinearGrid(size = (N, N), x = (-5, 5), y = (-5, 5), topology = (Periodic, Bounded, Flat))
    A_bcs = FieldBoundaryConditions(north = GradientBoundaryCondition(-0.05), south = GradientBoundaryCondition(-0.05))
    Ai(x, y, z) = -0.05y + 0.02exp(-(x^2 + (y - 4)^2))
    ui(x, y, z) = (y - 3.5) * exp(-(x^2 + (y - 3.5)^2))
    vi(x, y, z) = -x * exp(-(x^2 + (y - 3.5)^2))
    if form == :jacobian_wall
        include(joinpath(refroot, "jacobian_formulation", "sw_mhd_jacobian_functions.jl"))
        model = ShallowWaterModel(grid = grid, timestepper = :RungeKutta3, boundary_conditions = (A = A_bcs,),
            momentum_advection = WENO5(vector_invariant = VelocityStencil()), mass_advection = WENO5(), tracer_advection = WENO5(),
            gravitational_acceleration = 9.81, coriolis = FPlane(f = 1), tracers = (:A),
            forcing = (u = Forcing(lorentz_force_func_x, discrete_form = true), v = Forcing(lorentz_force_func_y, discrete_form = true)),
            formulation = VectorInvariantFormulation())
        set!(model, u = ui, v = vi, h = 1, A = Ai)
    else
        include(joinpath(refroot, "divergence_formulation", "sw_mhd_divergence_functions.jl"))
        model = ShallowWaterModel(grid = grid, timestepper = :RungeKutta3, boundary_conditions = (A = A_bcs,),
            momentum_advection = WENO5(), mass_advection = WENO5(), tracer_advection = WENO5(),
            gravitational_acceleration = 9.81, coriolis = FPlane(f = 1), tracers = (:A),
            forcing = (uh = Forcing(div_lorentz_x, discrete_form = true), vh = Forcing(div_lorentz_y, discrete_form = true)),
            formulation = ConservativeFormulation())
        set!(model, uh = ui, vh = vi, h = 1, A = Ai)
    end
    return model
end

function dump(model, form, N, k, outdir)
    names = form in (:jacobian, :jacobian_wall) ? (:u, :v, :h) : (:uh, :vh, :h)
    fields = (getproperty(model.solution, names[1]), getproperty(model.solution, names[2]), model.solution.h, model.tracers.A)
    for (f, tag) in zip(fields, ("u", "v", "h", "A"))
        write(joinpath(outdir, "golden_$(form)_$(N)_step$(k)_$(tag).f64"), Array(parent(f)))
    end
end

outdir = length(ARGS) > 0 ? ARGS[1] : "reference_dump"
mkpath(outdir)
for form in (:jacobian, :divergence), N in (64,)
    model = build(form, N)
    Oceananigans.TimeSteppers.update_state!(model)
    dump(model, form, N, 0, outdir)
    n = 0
    for k in (1, 10, 100, 1000)
        for _ in 1:(k - n)
            time_step!(model, 0.01)
        end
        n = k
        dump(model, form, N, k, outdir)
    end
end
for form in (:jacobian_wall, :divergence_wall), N in (64,)
    model = build_wall(form, N)
    Oceananigans.TimeSteppers.update_state!(model)
    dump(model, form, N, 0, outdir)          # step 0 already shows which halo cells upstream fills
    n = 0
    for k in (1, 10, 100)
        for _ in 1:(k - n)
            time_step!(model, 0.01)
        end
        n = k
        dump(model, form, N, k, outdir)
    end
end
println("Oceananigans ", pkgversion(Oceananigans), " -> ", outdir)
