/*
 * swmhd.h — C ABI of libswmhd_cuda.so: the B200 (sm_100a) replacement for the
 * per-RK3-substage hot path of writingindy/SWMHD.
 *
 * The reference has no FFI of its own.  Its only extension point is the Julia
 * closure contract  Forcing(f, discrete_form=true)  with
 *     f(i, j, k, grid, clock, model_fields)::Float64
 * (jacobian_formulation/SWMHD_example.jl:30-31,
 *  jacobian_formulation/sw_mhd_jacobian_functions.jl:20-26,
 *  divergence_formulation/divergence_sw_mhd.jl:28-29,
 *  divergence_formulation/sw_mhd_divergence_functions.jl:162-170),
 * which Oceananigans inlines into its tendency kernels.  A per-cell callback
 * cannot cross a C ABI at speed, so the boundary sits one level up, at
 * time_step!(model, dt): every entry point below replaces one call the
 * reference scripts make (directly or through Simulation/run!) on a
 * ShallowWaterModel.  INTEGRATION.md shows the Julia `ccall` stubs.
 *
 * Memory contract: every host buffer passed to set/get is the *parent array*
 * of an Oceananigans Field: (Nx+2Hx) x (Ny_f+2Hy) doubles, column-major,
 * i fastest, Hx=Hy=3.  Ny_f = Ny, except Ny+1 for the y-face field (v|vh)
 * when topo_y is Bounded.  0-based offset of logical (i,j), 1-based interior:
 *     (i + Hx - 1) + (Nx + 2Hx) * (j + Hy - 1)
 *
 * All functions return SWMHD_OK (0) or a negative error code; no exceptions
 * and no callbacks cross the ABI.  A context is not re-entrant.
 * There is NO CPU fallback: without a usable CUDA device swmhd_create fails.
 */
#ifndef SWMHD_ABI_H_
#define SWMHD_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWMHD_ABI_VERSION 2

enum {
    SWMHD_OK              =  0,
    SWMHD_ERR_ARG         = -1,  /* bad argument / unsupported option        */
    SWMHD_ERR_CUDA        = -2,  /* CUDA runtime error (see swmhd_last_error) */
    SWMHD_ERR_NONFINITE   = -3,  /* state contains NaN/Inf                    */
    SWMHD_ERR_NODEVICE    = -4,  /* no CUDA device: there is no CPU fallback  */
    SWMHD_ERR_STATE       = -5,  /* call sequence error                       */
    SWMHD_ERR_NCCL        = -6   /* NCCL error / libnccl.so.2 not loadable    */
};
#define SWMHD_MAX_GPUS 8

enum { SWMHD_PERIODIC = 0, SWMHD_BOUNDED = 1 };

/* formulation = which of the reference's two model set-ups is being run */
enum {
    SWMHD_JACOBIAN   = 0, /* VectorInvariantFormulation + lorentz_force_func_x/y
                             (SWMHD_example.jl:21-33)                          */
    SWMHD_DIVERGENCE = 1  /* ConservativeFormulation + div_lorentz_x/y
                             (divergence_sw_mhd.jl:19-31)                      */
};

/* field ids: solution (u|uh, v|vh, h) followed by the tracer A */
enum { SWMHD_U = 0, SWMHD_V = 1, SWMHD_H = 2, SWMHD_A = 3, SWMHD_NFIELDS = 4 };

/* arithmetic mode of the CUDA kernels */
enum {
    SWMHD_ARITH_FAST   = 0, /* FMA contraction, fused-division WENO weights:
                               <=1e-12 rel-L2 per step vs the strict form    */
    SWMHD_ARITH_STRICT = 1  /* no contraction, IEEE division, operation order
                               of the spec: bit-identical to oracle/         */
};

/* Ambiguity-register switches. Zero = default. */
enum {
    SWMHD_FLAG_WENO_JS        = 1 << 0, /* C1: JS weights instead of Z        */
    SWMHD_FLAG_PRESSURE_GHDH  = 1 << 1, /* C5: g*Ix(h)*dx(h) not dx(g h^2/2)  */
    SWMHD_FLAG_CDIVU_OVER_H   = 1 << 2, /* C6: A*div(uh,vh)/h                 */
    SWMHD_FLAG_DIAG_CENTRED   = 1 << 3, /* C11: centre-averaged squares       */
    /* C10 (Bounded-y wall semantics, SURVEY A.8 confidence L) — probed against the four published
       Bounded-y figures by tools/probe_c10.py, table in DESIGN.md section 2 */
    SWMHD_FLAG_BC_DEPTH1      = 1 << 4, /* no-flux / gradient BCs fill only the first halo row */
    SWMHD_FLAG_WALL_WENO3     = 1 << 5, /* one cell further from the wall than the centred
                                           fallback needs: WENO3 instead of centred 2nd order  */
    SWMHD_FLAG_V_MIRROR       = 1 << 6, /* v|vh beyond the wall: odd mirror instead of untouched */
    /* not in the register: what ELSE could explain the 64^2 Bounded-y deviation (same probe) */
    SWMHD_FLAG_TRACER_CEN2    = 1 << 7, /* A advected with centred 2nd order instead of WENO5     */
    SWMHD_FLAG_TRACER_CEN4    = 1 << 8  /* A advected with centred 4th order instead of WENO5     */
};

typedef struct swmhd_config {
    int32_t abi_version;   /* SWMHD_ABI_VERSION                               */
    int32_t Nx, Ny;        /* global interior size                            */
    int32_t Hx, Hy;        /* must be 3, 3 (WENO5 default halo)               */
    int32_t topo_x, topo_y;/* SWMHD_PERIODIC | SWMHD_BOUNDED (x: periodic only)*/
    int32_t formulation;   /* SWMHD_JACOBIAN | SWMHD_DIVERGENCE               */
    int32_t arith;         /* SWMHD_ARITH_FAST | SWMHD_ARITH_STRICT           */
    int32_t flags;         /* SWMHD_FLAG_*                                    */
    double  dx, dy;        /* Lx/Nx, Ly/Ny                                    */
    double  g, f;          /* gravitational_acceleration, FPlane f            */
    double  weno_eps;      /* C2: 1e-6                                        */
    double  h_ref;         /* h_i in the potential-energy diagnostic (1.0)    */
    /* GradientBoundaryCondition on A at south/north (Bounded-y only),
       SWMHD_example.jl:19 / divergence_sw_mhd.jl:17 (commented in both)     */
    int32_t A_gradient_bc; /* 0 = default no-flux, 1 = gradient               */
    int32_t device;        /* CUDA device ordinal                             */
    double  A_grad_south, A_grad_north;
    /* y-slab decomposition: this context owns global rows
       j in [slab_j0+1, slab_j0+slab_ny]; single GPU: 0, Ny.  The slab is
       stored exactly like a (Nx, slab_ny) Field with Hy=3 halos.             */
    int32_t slab_j0, slab_ny;
    int32_t rank, world;   /* position of the slab in the y ring              */
    /* Single-process multi-GPU: n_gpus > 1 (then world must be 1, slab_j0 = 0,
       slab_ny = Ny).  The context splits the Ny rows into n_gpus y-slabs, one
       per device in device_ids[] (south to north), creates one NCCL
       communicator per device (ncclCommInitAll) and exchanges the halo rows
       itself; host buffers of set/get are the GLOBAL parent arrays.
       n_gpus = 0 or 1: one GPU (`device`).                                   */
    int32_t n_gpus;
    int32_t device_ids[SWMHD_MAX_GPUS];
    int32_t reserved0;
} swmhd_config;

typedef struct swmhd_diag {
    double ke, me, pe, total;   /* SWMHD_example.jl:74-77, divergence_sw_mhd.jl:71-74 */
    double max_abs_u;           /* SWMHD_example.jl:57 (|uh/h| for DIVERGENCE, divergence_sw_mhd.jl:47,53) */
    double max_abs_A, min_h;    /* same lines                                  */
    double max_abs_div_hB;      /* max |dx(hBx)+dy(hBy)| at cell centres        */
    double sum_h;               /* mass, sum over interior                      */
    int32_t all_finite;         /* 0 if any of the four fields has NaN/Inf      */
    int32_t reserved;
} swmhd_diag;

typedef struct swmhd_ctx swmhd_ctx;

/* ShallowWaterModel(grid=..., ...) — SWMHD_example.jl:14-33, divergence_sw_mhd.jl:12-31 */
int  swmhd_create(const swmhd_config *cfg, swmhd_ctx **out);
void swmhd_destroy(swmhd_ctx *ctx);
const char *swmhd_last_error(const swmhd_ctx *ctx); /* ctx may be NULL: last create error */
int  swmhd_abi_version(void);

/* set!(model, u=..., v=..., h=..., A=...) — SWMHD_example.jl:41, divergence_sw_mhd.jl:38.
   host = parent array of the Field (this slab's rows), n = its length in doubles. */
int  swmhd_set_field(swmhd_ctx *ctx, int field, const double *host, size_t n);
/* interior(field)/parent(field) for callbacks and OutputWriters — SWMHD_example.jl:50-52,81-84 */
int  swmhd_get_field(swmhd_ctx *ctx, int field, double *host, size_t n);
size_t swmhd_field_len(const swmhd_ctx *ctx, int field);  /* doubles in the parent array */

/* update_state!(model) = fill_halo_regions! on solution and tracers (after set!) */
int  swmhd_fill_halos(swmhd_ctx *ctx);

/* time_step!(model, dt) x nsteps with RungeKutta3 — SWMHD_example.jl:23,42,97.  Blocking.
   Valid for one GPU, for n_gpus > 1 (single process) and for world > 1 once swmhd_comm_init has
   been called (one process per GPU; every rank makes the same calls): per substage the rows within
   one tile of a slab edge run first on a high-priority stream, their halo rows travel by
   ncclSend/ncclRecv (one group per substage) while the interior runs on the main stream. */
int  swmhd_step(swmhd_ctx *ctx, double dt, int nsteps);
/* same, but diagnostics of the state at the START of each step are produced by
   the stage-1 kernel at no extra HBM traffic and written to diags[0..nsteps) */
int  swmhd_step_diag(swmhd_ctx *ctx, double dt, int nsteps, swmhd_diag *diags);

/* set!(model, u, v, h, A) + time_step!(model, dt) (+ the diagnostics of the uploaded state when diag != NULL) in one call:
   the upload of the four parent arrays (page-locked for real overlap: swmhd_pin_host) is pipelined with stage 1 by row
   bands, so that only the first and last band of stage 1 run after the last byte has crossed the PCIe link.
   host[k] holds n_each >= swmhd_field_len(ctx, k) doubles; halos of the host arrays need not be filled.  Single slab. */
int  swmhd_upload_step(swmhd_ctx *ctx, const double *const host[4], size_t n_each, double dt, swmhd_diag *diag);

/* A batch of steps with a per-step dt: run!(simulation) clips dt to the next TimeInterval(0.1) output and to stop_time
   (upstream aligned_time_step; SWMHD_example.jl:42,81-84,97), so the step sizes between two output times are
   dt, dt, ..., remainder.  The host plans the sequence (the clock ticks (8/15, 2/15, 1/3) dt per stage, replayable in
   floating point) and the device runs it without returning in between.  diags may be NULL; else diags[0..nsteps). */
int  swmhd_step_seq(swmhd_ctx *ctx, const double *dts, int nsteps, swmhd_diag *diags);

/* one RK3 substage (stage = 1,2,3), including the halo fill that follows it */
int  swmhd_substage(swmhd_ctx *ctx, double dt, int stage);
/* calculate_tendencies!(model): G^n of the current state into four host parent arrays, each of
   capacity n_each doubles (>= the longest field: v|vh has one row more in a Bounded-y grid).
   Single slab, between steps only (SWMHD_ERR_STATE after stage 1 or 2 of a step: it overwrites G^-). */
int  swmhd_tendencies(swmhd_ctx *ctx, double *const G_host[4], size_t n_each);

/* the four energy means and the progress-callback reductions —
   SWMHD_example.jl:47-63,67-77; divergence_sw_mhd.jl:42-59,63-75 */
int  swmhd_diagnostics(swmhd_ctx *ctx, swmhd_diag *out);

/* The field writer's outputs (SWMHD_example.jl:67-69,81-84; divergence_sw_mhd.jl:63-66,77-82) computed on
   the device: u, v (velocities; uh/Ix(h), vh/Iy(h) for DIVERGENCE) and s = sqrt(u^2 + v^2) at (Face, Center),
   as parent arrays of u-, v- and u-shaped fields (halos filled).  A is swmhd_get_field(ctx, SWMHD_A).
   Call between steps only (it stages through the tendency buffers). Blocking. */
int  swmhd_get_outputs(swmhd_ctx *ctx, double *u_host, double *v_host, double *s_host);
/* The same writer (JLD2OutputWriter(model, (; u, v, A, s), schedule = TimeInterval(0.1)),
   SWMHD_example.jl:81-84; with_halos = true, divergence_sw_mhd.jl:79) without stalling the step loop:
   u, v, s and A of the CURRENT state are computed into staging buffers owned by the context and
   copied to the four host parent arrays on a copy stream; the call returns once that work is queued
   and swmhd_step may be called right away.  The host arrays are valid after swmhd_outputs_wait.
   For a truly asynchronous copy the host arrays must be page-locked (swmhd_pin_host). */
int  swmhd_get_outputs_async(swmhd_ctx *ctx, double *u_host, double *v_host, double *s_host, double *A_host);
int  swmhd_outputs_wait(swmhd_ctx *ctx);
int  swmhd_pin_host(void *ptr, size_t bytes);     /* cudaHostRegister / cudaHostUnregister */
int  swmhd_unpin_host(void *ptr);

/* model.clock: time and iteration as advanced by the RK3 stages */
double  swmhd_time(const swmhd_ctx *ctx);
int64_t swmhd_iteration(const swmhd_ctx *ctx);
int  swmhd_set_clock(swmhd_ctx *ctx, double time, int64_t iteration);

/* ---- y-slabs, one process per GPU (world > 1) ----------------------------------------------
   Rank 0 obtains an NCCL unique id, the host side distributes it (torch.distributed / MPI
   broadcast: plumbing), and every rank hands it to its context: swmhd_comm_init is collective
   (ncclCommInitRank(world, id, rank)).  From then on swmhd_fill_halos, swmhd_step[_diag],
   swmhd_substage and swmhd_diagnostics run the halo exchange and the cross-rank reductions
   inside the library.  libnccl.so.2 is dlopen'ed on first use (SWMHD_NCCL_LIB overrides the
   name), so a single-GPU user does not need NCCL. */
#define SWMHD_COMM_ID_BYTES 128
int  swmhd_comm_unique_id(void *id, size_t nbytes);                 /* nbytes >= SWMHD_COMM_ID_BYTES */
int  swmhd_comm_init(swmhd_ctx *ctx, const void *id, size_t nbytes);
/* (first row, rows) of slab `index` of `nslabs` for a grid of Ny rows: the decomposition used by n_gpus > 1 */
int  swmhd_split_rows(int Ny, int nslabs, int index, int *j0, int *ny);

/* ---- host-driven exchange (legacy: a host that owns its own transport, e.g. MPI in Julia) ----
   Per substage:  swmhd_substage_edges -> exchange rows -> swmhd_substage_interior ->
   swmhd_substage_finish.  Contexts with world > 1 and no swmhd_comm_init only.              */
int  swmhd_set_streams(swmhd_ctx *ctx, void *main_stream, void *edge_stream);
int  swmhd_substage_edges(swmhd_ctx *ctx, double dt, int stage);    /* rows within 3 of a slab edge  */
int  swmhd_substage_interior(swmhd_ctx *ctx, double dt, int stage); /* the rest, on the main stream   */
int  swmhd_substage_finish(swmhd_ctx *ctx, int stage);              /* swap buffers, tick clock        */
/* device pointer to the first of `nrows` contiguous parent rows of the NEW
   state of `field`: which = 0 south send rows, 1 north send rows,
   2 south halo (recv), 3 north halo (recv).  *row_doubles = Nx+6.           */
int  swmhd_exchange_rows(swmhd_ctx *ctx, int field, int which,
                         void **dev_ptr, int *nrows, size_t *row_doubles);
int  swmhd_sync(swmhd_ctx *ctx);
/* Test hook for pools without compute-sanitizer: with SWMHD_GUARD=1 in the environment at swmhd_create every device array
   sits between two 32 KB guard zones filled with a sentinel; SWMHD_OK = no kernel wrote outside its arrays so far. */
int  swmhd_check_guards(swmhd_ctx *ctx);
/* Slab diagnostics without a host round trip per step: arm a slot (0..1023) and the next stage-1
   substage (edges + interior) also evaluates the diagnostics of the state it starts from, fused in
   the substage kernels; read any number of slots later.  Values are slab partials scaled by the
   GLOBAL normalisation: sums add over ranks, max/min combine. */
int  swmhd_arm_diag(swmhd_ctx *ctx, int slot);
int  swmhd_get_diag_slots(swmhd_ctx *ctx, int first, int count, swmhd_diag *out);

/* bookkeeping for bench.py: kernels launched since create, and device-side
   duration of the last swmhd_step call measured with CUDA events (ms).     */
/* nsteps RK3 steps with a CUDA-event pair around every fused substage-kernel launch;
   out_ms[s] = mean device time of the stage-(s+1) kernel (roofline measurement). */
int  swmhd_step_profile(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3]);
/* same with the diagnostics fused into every stage-1 launch (what swmhd_step_diag runs) */
int  swmhd_step_profile_diag(swmhd_ctx *ctx, double dt, int nsteps, double out_ms[3]);
int64_t swmhd_launch_count(const swmhd_ctx *ctx);
double  swmhd_last_step_ms(const swmhd_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SWMHD_ABI_H_ */
