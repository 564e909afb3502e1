// Internal launch-parameter block shared by the CUDA translation units.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda.h>
#include <cuda_runtime.h>

namespace swmhd {

// RK3 constants of Oceananigans' RungeKutta3TimeStepper
constexpr double RK_GAMMA[3] = {8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0};
constexpr double RK_ZETA[3]  = {0.0, -17.0 / 60.0, -5.0 / 12.0};

struct KParams {
    int Nx, Ny;          // local interior size (Ny = rows of this y-slab)
    int P;               // row pitch of every parent array = Nx + 6
    int gj0, NyG;        // global row of local row j is gj0 + j; global Ny
    int by;              // topo_y == Bounded
    int tile_row0, tile_rows; // first tile-row and number of tile-rows of this launch
    int tiles_per_cta;        // consecutive tiles one CTA processes (set by the launcher)
    int rows[4];         // parent rows per field (Ny+6, v: +1 when Bounded-y)
    double dx, dy, rdx, rdy, inv_az, g, f, eps;
    double dt, gam, zet, dtgam;  // stage coefficients; dtgam = dt*gam (stage 1)
    double qrdxy;                // 0.25 / (dx dy): common factor of the telescoped Jacobian Lorentz force
    double dtzet;                // dt*zet (FAST arithmetic: U + (dtgam*Gn + dtzet*Gm) as two FMAs)
    double h_ref;                // h_i of the potential-energy diagnostic
    const double *Uo[4]; // state at the start of the substage (halos valid)
    double *Un[4];       // state after the substage (interior written)
    double *G[4];        // G^- on entry (stages 2,3), G^n on exit (stages 1,2)
    int use_tma;         // 1: tm[] are valid tensor maps of Uo[] (TMA tile loads)
    alignas(64) CUtensorMap tm[4];
    int use_rb;          // 1: tm_rb[] are valid (row-blocked kernel, substage_rb.cu: taller TMA box)
    int l2_ahead;        // L2 prefetch distance in tiles (CTAs in flight)
    int row_begin, row_end;   // 0-based cell rows [row_begin, row_end) of this launch (set by the rb launcher)
    alignas(64) CUtensorMap tm_rb[4];
    double *diag;        // per-CTA diagnostic partials [tiles][NDIAG] (stage-1 DIAG variant), or nullptr
};

struct HaloParams {
    int Nx, Ny, P;
    int by, first, last;   // Bounded-y; this slab touches the south / north wall
    int y_mode;            // 0: no y fill, 1: periodic wrap owned by this context, 2: wall BCs
    int grad;              // gradient BC on A
    double gs, gn, dy;     // A_grad_south, A_grad_north, dy
    int j_lo, j_hi;        // x-wrap parent rows (inclusive; empty if j_hi < j_lo)
    double *U[4];
    int rows[4];
};

struct DiagParams {
    int Nx, Ny, P, form;
    double dx, dy, g, h_ref;
    const double *U[4];
    double *partials;      // [nblocks][NDIAG]
    double *stage;         // [diag_stage_doubles()] scratch of the two-level final reduction
    int nblocks;
};

struct OutputParams {
    int Nx, Ny, P, form;
    const double *U[4];
    double *out_u, *out_v, *out_s;
};

constexpr int NDIAG = 9; // sums: ke, me, pe, sum_h [0..3]; maxima: |u|, |A|, -h, |div hB| [4..7]; nonfinite count [8]

// kernel launchers (one strict + one fast instantiation of substage_kernel.cu)
cudaError_t launch_substage_strict(const KParams &p, int form, int stage, cudaStream_t st);
cudaError_t launch_substage_fast(const KParams &p, int form, int stage, cudaStream_t st);
void substage_tile(int *tx, int *ty);
// row-blocked kernels (substage_rb.cu; FAST arithmetic, TMA): stages 1..3 of either formulation
cudaError_t launch_substage_rb(const KParams &p, int form, int stage, cudaStream_t st);
void substage_rb_tile(int *tx, int *ty);
int substage_rb_tiles_x(int form, int Nx);    // tiles per tile row (the divergence kernel owns 31 cells per tile row)
int substage_rb_stage_mask();                 // SWMHD_RB_STAGES: bit s-1 = stage s runs the row-blocked kernel (default 7)
cudaError_t launch_halo(const HaloParams &p, cudaStream_t st);
cudaError_t launch_diag(const DiagParams &p, double *out9, cudaStream_t st);
cudaError_t launch_diag_final(const double *partials, int nblocks, double *stage, double *out9, cudaStream_t st);
int diag_stage_doubles();
int diag_blocks(int Nx, int Ny);
cudaError_t launch_output(const OutputParams &p, cudaStream_t st);

} // namespace swmhd
