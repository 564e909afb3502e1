// substage_kernel.cu — the fused RK3-substage kernel of the SWMHD hot path (sm_100a).
//
// One launch = calculate_tendencies! + rk3_substep! + store_tendencies! of one
// substage for all four prognostic fields, with the reference's
// Lorentz-force closures inlined:
//   FORM 0  VectorInvariantFormulation + jacobian_formulation/sw_mhd_jacobian_functions.jl:1-26
//   FORM 1  ConservativeFormulation    + divergence_formulation/sw_mhd_divergence_functions.jl:1-170
// and, in the stage-1 DIAG variant, the energy / max / min / div(hB) diagnostics of
// SWMHD_example.jl:47-77 evaluated on the tile that is already staged (0 extra HBM bytes).
//
// This file is compiled twice:
//   -DSWMHD_STRICT=1 -fmad=false : operation order, IEEE divisions and rounding of the
//                                  specification -> bit-identical to oracle/swmhd_oracle.c
//   -DSWMHD_STRICT=0             : explicit FMAs, one-division WENO-Z weights, telescoped
//                                  difference-of-average operators, Newton reciprocals
//
// Tile scheme: a CTA owns TX x TY cells.
//   P0  stage u,v,h,A with a 3-cell halo in shared memory
//   P1  one balanced task list: derived staggered fields (zeta, velocity-stencil
//       averages, K, Bx, By | hBx, hBy, Bx, By, face h) and every face flux exactly once
//       (upwind side selected by sign: bit-identical to upwind_biased_product for finite data)
//   P2  per cell: difference the fluxes, vorticity advection, pressure, Coriolis,
//       Lorentz force, U += dt(gamma G^n + zeta G^-), store U_new and G^n.
// HBM traffic per cell and substage: 4-8 reads + 4-8 writes of doubles (algorithmic).
#include "kparams.h"
#include "device_prims.cuh"
#include <cstdlib>

#ifndef SWMHD_STRICT
#error "compile with -DSWMHD_STRICT=0|1"
#endif

namespace swmhd {
namespace {

#if SWMHD_STRICT
#define LAUNCH_NAME launch_substage_strict
#else
#define LAUNCH_NAME launch_substage_fast
#endif

#ifndef SWMHD_TPC_DEFAULT
#define SWMHD_TPC_DEFAULT 4
#endif
#ifndef SWMHD_MINB
#define SWMHD_MINB 3
#endif
#ifndef SWMHD_TX
#define SWMHD_TX 32
#define SWMHD_TY 8
#endif
constexpr int TX = SWMHD_TX, TY = SWMHD_TY, NT = TX * TY;
constexpr int W = TX + 6, HT = TY + 6, SZ = W * HT;        // raw tiles: [HT][W] (dense TMA box)
constexpr int SZP = (SZ * 8 + 127) / 128 * 16;             // raw tile padded to a multiple of 128 B (TMA dst alignment)
constexpr unsigned TILE_TX_BYTES = 4u * SZ * 8u;           // bytes one tile load brings in (4 fields)

// compact derived / flux arrays: (pitch, rows)
constexpr int ZP = TX + 5, ZR = TY + 5;     // ffc points a in [1,TX+5], b in [1,TY+5]
constexpr int CP = TX + 2, CR = TY + 2;     // ccc points a in [2,TX+3], b in [2,TY+3]
constexpr int XP = TX + 1, XR = TY;         // x-face-like a in [a0,a0+TX], b in [3,TY+2]
constexpr int YP = TX,     YR = TY + 1;     // y-face-like a in [3,TX+2],  b in [b0,b0+TY]
constexpr int BP = TX + 4, BR = TY + 4;     // divergence-form B fields a in [1,TX+4], b in [1,TY+4]

// ---------------------------------------------------------------------------
// arithmetic primitives
#if SWMHD_STRICT
__device__ __forceinline__ double fdiv(double a, double b) { return a / b; }
#define DIVDX(x) ((x) / p.dx)
#define DIVDY(x) ((x) / p.dy)
#define DIVAZ(x) ((x) / (p.dx * p.dy))
#define FXS p.dy
#define FYS p.dx
#else
__device__ __forceinline__ double fdiv(double a, double b) { return a * frcp(b); }
#define DIVDX(x) ((x) * p.rdx)
#define DIVDY(x) ((x) * p.rdy)
#define DIVAZ(x) ((x) * p.inv_az)
#define FXS 1.0
#define FYS 1.0
#endif

// ---- WENO5-Z.  a..e = psi[f-3..f+1] in upwind ("left") orientation ------------
#if SWMHD_STRICT
__device__ __forceinline__ void weno_beta(double a, double b, double c, double d, double e,
                                          double &b0, double &b1, double &b2) {
    double D0 = c - 2.0 * d + e, E0 = 3.0 * c - 4.0 * d + e;
    double D1 = b - 2.0 * c + d, E1 = b - d;
    double D2 = a - 2.0 * b + c, E2 = a - 4.0 * b + 3.0 * c;
    b0 = (13.0 / 12.0) * (D0 * D0) + 0.25 * (E0 * E0);
    b1 = (13.0 / 12.0) * (D1 * D1) + 0.25 * (E1 * E1);
    b2 = (13.0 / 12.0) * (D2 * D2) + 0.25 * (E2 * E2);
}
__device__ __forceinline__ double weno_blend(double a, double b, double c, double d, double e,
                                             double b0, double b1, double b2, double eps) {
    double p0 = (2.0 * c + 5.0 * d - e) / 6.0;
    double p1 = (-b + 5.0 * c + 2.0 * d) / 6.0;
    double p2 = (2.0 * a - 7.0 * b + 11.0 * c) / 6.0;
    double tau = fabs(b2 - b0);
    double r0 = tau / (b0 + eps), r1 = tau / (b1 + eps), r2 = tau / (b2 + eps);
    double a0 = 0.3 * (1.0 + r0 * r0);
    double a1 = 0.6 * (1.0 + r1 * r1);
    double a2 = 0.1 * (1.0 + r2 * r2);
    double sum = a0 + a1 + a2;
    double w0 = a0 / sum, w1 = a1 / sum, w2 = a2 / sum;
    return w0 * p0 + w1 * p1 + w2 * p2;
}
__device__ __forceinline__ double weno5(double a, double b, double c, double d, double e, double eps) {
    double b0, b1, b2;
    weno_beta(a, b, c, d, e, b0, b1, b2);
    return weno_blend(a, b, c, d, e, b0, b1, b2, eps);
}
// zeta with VelocityStencil smoothness: beta_k = (beta_k[ℑy u] + beta_k[ℑx v]) / 2
__device__ __forceinline__ double weno5_vs(const double *qz, const double *qu, const double *qv, int s, double eps) {
    double bu0, bu1, bu2, bv0, bv1, bv2;
    weno_beta(qu[0], qu[s], qu[2 * s], qu[3 * s], qu[4 * s], bu0, bu1, bu2);
    weno_beta(qv[0], qv[s], qv[2 * s], qv[3 * s], qv[4 * s], bv0, bv1, bv2);
    return weno_blend(qz[0], qz[s], qz[2 * s], qz[3 * s], qz[4 * s],
                      0.5 * (bu0 + bv0), 0.5 * (bu1 + bv1), 0.5 * (bu2 + bv2), eps);
}
#else
// FAST WENO5-Z in difference form (device_prims.cuh).
struct WenoDiff { double c, d1, d2, d3, d4; };
__device__ __forceinline__ WenoDiff weno_diffs(double a, double b, double c, double d, double e) {
    WenoDiff w; w.c = c; w.d1 = b - a; w.d2 = c - b; w.d3 = d - c; w.d4 = e - d; return w;
}
__device__ __forceinline__ void weno_beta_acc(const WenoDiff &w, double &c0, double &c1, double &c2) {
    weno_beta_acc4(w.d1, w.d2, w.d3, w.d4, c0, c1, c2);
}
__device__ __forceinline__ double weno_blend_c(const WenoDiff &w, double c0, double c1, double c2) {
    double num, den;
    weno_corr4(w.d1, w.d2, w.d3, w.d4, c0, c1, c2, num, den);
    return fma(num, frcp(den), w.c);
}
__device__ __forceinline__ double weno5(double a, double b, double c, double d, double e, double eps) {
    const WenoDiff w = weno_diffs(a, b, c, d, e);
    const double es = eps * (12.0 / 13.0);
    double c0 = es, c1 = es, c2 = es;
    weno_beta_acc(w, c0, c1, c2);
    return weno_blend_c(w, c0, c1, c2);
}
__device__ __forceinline__ double weno5_vs(const double *qz, const double *qu, const double *qv, int s, double eps) {
    const double es = eps * (24.0 / 13.0);      // beta = (beta_u + beta_v)/2: common scale 13/24
    double c0 = es, c1 = es, c2 = es;
    weno_beta_acc(weno_diffs(qu[0], qu[s], qu[2 * s], qu[3 * s], qu[4 * s]), c0, c1, c2);
    weno_beta_acc(weno_diffs(qv[0], qv[s], qv[2 * s], qv[3 * s], qv[4 * s]), c0, c1, c2);
    return weno_blend_c(weno_diffs(qz[0], qz[s], qz[2 * s], qz[3 * s], qz[4 * s]), c0, c1, c2);
}
#endif

__device__ __forceinline__ double sym4(double a, double b, double c, double d) {
#if SWMHD_STRICT
    return (7.0 * (b + c) - (a + d)) / 12.0;
#else
    return fma(7.0 / 12.0, b + c, (-1.0 / 12.0) * (a + d));
#endif
}
__device__ __forceinline__ double sym2(double b, double c) { return 0.5 * (b + c); }

// third-order biased interpolants of sw_mhd_divergence_functions.jl:25-35
__device__ __forceinline__ double third(double x2, double x5, double xm) { // (2*x2 + 5*x5 - xm)/6
#if SWMHD_STRICT
    return (2.0 * x2 + 5.0 * x5 - xm) / 6.0;
#else
    return fma(1.0 / 3.0, x2, fma(5.0 / 6.0, x5, (-1.0 / 6.0) * xm));
#endif
}
__device__ __forceinline__ double thirdR(double xm, double x5, double x2) { // (-xm + 5*x5 + 2*x2)/6
#if SWMHD_STRICT
    return (-xm + 5.0 * x5 + 2.0 * x2) / 6.0;
#else
    return fma(-1.0 / 6.0, xm, fma(5.0 / 6.0, x5, (1.0 / 3.0) * x2));
#endif
}

// ℑxy of four corner values, 0.5*(0.5*(a+b) + 0.5*(c+d))
__device__ __forceinline__ double avg4(double a, double b, double c, double d) {
#if SWMHD_STRICT
    return 0.5 * (0.5 * (a + b) + 0.5 * (c + d));
#else
    return 0.25 * ((a + b) + (c + d));
#endif
}

// Bounded-y wall buffer (oracle ybuf): footprint f-n..f+n-1 must stay in [1,hi]
__device__ __forceinline__ bool ybuf(int by, int f, int n, int hi) { return by && (f - n < 1 || f + n - 1 > hi); }

// The five samples of the upwind-biased WENO5 stencil of face f along a line with element
// stride st: left-biased (vel > 0) psi[f-3..f+1], right-biased the mirror psi[f+2..f-2].
// ctr points at psi[f]; returns vel * psi_upwind (== upwind_biased_product, other side * 0).
__device__ __forceinline__ double upwind_weno(const double *ctr, int st, double vel, double eps) {
    const bool pos = vel > 0.0;
    const double *q = pos ? ctr - 3 * st : ctr + 2 * st;
    const int s = pos ? st : -st;
    return vel * weno5(q[0], q[s], q[2 * s], q[3 * s], q[4 * s], eps);
}
// Bounded-y wall buffer variant: centred 2nd order instead of WENO5 (selected, not branched, so that
// several reconstructions of one thread stay in one basic block and their FP64 chains interleave)
__device__ __forceinline__ double upwind_weno_buf(const double *ctr, int st, double vel, double eps, bool buf) {
    const double w = upwind_weno(ctr, st, vel, eps);
    const double c2 = vel * sym2(ctr[-st], ctr[0]);
    return buf ? c2 : w;
}
__device__ __forceinline__ double upwind_sel(double vel, double L, double R) {
    return vel * (vel > 0.0 ? L : R);
}

#define RAW(arr, a, b) arr[(b) * W + (a)]

// ---------------------------------------------------------------------------
constexpr int NDG_ = YP * YR + XP * (TY + 2);   // diag scratch: sqBx (YP*YR) + sqBy (XP*(TY+2))
template <int FORM, bool DIAG> struct SmemLayout;
template <bool DIAG> struct SmemLayout<0, DIAG> {
    static constexpr int o_z = 0, o_ut = o_z + ZP * ZR, o_vt = o_ut + ZP * ZR;
    static constexpr int o_K = o_vt + ZP * ZR, o_Bx = o_K + CP * CR, o_By = o_Bx + CP * CR;
    static constexpr int o_Fxh = o_By + CP * CR, o_FxA = o_Fxh + XP * XR;
    static constexpr int o_Fyh = o_FxA + XP * XR, o_FyA = o_Fyh + YP * YR;
    static constexpr int o_dg = o_FyA + YP * YR;
    static constexpr int total = o_dg + (DIAG ? NDG_ : 0);
};
template <bool DIAG> struct SmemLayout<1, DIAG> {
    static constexpr int o_hBx = 0, o_hBy = o_hBx + BP * BR, o_Bx = o_hBy + BP * BR, o_By = o_Bx + BP * BR;
    static constexpr int o_hx = o_By + BP * BR, o_hy = o_hx + BP * BR, o_hff = o_hy + BP * BR;
    static constexpr int o_Fuu = o_hff + XP * YR, o_Lxx = o_Fuu + XP * XR, o_Fuv = o_Lxx + XP * XR, o_Lxy = o_Fuv + XP * XR;
    static constexpr int o_Tx = o_Lxy + XP * XR, o_uq = o_Tx + XP * XR;
    static constexpr int o_Fvu = o_uq + XP * XR, o_Lyx = o_Fvu + YP * YR, o_Fvv = o_Lyx + YP * YR, o_Lyy = o_Fvv + YP * YR;
    static constexpr int o_Ty = o_Lyy + YP * YR, o_vq = o_Ty + YP * YR;
    static constexpr int o_dg = o_vq + YP * YR;
    static constexpr int total = o_dg + (DIAG ? NDG_ : 0);
};

// STAGE 1,2,3; 0 = tendencies only (G written, U untouched).  DIAG only with STAGE 1.
// Persistent: gridDim.x CTAs loop over the tiles of the launch (x fastest, so concurrently
// resident CTAs share their halo columns/rows through L2).  NSTG = 2 double-buffers the raw
// tile: the TMA load of tile n+1 overlaps the arithmetic of tile n.  TMA = false is the
// plain-load path (odd Nx: the row pitch is not a multiple of 16 B).
template <int FORM, int STAGE, bool DIAG, bool TMA, int NSTG>
__global__ void __launch_bounds__(NT, SWMHD_MINB) substage_kernel(const __grid_constant__ KParams p) {
    // 128-byte aligned for the TMA destination; used directly (no integer round-up) so that the compiler
    // keeps the shared address space and emits LDS/STS instead of generic LD/ST.
    extern __shared__ __align__(128) unsigned char smem_bytes[];
    using L = SmemLayout<FORM, DIAG>;
    double *const raw0 = reinterpret_cast<double *>(smem_bytes);
    double *const smem = raw0 + NSTG * 4 * SZP;             // derived / flux arrays
    uint64_t *const mbar = reinterpret_cast<uint64_t *>(smem + L::total);
    const int tid = threadIdx.x;
    const int Nx = p.Nx, Ny = p.Ny, P = p.P;
    const double eps = p.eps;
    const int tiles_x = (Nx + TX - 1) / TX;
    const int ntiles = tiles_x * p.tile_rows;

    if constexpr (TMA) {
        if (tid == 0) {
            for (int s = 0; s < NSTG; s++) mbar_init(&mbar[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    auto issue_load = [&](int tile, int stage) {            // one elected thread
        const int c0 = (tile % tiles_x) * TX;                // parent column of local a = 0
        const int c1 = (p.tile_row0 + tile / tiles_x) * TY;  // parent row of local b = 0
        double *dst = raw0 + stage * 4 * SZP;
        mbar_expect_tx(&mbar[stage], TILE_TX_BYTES);
#pragma unroll
        for (int k = 0; k < 4; k++) tma_load_2d(dst + k * SZP, &p.tm[k], c0, c1, &mbar[stage]);
    };
    // A CTA owns `tpc` consecutive tiles [tile_lo, tile_hi); CTAs are launched dynamically by the
    // hardware, which keeps the resident CTAs of an SM out of phase (load / FP64 / store overlap).
    const int tpc = p.tiles_per_cta;
    const int tile_lo = blockIdx.x * tpc;
    const int tile_hi = min(ntiles, tile_lo + tpc);
    if constexpr (TMA && NSTG == 2) {
        if (tid == 0 && tile_lo < tile_hi) issue_load(tile_lo, 0);
    }
    if constexpr (TMA) {
        // pull the first tile of the CTA that will take over this slot into L2 now (CTAs are dispatched in
        // blockIdx order): its TMA wait then costs an L2 hit instead of an HBM round trip
        if (tid == 0 && p.l2_ahead > 0) {
            const int nxt = (blockIdx.x + p.l2_ahead) * tpc;
            if (nxt < ntiles) {
                const int c0 = (nxt % tiles_x) * TX, c1 = (p.tile_row0 + nxt / tiles_x) * TY;
#pragma unroll
                for (int k = 0; k < 4; k++) tma_prefetch_2d(&p.tm[k], c0, c1);
            }
        }
    }

    // own cell of this thread (tile-local)
    const int tx = tid % TX, ty = tid / TX;
    const int li = tx + 3, lj = ty + 3;
    double *s_sqBx = smem + L::o_dg, *s_sqBy = s_sqBx + YP * YR;   // DIAG only
    (void)s_sqBx; (void)s_sqBy;

  int it = 0;
  for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
    const int stg = (NSTG == 2) ? (it & 1) : 0;
    double *s_u = raw0 + stg * 4 * SZP, *s_v = s_u + SZP, *s_h = s_u + 2 * SZP, *s_A = s_u + 3 * SZP;
    const int i0 = (tile % tiles_x) * TX + 1;               // logical (1-based) first cell
    const int j0 = (p.tile_row0 + tile / tiles_x) * TY + 1;
    const int i = i0 + tx, j = j0 + ty;
    const bool active = (i <= Nx) && (j <= Ny);
    const int gj = p.gj0 + j;                               // global row (wall logic)
    const size_t gcell = (size_t)(i + 2) + (size_t)P * (size_t)(j + 2);

    // ---- P0: stage the four fields with a 3-cell halo -------------------------------
    if constexpr (TMA) {
        if (tid == 0) {
            if constexpr (NSTG == 2) {
                const int nxt = tile + 1;                    // prefetch: overlaps this tile's arithmetic
                if (nxt < tile_hi) issue_load(nxt, stg ^ 1);
            } else {
                issue_load(tile, 0);
            }
        }
    }
    // G^- of the own cell: issue the loads now, consume them after the tendencies
    double Gm0 = 0.0, Gm1 = 0.0, Gm2 = 0.0, Gm3 = 0.0;
    if constexpr (STAGE >= 2) {
        if (active) { Gm0 = p.G[0][gcell]; Gm1 = p.G[1][gcell]; Gm2 = p.G[2][gcell]; Gm3 = p.G[3][gcell]; }
    }
    if constexpr (TMA) {
        mbar_wait(&mbar[stg], (NSTG == 2) ? ((it >> 1) & 1) : (it & 1));
    } else {
        for (int t = tid; t < SZ; t += NT) {
            int a = t % W, b = t / W;
            int pi = i0 - 1 + a, pj = j0 - 1 + b;           // parent (0-based) column / row
            bool okx = pi < P;
            size_t g = (size_t)pi + (size_t)P * (size_t)pj;
            s_u[t] = (okx && pj < p.rows[0]) ? p.Uo[0][g] : 0.0;
            s_v[t] = (okx && pj < p.rows[1]) ? p.Uo[1][g] : 0.0;
            s_h[t] = (okx && pj < p.rows[2]) ? p.Uo[2][g] : 1.0;
            s_A[t] = (okx && pj < p.rows[3]) ? p.Uo[3][g] : 0.0;
        }
        __syncthreads();
    }

    double Gn0 = 0.0, Gn1 = 0.0, Gn2 = 0.0, Gn3 = 0.0;

    if constexpr (FORM == 0) {
        // ================= VectorInvariant + Jacobian Lorentz ======================
        double *s_z = smem + L::o_z, *s_ut = smem + L::o_ut, *s_vt = smem + L::o_vt;
        double *s_K = smem + L::o_K, *s_Bx = smem + L::o_Bx, *s_By = smem + L::o_By;
        double *s_Fxh = smem + L::o_Fxh, *s_FxA = smem + L::o_FxA, *s_Fyh = smem + L::o_Fyh, *s_FyA = smem + L::o_FyA;
#define Zf(arr, a, b) arr[((b) - 1) * ZP + (a) - 1]
#define Cc(arr, a, b) arr[((b) - 2) * CP + (a) - 2]
#define FX(arr, a, b) arr[((b) - 3) * XP + (a) - 3]
#define FY(arr, a, b) arr[((b) - 3) * YP + (a) - 3]

        // ---- A: light derived fields (zeta, velocity-stencil averages, K, Bx, By [, diag B^2]) ----
        constexpr int NZ = ZP * ZR, NC = CP * CR;
        constexpr int NDG = DIAG ? NDG_ : 0;
        auto task_zeta = [&](int q) {                               // zeta, ℑy u, ℑx v at ffc
            int a = 1 + q % ZP, b = 1 + q / ZP;
            double vc = RAW(s_v, a, b), vw = RAW(s_v, a - 1, b), uc = RAW(s_u, a, b), us = RAW(s_u, a, b - 1);
#if SWMHD_STRICT
            Zf(s_z, a, b) = DIVAZ((p.dy * vc - p.dy * vw) - (p.dx * uc - p.dx * us));
#else
            Zf(s_z, a, b) = fma(vc - vw, p.rdx, (us - uc) * p.rdy);
#endif
            Zf(s_ut, a, b) = 0.5 * (us + uc);
            Zf(s_vt, a, b) = 0.5 * (vw + vc);
        };
        auto task_ccc = [&](int q) {                                // K, Bx, By at ccc
            int a = 2 + q % CP, b = 2 + q / CP;
            double u0 = RAW(s_u, a, b), u1 = RAW(s_u, a + 1, b), v0 = RAW(s_v, a, b), v1 = RAW(s_v, a, b + 1);
            double hc = RAW(s_h, a, b);
#if SWMHD_STRICT
            Cc(s_K, a, b) = (0.5 * (u0 * u0 + u1 * u1) + 0.5 * (v0 * v0 + v1 * v1)) / 2.0;
            double Ac = RAW(s_A, a, b);
            double dyA0 = DIVDY(Ac - RAW(s_A, a, b - 1)), dyA1 = DIVDY(RAW(s_A, a, b + 1) - Ac);
            double dxA0 = DIVDX(Ac - RAW(s_A, a - 1, b)), dxA1 = DIVDX(RAW(s_A, a + 1, b) - Ac);
            Cc(s_Bx, a, b) = -(0.5 * (dyA0 + dyA1)) / hc;           // sw_mhd_jacobian_functions.jl:5-7
            Cc(s_By, a, b) = (0.5 * (dxA0 + dxA1)) / hc;            // :1-3
#else
            Cc(s_K, a, b) = 0.25 * (fma(u0, u0, u1 * u1) + fma(v0, v0, v1 * v1));
            double rh = frcp(hc);
            Cc(s_Bx, a, b) = ((RAW(s_A, a, b - 1) - RAW(s_A, a, b + 1)) * (0.5 * p.rdy)) * rh;
            Cc(s_By, a, b) = ((RAW(s_A, a + 1, b) - RAW(s_A, a - 1, b)) * (0.5 * p.rdx)) * rh;
#endif
        };
        // fixed passes (NZ = 481 and NC = 340 points over 256 threads): no dispatch loop, no mixed warps
        static_assert(NZ <= 2 * NT && NC <= 2 * NT, "phase A assumes at most two passes per list");
        task_zeta(tid);
        task_ccc(tid);
        if (tid + NT < NZ) task_zeta(tid + NT);
        if (tid + NT < NC) task_ccc(tid + NT);
        if constexpr (DIAG) {                                       // diagnostic B^2 at faces
            for (int q = tid; q < NDG; q += NT) {
                if (q < YP * YR) {                                  // (dyA / ℑy h)^2 at cfc, a in [3,TX+2], b in [3,TY+3]
                    int a = 3 + q % YP, b = 3 + q / YP;
                    double bx = fdiv(-DIVDY(RAW(s_A, a, b) - RAW(s_A, a, b - 1)), 0.5 * (RAW(s_h, a, b - 1) + RAW(s_h, a, b)));
                    s_sqBx[q] = bx * bx;
                } else {                                            // (dxA / ℑx h)^2 at fcc, a in [3,TX+3], b in [2,TY+3]
                    int r = q - YP * YR;
                    int a = 3 + r % XP, b = 2 + r / XP;
                    double by_ = fdiv(DIVDX(RAW(s_A, a, b) - RAW(s_A, a - 1, b)), 0.5 * (RAW(s_h, a - 1, b) + RAW(s_h, a, b)));
                    s_sqBy[r] = by_ * by_;
                }
            }
        }
        __syncthreads();

        // ---- B: all WENO5 work.  Every thread evaluates, in one straight-line block, the six
        // independent reconstructions of its own cell (vorticity along y and x with VelocityStencil
        // smoothness; h and A at the west and the south face), so their FP64 dependency chains
        // interleave.  Warps 0/1 add the tile's north row / east column of faces.
        double adv_u, adv_v, vhat, uhat;
        double own_u, own_v, own_fxh, own_fxA, own_fyh, own_fyA;   // carried in registers into phase C
        {
            vhat = avg4(RAW(s_v, li - 1, lj), RAW(s_v, li, lj), RAW(s_v, li - 1, lj + 1), RAW(s_v, li, lj + 1));
            uhat = avg4(RAW(s_u, li, lj - 1), RAW(s_u, li + 1, lj - 1), RAW(s_u, li, lj), RAW(s_u, li + 1, lj));
            const bool posv = vhat > 0.0, posu = uhat > 0.0;
            const int lfy = lj + 1, lfx = li + 1;                   // zeta to the centre j (along y) / i (along x)
            const int offy = ((posv ? lfy - 3 : lfy + 2) - 1) * ZP + li - 1;
            const int offx = (lj - 1) * ZP + (posu ? lfx - 3 : lfx + 2) - 1;
            const double zy = weno5_vs(s_z + offy, s_ut + offy, s_vt + offy, posv ? ZP : -ZP, eps);
            const double zx = weno5_vs(s_z + offx, s_ut + offx, s_vt + offx, posu ? 1 : -1, eps);
            const double z2 = sym2(Zf(s_z, li, lfy - 1), Zf(s_z, li, lfy));
            adv_u = vhat * (ybuf(p.by, gj + 1, 3, p.NyG + 1) ? z2 : zy);
            adv_v = uhat * zx;
            // west (fcc) and south (cfc) faces of the own cell: mass and tracer fluxes
            const double uw = RAW(s_u, li, lj), vs = RAW(s_v, li, lj);
            const bool buf = ybuf(p.by, gj, 3, p.NyG);
            const double fxh = upwind_weno(&RAW(s_h, li, lj), 1, uw, eps);
            const double fxA = upwind_weno(&RAW(s_A, li, lj), 1, uw, eps);
            const double fyh = upwind_weno_buf(&RAW(s_h, li, lj), W, vs, eps, buf);
            const double fyA = upwind_weno_buf(&RAW(s_A, li, lj), W, vs, eps, buf);
            // STRICT stores the fluxes Ax*u*c, Ay*v*c of the spec; FAST stores u*c, v*c and folds the
            // metric factors (Ax/Az = 1/dx, Ay/Az = 1/dy) into the divergence.
            own_u = uw; own_v = vs;
            own_fxh = FXS * fxh; own_fxA = FXS * fxA; own_fyh = FYS * fyh; own_fyA = FYS * fyA;
            FX(s_Fxh, li, lj) = own_fxh; FX(s_FxA, li, lj) = own_fxA;
            FY(s_Fyh, li, lj) = own_fyh; FY(s_FyA, li, lj) = own_fyA;
        }
        static_assert(TX <= 32 && TY <= 32, "leftover faces: north row on warp 0, east column on warp 1");
        if (tid < TX) {                                             // north row of y-faces, b = TY+3
            const int a = 3 + tid, b = TY + 3;
            const double vs = RAW(s_v, a, b);
            const bool buf = ybuf(p.by, p.gj0 + j0 + TY, 3, p.NyG);
            const double fyh = upwind_weno_buf(&RAW(s_h, a, b), W, vs, eps, buf);
            const double fyA = upwind_weno_buf(&RAW(s_A, a, b), W, vs, eps, buf);
            FY(s_Fyh, a, b) = FYS * fyh; FY(s_FyA, a, b) = FYS * fyA;
        } else if (tid >= 32 && tid < 32 + TY) {                    // east column of x-faces, a = TX+3
            const int a = TX + 3, b = 3 + (tid - 32);
            const double uw = RAW(s_u, a, b);
            const double fxh = upwind_weno(&RAW(s_h, a, b), 1, uw, eps);
            const double fxA = upwind_weno(&RAW(s_A, a, b), 1, uw, eps);
            FX(s_Fxh, a, b) = FXS * fxh; FX(s_FxA, a, b) = FXS * fxA;
        }
        __syncthreads();

        // ---- P2: tendencies of the own cell -------------------------------------------
        if (active) {
            // Gu at fcc
            {
                const double adv = adv_u;
                double dK = DIVDX(Cc(s_K, li, lj) - Cc(s_K, li - 1, lj));
                double pg = p.g * DIVDX(RAW(s_h, li, lj) - RAW(s_h, li - 1, lj));
                // lorentz_force_func_x — sw_mhd_jacobian_functions.jl:10-13,20-22
                double dxA = DIVDX(RAW(s_A, li, lj) - RAW(s_A, li - 1, lj));
                double hx = 0.5 * (RAW(s_h, li - 1, lj) + RAW(s_h, li, lj));
#if SWMHD_STRICT
#define DYBX(a, b) DIVDY(Cc(s_Bx, a, b) - Cc(s_Bx, a, (b) - 1))
#define DYA(a, b) DIVDY(RAW(s_A, a, b) - RAW(s_A, a, (b) - 1))
                double m1 = avg4(DYBX(li - 1, lj), DYBX(li, lj), DYBX(li - 1, lj + 1), DYBX(li, lj + 1));
                double m2 = avg4(DYA(li - 1, lj), DYA(li, lj), DYA(li - 1, lj + 1), DYA(li, lj + 1));
                double jac = dxA * m1 - m2 * DIVDX(Cc(s_Bx, li, lj) - Cc(s_Bx, li - 1, lj));
                double lor = fdiv(1.0, hx) * jac;
                Gn0 = (((adv - dK) - pg) + p.f * vhat) + lor;
#else
                // ℑxy(∂y F) telescopes to (F(i-1,j+1) + F(i,j+1) - F(i-1,j-1) - F(i,j-1)) / (4 dy)
                double m1 = ((Cc(s_Bx, li - 1, lj + 1) + Cc(s_Bx, li, lj + 1)) - (Cc(s_Bx, li - 1, lj - 1) + Cc(s_Bx, li, lj - 1))) * (0.25 * p.rdy);
                double m2 = ((RAW(s_A, li - 1, lj + 1) + RAW(s_A, li, lj + 1)) - (RAW(s_A, li - 1, lj - 1) + RAW(s_A, li, lj - 1))) * (0.25 * p.rdy);
                double jac = fma(dxA, m1, -(m2 * ((Cc(s_Bx, li, lj) - Cc(s_Bx, li - 1, lj)) * p.rdx)));
                Gn0 = fma(jac, frcp(hx), fma(p.f, vhat, (adv - dK) - pg));
#endif
            }
            // Gv at cfc (wall rows of a Bounded-y grid keep v = 0)
            if (!(p.by && gj < 2)) {
                const double adv = adv_v;
                double dK = DIVDY(Cc(s_K, li, lj) - Cc(s_K, li, lj - 1));
                double pg = p.g * DIVDY(RAW(s_h, li, lj) - RAW(s_h, li, lj - 1));
                // lorentz_force_func_y — sw_mhd_jacobian_functions.jl:15-18,24-26
                double dyA = DIVDY(RAW(s_A, li, lj) - RAW(s_A, li, lj - 1));
                double hy = 0.5 * (RAW(s_h, li, lj - 1) + RAW(s_h, li, lj));
#if SWMHD_STRICT
#define DXA(a, b) DIVDX(RAW(s_A, a, b) - RAW(s_A, (a) - 1, b))
#define DXBY(a, b) DIVDX(Cc(s_By, a, b) - Cc(s_By, (a) - 1, b))
                double m1 = avg4(DXA(li, lj - 1), DXA(li + 1, lj - 1), DXA(li, lj), DXA(li + 1, lj));
                double m2 = avg4(DXBY(li, lj - 1), DXBY(li + 1, lj - 1), DXBY(li, lj), DXBY(li + 1, lj));
                double jac = m1 * DIVDY(Cc(s_By, li, lj) - Cc(s_By, li, lj - 1)) - dyA * m2;
                double lor = fdiv(1.0, hy) * jac;
                Gn1 = (((-adv - dK) - pg) - p.f * uhat) + lor;
#else
                double m1 = ((RAW(s_A, li + 1, lj - 1) + RAW(s_A, li + 1, lj)) - (RAW(s_A, li - 1, lj - 1) + RAW(s_A, li - 1, lj))) * (0.25 * p.rdx);
                double m2 = ((Cc(s_By, li + 1, lj - 1) + Cc(s_By, li + 1, lj)) - (Cc(s_By, li - 1, lj - 1) + Cc(s_By, li - 1, lj))) * (0.25 * p.rdx);
                double jac = fma(m1, (Cc(s_By, li, lj) - Cc(s_By, li, lj - 1)) * p.rdy, -(dyA * m2));
                Gn1 = fma(jac, frcp(hy), fma(-p.f, uhat, (-adv - dK) - pg));
#endif
            }
            // Gh, GA at ccc
            {
#if SWMHD_STRICT
                Gn2 = -(p.inv_az * ((FX(s_Fxh, li + 1, lj) - own_fxh) + (FY(s_Fyh, li, lj + 1) - own_fyh)));
                double d = p.inv_az * ((FX(s_FxA, li + 1, lj) - own_fxA) + (FY(s_FyA, li, lj + 1) - own_fyA));
#else
                Gn2 = -fma(FX(s_Fxh, li + 1, lj) - own_fxh, p.rdx, (FY(s_Fyh, li, lj + 1) - own_fyh) * p.rdy);
                double d = fma(FX(s_FxA, li + 1, lj) - own_fxA, p.rdx, (FY(s_FyA, li, lj + 1) - own_fyA) * p.rdy);
#endif
#if SWMHD_STRICT
                double dv = p.inv_az * ((p.dy * RAW(s_u, li + 1, lj) - p.dy * own_u) +
                                        (p.dx * RAW(s_v, li, lj + 1) - p.dx * own_v));
#else
                double dv = fma(RAW(s_u, li + 1, lj) - own_u, p.rdx, (RAW(s_v, li, lj + 1) - own_v) * p.rdy);
#endif
                Gn3 = -d + RAW(s_A, li, lj) * dv;
            }
        }
    } else {
        // ================= Conservative + divergence-form Lorentz ====================
        double *s_hBx = smem + L::o_hBx, *s_hBy = smem + L::o_hBy, *s_Bx = smem + L::o_Bx, *s_By = smem + L::o_By;
        double *s_hx = smem + L::o_hx, *s_hy = smem + L::o_hy, *s_hff = smem + L::o_hff;
        double *s_Fuu = smem + L::o_Fuu, *s_Lxx = smem + L::o_Lxx, *s_Fuv = smem + L::o_Fuv, *s_Lxy = smem + L::o_Lxy;
        double *s_Tx = smem + L::o_Tx, *s_uq = smem + L::o_uq;
        double *s_Fvu = smem + L::o_Fvu, *s_Lyx = smem + L::o_Lyx, *s_Fvv = smem + L::o_Fvv, *s_Lyy = smem + L::o_Lyy;
        double *s_Ty = smem + L::o_Ty, *s_vq = smem + L::o_vq;
#define Bf(arr, a, b) arr[((b) - 1) * BP + (a) - 1]
#define HF(a, b) s_hff[((b) - 3) * XP + (a) - 3]
#define FXc(arr, a, b) arr[((b) - 3) * XP + (a) - 2]   /* ccc x-type: a in [2,TX+2] */
#define FX3(arr, a, b) arr[((b) - 3) * XP + (a) - 3]   /* a in [3,TX+3], b in [3,TY+2] */
#define FY3(arr, a, b) arr[((b) - 3) * YP + (a) - 3]   /* a in [3,TX+2], b in [3,TY+3] */
#define FYc(arr, a, b) arr[((b) - 2) * YP + (a) - 3]   /* ccc y-type: b in [2,TY+2] */

        // ---- P1a: hBx, hBy, Bx, By (sw_mhd_divergence_functions.jl:134-148), face and corner h ----
        for (int t = tid; t < BP * BR + XP * YR; t += NT) {
            if (t < BP * BR) {
                int a = 1 + t % BP, b = 1 + t / BP;
#if SWMHD_STRICT
#define DYA_(a_, b_) DIVDY(RAW(s_A, a_, b_) - RAW(s_A, a_, (b_) - 1))
#define DXA_(a_, b_) DIVDX(RAW(s_A, a_, b_) - RAW(s_A, (a_) - 1, b_))
                double hbx = -avg4(DYA_(a - 1, b), DYA_(a, b), DYA_(a - 1, b + 1), DYA_(a, b + 1));
                double hby = avg4(DXA_(a, b - 1), DXA_(a + 1, b - 1), DXA_(a, b), DXA_(a + 1, b));
#else
                double hbx = ((RAW(s_A, a - 1, b - 1) + RAW(s_A, a, b - 1)) - (RAW(s_A, a - 1, b + 1) + RAW(s_A, a, b + 1))) * (0.25 * p.rdy);
                double hby = ((RAW(s_A, a + 1, b - 1) + RAW(s_A, a + 1, b)) - (RAW(s_A, a - 1, b - 1) + RAW(s_A, a - 1, b))) * (0.25 * p.rdx);
#endif
                double hx = 0.5 * (RAW(s_h, a - 1, b) + RAW(s_h, a, b));
                double hy = 0.5 * (RAW(s_h, a, b - 1) + RAW(s_h, a, b));
                Bf(s_hBx, a, b) = hbx; Bf(s_hBy, a, b) = hby;
#if SWMHD_STRICT
                Bf(s_hx, a, b) = hx;   Bf(s_hy, a, b) = hy;
                Bf(s_Bx, a, b) = fdiv(hbx, hx);
                Bf(s_By, a, b) = fdiv(hby, hy);
#else
                const double rhx = frcp(hx), rhy = frcp(hy);        // FAST: s_hx / s_hy hold 1/ℑx h, 1/ℑy h
                Bf(s_hx, a, b) = rhx;  Bf(s_hy, a, b) = rhy;
                Bf(s_Bx, a, b) = hbx * rhx;
                Bf(s_By, a, b) = hby * rhy;
#endif
            } else {                                                // ℑxyᶠᶠᵃ h, a in [3,TX+3], b in [3,TY+3]
                int q = t - BP * BR;
                int a = 3 + q % XP, b = 3 + q / XP;
                HF(a, b) = avg4(RAW(s_h, a - 1, b - 1), RAW(s_h, a, b - 1), RAW(s_h, a - 1, b), RAW(s_h, a, b));
            }
        }
        __syncthreads();

        // ---- P1b: every flux once --------------------------------------------------------
        constexpr int NDG = DIAG ? NDG_ : 0;
        const int NyG = p.NyG, by = p.by;
        // Six flux families, one WENO5 each.  Every thread evaluates the six fluxes anchored at its own
        // cell in one straight-line block (independent FP64 chains interleave); warps 0/1 add the
        // tile's north row / east column.
        auto flux_uu = [&](int a, int b) {      // ccc: F_uu and Lxx (advective_lorentz_flux_hBx_bx :38-60)
            double ut = sym4(RAW(s_u, a - 1, b), RAW(s_u, a, b), RAW(s_u, a + 1, b), RAW(s_u, a + 2, b));
#if SWMHD_STRICT
            FXc(s_Fuu, a, b) = fdiv(p.dy * upwind_weno(&RAW(s_u, a + 1, b), 1, ut, eps), RAW(s_h, a, b));
#else
            const double mom_ = fdiv(p.dy * upwind_weno(&RAW(s_u, a + 1, b), 1, ut, eps), RAW(s_h, a, b));
#endif
            double ul = 0.5 * (Bf(s_hBx, a, b) + Bf(s_hBx, a + 1, b));
            double Lq = third(Bf(s_Bx, a + 1, b), Bf(s_Bx, a, b), Bf(s_Bx, a - 1, b));
            double Rq = thirdR(Bf(s_Bx, a + 2, b), Bf(s_Bx, a + 1, b), Bf(s_Bx, a, b));
#if SWMHD_STRICT
            FXc(s_Lxx, a, b) = p.dy * upwind_sel(ul, Lq, Rq);
#else
            FXc(s_Fuu, a, b) = p.dy * upwind_sel(ul, Lq, Rq) - mom_;      // FAST: Lorentz minus momentum flux, one array
#endif
        };
        auto flux_uv = [&](int a, int b) {      // ffc: F_uv and Lxy (advective_lorentz_flux_hBx_by :86-108)
            int gjf = p.gj0 + j0 + (b - 3);
            double u2 = sym2(RAW(s_u, a, b - 1), RAW(s_u, a, b));
            double u4 = sym4(RAW(s_u, a, b - 2), RAW(s_u, a, b - 1), RAW(s_u, a, b), RAW(s_u, a, b + 1));
            double ut = ybuf(by, gjf, 2, NyG) ? u2 : u4;
#if SWMHD_STRICT
            FX3(s_Fuv, a, b) = fdiv(p.dy * upwind_weno(&RAW(s_v, a, b), 1, ut, eps), HF(a, b));
#else
            const double mom_ = fdiv(p.dy * upwind_weno(&RAW(s_v, a, b), 1, ut, eps), HF(a, b));
#endif
            double ul = 0.5 * (Bf(s_hBx, a, b - 1) + Bf(s_hBx, a, b));
            double Lq = third(Bf(s_By, a, b), Bf(s_By, a - 1, b), Bf(s_By, a - 2, b));
            double Rq = thirdR(Bf(s_By, a + 1, b), Bf(s_By, a, b), Bf(s_By, a - 1, b));
#if SWMHD_STRICT
            FX3(s_Lxy, a, b) = p.dy * upwind_sel(ul, Lq, Rq);
#else
            FX3(s_Fuv, a, b) = p.dy * upwind_sel(ul, Lq, Rq) - mom_;      // FAST: Lorentz minus momentum flux, one array
#endif
        };
        auto flux_tx = [&](int a, int b) {      // fcc: tracer transport flux and uh/ℑx h
            double vel = RAW(s_u, a, b), hx = Bf(s_hx, a, b);
#if SWMHD_STRICT
            FX3(s_Tx, a, b) = fdiv(p.dy * upwind_weno(&RAW(s_A, a, b), 1, vel, eps), hx);
            FX3(s_uq, a, b) = fdiv(vel, hx);
#else
            FX3(s_Tx, a, b) = (p.dy * upwind_weno(&RAW(s_A, a, b), 1, vel, eps)) * hx;   // hx holds 1/ℑx h here
            FX3(s_uq, a, b) = vel * hx;
#endif
        };
        auto flux_vu = [&](int a, int b) {      // ffc: F_vu and Lyx (advective_lorentz_flux_hBy_bx :62-84 with edge branches)
            int gjf = p.gj0 + j0 + (b - 3);
            double vt = sym4(RAW(s_v, a - 2, b), RAW(s_v, a - 1, b), RAW(s_v, a, b), RAW(s_v, a + 1, b));
#if SWMHD_STRICT
            FY3(s_Fvu, a, b) = fdiv(p.dx * upwind_weno_buf(&RAW(s_u, a, b), W, vt, eps, ybuf(by, gjf, 3, NyG)), HF(a, b));
#else
            const double mom_ = fdiv(p.dx * upwind_weno_buf(&RAW(s_u, a, b), W, vt, eps, ybuf(by, gjf, 3, NyG)), HF(a, b));
#endif
            double vl = 0.5 * (Bf(s_hBy, a - 1, b) + Bf(s_hBy, a, b));
            double L3 = third(Bf(s_Bx, a, b), Bf(s_Bx, a, b - 1), Bf(s_Bx, a, b - 2));
            double R3 = thirdR(Bf(s_Bx, a, b + 1), Bf(s_Bx, a, b), Bf(s_Bx, a, b - 1));
            double L1 = Bf(s_Bx, a, b - 1), R1 = Bf(s_Bx, a, b);
            double Lq = L3, Rq = R3;
            if (by) {
                if (gjf == 1) { Lq = R1; Rq = R1; } else if (gjf == 2) { Lq = L1; Rq = R3; }
                else if (gjf == NyG) { Lq = L3; Rq = R1; } else if (gjf == NyG + 1) { Lq = L1; Rq = L1; }
            }
#if SWMHD_STRICT
            FY3(s_Lyx, a, b) = p.dx * upwind_sel(vl, Lq, Rq);
#else
            FY3(s_Fvu, a, b) = p.dx * upwind_sel(vl, Lq, Rq) - mom_;      // FAST: Lorentz minus momentum flux, one array
#endif
        };
        auto flux_vv = [&](int a, int b) {      // ccc: F_vv and Lyy (advective_lorentz_flux_hBy_by :110-132)
            int gjc = p.gj0 + j0 + (b - 3);                         // global cell row
            double v2 = sym2(RAW(s_v, a, b), RAW(s_v, a, b + 1));
            double v4 = sym4(RAW(s_v, a, b - 1), RAW(s_v, a, b), RAW(s_v, a, b + 1), RAW(s_v, a, b + 2));
            double vt = ybuf(by, gjc + 1, 2, NyG + 1) ? v2 : v4;
#if SWMHD_STRICT
            FYc(s_Fvv, a, b) = fdiv(p.dx * upwind_weno_buf(&RAW(s_v, a, b + 1), W, vt, eps, ybuf(by, gjc + 1, 3, NyG + 1)), RAW(s_h, a, b));
#else
            const double mom_ = fdiv(p.dx * upwind_weno_buf(&RAW(s_v, a, b + 1), W, vt, eps, ybuf(by, gjc + 1, 3, NyG + 1)), RAW(s_h, a, b));
#endif
            double vl = 0.5 * (Bf(s_hBy, a, b) + Bf(s_hBy, a, b + 1));
            double L3 = third(Bf(s_By, a, b + 1), Bf(s_By, a, b), Bf(s_By, a, b - 1));
            double R3 = thirdR(Bf(s_By, a, b + 2), Bf(s_By, a, b + 1), Bf(s_By, a, b));
            double L1 = Bf(s_By, a, b), R1 = Bf(s_By, a, b + 1);
            double Lq = L3, Rq = R3;
            if (by) {
                if (gjc == 0) { Lq = R1; Rq = R1; } else if (gjc == 1) { Lq = L1; Rq = R3; }
                else if (gjc == NyG - 1) { Lq = L3; Rq = R1; } else if (gjc == NyG) { Lq = L1; Rq = L1; }
            }
#if SWMHD_STRICT
            FYc(s_Lyy, a, b) = p.dx * upwind_sel(vl, Lq, Rq);
#else
            FYc(s_Fvv, a, b) = p.dx * upwind_sel(vl, Lq, Rq) - mom_;      // FAST: Lorentz minus momentum flux, one array
#endif
        };
        auto flux_ty = [&](int a, int b) {      // cfc: tracer transport flux and vh/ℑy h
            int gjf = p.gj0 + j0 + (b - 3);
            double vel = RAW(s_v, a, b), hy = Bf(s_hy, a, b);
#if SWMHD_STRICT
            FY3(s_Ty, a, b) = fdiv(p.dx * upwind_weno_buf(&RAW(s_A, a, b), W, vel, eps, ybuf(by, gjf, 3, NyG)), hy);
            FY3(s_vq, a, b) = fdiv(vel, hy);
#else
            FY3(s_Ty, a, b) = (p.dx * upwind_weno_buf(&RAW(s_A, a, b), W, vel, eps, ybuf(by, gjf, 3, NyG))) * hy;   // 1/ℑy h
            FY3(s_vq, a, b) = vel * hy;
#endif
        };
        flux_uu(li - 1, lj); flux_uv(li, lj); flux_tx(li, lj);
        flux_vu(li, lj); flux_vv(li, lj - 1); flux_ty(li, lj);
        if (tid < TX) {                                             // north row
            flux_vu(3 + tid, TY + 3); flux_vv(3 + tid, TY + 2); flux_ty(3 + tid, TY + 3);
        } else if (tid >= 32 && tid < 32 + TY) {                    // east column
            const int b = 3 + (tid - 32);
            flux_uu(TX + 2, b); flux_uv(TX + 3, b); flux_tx(TX + 3, b);
        }
        if constexpr (DIAG) {
            for (int q = tid; q < NDG; q += NT) {
                if (q < YP * YR) {
                    int a = 3 + q % YP, b = 3 + q / YP;
#if SWMHD_STRICT
                    double bx = fdiv(-DIVDY(RAW(s_A, a, b) - RAW(s_A, a, b - 1)), Bf(s_hy, a, b));
#else
                    double bx = -DIVDY(RAW(s_A, a, b) - RAW(s_A, a, b - 1)) * Bf(s_hy, a, b);
#endif
                    s_sqBx[q] = bx * bx;
                } else {
                    int r = q - YP * YR;
                    int a = 3 + r % XP, b = 2 + r / XP;
#if SWMHD_STRICT
                    double by_ = fdiv(DIVDX(RAW(s_A, a, b) - RAW(s_A, a - 1, b)), Bf(s_hx, a, b));
#else
                    double by_ = DIVDX(RAW(s_A, a, b) - RAW(s_A, a - 1, b)) * Bf(s_hx, a, b);
#endif
                    s_sqBy[r] = by_ * by_;
                }
            }
        }
        __syncthreads();

        // ---- P2 --------------------------------------------------------------------------
        if (active) {
            // d(g h^2 / 2): h = 1 + O(1e-9) makes this a cancellation; keep the products
            // un-contracted (no FMA) in both arithmetic modes so the rounding is symmetric.
            const double hg = 0.5 * p.g;
            double hc = RAW(s_h, li, lj), hw = RAW(s_h, li - 1, lj), hs = RAW(s_h, li, lj - 1);
            double Pc = __dmul_rn(hg, __dmul_rn(hc, hc));
            double Pw = __dmul_rn(hg, __dmul_rn(hw, hw)), Ps = __dmul_rn(hg, __dmul_rn(hs, hs));
            {   // Guh
                double dm = p.inv_az * ((FXc(s_Fuu, li, lj) - FXc(s_Fuu, li - 1, lj)) + (FY3(s_Fvu, li, lj + 1) - FY3(s_Fvu, li, lj)));
                double pg = DIVDX(__dsub_rn(Pc, Pw));
                double vhat = avg4(RAW(s_v, li - 1, lj), RAW(s_v, li, lj), RAW(s_v, li - 1, lj + 1), RAW(s_v, li, lj + 1));
#if SWMHD_STRICT
                double lor = p.inv_az * ((FXc(s_Lxx, li, lj) - FXc(s_Lxx, li - 1, lj)) + (FY3(s_Lyx, li, lj + 1) - FY3(s_Lyx, li, lj)));
                Gn0 = ((-dm - pg) + p.f * vhat) + lor;
#else
                Gn0 = fma(p.f, vhat, dm - pg);                      // dm already holds div(Lorentz - momentum flux)
#endif
            }
            if (!(p.by && gj < 2)) {   // Gvh
                double dm = p.inv_az * ((FX3(s_Fuv, li + 1, lj) - FX3(s_Fuv, li, lj)) + (FYc(s_Fvv, li, lj) - FYc(s_Fvv, li, lj - 1)));
                double pg = DIVDY(__dsub_rn(Pc, Ps));
                double uhat = avg4(RAW(s_u, li, lj - 1), RAW(s_u, li + 1, lj - 1), RAW(s_u, li, lj), RAW(s_u, li + 1, lj));
#if SWMHD_STRICT
                double lor = p.inv_az * ((FX3(s_Lxy, li + 1, lj) - FX3(s_Lxy, li, lj)) + (FYc(s_Lyy, li, lj) - FYc(s_Lyy, li, lj - 1)));
                Gn1 = ((-dm - pg) - p.f * uhat) + lor;
#else
                Gn1 = fma(-p.f, uhat, dm - pg);
#endif
            }
            {   // Gh (centred), GA
#if SWMHD_STRICT
                double dv = p.inv_az * ((p.dy * RAW(s_u, li + 1, lj) - p.dy * RAW(s_u, li, lj)) +
                                        (p.dx * RAW(s_v, li, lj + 1) - p.dx * RAW(s_v, li, lj)));
#else
                double dv = fma(RAW(s_u, li + 1, lj) - RAW(s_u, li, lj), p.rdx, (RAW(s_v, li, lj + 1) - RAW(s_v, li, lj)) * p.rdy);
#endif
                Gn2 = -dv;
                double d = p.inv_az * ((FX3(s_Tx, li + 1, lj) - FX3(s_Tx, li, lj)) + (FY3(s_Ty, li, lj + 1) - FY3(s_Ty, li, lj)));
                double cdiv = DIVDX(FX3(s_uq, li + 1, lj) - FX3(s_uq, li, lj)) + DIVDY(FY3(s_vq, li, lj + 1) - FY3(s_vq, li, lj));
                Gn3 = -d + RAW(s_A, li, lj) * cdiv;
            }
        }
    }

    // ---- RK3 substep + stores ---------------------------------------------------------
    if (active) {
        const double Gn[4] = {Gn0, Gn1, Gn2, Gn3};
        const double Gm[4] = {Gm0, Gm1, Gm2, Gm3};
        const double Uc[4] = {RAW(s_u, li, lj), RAW(s_v, li, lj), RAW(s_h, li, lj), RAW(s_A, li, lj)};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if constexpr (STAGE == 0) {
                p.G[k][gcell] = Gn[k];
            } else if constexpr (STAGE == 1) {
                p.Un[k][gcell] = Uc[k] + p.dtgam * Gn[k];
                p.G[k][gcell] = Gn[k];
            } else {
                p.Un[k][gcell] = Uc[k] + p.dt * (p.gam * Gn[k] + p.zet * Gm[k]);
                if constexpr (STAGE == 2) p.G[k][gcell] = Gn[k];
            }
        }
    }

    // ---- fused diagnostics of the state at the start of the step -------------
    if constexpr (DIAG) {
        double ke = 0, me = 0, pe = 0, sh = 0, mu = 0, mA = 0, mh = -INFINITY, md = 0, nf = 0;
        if (active) {
            auto sq = [](double x) { return x * x; };
            double hh = RAW(s_h, li, lj), aa = RAW(s_A, li, lj), uu = RAW(s_u, li, lj), vv = RAW(s_v, li, lj);
            // KE bracket u^2 + ℑxyᶠᶜᵃ(v^2) at fcc(i) and fcc(i+1), averaged to the centre
            auto keb = [&](int a) {
                return sq(RAW(s_u, a, lj)) + avg4(sq(RAW(s_v, a - 1, lj)), sq(RAW(s_v, a, lj)), sq(RAW(s_v, a - 1, lj + 1)), sq(RAW(s_v, a, lj + 1)));
            };
            double wgt = (FORM == 0) ? 0.5 * hh : 0.5 * fdiv(1.0, hh);
            ke = wgt * (0.5 * (keb(li) + keb(li + 1)));
            // ME bracket Bx^2 + ℑxyᶜᶠᵃ(By^2) at cfc(j) and cfc(j+1)
            auto SQBX = [&](int a, int b) { return s_sqBx[(b - 3) * YP + a - 3]; };
            auto SQBY = [&](int a, int b) { return s_sqBy[(b - 2) * XP + a - 3]; };
            auto meb = [&](int b) {
                return SQBX(li, b) + avg4(SQBY(li, b - 1), SQBY(li + 1, b - 1), SQBY(li, b), SQBY(li + 1, b));
            };
            me = (0.5 * hh) * (0.5 * (meb(lj) + meb(lj + 1)));
            double dh = hh - p.h_ref;
            pe = (0.5 * p.g) * (dh * dh);
            sh = hh;
            mu = (FORM == 0) ? fabs(uu) : fabs(fdiv(uu, 0.5 * (RAW(s_h, li - 1, lj) + hh)));
            mA = fabs(aa);
            mh = -hh;
            // div(hB) at ccc with hBx, hBy of sw_mhd_divergence_functions.jl:142-148 (pure round-off: ~1e-15)
#if SWMHD_STRICT
            auto DyA = [&](int a, int b) { return DIVDY(RAW(s_A, a, b) - RAW(s_A, a, b - 1)); };
            auto DxA = [&](int a, int b) { return DIVDX(RAW(s_A, a, b) - RAW(s_A, a - 1, b)); };
            auto hBx = [&](int a, int b) { return -avg4(DyA(a - 1, b), DyA(a, b), DyA(a - 1, b + 1), DyA(a, b + 1)); };
            auto hBy = [&](int a, int b) { return avg4(DxA(a, b - 1), DxA(a + 1, b - 1), DxA(a, b), DxA(a + 1, b)); };
            md = fabs(DIVDX(hBx(li + 1, lj) - hBx(li, lj)) + DIVDY(hBy(li, lj + 1) - hBy(li, lj)));
#else
            {   // telescoped ℑxy∂: hBx(a,b) = (A(a-1,b-1)+A(a,b-1)-A(a-1,b+1)-A(a,b+1))/(4dy), hBy likewise
                const double Amm = RAW(s_A, li - 1, lj - 1), A0m = RAW(s_A, li, lj - 1), Apm = RAW(s_A, li + 1, lj - 1);
                const double Am0 = RAW(s_A, li - 1, lj), Ap0 = RAW(s_A, li + 1, lj);
                const double Amp = RAW(s_A, li - 1, lj + 1), A0p = RAW(s_A, li, lj + 1), App = RAW(s_A, li + 1, lj + 1);
                const double hbx0 = (Amm + A0m) - (Amp + A0p), hbx1 = (A0m + Apm) - (A0p + App);
                const double hby0 = (Apm + Ap0) - (Amm + Am0), hby1 = (Ap0 + App) - (Am0 + Amp);
                md = fabs((hbx1 - hbx0) * (0.25 * p.rdy) * p.rdx + (hby1 - hby0) * (0.25 * p.rdx) * p.rdy);
            }
#endif
            if (!(isfinite(hh) && isfinite(aa) && isfinite(uu) && isfinite(vv))) nf = 1.0;
        }
        // 256 -> 32 through shared memory (one add per value and thread row), then one warp finishes
        // with shuffles: 3x fewer FP64 instructions than eight independent warp trees.
        // (the scratch aliases the derived/flux arrays, which are dead once every thread has finished phase C,
        //  so the DIAG variant keeps the 3 CTAs/SM of the plain one)
        static_assert(NDIAG * NT <= L::o_dg, "diag scratch must fit in the derived-array region");
        double (*red)[NT] = reinterpret_cast<double (*)[NT]>(smem);
        const double mine[NDIAG] = {ke, me, pe, sh, mu, mA, mh, md, nf};
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NDIAG; q++) red[q][tid] = mine[q];
        __syncthreads();
        if (tid < 32) {
            double r[NDIAG];
#pragma unroll
            for (int q = 0; q < NDIAG; q++) {
                const bool is_max = (q >= 4 && q <= 7);
                double acc = red[q][tid];
#pragma unroll
                for (int w = 1; w < NT / 32; w++) {
                    const double x = red[q][tid + 32 * w];
                    acc = is_max ? fmax(acc, x) : acc + x;
                }
                r[q] = is_max ? warp_max(acc) : warp_sum(acc);
            }
            if (tid == 0) {
                double *dst = p.diag + ((size_t)p.tile_row0 * tiles_x + tile) * NDIAG;   // slot 6 = max(-h) = -min h
#pragma unroll
                for (int q = 0; q < NDIAG; q++) dst[q] = r[q];
            }
        }
    }
    __syncthreads();   // end of tile: derived arrays and this raw stage may be overwritten
  }
}

template <int FORM, int STAGE, bool DIAG, bool TMA>
cudaError_t launch_cfg(const KParams &p, cudaStream_t st) {
    // One raw-tile stage: with one tile per CTA (the default, non-persistent launch) the other resident
    // CTAs hide the TMA latency; a second stage only costs registers and shared memory.
    // FORM 0: two raw-tile stages (the TMA load of the CTA's next tile overlaps the arithmetic of the
    // current one); FORM 1 needs its shared memory for 3 CTAs/SM and relies on the other CTAs.
    constexpr int NSTG = (TMA && FORM == 0) ? 2 : 1;
    constexpr size_t bytes = 128 + ((size_t)NSTG * 4 * SZP + SmemLayout<FORM, DIAG>::total) * sizeof(double) + 16;
    auto kern = substage_kernel<FORM, STAGE, DIAG, TMA, NSTG>;
    // per device (the shared-memory opt-in is a per-device attribute; a process may drive several devices)
    constexpr int MAX_DEVICES = 64;
    static int max_ctas_dev[MAX_DEVICES] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
    if (max_ctas_dev[dev] == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        int sms = 0, occ = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, bytes);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        max_ctas_dev[dev] = occ * sms;
    }
    const int max_ctas = max_ctas_dev[dev];
    const int ntiles = ((p.Nx + TX - 1) / TX) * p.tile_rows;
    static int tpc = -1;
    if (tpc < 0) { const char *ev = getenv("SWMHD_TPC"); tpc = ev ? atoi(ev) : SWMHD_TPC_DEFAULT; if (tpc < 1) tpc = 1; }
    KParams q = p;
    q.tiles_per_cta = (NSTG == 2) ? tpc : 1;
    const int grid = (ntiles + q.tiles_per_cta - 1) / q.tiles_per_cta;
    static int ahead = -2;
    if (ahead == -2) { const char *ev = getenv("SWMHD_L2_AHEAD"); ahead = ev ? atoi(ev) : -1; }
    q.l2_ahead = (ahead >= 0) ? ahead : max_ctas;               // CTAs in flight = distance to the slot's next CTA
    kern<<<grid, NT, bytes, st>>>(q);
    return cudaGetLastError();
}
template <int FORM, int STAGE, bool DIAG>
cudaError_t launch_one(const KParams &p, cudaStream_t st) {
    return p.use_tma ? launch_cfg<FORM, STAGE, DIAG, true>(p, st) : launch_cfg<FORM, STAGE, DIAG, false>(p, st);
}

} // namespace

cudaError_t LAUNCH_NAME(const KParams &p, int form, int stage, cudaStream_t st) {
    if (p.tile_rows <= 0) return cudaSuccess;
    const bool dg = (p.diag != nullptr);
    if (dg && stage != 1) return cudaErrorInvalidValue;
#if !SWMHD_STRICT
    // FAST arithmetic with TMA: the row-blocked kernels (substage_rb.cu); SWMHD_RB_STAGES masks stages
    // (bit s-1 = stage s, default 7 = all) for A/B runs against this file's kernel.
    if (stage >= 1 && p.use_tma && p.use_rb && ((substage_rb_stage_mask() >> (stage - 1)) & 1))
        return launch_substage_rb(p, form, stage, st);
#endif
    switch (form * 4 + stage) {
        case 0: return launch_one<0, 0, false>(p, st);
        case 1: return dg ? launch_one<0, 1, true>(p, st) : launch_one<0, 1, false>(p, st);
        case 2: return launch_one<0, 2, false>(p, st);
        case 3: return launch_one<0, 3, false>(p, st);
        case 4: return launch_one<1, 0, false>(p, st);
        case 5: return dg ? launch_one<1, 1, true>(p, st) : launch_one<1, 1, false>(p, st);
        case 6: return launch_one<1, 2, false>(p, st);
        case 7: return launch_one<1, 3, false>(p, st);
    }
    return cudaErrorInvalidValue;
}

#if SWMHD_STRICT
void substage_tile(int *tx, int *ty) { *tx = TX; *ty = TY; }
#else
int substage_rb_stage_mask() {
    static int mask = -1;
    if (mask < 0) { const char *e = getenv("SWMHD_RB_STAGES"); mask = e ? atoi(e) : 7; }
    return mask;
}
#endif

} // namespace swmhd
