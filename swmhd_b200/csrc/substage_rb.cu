// substage_rb.cu — row-blocked fused RK3-substage kernel (sm_100a), Jacobian formulation, FAST arithmetic.
//
// Same work per launch as substage_kernel.cu (calculate_tendencies! + rk3_substep! + store_tendencies!
// for all four fields, jacobian_formulation/sw_mhd_jacobian_functions.jl:1-26 inlined; in the
// stage-1 DIAG variant also the diagnostics of SWMHD_example.jl:47-77), different thread mapping:
//
//   * a CTA owns a 32 x RB_TY tile staged by TMA, a WARP owns RB_R consecutive rows of it and walks
//     them south to north, a THREAD one column of those rows.
//   * after the derived staggered fields (phase A, CTA-wide) warps never meet again: the flux through
//     a cell's north face is kept in registers for the next row (one extra evaluation per warp for the
//     south face of its first row), the east-face fluxes travel by warp shuffle, the tile's east column
//     of faces is a per-warp pre-pass.  One __syncthreads per tile (one more in the DIAG variant)
//     instead of three, and no flux arrays in shared memory.
//   * the next tile of the SM slot is pulled into L2 while this one computes (cp.async.bulk.prefetch).
//
// Measured (profiles/README.md): 0.714 / 0.752 / 0.726 ms per stage at 4096^2 against 0.756 / 0.782 /
// 0.743 ms of the one-thread-per-cell kernel; with the fused diagnostics 0.87 vs 1.06 ms.  A variant
// that also kept the y stencils in sliding register windows was slower (register moves and selects
// outweighed the saved shared-memory loads): see the history table there.
// Arithmetic per value is the FAST arithmetic of substage_kernel.cu (same operation order).
#include "kparams.h"
#include "device_prims.cuh"
#include <cstdlib>
#include <cmath>

namespace swmhd {
namespace {

#ifndef RB_TY
#define RB_TY 16
#endif
#ifndef RB_R
#define RB_R 4
#endif
#ifndef RB_MINB
#define RB_MINB 3
#endif
#ifndef RB_L2_PREFETCH
#define RB_L2_PREFETCH 1
#endif
constexpr int TX = 32, TYB = RB_TY, R = RB_R, NW = TYB / R, NT = 32 * NW;
static_assert(TYB % R == 0 && TYB % 8 == 0, "tile rows: multiple of the rows per warp and of the 8-row launch granularity");
static_assert((R & (R - 1)) == 0 && 2 * R <= 32, "rows per warp: power of two");
constexpr int W = TX + 6, HT = TYB + 6, SZ = W * HT;       // raw tiles: [HT][W] (dense TMA box)
constexpr int SZP = (SZ * 8 + 127) / 128 * 16;             // padded to a multiple of 128 B (TMA dst alignment)
constexpr unsigned TILE_TX_BYTES = 4u * SZ * 8u;
constexpr int ZP = TX + 5, ZR = TYB + 5;                   // ffc points a in [1,TX+5], b in [1,TYB+5]
constexpr int CP = TX + 2, CR = TYB + 2;                   // ccc points a in [2,TX+3], b in [2,TYB+3]
constexpr int NZ = ZP * ZR, NC = CP * CR;
constexpr int o_z = 0, o_ut = NZ, o_vt = 2 * NZ, o_K = 3 * NZ, o_Bx = o_K + NC, o_By = o_Bx + NC;
constexpr int DERIVED = o_By + NC;
constexpr size_t SMEM_BYTES = ((size_t)4 * SZP + DERIVED + 2 + NW * NDIAG) * sizeof(double);   // + mbarrier + DIAG partials

#define RAW(arr, a, b) arr[(b) * W + (a)]
#define Zf(arr, a, b) arr[((b) - 1) * ZP + (a) - 1]
#define Cc(arr, a, b) arr[((b) - 2) * CP + (a) - 2]

__device__ __forceinline__ double avg4(double a, double b, double c, double d) { return 0.25 * ((a + b) + (c + d)); }
__device__ __forceinline__ double sym2(double b, double c) { return 0.5 * (b + c); }
// Bounded-y wall buffer (oracle ybuf): footprint f-n..f+n-1 must stay in [1,hi]
__device__ __forceinline__ bool ybuf(int by, int f, int n, int hi) { return by && (f - n < 1 || f + n - 1 > hi); }

// ---- WENO5-Z from shared memory (x direction): pointer-selected upwind stencil --------------------
__device__ __forceinline__ double weno5_mem(const double *q, int s, double es) {
    const double a = q[0], b = q[s], c = q[2 * s], d = q[3 * s], e = q[4 * s];
    const double d1 = b - a, d2 = c - b, d3 = d - c, d4 = e - d;
    double c0 = es, c1 = es, c2 = es, num, den;
    weno_beta_acc4(d1, d2, d3, d4, c0, c1, c2);
    weno_corr4(d1, d2, d3, d4, c0, c1, c2, num, den);
    return fma(num, frcp(den), c);
}
// vel * psi_upwind at face f (ctr = &psi[f]): left-biased psi[f-3..f+1] for vel > 0, else the mirror
__device__ __forceinline__ double upwind_weno_mem(const double *ctr, double vel, double eps) {
    const bool pos = vel > 0.0;
    return vel * weno5_mem(pos ? ctr - 3 : ctr + 2, pos ? 1 : -1, eps * (12.0 / 13.0));
}
// first differences of the five upwind samples of a shared-memory line (pointer-selected side)
__device__ __forceinline__ void diffs_mem(const double *q, int s, double &d1, double &d2, double &d3, double &d4, double &c) {
    const double a = q[0], b = q[s], e = q[4 * s], d = q[3 * s];
    c = q[2 * s];
    d1 = b - a; d2 = c - b; d3 = d - c; d4 = e - d;
}
__device__ __forceinline__ void diffs_mem(const double *q, int s, double &d1, double &d2, double &d3, double &d4) {
    double c; diffs_mem(q, s, d1, d2, d3, d4, c);
}
// stencil of face f along a line of stride st (ctr = &psi[f]): left-biased psi[f-3..f+1] for pos, else the mirror
#define UPD(pos, ctr, st, n) diffs_mem((pos) ? (ctr) - 3 * (st) : (ctr) + 2 * (st), (pos) ? (st) : -(st), d1[n], d2[n], d3[n], d4[n])
#define UPDC(pos, ctr, st, n, cvar) diffs_mem((pos) ? (ctr) - 3 * (st) : (ctr) + 2 * (st), (pos) ? (st) : -(st), d1[n], d2[n], d3[n], d4[n], cvar)

// STAGE 1,2,3.  DIAG only with STAGE 1.
template <int STAGE, bool DIAG>
__global__ void __launch_bounds__(NT, RB_MINB) substage_rb_kernel(const __grid_constant__ KParams p) {
    // 128-byte aligned for the TMA destination; used directly so that the compiler keeps the shared
    // address space (LDS/STS instead of generic LD/ST).
    extern __shared__ __align__(128) unsigned char smem_bytes[];
    double *const s_u = reinterpret_cast<double *>(smem_bytes);
    double *const s_v = s_u + SZP, *const s_h = s_u + 2 * SZP, *const s_A = s_u + 3 * SZP;
    double *const smem = s_u + 4 * SZP;
    double *const s_z = smem + o_z, *const s_ut = smem + o_ut, *const s_vt = smem + o_vt;
    double *const s_K = smem + o_K, *const s_Bx = smem + o_Bx, *const s_By = smem + o_By;
    uint64_t *const mbar = reinterpret_cast<uint64_t *>(smem + DERIVED);
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int Nx = p.Nx, P = p.P;
    const double eps = p.eps;
    const int tiles_x = (Nx + TX - 1) / TX;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int row0 = p.row_begin + tile_y * TYB;            // 0-based first cell row = parent row of b = 0

    // ---- P0: stage u,v,h,A with a 3-cell halo: four TMA boxes of (TX+6) x (TYB+6) doubles ----------
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_fence_init();
        mbar_expect_tx(mbar, TILE_TX_BYTES);
#pragma unroll
        for (int k = 0; k < 4; k++) tma_load_2d(s_u + k * SZP, &p.tm_rb[k], tile_x * TX, row0, mbar);
#if RB_L2_PREFETCH
        // pull the tile that the CTA taking over this slot will load into L2 now (CTAs are dispatched in
        // blockIdx order): its TMA wait then costs an L2 hit instead of an HBM round trip
        const int nxt = blockIdx.x + p.l2_ahead;
        if (nxt < gridDim.x) {
#pragma unroll
            for (int k = 0; k < 4; k++) tma_prefetch_2d(&p.tm_rb[k], (nxt % tiles_x) * TX, p.row_begin + (nxt / tiles_x) * TYB);
        }
#endif
    }
    const int li = lane + 3, lj0 = 3 + wp * R;              // tile-local column / first row of this thread
    const int i = tile_x * TX + 1 + lane;                   // logical (1-based) column
    const int jc0 = row0 + 1 + wp * R;                      // logical (1-based, slab-local) first row
    const int gj0 = p.gj0 + jc0;                            // global row of the first cell (wall logic)
    double Gq0 = 0.0, Gq1 = 0.0, Gq2 = 0.0, Gq3 = 0.0;      // G^- of the row in flight (software-pipelined global loads)
    if constexpr (STAGE >= 2) {
        if ((i <= Nx) && (jc0 <= p.row_end)) {               // first row: in flight during the tile wait and phase A
            const size_t g0 = (size_t)(i + 2) + (size_t)P * (size_t)(jc0 + 2);
            Gq0 = p.G[0][g0]; Gq1 = p.G[1][g0]; Gq2 = p.G[2][g0]; Gq3 = p.G[3][g0];
        }
    }
    __syncthreads();                                        // barrier initialised before anybody polls it
    mbar_wait(mbar, 0);

    // ---- A: derived staggered fields, each point once per tile --------------------------------------
#pragma unroll 4
    for (int q = tid; q < NZ; q += NT) {                    // zeta, ℑy u, ℑx v at ffc
        const int a = 1 + q % ZP, b = 1 + q / ZP;
        const double vc = RAW(s_v, a, b), vw = RAW(s_v, a - 1, b), uc = RAW(s_u, a, b), us = RAW(s_u, a, b - 1);
        s_z[q] = fma(vc - vw, p.rdx, (us - uc) * p.rdy);
        s_ut[q] = 0.5 * (us + uc);
        s_vt[q] = 0.5 * (vw + vc);
    }
#pragma unroll 3
    for (int q = tid; q < NC; q += NT) {                    // K, Bx, By at ccc (sw_mhd_jacobian_functions.jl:1-7)
        const int a = 2 + q % CP, b = 2 + q / CP;
        const double u0 = RAW(s_u, a, b), u1 = RAW(s_u, a + 1, b), v0 = RAW(s_v, a, b), v1 = RAW(s_v, a, b + 1);
        s_K[q] = 0.25 * (fma(u0, u0, u1 * u1) + fma(v0, v0, v1 * v1));
        const double rh = frcp(RAW(s_h, a, b));
        s_Bx[q] = ((RAW(s_A, a, b - 1) - RAW(s_A, a, b + 1)) * (0.5 * p.rdy)) * rh;
        s_By[q] = ((RAW(s_A, a + 1, b) - RAW(s_A, a - 1, b)) * (0.5 * p.rdx)) * rh;
    }
    __syncthreads();

    // ---- B/C: warp-private from here on --------------------------------------------------------------

    // B0: the tile's east column of x faces for this warp's rows: lanes 0..R-1 take h, lanes R..2R-1 take A
    double eastF = 0.0;
    if (lane < 2 * R) {
        const int b = lj0 + (lane & (R - 1));
        const double *arr = (lane < R) ? s_h : s_A;
        eastF = upwind_weno_mem(&RAW(arr, TX + 3, b), RAW(s_u, TX + 3, b), eps);
    }

    // Row loop.  Iteration it = -1 only produces the fluxes through the south face of the warp's first
    // row; every other iteration evaluates the NORTH face lj+1 of its row and keeps it for the next one.
    double vC = RAW(s_v, li, lj0 - 1), vW = RAW(s_v, li - 1, lj0 - 1);
    double fyh_s = 0.0, fyA_s = 0.0;
    double dg[NDIAG];
    if constexpr (DIAG) {
#pragma unroll
        for (int q = 0; q < NDIAG; q++) dg[q] = 0.0;
        dg[6] = -INFINITY;
    }
    (void)dg;

#pragma unroll 1
    for (int it = -1; it < R; ++it) {
        const int lj = lj0 + it;
        const int j = jc0 + it, gj = gj0 + it;
        const bool active = (i <= Nx) && (j <= p.row_end);
        const size_t gcell = (size_t)(i + 2) + (size_t)P * (size_t)(j + 2);
        // G^- of this row was requested one iteration ago (row 0: before the tile wait); request the next row
        const double Gm0 = Gq0, Gm1 = Gq1, Gm2 = Gq2, Gm3 = Gq3;
        if constexpr (STAGE >= 2) {
            if (it >= 0 && it + 1 < R && (i <= Nx) && (j + 1 <= p.row_end)) {
                const size_t gn = gcell + (size_t)P;
                Gq0 = p.G[0][gn]; Gq1 = p.G[1][gn]; Gq2 = p.G[2][gn]; Gq3 = p.G[3][gn];
            }
        }
        const double vN = RAW(s_v, li, lj + 1), vWn = RAW(s_v, li - 1, lj + 1);
        const bool posN = vN > 0.0, bufN = ybuf(p.by, gj + 1, 3, p.NyG);
        const double es1 = eps * (12.0 / 13.0), es2 = eps * (24.0 / 13.0);
        double fyh_n, fyA_n;
        const double *const ph = &RAW(s_h, li, lj + 1), *const pA = &RAW(s_A, li, lj + 1);   // north face lj+1
        if (it < 0) {
            // south face of the warp's first row: h and A fluxes
            double d1[2], d2[2], d3[2], d4[2], c0[2] = {es1, es1}, c1[2] = {es1, es1}, c2[2] = {es1, es1}, num[2], den[2], rc[2];
            double ch, cA;
            UPDC(posN, ph, W, 0, ch); UPDC(posN, pA, W, 1, cA);
            beta_acc_n<2>(d1, d2, d3, d4, c0, c1, c2);
            corr_n<2>(d1, d2, d3, d4, c0, c1, c2, num, den);
            rcp_n<2>(den, rc);
            const double wh_ = vN * fma(num[0], rc[0], ch);
            const double wA_ = vN * fma(num[1], rc[1], cA);
            fyh_n = bufN ? vN * sym2(ph[-W], ph[0]) : wh_;
            fyA_n = bufN ? vN * sym2(pA[-W], pA[0]) : wA_;
        } else {
            const double vhat = avg4(vW, vC, vWn, vN);
            const double uw = RAW(s_u, li, lj), ue = RAW(s_u, li + 1, lj);
            const double uhat = avg4(RAW(s_u, li, lj - 1), RAW(s_u, li + 1, lj - 1), uw, ue);
            const bool posv = vhat > 0.0, posu = uhat > 0.0, posx = uw > 0.0;
            const int offx = (lj - 1) * ZP + (posu ? li - 2 : li + 3) - 1;   // zeta to the centre i along x
            const int sx = posu ? 1 : -1;
            const double *qh = posx ? &RAW(s_h, li - 3, lj) : &RAW(s_h, li + 2, lj);
            const double *qA = posx ? &RAW(s_A, li - 3, lj) : &RAW(s_A, li + 2, lj);
            const int sp = posx ? 1 : -1;
            const double *const pz = &Zf(s_z, li, lj + 1), *const pu = &Zf(s_ut, li, lj + 1), *const pv = &Zf(s_vt, li, lj + 1);
            double rc[8], num2[2], num4[4], cz0, cz1, ch0, cA1, ch2, cA3;
            {   // vorticity pair: [0] to the centre of row lj along y, [1] to the centre i along x;
                // VelocityStencil smoothness: beta accumulated over ℑy u and ℑx v
                double d1[2], d2[2], d3[2], d4[2], c0[2] = {es2, es2}, c1[2] = {es2, es2}, c2[2] = {es2, es2}, den[2];
                UPD(posv, pu, ZP, 0); diffs_mem(s_ut + offx, sx, d1[1], d2[1], d3[1], d4[1]);
                beta_acc_n<2>(d1, d2, d3, d4, c0, c1, c2);
                UPD(posv, pv, ZP, 0); diffs_mem(s_vt + offx, sx, d1[1], d2[1], d3[1], d4[1]);
                beta_acc_n<2>(d1, d2, d3, d4, c0, c1, c2);
                UPDC(posv, pz, ZP, 0, cz0); diffs_mem(s_z + offx, sx, d1[1], d2[1], d3[1], d4[1], cz1);
                corr_n<2>(d1, d2, d3, d4, c0, c1, c2, num2, den);
                rc[0] = den[0]; rc[1] = den[1];
            }
            {   // flux quartet: h, A through the north face, h, A through the west face
                double d1[4], d2[4], d3[4], d4[4], c0[4] = {es1, es1, es1, es1}, c1[4] = {es1, es1, es1, es1}, c2[4] = {es1, es1, es1, es1}, den[4];
                UPDC(posN, ph, W, 0, ch0); UPDC(posN, pA, W, 1, cA1);
                diffs_mem(qh, sp, d1[2], d2[2], d3[2], d4[2], ch2);
                diffs_mem(qA, sp, d1[3], d2[3], d3[3], d4[3], cA3);
                beta_acc_n<4>(d1, d2, d3, d4, c0, c1, c2);
                corr_n<4>(d1, d2, d3, d4, c0, c1, c2, num4, den);
                rc[2] = den[0]; rc[3] = den[1]; rc[4] = den[2]; rc[5] = den[3];
            }
            const double hc = ph[-W], hw_ = RAW(s_h, li - 1, lj), hs = RAW(s_h, li, lj - 1);
            rc[6] = 0.5 * (hw_ + hc);                                   // ℑx h
            rc[7] = 0.5 * (hs + hc);                                    // ℑy h
            {
                double x[8];
#pragma unroll
                for (int n = 0; n < 8; n++) x[n] = rc[n];
                rcp_n<8>(x, rc);
            }
            const double zy = fma(num2[0], rc[0], cz0);
            const double zx = fma(num2[1], rc[1], cz1);
            const double wh_ = vN * fma(num4[0], rc[2], ch0);
            const double wA_ = vN * fma(num4[1], rc[3], cA1);
            const double fxh = uw * fma(num4[2], rc[4], ch2);
            const double fxA = uw * fma(num4[3], rc[5], cA3);
            fyh_n = bufN ? vN * sym2(hc, ph[0]) : wh_;
            fyA_n = bufN ? vN * sym2(pA[-W], pA[0]) : wA_;
            const double adv_u = vhat * (ybuf(p.by, gj + 1, 3, p.NyG + 1) ? sym2(pz[-ZP], pz[0]) : zy);
            const double adv_v = uhat * zx;
            // east faces: the neighbour lane's west face; lane 31 takes the pre-pass value
            double fxh_e = __shfl_down_sync(0xffffffffu, fxh, 1), fxA_e = __shfl_down_sync(0xffffffffu, fxA, 1);
            const double eh = __shfl_sync(0xffffffffu, eastF, it), eA = __shfl_sync(0xffffffffu, eastF, R + it);
            if (lane == 31) { fxh_e = eh; fxA_e = eA; }

            double Gn0, Gn1, Gn2, Gn3;
            const double Ac = pA[-W], An = pA[0], As = RAW(s_A, li, lj - 1);
            const double Aw = RAW(s_A, li - 1, lj), Awn = RAW(s_A, li - 1, lj + 1), Aws = RAW(s_A, li - 1, lj - 1);
            {   // Gu at fcc — lorentz_force_func_x, sw_mhd_jacobian_functions.jl:10-13,20-22 — and
                // Gv at cfc — lorentz_force_func_y, :15-18,24-26 — advanced together
                const double Kc = Cc(s_K, li, lj);
                const double dKx = (Kc - Cc(s_K, li - 1, lj)) * p.rdx;
                const double dKy = (Kc - Cc(s_K, li, lj - 1)) * p.rdy;
                const double pgx = p.g * ((hc - hw_) * p.rdx);
                const double pgy = p.g * ((hc - hs) * p.rdy);
                const double dxA = (Ac - Aw) * p.rdx;
                const double dyA = (Ac - As) * p.rdy;
                // ℑxy(∂y F) telescopes to (F(i-1,j+1) + F(i,j+1) - F(i-1,j-1) - F(i,j-1)) / (4 dy)
                const double m1x = ((Cc(s_Bx, li - 1, lj + 1) + Cc(s_Bx, li, lj + 1)) - (Cc(s_Bx, li - 1, lj - 1) + Cc(s_Bx, li, lj - 1))) * (0.25 * p.rdy);
                const double m1y = ((RAW(s_A, li + 1, lj - 1) + RAW(s_A, li + 1, lj)) - (Aws + Aw)) * (0.25 * p.rdx);
                const double m2x = ((Awn + An) - (Aws + As)) * (0.25 * p.rdy);
                const double m2y = ((Cc(s_By, li + 1, lj - 1) + Cc(s_By, li + 1, lj)) - (Cc(s_By, li - 1, lj - 1) + Cc(s_By, li - 1, lj))) * (0.25 * p.rdx);
                const double jacx = fma(dxA, m1x, -(m2x * ((Cc(s_Bx, li, lj) - Cc(s_Bx, li - 1, lj)) * p.rdx)));
                const double jacy = fma(m1y, (Cc(s_By, li, lj) - Cc(s_By, li, lj - 1)) * p.rdy, -(dyA * m2y));
                Gn0 = fma(jacx, rc[6], fma(p.f, vhat, (adv_u - dKx) - pgx));
                Gn1 = fma(jacy, rc[7], fma(-p.f, uhat, (-adv_v - dKy) - pgy));
                if (p.by && gj < 2) Gn1 = 0.0;                          // wall rows of a Bounded-y grid keep v = 0
            }
            {   // Gh, GA at ccc: flux divergences (metric factors folded in: Ax/Az = 1/dx, Ay/Az = 1/dy)
                Gn2 = -fma(fxh_e - fxh, p.rdx, (fyh_n - fyh_s) * p.rdy);
                const double d = fma(fxA_e - fxA, p.rdx, (fyA_n - fyA_s) * p.rdy);
                const double dv = fma(ue - uw, p.rdx, (vN - vC) * p.rdy);
                Gn3 = -d + Ac * dv;
            }
            // ---- RK3 substep + stores --------------------------------------------------------------
            if (active) {
                const double Gn[4] = {Gn0, Gn1, Gn2, Gn3};
                const double Gm[4] = {Gm0, Gm1, Gm2, Gm3};
                const double Uc[4] = {uw, vC, hc, Ac};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if constexpr (STAGE == 1) {
                        p.Un[k][gcell] = Uc[k] + p.dtgam * Gn[k];
                        p.G[k][gcell] = Gn[k];
                    } else {
                        p.Un[k][gcell] = Uc[k] + p.dt * (p.gam * Gn[k] + p.zet * Gm[k]);
                        if constexpr (STAGE == 2) p.G[k][gcell] = Gn[k];
                    }
                }
            }
        }
        // slide to the next row
        fyh_s = fyh_n; fyA_s = fyA_n; vC = vN; vW = vWn;
    }

    // ---- fused diagnostics of the state at the start of the step --------------------------
    // Warp-private like the rest: every thread evaluates the squared face fields of its own column for
    // its R cells from the raw tile (2R+3 reciprocals advanced together), the east neighbours come by
    // shuffle, the tile's east column from lanes 0..R+1.
    if constexpr (DIAG) {
        double *const s_red = smem + DERIVED + 2;            // NW x NDIAG partials (behind the mbarrier word)
        constexpr int NR = R + 2;                            // rows b = lj0-1 .. lj0+R, index k = b - (lj0-1)
        auto sq = [](double x) { return x * x; };
        double Ac[NR], Aw[NR], Ae[NR], hcol[NR], sqy[NR], sqx[NR];
        {
            double den[2 * NR - 1], rc[2 * NR - 1], hwst[NR];
#pragma unroll
            for (int k = 0; k < NR; k++) {
                const int b = lj0 - 1 + k;
                Ac[k] = RAW(s_A, li, b); Aw[k] = RAW(s_A, li - 1, b); Ae[k] = RAW(s_A, li + 1, b);
                hcol[k] = RAW(s_h, li, b); hwst[k] = RAW(s_h, li - 1, b);
            }
#pragma unroll
            for (int k = 0; k < NR; k++) den[k] = 0.5 * (hwst[k] + hcol[k]);                 // ℑx h at fcc(li, b)
#pragma unroll
            for (int k = 1; k < NR; k++) den[NR + k - 1] = 0.5 * (hcol[k - 1] + hcol[k]);     // ℑy h at cfc(li, b)
            rcp_n<2 * NR - 1>(den, rc);
#pragma unroll
            for (int k = 0; k < NR; k++) sqy[k] = sq(((Ac[k] - Aw[k]) * p.rdx) * rc[k]);      // (dxA / ℑx h)^2
            sqx[0] = 0.0;
#pragma unroll
            for (int k = 1; k < NR; k++) sqx[k] = sq(-((Ac[k] - Ac[k - 1]) * p.rdy) * rc[NR + k - 1]);   // (dyA / ℑy h)^2
        }
        double uu[R], vc[R + 1], kb[R];
        {
            double vw2[R + 1], vc2[R + 1];
#pragma unroll
            for (int k = 0; k <= R; k++) { vc[k] = RAW(s_v, li, lj0 + k); vc2[k] = sq(vc[k]); vw2[k] = sq(RAW(s_v, li - 1, lj0 + k)); }
#pragma unroll
            for (int r = 0; r < R; r++) {                    // KE bracket u^2 + ℑxyᶠᶜᵃ(v^2) at fcc(li)
                uu[r] = RAW(s_u, li, lj0 + r);
                kb[r] = sq(uu[r]) + avg4(vw2[r], vc2[r], vw2[r + 1], vc2[r + 1]);
            }
        }
        // the tile's east column (a = TX+3), lane l takes row k = l
        double e_sqy = 0.0, e_kb = 0.0;
        if (lane < NR) {
            const int b = lj0 - 1 + lane, a = TX + 3;
            e_sqy = sq(((RAW(s_A, a, b) - RAW(s_A, a - 1, b)) * p.rdx) * frcp(0.5 * (RAW(s_h, a - 1, b) + RAW(s_h, a, b))));
            if (lane < R) {
                const int bb = lj0 + lane;
                e_kb = sq(RAW(s_u, a, bb)) + avg4(sq(RAW(s_v, a - 1, bb)), sq(RAW(s_v, a, bb)), sq(RAW(s_v, a - 1, bb + 1)), sq(RAW(s_v, a, bb + 1)));
            }
        }
        double sqye[NR], kbe[R];
#pragma unroll
        for (int k = 0; k < NR; k++) {
            sqye[k] = __shfl_down_sync(0xffffffffu, sqy[k], 1);
            const double t = __shfl_sync(0xffffffffu, e_sqy, k);
            if (lane == 31) sqye[k] = t;
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            kbe[r] = __shfl_down_sync(0xffffffffu, kb[r], 1);
            const double t = __shfl_sync(0xffffffffu, e_kb, r);
            if (lane == 31) kbe[r] = t;
        }
        double mb[R + 1];                                    // ME bracket Bx^2 + ℑxyᶜᶠᵃ(By^2) at cfc(li, b), b = lj0 .. lj0+R
#pragma unroll
        for (int k = 1; k < NR; k++) mb[k - 1] = sqx[k] + avg4(sqy[k - 1], sqye[k - 1], sqy[k], sqye[k]);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int j = jc0 + r, k = r + 1;
            if ((i <= Nx) && (j <= p.row_end)) {
                const double hh = hcol[k], aa = Ac[k];
                dg[0] += (0.5 * hh) * (0.5 * (kb[r] + kbe[r]));
                dg[1] += (0.5 * hh) * (0.5 * (mb[r] + mb[r + 1]));
                const double dh = hh - p.h_ref;
                dg[2] += (0.5 * p.g) * (dh * dh);
                dg[3] += hh;
                dg[4] = fmax(dg[4], fabs(uu[r]));
                dg[5] = fmax(dg[5], fabs(aa));
                dg[6] = fmax(dg[6], -hh);
                {   // div(hB) at ccc, telescoped ℑxy∂ (pure round-off)
                    const double hbx0 = (Aw[k - 1] + Ac[k - 1]) - (Aw[k + 1] + Ac[k + 1]), hbx1 = (Ac[k - 1] + Ae[k - 1]) - (Ac[k + 1] + Ae[k + 1]);
                    const double hby0 = (Ae[k - 1] + Ae[k]) - (Aw[k - 1] + Aw[k]), hby1 = (Ae[k] + Ae[k + 1]) - (Aw[k] + Aw[k + 1]);
                    dg[7] = fmax(dg[7], fabs((hbx1 - hbx0) * (0.25 * p.rdy) * p.rdx + (hby1 - hby0) * (0.25 * p.rdx) * p.rdy));
                }
                if (!(isfinite(hh) && isfinite(aa) && isfinite(uu[r]) && isfinite(vc[r]))) dg[8] += 1.0;
            }
        }
        // fixed-order reduction: R cells per thread (above), warp tree, then the warps in order
#pragma unroll
        for (int q = 0; q < NDIAG; q++) {
            const bool is_max = (q >= 4 && q <= 7);
            const double x = is_max ? warp_max(dg[q]) : warp_sum(dg[q]);
            if (lane == 0) s_red[wp * NDIAG + q] = x;
        }
        __syncthreads();
        if (tid < NDIAG) {
            const bool is_max = (tid >= 4 && tid <= 7);
            double acc = s_red[tid];
            for (int w2 = 1; w2 < NW; w2++) { const double x = s_red[w2 * NDIAG + tid]; acc = is_max ? fmax(acc, x) : acc + x; }
            // partial slots are indexed by 8-row tile rows (the launch granularity of the host side): this
            // tile fills its first slot and neutral elements into the others it covers
            const int tr8 = row0 / 8;
            p.diag[((size_t)tr8 * tiles_x + tile_x) * NDIAG + tid] = acc;
            for (int s = 1; s < TYB / 8; s++)
                if (row0 + 8 * s < p.row_end) p.diag[((size_t)(tr8 + s) * tiles_x + tile_x) * NDIAG + tid] = (tid == 6) ? -INFINITY : 0.0;
        }
    }
}

template <int STAGE, bool DIAG>
cudaError_t launch_rb(const KParams &p, cudaStream_t st) {
    auto kern = substage_rb_kernel<STAGE, DIAG>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles_x = (p.Nx + TX - 1) / TX;
    const int tiles_y = (p.row_end - p.row_begin + TYB - 1) / TYB;
    static int ahead = -1;
    if (ahead < 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        const char *env = getenv("SWMHD_L2_AHEAD");
        ahead = env ? atoi(env) : (occ < 1 ? 1 : occ) * sms;   // CTAs in flight = distance to the slot's next tile
    }
    KParams q = p;
    q.l2_ahead = ahead;
    kern<<<tiles_x * tiles_y, NT, SMEM_BYTES, st>>>(q);
    return cudaGetLastError();
}

} // namespace

void substage_rb_tile(int *tx, int *ty) { *tx = TX; *ty = TYB; }

cudaError_t launch_substage_rb(const KParams &p0, int stage, cudaStream_t st) {
    if (p0.tile_rows <= 0) return cudaSuccess;
    if (!p0.use_rb) return cudaErrorInvalidValue;
    int tx8, ty8;
    substage_tile(&tx8, &ty8);                              // launch granularity of the host side (tile rows)
    KParams p = p0;
    p.row_begin = p0.tile_row0 * ty8;
    p.row_end = (p0.tile_row0 + p0.tile_rows) * ty8;
    if (p.row_end > p.Ny) p.row_end = p.Ny;
    if (p.row_end <= p.row_begin) return cudaSuccess;
    const bool dg = (p.diag != nullptr);
    if (dg && stage != 1) return cudaErrorInvalidValue;
    switch (stage) {
        case 1: return dg ? launch_rb<1, true>(p, st) : launch_rb<1, false>(p, st);
        case 2: return launch_rb<2, false>(p, st);
        case 3: return launch_rb<3, false>(p, st);
    }
    return cudaErrorInvalidValue;
}

} // namespace swmhd
