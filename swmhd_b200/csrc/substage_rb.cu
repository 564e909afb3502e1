// substage_rb.cu — row-blocked fused RK3-substage kernels (sm_100a), FAST arithmetic, both formulations.
//
// Same work per launch as substage_kernel.cu (calculate_tendencies! + rk3_substep! + store_tendencies!
// for all four fields with the reference's Lorentz closures inlined — substage_rb_kernel:
// jacobian_formulation/sw_mhd_jacobian_functions.jl:1-26; substage_rbd_kernel:
// divergence_formulation/sw_mhd_divergence_functions.jl:1-170 — and in the stage-1 DIAG variants the
// diagnostics of SWMHD_example.jl:47-77 / divergence_sw_mhd.jl:42-75), different thread mapping:
//
//   * a CTA owns a 32 x RB_TY tile staged by TMA, a WARP owns RB_R consecutive rows of it and walks
//     them south to north, a THREAD one column of those rows.
//   * after the derived staggered fields (phase A, CTA-wide) warps never meet again: the fluxes through
//     a cell's north face are kept in registers for the next row (one extra evaluation per warp for the
//     south face of its first row), the east-face fluxes travel by warp shuffle (Jacobian: the tile's
//     east column of faces is a per-warp pre-pass; divergence: lane 31 only supplies them).
//     One __syncthreads per tile (one more in the DIAG variants) instead of three, no flux arrays.
//   * phase A keeps only what several threads share and what is expensive: anything a thread can rebuild
//     from values it holds anyway (kinetic energy, reciprocals of face depths) stays in registers, so that
//     the tile fits four times into an SM (55 KB of shared memory, <= 128 registers: 16 warps per SM).
//   * the next tile of the SM slot is pulled into L2 while this one computes (cp.async.bulk.prefetch).
//
// Measured at 4096^2 (profiles/README.md): Jacobian 0.67 / 0.71 / 0.70 ms per stage against 0.756 / 0.782 /
// 0.743 ms of the one-thread-per-cell kernel, divergence 0.74 / 0.77 / 0.76 against 0.93 / 0.94 / 0.88 ms.
// Variants that were measured and dropped (y stencils in sliding register windows, unrolled row loop, other
// tile shapes) are in the history table there.
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
#ifndef RB_MINB            // CTAs per SM the Jacobian kernel is compiled for (<= 128 registers, 55.6 KB of shared memory)
#define RB_MINB 4
#endif
#ifndef RB_MINB_D          // the divergence kernel: 55.4 KB of shared memory
#define RB_MINB_D 4
#endif
// Register cap of the variants WITHOUT fused diagnostics: compiled for 3 CTAs/SM (<= 168 registers) ptxas still ends at
// 114-126 registers — 4 CTAs/SM stay resident (tests/test_build_properties.py checks <= 128) — but schedules with less
// rematerialisation: 1.909 vs 1.940 ms (Jacobian), 2.014 vs 2.037 ms (divergence) per step at 4096^2.  The DIAG variants
// need the cap of 4 (they would take 142-152 registers and lose the fourth CTA).
#ifndef RB_MINB_PLAIN
#define RB_MINB_PLAIN 3
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
// ffc points a in [1,TX+5], b in [1,TYB+5], stored with the pitch of the raw tile (one unused column): point q of the
// array is raw index q + W + 1, so phase A walks both with one counter, and a warp's 32 consecutive points never
// straddle a hole (the row wrap of a 37-wide array cost an extra shared-memory wavefront per straddling half-warp)
constexpr int ZP = TX + 6, ZR = TYB + 5;
constexpr int CP = TX + 2, CR = TYB + 2;                   // ccc points a in [2,TX+3], b in [2,TYB+3]
constexpr int NZ = ZP * ZR, NC = CP * CR;
constexpr int o_z = 0, o_ut = NZ, o_vt = 2 * NZ, o_Bx = 3 * NZ, o_By = o_Bx + NC;
constexpr int DERIVED = o_By + NC;
// per warp: the tile's east column of faces, read by lane 31 only: R h fluxes, R A fluxes; DIAG variant: R+2 squared
// By values and R kinetic-energy brackets.  The warp's DIAG partials (NDIAG doubles) reuse its slice at the end.
constexpr int NE = 4 * R + 2;
static_assert(NE >= NDIAG, "the per-warp scratch also holds the warp's diagnostic partials");
constexpr size_t SMEM_BYTES = ((size_t)4 * SZP + DERIVED + 2 + NW * NE) * sizeof(double);   // + mbarrier + per-warp scratch
static_assert(RB_MINB * (SMEM_BYTES + 1024) <= 228 * 1024, "Jacobian kernel: shared memory of RB_MINB CTAs per SM (1 KB reserved per CTA)");

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
    double d1[1] = {b - a}, d2[1] = {c - b}, d3[1] = {d - c}, d4[1] = {e - d};
    double c0[1] = {es}, c1[1] = {es}, c2[1] = {es}, num[1], den[1], rc[1];
    beta_acc_n<1>(d1, d2, d3, d4, c0, c1, c2);
    corr10_n<1>(d1, d2, d3, d4, c0, c1, c2, num, den);
    rcp_mix_n<1, 1>(den, rc);
    return fma(num[0], rc[0], c);
}
// vel * psi_upwind at face f (ctr = &psi[f]): left-biased psi[f-3..f+1] for vel > 0, else the mirror
__device__ __forceinline__ double upwind_weno_mem(const double *ctr, double vel, double eps) {
    const bool pos = gt0(vel);
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
__global__ void __launch_bounds__(NT, DIAG ? RB_MINB : RB_MINB_PLAIN) substage_rb_kernel(const __grid_constant__ KParams p) {
    // 128-byte aligned for the TMA destination; used directly so that the compiler keeps the shared
    // address space (LDS/STS instead of generic LD/ST).
    extern __shared__ __align__(128) unsigned char smem_bytes[];
    double *const s_u = reinterpret_cast<double *>(smem_bytes);
    double *const s_v = s_u + SZP, *const s_h = s_u + 2 * SZP, *const s_A = s_u + 3 * SZP;
    double *const smem = s_u + 4 * SZP;
    double *const s_z = smem + o_z, *const s_ut = smem + o_ut, *const s_vt = smem + o_vt;
    double *const s_Bx = smem + o_Bx, *const s_By = smem + o_By;
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
    static_assert(ZP == W, "derived ffc arrays share the pitch of the raw tile");
#pragma unroll 4
    for (int q = tid; q < NZ; q += NT) {                    // zeta, ℑy u, ℑx v at ffc; raw index of point q is q + W + 1
        const int r = q + W + 1;
        const double vc = s_v[r], vw = s_v[r - 1], uc = s_u[r], us = s_u[r - W];
        s_z[q] = fma(vc - vw, p.rdx, (us - uc) * p.rdy);
        s_ut[q] = 0.5 * (us + uc);
        s_vt[q] = 0.5 * (vw + vc);
    }
    {   // Bx, By at ccc (sw_mhd_jacobian_functions.jl:1-7)
        constexpr int DA = NT % CP, DB = NT / CP;
        int a = tid % CP, r = (tid / CP + 2) * W + a + 2;
        const double hry = 0.5 * p.rdy, hrx = 0.5 * p.rdx;
#pragma unroll 3
        for (int q = tid; q < NC; q += NT) {
            const double rh = frcp(s_h[r]);
            s_Bx[q] = ((s_A[r - W] - s_A[r + W]) * hry) * rh;
            s_By[q] = ((s_A[r + 1] - s_A[r - 1]) * hrx) * rh;
            a += DA; r += DB * W + DA;
            if (a >= CP) { a -= CP; r += W - CP; }
        }
    }
    __syncthreads();

    // ---- B/C: warp-private from here on --------------------------------------------------------------

    // warps of a ragged last tile row (or of an 8-row edge strip) that own no cell skip the row walk
    const bool warp_has_rows = (jc0 <= p.row_end);

    // B0: the tile's east column (a = TX+3) for this warp's rows, into the warp's scratch slice (lane 31 reads it in
    // the row walk): lanes 0..R-1 the h fluxes, lanes R..2R-1 the A fluxes; DIAG: lanes 2R..3R+1 the squared By of
    // rows lj0-1..lj0+R, lanes 3R+2..4R+1 the kinetic-energy brackets of rows lj0..lj0+R-1
    double *const s_e = smem + DERIVED + 2 + wp * NE;
    if (warp_has_rows) {
        constexpr int a = TX + 3;
        if (lane < 2 * R) {
            const int b = lj0 + (lane & (R - 1));
            const double *arr = (lane < R) ? s_h : s_A;
            s_e[lane] = upwind_weno_mem(&RAW(arr, a, b), RAW(s_u, a, b), eps);
        }
        if constexpr (DIAG) {
            if (lane >= 2 * R && lane < 3 * R + 2) {
                const int b = lj0 - 1 + (lane - 2 * R);
                const double by_ = ((RAW(s_A, a, b) - RAW(s_A, a - 1, b)) * p.rdx) * frcp(0.5 * (RAW(s_h, a - 1, b) + RAW(s_h, a, b)));
                s_e[lane] = by_ * by_;
            } else if (lane >= 3 * R + 2 && lane < 4 * R + 2) {
                const int b = lj0 + (lane - (3 * R + 2));
                const double v00 = RAW(s_v, a - 1, b), v10 = RAW(s_v, a, b), v01 = RAW(s_v, a - 1, b + 1), v11 = RAW(s_v, a, b + 1), ua = RAW(s_u, a, b);
                s_e[lane] = fma(ua, ua, avg4(v00 * v00, v10 * v10, v01 * v01, v11 * v11));
            }
        }
        __syncwarp();
    }

    // Row loop.  Iteration it = -1 only produces the fluxes through the south face of the warp's first
    // row; every other iteration evaluates the NORTH face lj+1 of its row and keeps it for the next one.
    double vC = RAW(s_v, li, lj0 - 1), vW = RAW(s_v, li - 1, lj0 - 1);
    double fyh_s = 0.0, fyA_s = 0.0, K_s = 0.0;
    const double es1 = eps * (12.0 / 13.0), es2 = eps * (24.0 / 13.0);
    const bool col_ok = (i <= Nx);
    // DIAG: energy sums accumulated in the row walk (they need the reciprocal face depths it holds anyway);
    // sqy_s = By^2 at fcc (li, lj-1), meb_s = ME bracket at cfc (li, lj-1): one row behind
    double ke_acc = 0.0, me_acc = 0.0, sqy_s = 0.0, meb_s = 0.0;
    (void)ke_acc; (void)me_acc; (void)sqy_s; (void)meb_s;
    size_t gcell = (size_t)(i + 2) + (size_t)P * (size_t)(jc0 + 1);   // cell (i, jc0 - 1): advanced by one row per iteration

#pragma unroll 1
    for (int it = warp_has_rows ? -1 : R; it < R; ++it) {
        const int lj = lj0 + it;
        const int j = jc0 + it, gj = gj0 + it;
        const bool active = col_ok && (j <= p.row_end);
        const double vN = RAW(s_v, li, lj + 1), vWn = RAW(s_v, li - 1, lj + 1);
        const bool posN = gt0(vN);
        double fyh_n, fyA_n;
        const double *const ph = &RAW(s_h, li, lj + 1), *const pA = &RAW(s_A, li, lj + 1);   // north face lj+1
        // kinetic energy K at ccc (li, lj), in registers: the row walk keeps it as the next row's K(li, lj-1)
        // (a phase-A array of K would push the tile over the shared memory of 4 CTAs per SM)
        const double uw = RAW(s_u, li, lj), ue = RAW(s_u, li + 1, lj);
        const double Kc = 0.25 * (fma(uw, uw, ue * ue) + fma(vC, vC, vN * vN));
        if (it < 0) {
            // south face of the warp's first row: h and A fluxes
            double d1[2], d2[2], d3[2], d4[2], c0[2] = {es1, es1}, c1[2] = {es1, es1}, c2[2] = {es1, es1}, num[2], den[2], rc[2];
            double ch, cA;
            UPDC(posN, ph, W, 0, ch); UPDC(posN, pA, W, 1, cA);
            beta_acc_n<2>(d1, d2, d3, d4, c0, c1, c2);
            corr10_n<2>(d1, d2, d3, d4, c0, c1, c2, num, den);
            if constexpr (DIAG) {   // By^2 at fcc (li, lj0-1) seeds the magnetic-energy bracket of the first row
                double x3[3] = {den[0], den[1], 0.5 * (RAW(s_h, li - 1, lj) + ph[-W])}, r3[3];
                rcp_mix_n<3, 2>(x3, r3);
                rc[0] = r3[0]; rc[1] = r3[1];
                const double by_ = ((pA[-W] - RAW(s_A, li - 1, lj)) * p.rdx) * r3[2];
                sqy_s = by_ * by_;
            } else {
                rcp_mix_n<2, 2>(den, rc);
            }
            fyh_n = vN * fma(num[0], rc[0], ch);
            fyA_n = vN * fma(num[1], rc[1], cA);
        } else {
            const double vhat = avg4(vW, vC, vWn, vN);
            const double uhat = avg4(RAW(s_u, li, lj - 1), RAW(s_u, li + 1, lj - 1), uw, ue);
            const bool posv = gt0(vhat), posu = gt0(uhat), posx = gt0(uw);
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
                corr10_n<2>(d1, d2, d3, d4, c0, c1, c2, num2, den);
                rc[0] = den[0]; rc[1] = den[1];
            }
            {   // flux quartet: h, A through the north face, h, A through the west face
                double d1[4], d2[4], d3[4], d4[4], c0[4] = {es1, es1, es1, es1}, c1[4] = {es1, es1, es1, es1}, c2[4] = {es1, es1, es1, es1}, den[4];
                UPDC(posN, ph, W, 0, ch0); UPDC(posN, pA, W, 1, cA1);
                diffs_mem(qh, sp, d1[2], d2[2], d3[2], d4[2], ch2);
                diffs_mem(qA, sp, d1[3], d2[3], d3[3], d4[3], cA3);
                beta_acc_n<4>(d1, d2, d3, d4, c0, c1, c2);
                corr10_n<4>(d1, d2, d3, d4, c0, c1, c2, num4, den);
                rc[2] = den[0]; rc[3] = den[1]; rc[4] = den[2]; rc[5] = den[3];
            }
            const double hc = ph[-W], hw_ = RAW(s_h, li - 1, lj), hs = RAW(s_h, li, lj - 1);
            rc[6] = 0.5 * (hw_ + hc);                                   // ℑx h
            rc[7] = 0.5 * (hs + hc);                                    // ℑy h
            {
                double x[8];
#pragma unroll
                for (int n = 0; n < 8; n++) x[n] = rc[n];
                rcp_mix_n<8, 6>(x, rc);                                 // six WENO denominators, two face depths
            }
            fyh_n = vN * fma(num4[0], rc[2], ch0);
            fyA_n = vN * fma(num4[1], rc[3], cA1);
            const double fxh = uw * fma(num4[2], rc[4], ch2);
            const double fxA = uw * fma(num4[3], rc[5], cA3);
            double adv_u = vhat * fma(num2[0], rc[0], cz0);
            const double adv_v = uhat * fma(num2[1], rc[1], cz1);
            // east faces: the neighbour lane's west face; lane 31 takes the pre-pass value
            double fxh_e = __shfl_down_sync(0xffffffffu, fxh, 1), fxA_e = __shfl_down_sync(0xffffffffu, fxA, 1);
            if (lane == 31) { fxh_e = s_e[it]; fxA_e = s_e[R + it]; }

            double Gn0, Gn1, Gn2, Gn3;
            const double Ac = pA[-W], An = pA[0], As = RAW(s_A, li, lj - 1);
            const double Aw = RAW(s_A, li - 1, lj), Awn = RAW(s_A, li - 1, lj + 1), Aws = RAW(s_A, li - 1, lj - 1);
            {   // Gu at fcc — lorentz_force_func_x, sw_mhd_jacobian_functions.jl:10-13,20-22 — and
                // Gv at cfc — lorentz_force_func_y, :15-18,24-26 — advanced together.
                // ℑxy(∂y F) telescopes to (F(i-1,j+1) + F(i,j+1) - F(i-1,j-1) - F(i,j-1)) / (4 dy); the common
                // factor 1/(4 dx dy) of both Jacobians multiplies the reciprocal depth once.
                const double uww = RAW(s_u, li - 1, lj);
                const double Kw = 0.25 * (fma(uww, uww, uw * uw) + fma(vW, vW, vWn * vWn)); // K at ccc (li-1, lj)
                const double S1x = (Cc(s_Bx, li - 1, lj + 1) + Cc(s_Bx, li, lj + 1)) - (Cc(s_Bx, li - 1, lj - 1) + Cc(s_Bx, li, lj - 1));
                const double S1y = (RAW(s_A, li + 1, lj - 1) + RAW(s_A, li + 1, lj)) - (Aws + Aw);
                const double S2x = (Awn + An) - (Aws + As);
                const double S2y = (Cc(s_By, li + 1, lj - 1) + Cc(s_By, li + 1, lj)) - (Cc(s_By, li - 1, lj - 1) + Cc(s_By, li - 1, lj));
                const double jx = fma(Ac - Aw, S1x, -(S2x * (Cc(s_Bx, li, lj) - Cc(s_Bx, li - 1, lj))));
                const double jy = fma(S1y, Cc(s_By, li, lj) - Cc(s_By, li, lj - 1), -((Ac - As) * S2y));
                // -dx(K + g h), -dy(K + g h)
                const double tx = fma(p.g, hc - hw_, Kc - Kw);
                const double ty = fma(p.g, hc - hs, Kc - K_s);
                // Bounded-y: centred second order within the wall buffers (oracle ybuf), v = 0 on the wall rows
                if (p.by && ybuf(1, gj + 1, 3, p.NyG + 1)) adv_u = vhat * sym2(pz[-ZP], pz[0]);
                Gn0 = fma(jx, rc[6] * p.qrdxy, fma(p.f, vhat, fma(-p.rdx, tx, adv_u)));
                Gn1 = fma(jy, rc[7] * p.qrdxy, fma(-p.f, uhat, fma(-p.rdy, ty, -adv_v)));
                if (p.by && gj < 2) Gn1 = 0.0;
            }
            if (p.by && ybuf(1, gj + 1, 3, p.NyG)) { fyh_n = vN * sym2(hc, ph[0]); fyA_n = vN * sym2(Ac, An); }
            {   // Gh, GA at ccc: flux divergences (metric factors folded in: Ax/Az = 1/dx, Ay/Az = 1/dy)
                Gn2 = -fma(fxh_e - fxh, p.rdx, (fyh_n - fyh_s) * p.rdy);
                const double d = fma(fxA_e - fxA, p.rdx, (fyA_n - fyA_s) * p.rdy);
                const double dv = fma(ue - uw, p.rdx, (vN - vC) * p.rdy);
                Gn3 = fma(Ac, dv, -d);
            }
            // ---- RK3 substep + stores --------------------------------------------------------------
            if (active) {
                const double Gn[4] = {Gn0, Gn1, Gn2, Gn3};
                const double Gm[4] = {Gq0, Gq1, Gq2, Gq3};
                const double Uc[4] = {uw, vC, hc, Ac};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if constexpr (STAGE == 1) {
                        p.Un[k][gcell] = fma(p.dtgam, Gn[k], Uc[k]);
                        p.G[k][gcell] = Gn[k];
                    } else {
                        p.Un[k][gcell] = fma(p.dtgam, Gn[k], fma(p.dtzet, Gm[k], Uc[k]));
                        if constexpr (STAGE == 2) p.G[k][gcell] = Gn[k];
                    }
                }
            }
            if constexpr (DIAG) {
                // KE bracket u^2 + ℑxyᶠᶜᵃ(v^2) at fcc (li, lj); ME bracket Bx^2 + ℑxyᶜᶠᵃ(By^2) at cfc (li, lj) with
                // Bx = -dyA / ℑy h at cfc, By = dxA / ℑx h at fcc (SWMHD_example.jl:70-75); the east
                // neighbours come from lane+1, lane 31 takes the tile's east column from the warp's scratch slice
                const double byf = ((Ac - Aw) * p.rdx) * rc[6], bxf = ((Ac - As) * p.rdy) * rc[7];
                const double sqy = byf * byf;
                const double csum = sqy_s + sqy;
                const double keb = fma(uw, uw, avg4(vW * vW, vC * vC, vWn * vWn, vN * vN));
                double csum_e = __shfl_down_sync(0xffffffffu, csum, 1), keb_e = __shfl_down_sync(0xffffffffu, keb, 1);
                if (lane == 31) { csum_e = s_e[2 * R + it] + s_e[2 * R + it + 1]; keb_e = s_e[3 * R + 2 + it]; }
                const double meb = fma(0.25, csum + csum_e, bxf * bxf);
                if (it >= 1 && col_ok && (j - 1 <= p.row_end)) me_acc = fma(0.25 * hs, meb_s + meb, me_acc);   // cell (li, lj-1): (h/2) (meb_s + meb)/2
                if (active) ke_acc = fma(0.25 * hc, keb + keb_e, ke_acc);               // cell (li, lj):   (h/2) (keb + keb_e)/2
                meb_s = meb; sqy_s = sqy;
            }
            // G^- of the next row: requested now, consumed at the end of the next iteration (the first row's was
            // requested before the tile wait)
            if constexpr (STAGE >= 2) {
                if (it + 1 < R && col_ok && (j + 1 <= p.row_end)) {
                    const size_t gn = gcell + (size_t)P;
                    Gq0 = p.G[0][gn]; Gq1 = p.G[1][gn]; Gq2 = p.G[2][gn]; Gq3 = p.G[3][gn];
                }
            }
        }
        if (it < 0 && p.by && ybuf(1, gj + 1, 3, p.NyG)) {     // south face of the first row inside a wall buffer
            fyh_n = vN * sym2(ph[-W], ph[0]); fyA_n = vN * sym2(pA[-W], pA[0]);
        }
        // slide to the next row
        fyh_s = fyh_n; fyA_s = fyA_n; K_s = Kc; vC = vN; vW = vWn;
        gcell += (size_t)P;
    }

    // ---- fused diagnostics of the state at the start of the step --------------------------
    // KE and ME were accumulated in the row walk.  Left: the magnetic-energy bracket north of the warp's last row
    // (two more reciprocal depths), and the quantities that need no reciprocal (PE, mass, extrema, div(hB), finite
    // flag), evaluated per thread for its R cells from the raw tile.
    if constexpr (DIAG) {
        double sums[4] = {0.0, 0.0, 0.0, 0.0};                // ke, me, pe, sum h
        const unsigned long long kneut = ord_key(-INFINITY);
        unsigned long long kmax[4] = {ord_key(0.0), ord_key(0.0), kneut, ord_key(0.0)};   // |u|, |A|, -h, |div hB| as ordered keys
        bool bad = false;
        if (warp_has_rows) {
            {   // ME of the cell (li, lj0+R-1): bracket at cfc (li, lj0+R)
                const int lt = lj0 + R;
                const double hN = RAW(s_h, li, lt), hC = RAW(s_h, li, lt - 1), AN = RAW(s_A, li, lt);
                double x2[2] = {0.5 * (RAW(s_h, li - 1, lt) + hN), 0.5 * (hC + hN)}, r2[2];
                rcp_n<2>(x2, r2);
                const double byf = ((AN - RAW(s_A, li - 1, lt)) * p.rdx) * r2[0], bxf = ((AN - RAW(s_A, li, lt - 1)) * p.rdy) * r2[1];
                const double csum = fma(byf, byf, sqy_s);
                double csum_e = __shfl_down_sync(0xffffffffu, csum, 1);
                if (lane == 31) csum_e = s_e[3 * R] + s_e[3 * R + 1];
                const double meb = fma(0.25, csum + csum_e, bxf * bxf);
                if (col_ok && (jc0 + R - 1 <= p.row_end)) me_acc = fma(0.25 * hC, meb_s + meb, me_acc);
            }
            sums[0] = ke_acc; sums[1] = me_acc;
            double Am[3], A0[3], Ap[3];                      // A at columns li-1, li, li+1 of rows lj-1, lj, lj+1 (sliding)
#pragma unroll
            for (int c = 0; c < 3; c++) { Am[c] = RAW(s_A, li - 1 + c, lj0 - 1); A0[c] = RAW(s_A, li - 1 + c, lj0); }
#pragma unroll
            for (int r = 0; r < R; r++) {            // branch-free: cells beyond the launch contribute neutral elements
                const int lj = lj0 + r;
#pragma unroll
                for (int c = 0; c < 3; c++) Ap[c] = RAW(s_A, li - 1 + c, lj + 1);
                const bool ok = col_ok && (jc0 + r <= p.row_end);
                const double hh = RAW(s_h, li, lj), uu = RAW(s_u, li, lj), vv = RAW(s_v, li, lj), aa = A0[1];
                const double dh = hh - p.h_ref;
                sums[2] = fma(ok ? (0.5 * p.g) * dh : 0.0, dh, sums[2]);
                sums[3] += ok ? hh : 0.0;
                kmax[0] = max(kmax[0], ok ? ord_key(fabs(uu)) : kneut);
                kmax[1] = max(kmax[1], ok ? ord_key(fabs(aa)) : kneut);
                kmax[2] = max(kmax[2], ok ? ord_key(-hh) : kneut);
                {   // div(hB) at ccc, telescoped ℑxy∂ (pure round-off)
                    const double hbx0 = (Am[0] + Am[1]) - (Ap[0] + Ap[1]), hbx1 = (Am[1] + Am[2]) - (Ap[1] + Ap[2]);
                    const double hby0 = (Am[2] + A0[2]) - (Am[0] + A0[0]), hby1 = (A0[2] + Ap[2]) - (A0[0] + Ap[0]);
                    const double dv = fabs((hbx1 - hbx0) * (0.25 * p.rdy) * p.rdx + (hby1 - hby0) * (0.25 * p.rdx) * p.rdy);
                    kmax[3] = max(kmax[3], ok ? ord_key(dv) : kneut);
                }
                // finite <=> exponent field below 0x7ff: one integer test on the OR of the four high words' exponent overflow
                const unsigned eh = (unsigned)__double2hiint(hh) & 0x7ff00000u, ea = (unsigned)__double2hiint(aa) & 0x7ff00000u;
                const unsigned eu = (unsigned)__double2hiint(uu) & 0x7ff00000u, ev = (unsigned)__double2hiint(vv) & 0x7ff00000u;
                bad = bad || (ok && (eh == 0x7ff00000u || ea == 0x7ff00000u || eu == 0x7ff00000u || ev == 0x7ff00000u));
#pragma unroll
                for (int c = 0; c < 3; c++) { Am[c] = A0[c]; A0[c] = Ap[c]; }
            }
        }
        // fixed-order reduction: R cells per thread (above), the warp (four sums by reduce-scatter, four extrema by REDUX
        // on ordered keys, the finite flag by vote), then the warps in order
        double *const s_red = smem + DERIVED + 2;            // warp w's partials at the start of its scratch slice
        __syncwarp();
        {
            const double tot = warp_sum4(sums, lane);
            unsigned long long km[4];
#pragma unroll
            for (int q = 0; q < 4; q++) km[q] = warp_max_key(kmax[q]);
            const bool anybad = __any_sync(0xffffffffu, bad);
            if ((lane & 7) == 0) s_red[wp * NE + (lane >> 3)] = tot;
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) s_red[wp * NE + 4 + q] = ord_val(km[q]);
                s_red[wp * NE + 8] = anybad ? 1.0 : 0.0;
            }
        }
        __syncthreads();
        if (tid < NDIAG) {
            const bool is_max = (tid >= 4 && tid <= 7);
            double acc = s_red[tid];
            for (int w2 = 1; w2 < NW; w2++) { const double x = s_red[w2 * NE + tid]; acc = is_max ? fmax(acc, x) : acc + x; }
            // partial slots are indexed by 8-row tile rows (the launch granularity of the host side): this
            // tile fills its first slot and neutral elements into the others it covers
            const int tr8 = row0 / 8;
            p.diag[((size_t)tr8 * tiles_x + tile_x) * NDIAG + tid] = acc;
            for (int s2 = 1; s2 < TYB / 8; s2++)
                if (row0 + 8 * s2 < p.row_end) p.diag[((size_t)(tr8 + s2) * tiles_x + tile_x) * NDIAG + tid] = (tid == 6) ? -INFINITY : 0.0;
        }
    }
}

// =====================================================================================================
// Divergence (conservative) formulation, FAST arithmetic: ConservativeFormulation tendencies
// with div_lorentz_x/y of divergence_formulation/sw_mhd_divergence_functions.jl:1-170 inlined.
//
// Same warp-private row walk.  Per row a thread evaluates the three x-type fluxes anchored at its column
// (F_uu - Lxx at the ccc point li-1, F_uv - Lxy at the ffc point li, the tracer transport flux at fcc li)
// and the three y-type fluxes of the row's NORTH side (F_vu - Lyx at ffc (li, lj+1), F_vv - Lyy at ccc
// (li, lj), the tracer flux at cfc (li, lj+1)), which stay in registers as the next row's south side.
// The east neighbours of the x-type fluxes come from lane+1 by shuffle; lane 31 only supplies them:
// a tile is 31 cells wide (32 faces), which costs 3 % of the lanes instead of a pre-pass over three
// flux families.
constexpr int TXD = TX - 1;                                // cells per tile row in the divergence kernel
// hBx, hBy, Bx, By: a in [1,TX+4], b in [1,TYB+4]; 1/h at ccc: a in [2,TX+3], b in [2,TYB+3].  All at the pitch of the raw
// tile (unused columns at the end of each row): point t of an array is raw index t + const, no holes inside a warp's 32 points
constexpr int BP = TX + 6, BR = TYB + 4;
constexpr int RP = TX + 6, RR = TYB + 2;
constexpr int NB = BP * BR, NRH = RP * RR;
constexpr int DERIVED_D = 4 * NB + NRH;                    // 55.4 KB with the raw tile: 4 CTAs per SM
constexpr size_t SMEM_BYTES_D = ((size_t)4 * SZP + DERIVED_D + 2 + NW * NDIAG) * sizeof(double);
static_assert(RB_MINB_D * (SMEM_BYTES_D + 1024) <= 228 * 1024, "divergence kernel: shared memory of RB_MINB_D CTAs per SM (1 KB reserved per CTA)");
#define Bf(arr, a, b) arr[((b) - 1) * BP + (a) - 1]
#define RH(a, b) s_rh[((b) - 2) * RP + (a) - 2]

__device__ __forceinline__ double sym4(double a, double b, double c, double d) {
    return fma(7.0 / 12.0, b + c, (-1.0 / 12.0) * (a + d));
}
// third-order biased interpolants of sw_mhd_divergence_functions.jl:25-35
__device__ __forceinline__ double third(double x2, double x5, double xm) {      // (2*x2 + 5*x5 - xm)/6
    return fma(1.0 / 3.0, x2, fma(5.0 / 6.0, x5, (-1.0 / 6.0) * xm));
}
__device__ __forceinline__ double thirdR(double xm, double x5, double x2) {     // (-xm + 5*x5 + 2*x2)/6
    return fma(-1.0 / 6.0, xm, fma(5.0 / 6.0, x5, (1.0 / 3.0) * x2));
}
__device__ __forceinline__ double upwind_sel(double vel, double L, double Rr) { return vel * (vel > 0.0 ? L : Rr); }

template <int STAGE, bool DIAG>
__global__ void __launch_bounds__(NT, DIAG ? RB_MINB_D : RB_MINB_PLAIN) substage_rbd_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) unsigned char smem_bytes[];
    double *const s_u = reinterpret_cast<double *>(smem_bytes);
    double *const s_v = s_u + SZP, *const s_h = s_u + 2 * SZP, *const s_A = s_u + 3 * SZP;
    double *const smem = s_u + 4 * SZP;
    double *const s_hBx = smem, *const s_hBy = smem + NB, *const s_Bx = smem + 2 * NB, *const s_By = smem + 3 * NB;
    double *const s_rh = smem + 4 * NB;                                    // 1/h at ccc (each used by F_uu and F_vv)
    uint64_t *const mbar = reinterpret_cast<uint64_t *>(smem + DERIVED_D);
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int Nx = p.Nx, P = p.P;
    const double eps = p.eps;
    const int tiles_x = (Nx + TXD - 1) / TXD;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int row0 = p.row_begin + tile_y * TYB;            // 0-based first cell row = parent row of b = 0

    // ---- P0: stage uh,vh,h,A with a 3-cell halo (same boxes as the Jacobian kernel) ------------------
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_fence_init();
        mbar_expect_tx(mbar, TILE_TX_BYTES);
#pragma unroll
        for (int k = 0; k < 4; k++) tma_load_2d(s_u + k * SZP, &p.tm_rb[k], (tile_x * TXD) & ~1, row0, mbar);
#if RB_L2_PREFETCH
        const int nxt = blockIdx.x + p.l2_ahead;             // the tile of the CTA that takes over this slot, into L2
        if (nxt < gridDim.x) {
#pragma unroll
            for (int k = 0; k < 4; k++) tma_prefetch_2d(&p.tm_rb[k], ((nxt % tiles_x) * TXD) & ~1, p.row_begin + (nxt / tiles_x) * TYB);
        }
#endif
    }
    // the box starts at an even parent column (16-byte aligned source address; an odd start raised an
    // illegal-instruction fault): tiles with an odd first column sit one column further right in the box
    const int li = lane + 3 + ((tile_x * TXD) & 1), lj0 = 3 + wp * R;   // tile-local column / first row of this thread
    const int i = tile_x * TXD + 1 + lane;                  // logical (1-based) column
    const bool own = (lane < TXD) && (i <= Nx);             // lane 31 only supplies east-neighbour fluxes
    const int jc0 = row0 + 1 + wp * R;                      // logical (1-based, slab-local) first row
    const int gj0 = p.gj0 + jc0;                            // global row of the first cell (wall logic)
    const int NyG = p.NyG, by = p.by;
    double Gq0 = 0.0, Gq1 = 0.0, Gq2 = 0.0, Gq3 = 0.0;      // G^- of the row in flight
    if constexpr (STAGE >= 2) {
        if (own && (jc0 <= p.row_end)) {
            const size_t g0 = (size_t)(i + 2) + (size_t)P * (size_t)(jc0 + 2);
            Gq0 = p.G[0][g0]; Gq1 = p.G[1][g0]; Gq2 = p.G[2][g0]; Gq3 = p.G[3][g0];
        }
    }
    __syncthreads();                                        // barrier initialised before anybody polls it
    mbar_wait(mbar, 0);

    // ---- A: hBx, hBy, Bx, By (sw_mhd_divergence_functions.jl:134-148) and the reciprocal depths ---------
    static_assert(BP == W && RP == W, "derived arrays share the pitch of the raw tile");
    {
        const double qy = 0.25 * p.rdy, qx = 0.25 * p.rdx;
#pragma unroll 2
        for (int t = tid; t < NB; t += NT) {                // raw index of point t is t + W + 1
            const int r = t + W + 1;
            // telescoped ℑxy∂: hBx = -(ℑxy ∂y A), hBy = ℑxy ∂x A
            const double Asw = s_A[r - W - 1], As = s_A[r - W], Aw = s_A[r - 1];
            const double hbx = ((Asw + As) - (s_A[r + W - 1] + s_A[r + W])) * qy;
            const double hby = ((s_A[r - W + 1] + s_A[r + 1]) - (Asw + Aw)) * qx;
            const double hc = s_h[r];
            double x2[2] = {0.5 * (s_h[r - 1] + hc), 0.5 * (s_h[r - W] + hc)}, r2[2];
            rcp_n<2>(x2, r2);
            s_hBx[t] = hbx; s_hBy[t] = hby;
            s_Bx[t] = hbx * r2[0]; s_By[t] = hby * r2[1];
        }
    }
    for (int q = tid; q < NRH; q += NT) s_rh[q] = frcp(s_h[q + 2 * W + 2]);     // raw index of point q is q + 2 W + 2
    __syncthreads();

    // ---- B/C: warp-private row walk ----------------------------------------------------------------------
    const double es = eps * (12.0 / 13.0);
    double Fvu_s = 0.0, Fvv_s = 0.0, Ty_s = 0.0, vq_s = 0.0;   // south side of the current row (from the previous iteration)
    double rhff_s = 0.0;                                       // 1/ℑxy h at the ffc point (li, lj), likewise
    const bool warp_has_rows = (jc0 <= p.row_end);             // ragged last tile row / 8-row edge strips
    size_t gcell = (size_t)(i + 2) + (size_t)P * (size_t)(jc0 + 1);   // cell (i, jc0 - 1): one row further per iteration
    // DIAG: the sums and the maximum that need reciprocal face depths are accumulated in the row walk, which holds them
    // anyway.  sqx_s = Bx^2 at cfc (li, lj) (from the previous iteration's north side), sqy_s = By^2 at fcc (li, lj-1),
    // meb_s = ME bracket at cfc (li, lj-1).
    double ke_acc = 0.0, me_acc = 0.0, mu_acc = 0.0, sqx_s = 0.0, sqy_s = 0.0, meb_s = 0.0;
    (void)ke_acc; (void)me_acc; (void)mu_acc; (void)sqx_s; (void)sqy_s; (void)meb_s;

#pragma unroll 1
    for (int it = warp_has_rows ? -1 : R; it < R; ++it) {
        const int lj = lj0 + it, ln = lj + 1;               // current row, north face row
        const int j = jc0 + it, gj = gj0 + it;
        const bool active = own && (j <= p.row_end);
        // ---- north side: F_vu - Lyx at ffc (li, ln), F_vv - Lyy at ccc (li, lj), tracer flux at cfc (li, ln) ----
        double Fvu_n, Fvv_c, Ty_n, vq_n, rhff_n, sqx_n = 0.0;
        (void)sqx_n;
        {
            const double vln = RAW(s_v, li, ln);
            double vel[3];
            vel[0] = sym4(RAW(s_v, li - 2, ln), RAW(s_v, li - 1, ln), vln, RAW(s_v, li + 1, ln));
            const double vc = RAW(s_v, li, lj);
            vel[1] = sym4(RAW(s_v, li, lj - 1), vc, vln, RAW(s_v, li, lj + 2));
            if (by && ybuf(1, gj + 1, 2, NyG + 1)) vel[1] = sym2(vc, vln);
            vel[2] = vln;
            const double *const cu = &RAW(s_u, li, ln), *const cv = &RAW(s_v, li, ln), *const cA = &RAW(s_A, li, ln);
            double d1[3], d2[3], d3[3], d4[3], c0[3] = {es, es, es}, c1[3] = {es, es, es}, c2[3] = {es, es, es}, num[3], den[3], cc[3];
            UPDC(gt0(vel[0]), cu, W, 0, cc[0]); UPDC(gt0(vel[1]), cv, W, 1, cc[1]); UPDC(gt0(vel[2]), cA, W, 2, cc[2]);
            beta_acc_n<3>(d1, d2, d3, d4, c0, c1, c2);
            corr10_n<3>(d1, d2, d3, d4, c0, c1, c2, num, den);
            // the three WENO denominators and the face depths ℑy h at cfc (li, ln), ℑxy h at ffc (li, ln): five reciprocals together
            double x5[5], rc[5];
            {
                const double h00 = RAW(s_h, li, lj), h01 = RAW(s_h, li, ln);
                x5[0] = den[0]; x5[1] = den[1]; x5[2] = den[2];
                x5[3] = 0.5 * (h00 + h01);
                x5[4] = avg4(RAW(s_h, li - 1, lj), h00, RAW(s_h, li - 1, ln), h01);
            }
            rcp_mix_n<5, 3>(x5, rc);
            const double rhy_n = rc[3];
            rhff_n = rc[4];
            double fl[3];
#pragma unroll
            for (int n = 0; n < 3; n++) fl[n] = vel[n] * fma(num[n], rc[n], cc[n]);
            const double B0u = Bf(s_Bx, li, ln), Bmu = Bf(s_Bx, li, ln - 1);
            const double L3u = third(B0u, Bmu, Bf(s_Bx, li, ln - 2)), R3u = thirdR(Bf(s_Bx, li, ln + 1), B0u, Bmu);
            const double B0v = Bf(s_By, li, lj), Bpv = Bf(s_By, li, ln);
            const double L3v = third(Bpv, B0v, Bf(s_By, li, lj - 1)), R3v = thirdR(Bf(s_By, li, lj + 2), Bpv, B0v);
            double Lqu = L3u, Rqu = R3u, Lqv = L3v, Rqv = R3v;
            if (by) {
                // Bounded-y wall buffers: centred second order (oracle ybuf) and the edge branches of
                // advective_lorentz_flux_hBy_bx / hBy_by (sw_mhd_divergence_functions.jl:62-84, 110-132)
                if (ybuf(1, gj + 1, 3, NyG)) { fl[0] = vel[0] * sym2(cu[-W], cu[0]); fl[2] = vel[2] * sym2(cA[-W], cA[0]); }
                if (ybuf(1, gj + 1, 3, NyG + 1)) fl[1] = vel[1] * sym2(cv[-W], cv[0]);
                const int gjf = gj + 1;
                if (gjf == 1) { Lqu = B0u; Rqu = B0u; } else if (gjf == 2) { Lqu = Bmu; Rqu = R3u; }
                else if (gjf == NyG) { Lqu = L3u; Rqu = B0u; } else if (gjf == NyG + 1) { Lqu = Bmu; Rqu = Bmu; }
                if (gj == 0) { Lqv = Bpv; Rqv = Bpv; } else if (gj == 1) { Lqv = B0v; Rqv = R3v; }
                else if (gj == NyG - 1) { Lqv = L3v; Rqv = Bpv; } else if (gj == NyG) { Lqv = B0v; Rqv = B0v; }
            }
            {   // F_vu - Lyx (advective_lorentz_flux_hBy_bx, :62-84)
                const double mom = (p.dx * fl[0]) * rhff_n;
                const double vl = 0.5 * (Bf(s_hBy, li - 1, ln) + Bf(s_hBy, li, ln));
                Fvu_n = p.dx * upwind_sel(vl, Lqu, Rqu) - mom;
            }
            {   // F_vv - Lyy (advective_lorentz_flux_hBy_by, :110-132)
                const double mom = (p.dx * fl[1]) * RH(li, lj);
                const double vl = 0.5 * (Bf(s_hBy, li, lj) + Bf(s_hBy, li, ln));
                Fvv_c = p.dx * upwind_sel(vl, Lqv, Rqv) - mom;
            }
            {   // tracer transport flux and vh/ℑy h at cfc (li, ln)
                Ty_n = (p.dx * fl[2]) * rhy_n;
                vq_n = vln * rhy_n;
            }
            if constexpr (DIAG) {   // Bx^2 = (dyA / ℑy h)^2 at cfc (li, ln): the next row's sqx_s
                const double bxf = ((cA[0] - cA[-W]) * p.rdy) * rhy_n;
                sqx_n = bxf * bxf;
                if (it < 0) {       // By^2 at fcc (li, lj0-1) seeds the bracket of the first row
                    const double byf = ((cA[-W] - RAW(s_A, li - 1, lj)) * p.rdx) * frcp(0.5 * (RAW(s_h, li - 1, lj) + RAW(s_h, li, lj)));
                    sqy_s = byf * byf;
                }
            }
        }
        if (it >= 0) {
            // ---- x side: F_uu - Lxx at ccc li-1, F_uv - Lxy at ffc li, tracer flux at fcc li ---------------------
            const double uc = RAW(s_u, li, lj), ue = RAW(s_u, li + 1, lj), us = RAW(s_u, li, lj - 1);
            double Fuu_w, Fuv_w, Tx_w, uq_w, rhx_d;
            {
                double vel[3];
                vel[0] = sym4(RAW(s_u, li - 2, lj), RAW(s_u, li - 1, lj), uc, ue);
                vel[1] = sym4(RAW(s_u, li, lj - 2), us, uc, RAW(s_u, li, lj + 1));
                if (by && ybuf(1, gj, 2, NyG)) vel[1] = sym2(us, uc);
                vel[2] = uc;
                const double *const cu = &RAW(s_u, li, lj), *const cv = &RAW(s_v, li, lj), *const cA = &RAW(s_A, li, lj);
                double d1[3], d2[3], d3[3], d4[3], c0[3] = {es, es, es}, c1[3] = {es, es, es}, c2[3] = {es, es, es}, num[3], den[3], cc[3];
                UPDC(gt0(vel[0]), cu, 1, 0, cc[0]); UPDC(gt0(vel[1]), cv, 1, 1, cc[1]); UPDC(gt0(vel[2]), cA, 1, 2, cc[2]);
                beta_acc_n<3>(d1, d2, d3, d4, c0, c1, c2);
                corr10_n<3>(d1, d2, d3, d4, c0, c1, c2, num, den);
                double x4[4] = {den[0], den[1], den[2], 0.5 * (RAW(s_h, li - 1, lj) + RAW(s_h, li, lj))}, rc[4];   // + ℑx h at fcc (li, lj)
                rcp_mix_n<4, 3>(x4, rc);
                double fl[3];
#pragma unroll
                for (int n = 0; n < 3; n++) fl[n] = vel[n] * fma(num[n], rc[n], cc[n]);
                {   // F_uu - Lxx (advective_lorentz_flux_hBx_bx, :38-60) at the ccc point li-1
                    const double mom = (p.dy * fl[0]) * RH(li - 1, lj);
                    const double ul = 0.5 * (Bf(s_hBx, li - 1, lj) + Bf(s_hBx, li, lj));
                    const double B0 = Bf(s_Bx, li - 1, lj), Bp = Bf(s_Bx, li, lj);
                    const double Lq = third(Bp, B0, Bf(s_Bx, li - 2, lj)), Rq = thirdR(Bf(s_Bx, li + 1, lj), Bp, B0);
                    Fuu_w = p.dy * upwind_sel(ul, Lq, Rq) - mom;
                }
                {   // F_uv - Lxy (advective_lorentz_flux_hBx_by, :86-108) at the ffc point (li, lj)
                    const double mom = (p.dy * fl[1]) * rhff_s;
                    const double ul = 0.5 * (Bf(s_hBx, li, lj - 1) + Bf(s_hBx, li, lj));
                    const double B0 = Bf(s_By, li, lj), Bm = Bf(s_By, li - 1, lj);
                    const double Lq = third(B0, Bm, Bf(s_By, li - 2, lj)), Rq = thirdR(Bf(s_By, li + 1, lj), B0, Bm);
                    Fuv_w = p.dy * upwind_sel(ul, Lq, Rq) - mom;
                }
                {   // tracer transport flux and uh/ℑx h at fcc (li, lj)
                    const double rhx = rc[3];
                    Tx_w = (p.dy * fl[2]) * rhx;
                    uq_w = uc * rhx;
                    rhx_d = rhx;
                }
            }
            const double Fuu_e = __shfl_down_sync(0xffffffffu, Fuu_w, 1), Fuv_e = __shfl_down_sync(0xffffffffu, Fuv_w, 1);
            const double Tx_e = __shfl_down_sync(0xffffffffu, Tx_w, 1), uq_e = __shfl_down_sync(0xffffffffu, uq_w, 1);
            if constexpr (DIAG) {
                // ---- energy sums and max |uh / ℑx h| (divergence_sw_mhd.jl:47,53,63-75), row lj ----------
                // KE bracket uh^2 + ℑxyᶠᶜᵃ(vh^2) at fcc (li, lj), ME bracket Bx^2 + ℑxyᶜᶠᵃ(By^2) at cfc (li, lj); uq_w / uc is
                // 1/ℑx h at fcc (li, lj); lane 31 is the column beyond the tile and supplies the east neighbours
                const double Ac_ = RAW(s_A, li, lj), hs_ = RAW(s_h, li, lj - 1);
                const double byf = ((Ac_ - RAW(s_A, li - 1, lj)) * p.rdx) * rhx_d;
                const double sqy = byf * byf, csum = sqy_s + sqy;
                const double v00 = RAW(s_v, li - 1, lj), v10 = RAW(s_v, li, lj), v01 = RAW(s_v, li - 1, ln), v11 = RAW(s_v, li, ln);
                const double keb = fma(uc, uc, avg4(v00 * v00, v10 * v10, v01 * v01, v11 * v11));
                const double csum_e = __shfl_down_sync(0xffffffffu, csum, 1), keb_e = __shfl_down_sync(0xffffffffu, keb, 1);
                const double meb = fma(0.25, csum + csum_e, sqx_s);
                if (it >= 1 && own && (j - 1 <= p.row_end)) me_acc = fma(0.25 * hs_, meb_s + meb, me_acc);   // cell (li, lj-1)
                if (active) {
                    ke_acc = fma(0.25 * RH(li, lj), keb + keb_e, ke_acc);                                     // (1/h)/2 (keb + keb_e)/2
                    mu_acc = fmax(mu_acc, fabs(uq_w));
                }
                meb_s = meb; sqy_s = sqy;
            }

            // ---- tendencies ----------------------------------------------------------------------
            // d(g h^2 / 2): h = 1 + O(1e-9) makes this a cancellation; keep the products un-contracted
            const double hg = 0.5 * p.g;
            const double hc = RAW(s_h, li, lj), hw_ = RAW(s_h, li - 1, lj), hs = RAW(s_h, li, lj - 1);
            const double Pc = __dmul_rn(hg, __dmul_rn(hc, hc)), Pw = __dmul_rn(hg, __dmul_rn(hw_, hw_)), Ps = __dmul_rn(hg, __dmul_rn(hs, hs));
            const double vc = RAW(s_v, li, lj), vn = RAW(s_v, li, ln);
            const double Ac = RAW(s_A, li, lj);
            double Gn0, Gn1 = 0.0, Gn2, Gn3;
            {   // Guh (dm holds div(Lorentz - momentum flux))
                const double dm = p.inv_az * ((Fuu_e - Fuu_w) + (Fvu_n - Fvu_s));
                const double pg = __dsub_rn(Pc, Pw) * p.rdx;
                const double vhat = avg4(RAW(s_v, li - 1, lj), vc, RAW(s_v, li - 1, ln), vn);
                Gn0 = fma(p.f, vhat, dm - pg);
            }
            if (!(by && gj < 2)) {   // Gvh (wall rows of a Bounded-y grid keep vh = 0)
                const double dm = p.inv_az * ((Fuv_e - Fuv_w) + (Fvv_c - Fvv_s));
                const double pg = __dsub_rn(Pc, Ps) * p.rdy;
                const double uhat = avg4(us, RAW(s_u, li + 1, lj - 1), uc, ue);
                Gn1 = fma(-p.f, uhat, dm - pg);
            }
            {   // Gh (centred), GA
                Gn2 = -fma(ue - uc, p.rdx, (vn - vc) * p.rdy);
                const double d = p.inv_az * ((Tx_e - Tx_w) + (Ty_n - Ty_s));
                const double cdiv = fma(uq_e - uq_w, p.rdx, (vq_n - vq_s) * p.rdy);
                Gn3 = fma(Ac, cdiv, -d);
            }
            if (active) {
                const double Gn[4] = {Gn0, Gn1, Gn2, Gn3};
                const double Gm[4] = {Gq0, Gq1, Gq2, Gq3};
                const double Uc[4] = {uc, vc, hc, Ac};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if constexpr (STAGE == 1) {
                        p.Un[k][gcell] = fma(p.dtgam, Gn[k], Uc[k]);
                        p.G[k][gcell] = Gn[k];
                    } else {
                        p.Un[k][gcell] = fma(p.dtgam, Gn[k], fma(p.dtzet, Gm[k], Uc[k]));
                        if constexpr (STAGE == 2) p.G[k][gcell] = Gn[k];
                    }
                }
            }
            if constexpr (STAGE >= 2) {     // G^- of the next row: requested now, consumed at the end of the next iteration
                if (it + 1 < R && own && (j + 1 <= p.row_end)) {
                    const size_t gn = gcell + (size_t)P;
                    Gq0 = p.G[0][gn]; Gq1 = p.G[1][gn]; Gq2 = p.G[2][gn]; Gq3 = p.G[3][gn];
                }
            }
        }
        Fvu_s = Fvu_n; Fvv_s = Fvv_c; Ty_s = Ty_n; vq_s = vq_n; rhff_s = rhff_n;
        if constexpr (DIAG) sqx_s = sqx_n;
        gcell += (size_t)P;
    }

    // ---- fused diagnostics of the state at the start of the step ------------------------------
    // KE, ME and max |uh / ℑx h| were accumulated in the row walk.  Left: the magnetic-energy bracket north of the warp's
    // last row (one more reciprocal depth) and the quantities that need no reciprocal.
    if constexpr (DIAG) {
        double sums[4] = {0.0, 0.0, 0.0, 0.0};                // ke, me, pe, sum h
        const unsigned long long kneut = ord_key(-INFINITY);
        unsigned long long kmax[4] = {ord_key(0.0), ord_key(0.0), kneut, ord_key(0.0)};   // |u|, |A|, -h, |div hB| as ordered keys
        bool bad = false;
        if (warp_has_rows) {
            {   // ME of the cell (li, lj0+R-1): bracket at cfc (li, lj0+R); its Bx^2 is sqx_s
                const int lt = lj0 + R;
                const double byf = ((RAW(s_A, li, lt) - RAW(s_A, li - 1, lt)) * p.rdx) * frcp(0.5 * (RAW(s_h, li - 1, lt) + RAW(s_h, li, lt)));
                const double csum = fma(byf, byf, sqy_s);
                const double csum_e = __shfl_down_sync(0xffffffffu, csum, 1);
                const double meb = fma(0.25, csum + csum_e, sqx_s);
                if (own && (jc0 + R - 1 <= p.row_end)) me_acc = fma(0.25 * RAW(s_h, li, lt - 1), meb_s + meb, me_acc);
            }
            sums[0] = ke_acc; sums[1] = me_acc;
            kmax[0] = ord_key(mu_acc);
            double Am[3], A0[3], Ap[3];                      // A at columns li-1, li, li+1 of rows lj-1, lj, lj+1 (sliding)
#pragma unroll
            for (int c = 0; c < 3; c++) { Am[c] = RAW(s_A, li - 1 + c, lj0 - 1); A0[c] = RAW(s_A, li - 1 + c, lj0); }
#pragma unroll
            for (int r = 0; r < R; r++) {            // branch-free: cells beyond the launch contribute neutral elements
                const int lj = lj0 + r;
#pragma unroll
                for (int c = 0; c < 3; c++) Ap[c] = RAW(s_A, li - 1 + c, lj + 1);
                const bool ok = own && (jc0 + r <= p.row_end);
                const double hh = RAW(s_h, li, lj), uu = RAW(s_u, li, lj), vv = RAW(s_v, li, lj), aa = A0[1];
                const double dh = hh - p.h_ref;
                sums[2] = fma(ok ? (0.5 * p.g) * dh : 0.0, dh, sums[2]);
                sums[3] += ok ? hh : 0.0;
                kmax[1] = max(kmax[1], ok ? ord_key(fabs(aa)) : kneut);
                kmax[2] = max(kmax[2], ok ? ord_key(-hh) : kneut);
                {   // div(hB) at ccc, telescoped ℑxy∂ (pure round-off)
                    const double hbx0 = (Am[0] + Am[1]) - (Ap[0] + Ap[1]), hbx1 = (Am[1] + Am[2]) - (Ap[1] + Ap[2]);
                    const double hby0 = (Am[2] + A0[2]) - (Am[0] + A0[0]), hby1 = (A0[2] + Ap[2]) - (A0[0] + Ap[0]);
                    const double dv = fabs((hbx1 - hbx0) * (0.25 * p.rdy) * p.rdx + (hby1 - hby0) * (0.25 * p.rdx) * p.rdy);
                    kmax[3] = max(kmax[3], ok ? ord_key(dv) : kneut);
                }
                // finite <=> exponent field below 0x7ff: one integer test on the OR of the four high words' exponent overflow
                const unsigned eh = (unsigned)__double2hiint(hh) & 0x7ff00000u, ea = (unsigned)__double2hiint(aa) & 0x7ff00000u;
                const unsigned eu = (unsigned)__double2hiint(uu) & 0x7ff00000u, ev = (unsigned)__double2hiint(vv) & 0x7ff00000u;
                bad = bad || (ok && (eh == 0x7ff00000u || ea == 0x7ff00000u || eu == 0x7ff00000u || ev == 0x7ff00000u));
#pragma unroll
                for (int c = 0; c < 3; c++) { Am[c] = A0[c]; A0[c] = Ap[c]; }
            }
        }
        // fixed-order reduction: R cells per thread (above), the warp (four sums by reduce-scatter, four extrema by REDUX
        // on ordered keys, the finite flag by vote), then the warps in order
        double *const s_red = smem + DERIVED_D + 2;            // warp w's partials at the start of its scratch slice
        __syncwarp();
        {
            const double tot = warp_sum4(sums, lane);
            unsigned long long km[4];
#pragma unroll
            for (int q = 0; q < 4; q++) km[q] = warp_max_key(kmax[q]);
            const bool anybad = __any_sync(0xffffffffu, bad);
            if ((lane & 7) == 0) s_red[wp * NDIAG + (lane >> 3)] = tot;
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) s_red[wp * NDIAG + 4 + q] = ord_val(km[q]);
                s_red[wp * NDIAG + 8] = anybad ? 1.0 : 0.0;
            }
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
            for (int s2 = 1; s2 < TYB / 8; s2++)
                if (row0 + 8 * s2 < p.row_end) p.diag[((size_t)(tr8 + s2) * tiles_x + tile_x) * NDIAG + tid] = (tid == 6) ? -INFINITY : 0.0;
        }
    }
}

// Per-device launch configuration: the dynamic-shared-memory opt-in is a per-device function attribute and
// the L2 prefetch distance depends on the device's SM count, so neither may be cached per process (a
// context may live on any device, several devices may be driven by one process).
constexpr int MAX_DEVICES = 64;
template <int FORM, int STAGE, bool DIAG>
cudaError_t launch_rb(const KParams &p, cudaStream_t st) {
    auto kern = (FORM == 0) ? substage_rb_kernel<STAGE, DIAG> : substage_rbd_kernel<STAGE, DIAG>;
    constexpr size_t bytes = (FORM == 0) ? SMEM_BYTES : SMEM_BYTES_D;
    constexpr int cells_x = (FORM == 0) ? TX : TXD;
    static int ahead[MAX_DEVICES] = {};                     // 0 = this device is not configured yet
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
    if (ahead[dev] == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        int sms = 0, occ = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, bytes);
        if (e != cudaSuccess) return e;
        const char *env = getenv("SWMHD_L2_AHEAD");
        const int a = env ? atoi(env) : (occ < 1 ? 1 : occ) * sms;   // CTAs in flight = distance to the slot's next tile
        ahead[dev] = a < 1 ? 1 : a;
    }
    const int tiles_x = (p.Nx + cells_x - 1) / cells_x;
    const int tiles_y = (p.row_end - p.row_begin + TYB - 1) / TYB;
    KParams q = p;
    q.l2_ahead = ahead[dev];
    kern<<<tiles_x * tiles_y, NT, bytes, st>>>(q);
    return cudaGetLastError();
}

} // namespace

void substage_rb_tile(int *tx, int *ty) { *tx = TX; *ty = TYB; }
int substage_rb_tiles_x(int form, int Nx) { const int c = (form == 0) ? TX : TXD; return (Nx + c - 1) / c; }

cudaError_t launch_substage_rb(const KParams &p0, int form, int stage, cudaStream_t st) {
    if (p0.tile_rows <= 0) return cudaSuccess;
    if (!p0.use_rb) return cudaErrorInvalidValue;
    int tx8, ty8;
    substage_tile(&tx8, &ty8);                              // launch granularity of the host side (tile rows)
    if (ty8 != 8) return cudaErrorInvalidValue;
    KParams p = p0;
    p.row_begin = p0.tile_row0 * ty8;
    p.row_end = (p0.tile_row0 + p0.tile_rows) * ty8;
    if (p.row_end > p.Ny) p.row_end = p.Ny;
    if (p.row_end <= p.row_begin) return cudaSuccess;
    const bool dg = (p.diag != nullptr);
    if (dg && stage != 1) return cudaErrorInvalidValue;
    switch (form * 4 + stage) {
        case 1: return dg ? launch_rb<0, 1, true>(p, st) : launch_rb<0, 1, false>(p, st);
        case 2: return launch_rb<0, 2, false>(p, st);
        case 3: return launch_rb<0, 3, false>(p, st);
        case 5: return dg ? launch_rb<1, 1, true>(p, st) : launch_rb<1, 1, false>(p, st);
        case 6: return launch_rb<1, 2, false>(p, st);
        case 7: return launch_rb<1, 3, false>(p, st);
    }
    return cudaErrorInvalidValue;
}

} // namespace swmhd
