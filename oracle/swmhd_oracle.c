/*
 * swmhd_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C, FP64, no-FMA-contraction restatement of one RK3 time step of the
 * reference's two model set-ups:
 *   JACOBIAN   : jacobian_formulation/SWMHD_example.jl:21-33
 *                + jacobian_formulation/sw_mhd_jacobian_functions.jl:1-26
 *   DIVERGENCE : divergence_formulation/divergence_sw_mhd.jl:19-31
 *                + divergence_formulation/sw_mhd_divergence_functions.jl:1-170
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.  Nothing under
 * swmhd_b200/ imports it; the product path has no CPU fallback.
 *
 * PARITY STATUS: *** parity unpinned against upstream Oceananigans ***
 * The two *_functions.jl files above are reference-owned and are restated
 * line by line (each function below cites its lines).  Everything else on the
 * step (WENO5, vector-invariant / conservative shallow-water tendencies,
 * RK3, halo filling, AbstractOperations diagnostics) lives in Oceananigans.jl,
 * an un-vendored, un-pinned dependency (no Project.toml/Manifest.toml in the
 * reference; inferred v0.76.x).  Julia and Oceananigans are not
 * installable here, so that part follows the recalled specification in
 * SURVEY.md Appendix A, and is pinned only by
 *   (i)   the closed-form answers of the reference's own operator scripts
 *         (test_formulations.jl:12-18, MHD_visualize.jl:8-24),
 *   (ii)  all twelve published energy plots of the reference (energy_plots/:
 *         3 initial conditions x 64^2/128^2 x both formulations, Bounded-y
 *         included), machine-digitised every half time unit into
 *         tests/golden/published_traces.json: 1e-5 absolute in the low-B
 *         runs, 0.1-0.4 % in the others (tests/test_published_traces.py), and
 *   (iii) the survey's independent scratch-restatement checksums (B.5).
 * julia/dump_reference.jl writes real-upstream golden files wherever Julia +
 * Oceananigans exist.
 *
 * Arithmetic conventions (mirrored exactly by the STRICT CUDA kernels):
 *   - compiled with -ffp-contract=off: every * and + rounds separately;
 *   - expressions are evaluated left to right exactly as written here;
 *   - "inv_az * (...)" multiplies by the rounded reciprocal 1/(dx*dy);
 *     "/ dx" is a true division.
 *
 * Indexing: logical 1-based (i,j) with halo 3 -> parent offset
 *   (i+2) + (Nx+6)*(j+2)   (parent of Oceananigans' OffsetArray).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../include/swmhd.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define H3 3

typedef struct {
    int Nx, Ny, P;          /* P = Nx + 6                                     */
    int by;                 /* topo_y == Bounded                               */
    int bx;                 /* topo_x == Bounded (Lorentz edge branches only)  */
    int form, flags;
    double dx, dy, g, f, eps, inv_az;
    const double *u, *v, *h, *A;   /* u|uh, v|vh, h, A (haloed parents)        */
} S;

static inline size_t IX(const S *s, int i, int j) {
    return (size_t)(i + 2) + (size_t)s->P * (size_t)(j + 2);
}
static inline double fU(const S *s, int i, int j) { return s->u[IX(s, i, j)]; }
static inline double fV(const S *s, int i, int j) { return s->v[IX(s, i, j)]; }
static inline double fH(const S *s, int i, int j) { return s->h[IX(s, i, j)]; }
static inline double fA(const S *s, int i, int j) { return s->A[IX(s, i, j)]; }

static void S_init(S *s, const swmhd_config *c, const double *u, const double *v,
                   const double *h, const double *A) {
    s->Nx = c->Nx; s->Ny = c->Ny; s->P = c->Nx + 2 * H3;
    s->by = (c->topo_y == SWMHD_BOUNDED);
    s->bx = (c->topo_x == SWMHD_BOUNDED);
    s->form = c->formulation; s->flags = c->flags;
    s->dx = c->dx; s->dy = c->dy; s->g = c->g; s->f = c->f; s->eps = c->weno_eps;
    s->inv_az = 1.0 / (c->dx * c->dy);
    s->u = u; s->v = v; s->h = h; s->A = A;
}

static int rows_of(const swmhd_config *c, int field) {
    return c->Ny + 2 * H3 + ((field == SWMHD_V && c->topo_y == SWMHD_BOUNDED) ? 1 : 0);
}
/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs of bench.py ask for all cores explicitly */
int swmhd_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

size_t swmhd_oracle_field_len(const swmhd_config *c, int field) {
    return (size_t)(c->Nx + 2 * H3) * (size_t)rows_of(c, field);
}

/* ------------------------------------------------------------------------- */
/* WENO5.  q[0..4] = psi[f-3], psi[f-2], psi[f-1], psi[f], psi[f+1]
 * is the LEFT-biased sample set of face f; the right-biased value is the same
 * function applied to the mirror psi[f+2], psi[f+1], psi[f], psi[f-1], psi[f-2]. */
static inline void weno_beta(const double q[5], double b[3]) {
    const double a = q[0], bb = q[1], c = q[2], d = q[3], e = q[4];
    double D0 = c - 2.0 * d + e,  E0 = 3.0 * c - 4.0 * d + e;   /* (f-1,f,f+1)   C=3/10 */
    double D1 = bb - 2.0 * c + d, E1 = bb - d;                  /* (f-2,f-1,f)   C=3/5  */
    double D2 = a - 2.0 * bb + c, E2 = a - 4.0 * bb + 3.0 * c;  /* (f-3,f-2,f-1) C=1/10 */
    b[0] = (13.0 / 12.0) * (D0 * D0) + 0.25 * (E0 * E0);
    b[1] = (13.0 / 12.0) * (D1 * D1) + 0.25 * (E1 * E1);
    b[2] = (13.0 / 12.0) * (D2 * D2) + 0.25 * (E2 * E2);
}

static inline double weno_blend(const S *s, const double q[5], const double b[3]) {
    const double a = q[0], bb = q[1], c = q[2], d = q[3], e = q[4];
    double p0 = (2.0 * c + 5.0 * d - e) / 6.0;
    double p1 = (-bb + 5.0 * c + 2.0 * d) / 6.0;
    double p2 = (2.0 * a - 7.0 * bb + 11.0 * c) / 6.0;
    double a0, a1, a2;
    if (s->flags & SWMHD_FLAG_WENO_JS) {
        a0 = 0.3 / ((b[0] + s->eps) * (b[0] + s->eps));
        a1 = 0.6 / ((b[1] + s->eps) * (b[1] + s->eps));
        a2 = 0.1 / ((b[2] + s->eps) * (b[2] + s->eps));
    } else { /* Z weights, tau5 = |beta2 - beta0| */
        double tau = fabs(b[2] - b[0]);
        double r0 = tau / (b[0] + s->eps), r1 = tau / (b[1] + s->eps), r2 = tau / (b[2] + s->eps);
        a0 = 0.3 * (1.0 + r0 * r0);
        a1 = 0.6 * (1.0 + r1 * r1);
        a2 = 0.1 * (1.0 + r2 * r2);
    }
    double sum = a0 + a1 + a2;
    double w0 = a0 / sum, w1 = a1 / sum, w2 = a2 / sum;
    return w0 * p0 + w1 * p1 + w2 * p2;
}

static inline double weno(const S *s, const double q[5]) {
    double b[3];
    weno_beta(q, b);
    return weno_blend(s, q, b);
}

/* centred 4th-order symmetric interpolant that accompanies WENO5 (A.3):
 * a..d = psi[f-2], psi[f-1], psi[f], psi[f+1]                                 */
static inline double sym4(double a, double b, double c, double d) {
    return (7.0 * (b + c) - (a + d)) / 12.0;
}
/* centred 2nd order, used inside the Bounded-y wall buffer (A.8) */
static inline double sym2(double b, double c) { return 0.5 * (b + c); }

/* upwind_biased_product — sw_mhd_divergence_functions.jl:3 (verbatim copy of upstream's) */
static inline double upwind(double ut, double L, double R) {
    return ((ut + fabs(ut)) * L + (ut - fabs(ut)) * R) / 2.0;
}

typedef double (*fn2)(const S *, int, int);

static inline void line_x(const S *s, fn2 F, int f, int j, double q[5], double r[5]) {
    q[0] = F(s, f - 3, j); q[1] = F(s, f - 2, j); q[2] = F(s, f - 1, j); q[3] = F(s, f, j); q[4] = F(s, f + 1, j);
    r[0] = F(s, f + 2, j); r[1] = q[4]; r[2] = q[3]; r[3] = q[2]; r[4] = q[1];
}
static inline void line_y(const S *s, fn2 F, int i, int f, double q[5], double r[5]) {
    q[0] = F(s, i, f - 3); q[1] = F(s, i, f - 2); q[2] = F(s, i, f - 1); q[3] = F(s, i, f); q[4] = F(s, i, f + 1);
    r[0] = F(s, i, f + 2); r[1] = q[4]; r[2] = q[3]; r[3] = q[2]; r[4] = q[1];
}

/* Bounded-y wall buffer (confidence L): a y-reconstruction whose
 * footprint f-n..f+n-1 leaves the valid element range [lo,hi] of its line
 * falls back to centred 2nd order.  Centre-located lines (h, A, u|uh) are
 * valid on 1..Ny, face-located lines (v|vh, zeta) on 1..Ny+1.                 */
static inline int ybuf(const S *s, int f, int n, int lo, int hi) {
    return s->by && (f - n < lo || f + n - 1 > hi);
}
#define YC_HI(s) ((s)->Ny)
#define YF_HI(s) ((s)->Ny + 1)

/* WENO3 (C10 probe): left-biased value at the face between q[1] and q[2] from q[0..2] = psi[f-2], psi[f-1], psi[f];
 * candidates (psi[f-1]+psi[f])/2 (C = 2/3) and (-psi[f-2]+3 psi[f-1])/2 (C = 1/3), Z or JS weights like WENO5. */
static inline double weno3(const S *s, double a, double b, double c, double b0, double b1) {
    double p0 = 0.5 * (b + c), p1 = 0.5 * (3.0 * b - a);
    double a0, a1;
    if (s->flags & SWMHD_FLAG_WENO_JS) {
        a0 = (2.0 / 3.0) / ((b0 + s->eps) * (b0 + s->eps));
        a1 = (1.0 / 3.0) / ((b1 + s->eps) * (b1 + s->eps));
    } else {
        double tau = fabs(b0 - b1), r0 = tau / (b0 + s->eps), r1 = tau / (b1 + s->eps);
        a0 = (2.0 / 3.0) * (1.0 + r0 * r0);
        a1 = (1.0 / 3.0) * (1.0 + r1 * r1);
    }
    return (a0 * p0 + a1 * p1) / (a0 + a1);
}
static inline void weno3_beta(double a, double b, double c, double *b0, double *b1) {
    *b0 = (c - b) * (c - b); *b1 = (b - a) * (b - a);
}

/* left/right WENO5 values of F at x-face f of row j */
static inline void weno_x(const S *s, fn2 F, int f, int j, double *L, double *R) {
    double q[5], r[5];
    line_x(s, F, f, j, q, r);
    *L = weno(s, q); *R = weno(s, r);
}
/* left/right WENO5 values of F at y-face f of column i; valid elements lo..hi */
static inline void weno_y(const S *s, fn2 F, int i, int f, int hi, double *L, double *R) {
    if (ybuf(s, f, 3, 1, hi)) {
        if ((s->flags & SWMHD_FLAG_WALL_WENO3) && !ybuf(s, f, 2, 1, hi)) {
            double a = F(s, i, f - 2), b = F(s, i, f - 1), c = F(s, i, f), d = F(s, i, f + 1), b0, b1;
            weno3_beta(a, b, c, &b0, &b1); *L = weno3(s, a, b, c, b0, b1);
            weno3_beta(d, c, b, &b0, &b1); *R = weno3(s, d, c, b, b0, b1);
            return;
        }
        *L = *R = sym2(F(s, i, f - 1), F(s, i, f)); return;
    }
    double q[5], r[5];
    line_y(s, F, i, f, q, r);
    *L = weno(s, q); *R = weno(s, r);
}
static inline double sym_x(const S *s, fn2 F, int f, int j) {
    return sym4(F(s, f - 2, j), F(s, f - 1, j), F(s, f, j), F(s, f + 1, j));
}
static inline double sym_y(const S *s, fn2 F, int i, int f, int hi) {
    if (ybuf(s, f, 2, 1, hi)) return sym2(F(s, i, f - 1), F(s, i, f));
    return sym4(F(s, i, f - 2), F(s, i, f - 1), F(s, i, f), F(s, i, f + 1));
}

/* ------------------------------------------------------------------------- */
/* A.2 operators                                                               */
static inline double dxA(const S *s, int i, int j) { return (fA(s, i, j) - fA(s, i - 1, j)) / s->dx; } /* ∂xᶠᶜᶜ A */
static inline double dyA(const S *s, int i, int j) { return (fA(s, i, j) - fA(s, i, j - 1)) / s->dy; } /* ∂yᶜᶠᶜ A */
static inline double ixf_h(const S *s, int i, int j) { return 0.5 * (fH(s, i - 1, j) + fH(s, i, j)); } /* ℑxᶠᵃᵃ h */
static inline double iyf_h(const S *s, int i, int j) { return 0.5 * (fH(s, i, j - 1) + fH(s, i, j)); } /* ℑyᵃᶠᵃ h */
/* ℑxyᶠᶜᵃ F = ℑyᵃᶜᵃ(ℑxᶠᵃᵃ F);  ℑxyᶜᶠᵃ F = ℑyᵃᶠᵃ(ℑxᶜᵃᵃ F);  ℑxyᶠᶠᵃ F = ℑyᵃᶠᵃ(ℑxᶠᵃᵃ F) */
static inline double ixy_fc(const S *s, fn2 F, int i, int j) {
    return 0.5 * (0.5 * (F(s, i - 1, j) + F(s, i, j)) + 0.5 * (F(s, i - 1, j + 1) + F(s, i, j + 1)));
}
static inline double ixy_cf(const S *s, fn2 F, int i, int j) {
    return 0.5 * (0.5 * (F(s, i, j - 1) + F(s, i + 1, j - 1)) + 0.5 * (F(s, i, j) + F(s, i + 1, j)));
}
static inline double ixy_ff(const S *s, fn2 F, int i, int j) {
    return 0.5 * (0.5 * (F(s, i - 1, j - 1) + F(s, i, j - 1)) + 0.5 * (F(s, i - 1, j) + F(s, i, j)));
}

/* ------------------------------------------------------------------------- */
/* Jacobian-form Lorentz force — jacobian_formulation/sw_mhd_jacobian_functions.jl */
static inline double jBy(const S *s, int i, int j) { /* :1-3   By at ccc */
    return (0.5 * (dxA(s, i, j) + dxA(s, i + 1, j))) / fH(s, i, j);
}
static inline double jBx(const S *s, int i, int j) { /* :5-7   Bx at ccc */
    return -(0.5 * (dyA(s, i, j) + dyA(s, i, j + 1))) / fH(s, i, j);
}
static inline double dy_jBx(const S *s, int i, int j) { return (jBx(s, i, j) - jBx(s, i, j - 1)) / s->dy; }
static inline double dx_jBy(const S *s, int i, int j) { return (jBy(s, i, j) - jBy(s, i - 1, j)) / s->dx; }

static inline double jacobian_x(const S *s, int i, int j) { /* :10-13, fcc */
    double t1 = dxA(s, i, j) * ixy_fc(s, dy_jBx, i, j);
    double t2 = ixy_fc(s, dyA, i, j) * ((jBx(s, i, j) - jBx(s, i - 1, j)) / s->dx);
    return t1 - t2;
}
static inline double jacobian_y(const S *s, int i, int j) { /* :15-18, cfc */
    double t1 = ixy_cf(s, dxA, i, j) * ((jBy(s, i, j) - jBy(s, i, j - 1)) / s->dy);
    double t2 = dyA(s, i, j) * ixy_cf(s, dx_jBy, i, j);
    return t1 - t2;
}
static inline double lorentz_force_func_x(const S *s, int i, int j) { /* :20-22 */
    return (1.0 / ixf_h(s, i, j)) * jacobian_x(s, i, j);
}
static inline double lorentz_force_func_y(const S *s, int i, int j) { /* :24-26 */
    return (1.0 / iyf_h(s, i, j)) * jacobian_y(s, i, j);
}

/* ------------------------------------------------------------------------- */
/* Divergence-form Lorentz force — divergence_formulation/sw_mhd_divergence_functions.jl */
static inline double hBx(const S *s, int i, int j) { return -ixy_fc(s, dyA, i, j); }          /* :142-144 fcc */
static inline double hBy(const S *s, int i, int j) { return ixy_cf(s, dxA, i, j); }           /* :146-148 cfc */
static inline double dBx(const S *s, int i, int j) { return hBx(s, i, j) / ixf_h(s, i, j); }  /* :134-136 fcc */
static inline double dBy(const S *s, int i, int j) { return hBy(s, i, j) / iyf_h(s, i, j); }  /* :138-140 cfc */

/* third-order biased interpolants :25-35 and first-order ones :12-22, written
 * for the ᶠ variant at face index (i or j); the ᶜ variants are these at +1.  */
static inline double L3x(const S *s, fn2 F, int i, int j) { return (2.0 * F(s, i, j) + 5.0 * F(s, i - 1, j) - F(s, i - 2, j)) / 6.0; }
static inline double R3x(const S *s, fn2 F, int i, int j) { return (-F(s, i + 1, j) + 5.0 * F(s, i, j) + 2.0 * F(s, i - 1, j)) / 6.0; }
static inline double L3y(const S *s, fn2 F, int i, int j) { return (2.0 * F(s, i, j) + 5.0 * F(s, i, j - 1) - F(s, i, j - 2)) / 6.0; }
static inline double R3y(const S *s, fn2 F, int i, int j) { return (-F(s, i, j + 1) + 5.0 * F(s, i, j) + 2.0 * F(s, i, j - 1)) / 6.0; }
static inline double L1x(const S *s, fn2 F, int i, int j) { return F(s, i - 1, j); }
static inline double R1x(const S *s, fn2 F, int i, int j) { return F(s, i, j); }
static inline double L1y(const S *s, fn2 F, int i, int j) { return F(s, i, j - 1); }
static inline double R1y(const S *s, fn2 F, int i, int j) { return F(s, i, j); }

static inline double flux_hBx_bx(const S *s, int i, int j) { /* :38-60, ccc */
    double ut = 0.5 * (hBx(s, i, j) + hBx(s, i + 1, j));
    double L, R;
    if (s->bx && i == 0)               { L = R1x(s, dBx, i + 1, j); R = R1x(s, dBx, i + 1, j); }
    else if (s->bx && i == 1)          { L = L1x(s, dBx, i + 1, j); R = R3x(s, dBx, i + 1, j); }
    else if (s->bx && i == s->Nx - 1)  { L = L3x(s, dBx, i + 1, j); R = R1x(s, dBx, i + 1, j); }
    else if (s->bx && i == s->Nx)      { L = L1x(s, dBx, i + 1, j); R = L1x(s, dBx, i + 1, j); }
    else                               { L = L3x(s, dBx, i + 1, j); R = R3x(s, dBx, i + 1, j); }
    return s->dy * upwind(ut, L, R); /* Axᶜᶜᶜ = Δy */
}
static inline double flux_hBy_bx(const S *s, int i, int j) { /* :62-84, ffc */
    double vt = 0.5 * (hBy(s, i - 1, j) + hBy(s, i, j));
    double L, R;
    if (s->by && j == 1)               { L = R1y(s, dBx, i, j); R = R1y(s, dBx, i, j); }
    else if (s->by && j == 2)          { L = L1y(s, dBx, i, j); R = R3y(s, dBx, i, j); }
    else if (s->by && j == s->Ny)      { L = L3y(s, dBx, i, j); R = R1y(s, dBx, i, j); }
    else if (s->by && j == s->Ny + 1)  { L = L1y(s, dBx, i, j); R = L1y(s, dBx, i, j); }
    else                               { L = L3y(s, dBx, i, j); R = R3y(s, dBx, i, j); }
    return s->dx * upwind(vt, L, R); /* Ayᶠᶠᶜ = Δx */
}
static inline double flux_hBx_by(const S *s, int i, int j) { /* :86-108, ffc */
    double ut = 0.5 * (hBx(s, i, j - 1) + hBx(s, i, j));
    double L, R;
    if (s->bx && i == 1)               { L = R1x(s, dBy, i, j); R = R1x(s, dBy, i, j); }
    else if (s->bx && i == 2)          { L = L1x(s, dBy, i, j); R = R3x(s, dBy, i, j); }
    else if (s->bx && i == s->Nx)      { L = L3x(s, dBy, i, j); R = R1x(s, dBy, i, j); }
    else if (s->bx && i == s->Nx + 1)  { L = L1x(s, dBy, i, j); R = L1x(s, dBy, i, j); }
    else                               { L = L3x(s, dBy, i, j); R = R3x(s, dBy, i, j); }
    return s->dy * upwind(ut, L, R); /* Axᶠᶠᶜ = Δy */
}
static inline double flux_hBy_by(const S *s, int i, int j) { /* :110-132, ccc */
    double vt = 0.5 * (hBy(s, i, j) + hBy(s, i, j + 1));
    double L, R;
    if (s->by && j == 0)               { L = R1y(s, dBy, i, j + 1); R = R1y(s, dBy, i, j + 1); }
    else if (s->by && j == 1)          { L = L1y(s, dBy, i, j + 1); R = R3y(s, dBy, i, j + 1); }
    else if (s->by && j == s->Ny - 1)  { L = L3y(s, dBy, i, j + 1); R = R1y(s, dBy, i, j + 1); }
    else if (s->by && j == s->Ny)      { L = L1y(s, dBy, i, j + 1); R = L1y(s, dBy, i, j + 1); }
    else                               { L = L3y(s, dBy, i, j + 1); R = R3y(s, dBy, i, j + 1); }
    return s->dx * upwind(vt, L, R); /* Ayᶜᶜᶜ = Δx */
}
static inline double div_lorentz_x(const S *s, int i, int j) { /* :162-165, fcc */
    return s->inv_az * ((flux_hBx_bx(s, i, j) - flux_hBx_bx(s, i - 1, j)) +
                        (flux_hBy_bx(s, i, j + 1) - flux_hBy_bx(s, i, j)));
}
static inline double div_lorentz_y(const S *s, int i, int j) { /* :167-170, cfc */
    return s->inv_az * ((flux_hBx_by(s, i + 1, j) - flux_hBx_by(s, i, j)) +
                        (flux_hBy_by(s, i, j) - flux_hBy_by(s, i, j - 1)));
}

/* ------------------------------------------------------------------------- */
/* VectorInvariantFormulation tendencies — SWMHD_example.jl:21-33  */
static inline double zeta(const S *s, int i, int j) { /* ζ₃ᶠᶠᶜ */
    return ((s->dy * fV(s, i, j) - s->dy * fV(s, i - 1, j)) -
            (s->dx * fU(s, i, j) - s->dx * fU(s, i, j - 1))) / (s->dx * s->dy);
}
static inline double ut_ff(const S *s, int i, int j) { return 0.5 * (fU(s, i, j - 1) + fU(s, i, j)); } /* ℑyᵃᶠᵃ u */
static inline double vt_ff(const S *s, int i, int j) { return 0.5 * (fV(s, i - 1, j) + fV(s, i, j)); } /* ℑxᶠᵃᵃ v */
static inline double Kin(const S *s, int i, int j) { /* (ℑxᶜ u² + ℑyᶜ v²)/2, ccc */
    double u0 = fU(s, i, j), u1 = fU(s, i + 1, j), v0 = fV(s, i, j), v1 = fV(s, i, j + 1);
    return (0.5 * (u0 * u0 + u1 * u1) + 0.5 * (v0 * v0 + v1 * v1)) / 2.0;
}
/* ζ reconstruction with VelocityStencil smoothness: β_k = ½(β_k[ℑy u] + β_k[ℑx v]) */
static inline double weno_vs(const S *s, const double qz[5], const double qu[5], const double qv[5]) {
    double bu[3], bv[3], b[3];
    weno_beta(qu, bu); weno_beta(qv, bv);
    for (int k = 0; k < 3; k++) b[k] = 0.5 * (bu[k] + bv[k]);
    return weno_blend(s, qz, b);
}

static double Gu_vi(const S *s, int i, int j) { /* fcc */
    double vhat = ixy_fc(s, fV, i, j);
    double zL, zR;
    int f = j + 1; /* ζ along y, to the centre j  (→c convention = face f=j+1) */
    if (ybuf(s, f, 3, 1, YF_HI(s)) && (s->flags & SWMHD_FLAG_WALL_WENO3) && !ybuf(s, f, 2, 1, YF_HI(s))) {
        /* WENO3 with VelocityStencil smoothness: beta_k = (beta_k[ℑy u] + beta_k[ℑx v]) / 2 */
        double z[4], uu[4], vv[4], bu0, bu1, bv0, bv1;
        for (int k = 0; k < 4; k++) { z[k] = zeta(s, i, f - 2 + k); uu[k] = ut_ff(s, i, f - 2 + k); vv[k] = vt_ff(s, i, f - 2 + k); }
        weno3_beta(uu[0], uu[1], uu[2], &bu0, &bu1); weno3_beta(vv[0], vv[1], vv[2], &bv0, &bv1);
        zL = weno3(s, z[0], z[1], z[2], 0.5 * (bu0 + bv0), 0.5 * (bu1 + bv1));
        weno3_beta(uu[3], uu[2], uu[1], &bu0, &bu1); weno3_beta(vv[3], vv[2], vv[1], &bv0, &bv1);
        zR = weno3(s, z[3], z[2], z[1], 0.5 * (bu0 + bv0), 0.5 * (bu1 + bv1));
    } else if (ybuf(s, f, 3, 1, YF_HI(s))) {
        zL = zR = sym2(zeta(s, i, f - 1), zeta(s, i, f));
    } else {
        double qz[5], rz[5], qu[5], ru[5], qv[5], rv[5];
        line_y(s, zeta, i, f, qz, rz); line_y(s, ut_ff, i, f, qu, ru); line_y(s, vt_ff, i, f, qv, rv);
        zL = weno_vs(s, qz, qu, qv); zR = weno_vs(s, rz, ru, rv);
    }
    double adv = upwind(vhat, zL, zR);
    double dK = (Kin(s, i, j) - Kin(s, i - 1, j)) / s->dx;
    double pg = s->g * ((fH(s, i, j) - fH(s, i - 1, j)) / s->dx);
    return (((adv - dK) - pg) + s->f * vhat) + lorentz_force_func_x(s, i, j);
}
static double Gv_vi(const S *s, int i, int j) { /* cfc */
    double uhat = ixy_cf(s, fU, i, j);
    double qz[5], rz[5], qu[5], ru[5], qv[5], rv[5];
    int f = i + 1; /* ζ along x, to the centre i */
    line_x(s, zeta, f, j, qz, rz); line_x(s, ut_ff, f, j, qu, ru); line_x(s, vt_ff, f, j, qv, rv);
    double zL = weno_vs(s, qz, qu, qv), zR = weno_vs(s, rz, ru, rv);
    double adv = upwind(uhat, zL, zR);
    double dK = (Kin(s, i, j) - Kin(s, i, j - 1)) / s->dy;
    double pg = s->g * ((fH(s, i, j) - fH(s, i, j - 1)) / s->dy);
    return (((-adv - dK) - pg) - s->f * uhat) + lorentz_force_func_y(s, i, j);
}
/* advective flux of centre-located c by (u,v): fcc and cfc */
/* probe only: the tracer A reconstructed with a centred interpolant (both upwind sides equal) */
static inline int tracer_centred(const S *s, fn2 C) {
    return C == fA && (s->flags & (SWMHD_FLAG_TRACER_CEN2 | SWMHD_FLAG_TRACER_CEN4));
}
static inline void rec_x(const S *s, fn2 C, int i, int j, double *L, double *R) {
    if (tracer_centred(s, C)) {
        *L = *R = (s->flags & SWMHD_FLAG_TRACER_CEN2) ? sym2(C(s, i - 1, j), C(s, i, j)) : sym_x(s, C, i, j);
        return;
    }
    weno_x(s, C, i, j, L, R);
}
static inline void rec_y(const S *s, fn2 C, int i, int j, double *L, double *R) {
    if (tracer_centred(s, C)) {
        *L = *R = (s->flags & SWMHD_FLAG_TRACER_CEN2) ? sym2(C(s, i, j - 1), C(s, i, j)) : sym_y(s, C, i, j, YC_HI(s));
        return;
    }
    weno_y(s, C, i, j, YC_HI(s), L, R);
}
static inline double adv_flux_x(const S *s, fn2 C, int i, int j) {
    double L, R; rec_x(s, C, i, j, &L, &R);
    return s->dy * upwind(fU(s, i, j), L, R);
}
static inline double adv_flux_y(const S *s, fn2 C, int i, int j) {
    double L, R; rec_y(s, C, i, j, &L, &R);
    return s->dx * upwind(fV(s, i, j), L, R);
}
static inline double div_xy(const S *s, int i, int j) { /* div_xyᶜᶜᶜ(u|uh, v|vh) */
    return s->inv_az * ((s->dy * fU(s, i + 1, j) - s->dy * fU(s, i, j)) +
                        (s->dx * fV(s, i, j + 1) - s->dx * fV(s, i, j)));
}
static double Gh_vi(const S *s, int i, int j) {
    return -(s->inv_az * ((adv_flux_x(s, fH, i + 1, j) - adv_flux_x(s, fH, i, j)) +
                          (adv_flux_y(s, fH, i, j + 1) - adv_flux_y(s, fH, i, j))));
}
static double GA_vi(const S *s, int i, int j) {
    double d = s->inv_az * ((adv_flux_x(s, fA, i + 1, j) - adv_flux_x(s, fA, i, j)) +
                            (adv_flux_y(s, fA, i, j + 1) - adv_flux_y(s, fA, i, j)));
    return -d + fA(s, i, j) * div_xy(s, i, j);
}

/* ------------------------------------------------------------------------- */
/* ConservativeFormulation tendencies — divergence_sw_mhd.jl:19-31 */
static inline double h_ff(const S *s, int i, int j) { return ixy_ff(s, fH, i, j); }
static inline double F_uu(const S *s, int i, int j) { /* ccc */
    double ut = sym_x(s, fU, i + 1, j), L, R;
    weno_x(s, fU, i + 1, j, &L, &R);
    return s->dy * upwind(ut, L, R) / fH(s, i, j);
}
static inline double F_vu(const S *s, int i, int j) { /* ffc: vh carries uh across y-face j */
    double vt = sym_x(s, fV, i, j), L, R;
    weno_y(s, fU, i, j, YC_HI(s), &L, &R);
    return s->dx * upwind(vt, L, R) / h_ff(s, i, j);
}
static inline double F_uv(const S *s, int i, int j) { /* ffc: uh carries vh across x-face i */
    double ut = sym_y(s, fU, i, j, YC_HI(s)), L, R;
    weno_x(s, fV, i, j, &L, &R);
    return s->dy * upwind(ut, L, R) / h_ff(s, i, j);
}
static inline double F_vv(const S *s, int i, int j) { /* ccc */
    double vt = sym_y(s, fV, i, j + 1, YF_HI(s)), L, R;
    weno_y(s, fV, i, j + 1, YF_HI(s), &L, &R);
    return s->dx * upwind(vt, L, R) / fH(s, i, j);
}
static inline double half_g_h2(const S *s, int i, int j) { double h = fH(s, i, j); return (0.5 * s->g) * (h * h); }
static inline double pgrad_x(const S *s, int i, int j) {
    if (s->flags & SWMHD_FLAG_PRESSURE_GHDH)
        return s->g * ixf_h(s, i, j) * ((fH(s, i, j) - fH(s, i - 1, j)) / s->dx);
    return (half_g_h2(s, i, j) - half_g_h2(s, i - 1, j)) / s->dx;
}
static inline double pgrad_y(const S *s, int i, int j) {
    if (s->flags & SWMHD_FLAG_PRESSURE_GHDH)
        return s->g * iyf_h(s, i, j) * ((fH(s, i, j) - fH(s, i, j - 1)) / s->dy);
    return (half_g_h2(s, i, j) - half_g_h2(s, i, j - 1)) / s->dy;
}
static double Guh_c(const S *s, int i, int j) { /* fcc */
    double dm = s->inv_az * ((F_uu(s, i, j) - F_uu(s, i - 1, j)) + (F_vu(s, i, j + 1) - F_vu(s, i, j)));
    return ((-dm - pgrad_x(s, i, j)) + s->f * ixy_fc(s, fV, i, j)) + div_lorentz_x(s, i, j);
}
static double Gvh_c(const S *s, int i, int j) { /* cfc */
    double dm = s->inv_az * ((F_uv(s, i + 1, j) - F_uv(s, i, j)) + (F_vv(s, i, j) - F_vv(s, i, j - 1)));
    return ((-dm - pgrad_y(s, i, j)) - s->f * ixy_cf(s, fU, i, j)) + div_lorentz_y(s, i, j);
}
static double Gh_c(const S *s, int i, int j) { return -div_xy(s, i, j); } /* centred (C7) */
static inline double tr_flux_x(const S *s, int i, int j) { /* transport tracer flux / ℑxᶠ h */
    double L, R; rec_x(s, fA, i, j, &L, &R);
    return s->dy * upwind(fU(s, i, j), L, R) / ixf_h(s, i, j);
}
static inline double tr_flux_y(const S *s, int i, int j) {
    double L, R; rec_y(s, fA, i, j, &L, &R);
    return s->dx * upwind(fV(s, i, j), L, R) / iyf_h(s, i, j);
}
static inline double u_of(const S *s, int i, int j) { return fU(s, i, j) / ixf_h(s, i, j); } /* uh/ℑxᶠh */
static inline double v_of(const S *s, int i, int j) { return fV(s, i, j) / iyf_h(s, i, j); }
static double GA_c(const S *s, int i, int j) {
    double d = s->inv_az * ((tr_flux_x(s, i + 1, j) - tr_flux_x(s, i, j)) +
                            (tr_flux_y(s, i, j + 1) - tr_flux_y(s, i, j)));
    double cdiv;
    if (s->flags & SWMHD_FLAG_CDIVU_OVER_H) cdiv = div_xy(s, i, j) / fH(s, i, j);
    else cdiv = (u_of(s, i + 1, j) - u_of(s, i, j)) / s->dx + (v_of(s, i, j + 1) - v_of(s, i, j)) / s->dy;
    return -d + fA(s, i, j) * cdiv;
}

/* ------------------------------------------------------------------------- */
/* halo filling — upstream fill_halo_regions!, x then y          */
static void fill_halo_field(const swmhd_config *c, double *a, int field) {
    const int Nx = c->Nx, Ny = c->Ny, P = Nx + 6;
    const int rows = rows_of(c, field);
    /* x: periodic, over every row of the parent */
    for (int r = 0; r < rows; r++) {
        double *row = a + (size_t)P * r;
        for (int k = 0; k < 3; k++) { row[k] = row[Nx + k]; row[Nx + 3 + k] = row[3 + k]; }
    }
#define ROW(j) (a + (size_t)P * ((j) + 2))
    if (c->topo_y == SWMHD_PERIODIC) {
        for (int k = 1; k <= 3; k++) {
            memcpy(ROW(1 - k), ROW(Ny + 1 - k), sizeof(double) * P);
            memcpy(ROW(Ny + k), ROW(k), sizeof(double) * P);
        }
    } else if (field == SWMHD_V) {           /* impenetrable walls: v = 0 on j=1, Ny+1 */
        memset(ROW(1), 0, sizeof(double) * P);
        memset(ROW(Ny + 1), 0, sizeof(double) * P);
        if (c->flags & SWMHD_FLAG_V_MIRROR)   /* C10 probe: odd mirror beyond the wall instead of untouched cells */
            for (int k = 1; k <= 3; k++)
                for (int x = 0; x < P; x++) {
                    ROW(1 - k)[x] = -ROW(1 + k)[x];
                    if (k <= 2) ROW(Ny + 1 + k)[x] = -ROW(Ny + 1 - k)[x];
                }
    } else {                                  /* centre-located in y: mirror (+ gradient on A) */
        const int grad = (field == SWMHD_A && c->A_gradient_bc);
        const int depth = (c->flags & SWMHD_FLAG_BC_DEPTH1) ? 1 : 3;   /* C10 probe: deeper halo rows untouched */
        for (int k = 1; k <= depth; k++) {
            double os = grad ? c->A_grad_south * (double)(2 * k - 1) * c->dy : 0.0;
            double on = grad ? c->A_grad_north * (double)(2 * k - 1) * c->dy : 0.0;
            for (int x = 0; x < P; x++) {
                ROW(1 - k)[x] = ROW(k)[x] - os;
                ROW(Ny + k)[x] = ROW(Ny + 1 - k)[x] + on;
            }
        }
    }
#undef ROW
}

int swmhd_oracle_fill_halos(const swmhd_config *c, double *u, double *v, double *h, double *A) {
    fill_halo_field(c, u, SWMHD_U); fill_halo_field(c, v, SWMHD_V);
    fill_halo_field(c, h, SWMHD_H); fill_halo_field(c, A, SWMHD_A);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* calculate_tendencies!: G arrays have the parent layout of their field, only
 * the interior is written.                                                    */
int swmhd_oracle_tendencies(const swmhd_config *c, const double *u, const double *v, const double *h,
                            const double *A, double *Gu, double *Gv, double *Gh, double *GA) {
    S s_; S_init(&s_, c, u, v, h, A);
    const S *s = &s_;
    const int Nx = c->Nx, Ny = c->Ny;
    const int jv0 = s->by ? 2 : 1; /* Bounded-y: wall rows of v stay 0 */
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= Ny; j++) {
        for (int i = 1; i <= Nx; i++) {
            size_t k = IX(s, i, j);
            if (c->formulation == SWMHD_JACOBIAN) {
                Gu[k] = Gu_vi(s, i, j);
                Gv[k] = (j >= jv0) ? Gv_vi(s, i, j) : 0.0;
                Gh[k] = Gh_vi(s, i, j);
                GA[k] = GA_vi(s, i, j);
            } else {
                Gu[k] = Guh_c(s, i, j);
                Gv[k] = (j >= jv0) ? Gvh_c(s, i, j) : 0.0;
                Gh[k] = Gh_c(s, i, j);
                GA[k] = GA_c(s, i, j);
            }
        }
    }
    return 0;
}

/* reference-owned forcing closures alone (for the known-answer scripts) */
int swmhd_oracle_lorentz(const swmhd_config *c, const double *h, const double *A, double *Fx, double *Fy) {
    S s_; S_init(&s_, c, NULL, NULL, h, A);
    const S *s = &s_;
    for (int j = 1; j <= c->Ny; j++)
        for (int i = 1; i <= c->Nx; i++) {
            size_t k = IX(s, i, j);
            if (c->formulation == SWMHD_JACOBIAN) { Fx[k] = lorentz_force_func_x(s, i, j); Fy[k] = lorentz_force_func_y(s, i, j); }
            else { Fx[k] = div_lorentz_x(s, i, j); Fy[k] = div_lorentz_y(s, i, j); }
        }
    return 0;
}

/* 1-D WENO5 for unit tests: left/right value at faces f = 3 .. n-3 of a line */
int swmhd_oracle_weno_line(const swmhd_config *c, const double *psi, int n, double *L, double *R) {
    S s_; memset(&s_, 0, sizeof s_); s_.eps = c->weno_eps; s_.flags = c->flags;
    for (int f = 3; f <= n - 3; f++) {
        double q[5] = { psi[f - 3], psi[f - 2], psi[f - 1], psi[f], psi[f + 1] };
        double r[5] = { psi[f + 2], psi[f + 1], psi[f], psi[f - 1], psi[f - 2] };
        L[f] = weno(&s_, q); R[f] = weno(&s_, r);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* RK3                                                            */
static const double RK_GAMMA[3] = { 8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0 };
static const double RK_ZETA[3]  = { 0.0, -17.0 / 60.0, -5.0 / 12.0 };

/* one substage in place; Gn receives G^n, Gm holds G^- (ignored for stage 1) */
int swmhd_oracle_substage(const swmhd_config *c, double *u, double *v, double *h, double *A,
                          double *const Gn[4], double *const Gm[4], double dt, int stage) {
    if (stage < 1 || stage > 3) return -1;
    swmhd_oracle_tendencies(c, u, v, h, A, Gn[0], Gn[1], Gn[2], Gn[3]);
    double *Uf[4] = { u, v, h, A };
    const int Nx = c->Nx, Ny = c->Ny, P = Nx + 6;
    const double gam = RK_GAMMA[stage - 1], zet = RK_ZETA[stage - 1];
    for (int fld = 0; fld < 4; fld++) {
        double *Uq = Uf[fld]; const double *gn = Gn[fld], *gm = Gm[fld];
#pragma omp parallel for schedule(static)
        for (int j = 1; j <= Ny; j++)
            for (int i = 1; i <= Nx; i++) {
                size_t k = (size_t)(i + 2) + (size_t)P * (j + 2);
                if (stage == 1) Uq[k] = Uq[k] + dt * gam * gn[k];
                else            Uq[k] = Uq[k] + dt * (gam * gn[k] + zet * gm[k]);
            }
    }
    swmhd_oracle_fill_halos(c, u, v, h, A);
    return 0;
}

/* nsteps full RK3 steps in place; halos are valid on entry and on return.
 * clock (may be NULL): [0]=time, [1]=iteration, advanced like upstream's tick!. */
int swmhd_oracle_step(const swmhd_config *c, double *u, double *v, double *h, double *A,
                      double dt, int nsteps, double *clock) {
    size_t len[4]; double *Ga[4], *Gb[4];
    for (int k = 0; k < 4; k++) {
        len[k] = swmhd_oracle_field_len(c, k);
        Ga[k] = (double *)calloc(len[k], sizeof(double));
        Gb[k] = (double *)calloc(len[k], sizeof(double));
        if (!Ga[k] || !Gb[k]) return -2;
    }
    for (int n = 0; n < nsteps; n++) {
        double **Gn = Ga, **Gm = Gb;
        for (int stage = 1; stage <= 3; stage++) {
            swmhd_oracle_substage(c, u, v, h, A, Gn, Gm, dt, stage);
            double **t = Gn; Gn = Gm; Gm = t;   /* store_tendencies!: G⁻ ← Gⁿ */
            if (clock) {
                double sdt = (stage == 1) ? RK_GAMMA[0] * dt : (RK_GAMMA[stage - 1] + RK_ZETA[stage - 1]) * dt;
                clock[0] += sdt;
            }
        }
        if (clock) clock[1] += 1.0;
    }
    for (int k = 0; k < 4; k++) { free(Ga[k]); free(Gb[k]); }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* diagnostics — SWMHD_example.jl:47-63,67-77; divergence_sw_mhd.jl:42-59,63-75 */
static inline double sqU(const S *s, int i, int j) { double x = fU(s, i, j); return x * x; }
static inline double sqV(const S *s, int i, int j) { double x = fV(s, i, j); return x * x; }
/* diagnostic B: B_x = -∂y(A)/h at cfc (h → ℑyᶠ h), B_y = ∂x(A)/h at fcc (h → ℑxᶠ h) */
static inline double sq_dgBx(const S *s, int i, int j) { double b = -dyA(s, i, j) / iyf_h(s, i, j); return b * b; }
static inline double sq_dgBy(const S *s, int i, int j) { double b = dxA(s, i, j) / ixf_h(s, i, j); return b * b; }
static inline double ke_bracket_fc(const S *s, int i, int j) { return sqU(s, i, j) + ixy_fc(s, sqV, i, j); }   /* at fcc */
static inline double me_bracket_cf(const S *s, int i, int j) { return sq_dgBx(s, i, j) + ixy_cf(s, sq_dgBy, i, j); } /* at cfc */

int swmhd_oracle_diagnostics(const swmhd_config *c, const double *u, const double *v, const double *h,
                             const double *A, swmhd_diag *out) {
    S s_; S_init(&s_, c, u, v, h, A);
    const S *s = &s_;
    const int Nx = c->Nx, Ny = c->Ny;
    double ke = 0, me = 0, pe = 0, sumh = 0, maxu = 0, maxA = 0, minh = INFINITY, maxdiv = 0;
    int finite = 1;
    for (int j = 1; j <= Ny; j++) {
        double rke = 0, rme = 0, rpe = 0, rh = 0;
        for (int i = 1; i <= Nx; i++) {
            double hh = fH(s, i, j), aa = fA(s, i, j), uu = fU(s, i, j), vv = fV(s, i, j);
            double half_w = (c->formulation == SWMHD_JACOBIAN) ? 0.5 * hh : 0.5 * (1.0 / hh);
            double kb, mb;
            if (c->flags & SWMHD_FLAG_DIAG_CENTRED) {
                kb = 0.5 * (sqU(s, i, j) + sqU(s, i + 1, j)) + 0.5 * (sqV(s, i, j) + sqV(s, i, j + 1));
                mb = 0.5 * (sq_dgBx(s, i, j) + sq_dgBx(s, i, j + 1)) + 0.5 * (sq_dgBy(s, i, j) + sq_dgBy(s, i + 1, j));
            } else {
                kb = 0.5 * (ke_bracket_fc(s, i, j) + ke_bracket_fc(s, i + 1, j));   /* ℑxᶜ of the fcc bracket */
                mb = 0.5 * (me_bracket_cf(s, i, j) + me_bracket_cf(s, i, j + 1));   /* ℑyᶜ of the cfc bracket */
            }
            rke += half_w * kb;
            rme += (0.5 * hh) * mb;
            double dh = hh - c->h_ref;
            rpe += (0.5 * c->g) * (dh * dh);
            rh += hh;
            double speed = (c->formulation == SWMHD_JACOBIAN) ? fabs(uu) : fabs(uu / ixf_h(s, i, j));
            if (speed > maxu) maxu = speed;
            if (fabs(aa) > maxA) maxA = fabs(aa);
            if (hh < minh) minh = hh;
            double dv = (hBx(s, i + 1, j) - hBx(s, i, j)) / s->dx + (hBy(s, i, j + 1) - hBy(s, i, j)) / s->dy;
            if (fabs(dv) > maxdiv) maxdiv = fabs(dv);
            if (!isfinite(hh) || !isfinite(aa) || !isfinite(uu) || !isfinite(vv)) finite = 0;
        }
        ke += rke; me += rme; pe += rpe; sumh += rh;
    }
    const double n = (double)Nx * (double)Ny, Lx = Nx * c->dx, Ly = Ny * c->dy;
    out->ke = ke / n * Lx * Ly; out->me = me / n * Lx * Ly; out->pe = pe / n * Lx * Ly;
    out->total = out->ke + out->me + out->pe;
    out->max_abs_u = maxu; out->max_abs_A = maxA; out->min_h = minh; out->max_abs_div_hB = maxdiv;
    out->sum_h = sumh; out->all_finite = finite; out->reserved = 0;
    return 0;
}
