// device_prims.cuh — device helpers shared by the substage kernels (sm_100a).
//   * TMA (cp.async.bulk.tensor) + mbarrier wrappers
//   * warp reductions
//   * FAST-arithmetic pieces: Newton-refined reciprocal, WENO5-Z in difference form
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace swmhd {

__device__ __forceinline__ double warp_sum(double x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_max(double x) {
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_down_sync(0xffffffffu, x, o));
    return x;
}

// ---- TMA + mbarrier (sm_90+/sm_100a): cp.async.bulk.tensor into shared memory ------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// L2 prefetch of a TMA box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}

// ---- FAST arithmetic ------------------------------------------------------------------------------
// 1/x: rcp.approx (2^-23) + two Newton steps (full double precision for normal x)
__device__ __forceinline__ double frcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// WENO5-Z in difference form.  With the first differences d1..d4 of the five upwind
// samples (a,b,c,d,e) the second differences and the E terms cost one operation each, and every
// candidate is c + X_k/6 with X_0 = 4 d3 - d4, X_1 = d2 + 2 d3, X_2 = 5 d2 - 2 d1, so that
//     sum(w_k p_k) = c + sum(alpha_k C_k X_k / 6) / sum(C_k alpha_k):   a correction to the upwind-side
// cell value instead of a blend of O(1) numbers.
// alpha_k = 1 + tau^2/c_k^2 (c_k = beta_k + eps) is multiplied through by prod(c_k^2): one division.
//
// acc_k += D_k^2 + r E_k^2 with r = (1/4)/(13/12) = 3/13: the smoothness indicators divided by 13/12.
// The Z weights depend only on ratios tau/c_k, so a common scale of (beta_k + eps) is free: eps is scaled too.
// The accumulators are invariant under (d1,d2,d3,d4) -> -(d1,d2,d3,d4).
__device__ __forceinline__ void weno_beta_acc4(double d1, double d2, double d3, double d4,
                                               double &c0, double &c1, double &c2) {
    constexpr double r = 3.0 / 13.0;
    const double D0 = d4 - d3, E0 = fma(-3.0, d3, d4);      // (c,d,e): c-2d+e, 3c-4d+e
    const double D1 = d3 - d2, E1 = d2 + d3;                // (b,c,d): b-2c+d, -(b-d)
    const double D2 = d2 - d1, E2 = fma(3.0, d2, -d1);      // (a,b,c): a-2b+c, a-4b+3c
    c0 = fma(D0, D0, fma(r * E0, E0, c0));
    c1 = fma(D1, D1, fma(r * E1, E1, c1));
    c2 = fma(D2, D2, fma(r * E2, E2, c2));
}
// numerator and denominator of the correction: reconstruction = c + num/den (odd in d1..d4 / even)
__device__ __forceinline__ void weno_corr4(double d1, double d2, double d3, double d4,
                                           double c0, double c1, double c2, double &num, double &den) {
    const double tau = c2 - c0, t2 = tau * tau;
    const double s0 = c0 * c0, s1 = c1 * c1, s2 = c2 * c2;
    const double q2 = s0 * s1, q0 = s1 * s2, q1 = s0 * s2, S = q2 * s2;
    const double a0 = fma(t2, q0, S), a1 = fma(t2, q1, S), a2 = fma(t2, q2, S);
    const double Y0 = fma(0.2, d3, -0.05 * d4);                 // 0.3/6 (4 d3 - d4)
    const double Y1 = fma(0.2, d3, 0.1 * d2);                   // 0.6/6 (d2 + 2 d3)
    const double Y2 = fma(1.0 / 12.0, d2, (-1.0 / 30.0) * d1);  // 0.1/6 (5 d2 - 2 d1)
    num = fma(a0, Y0, fma(a1, Y1, a2 * Y2));
    den = fma(0.3, a0, fma(0.6, a1, 0.1 * a2));
}

// ---- batched forms: N independent reconstructions advanced in lock-step ---------------------------
// The FP64 pipe of an SM sustains its peak only when a warp offers >= 4 independent instructions in a
// row (tools/micro/fp64_pipe.cu: 2.17 cycles per warp instruction at ILP 4, 3.0 at ILP 1 with any
// number of warps; dependent-issue latency 8 cycles).  ptxas keeps the source order of independent
// chains, so the chains are interleaved here, statement by statement.  Arithmetic per chain is that of
// weno_beta_acc4 / weno_corr4 / frcp.
#define FORN _Pragma("unroll") for (int n = 0; n < N; n++)
template <int N>
__device__ __forceinline__ void beta_acc_n(const double (&d1)[N], const double (&d2)[N], const double (&d3)[N], const double (&d4)[N],
                                           double (&c0)[N], double (&c1)[N], double (&c2)[N]) {
    constexpr double r = 3.0 / 13.0;
    double D0[N], E0[N], D1[N], E1[N], D2[N], E2[N];
    FORN D0[n] = d4[n] - d3[n];
    FORN E0[n] = fma(-3.0, d3[n], d4[n]);
    FORN D1[n] = d3[n] - d2[n];
    FORN E1[n] = d2[n] + d3[n];
    FORN D2[n] = d2[n] - d1[n];
    FORN E2[n] = fma(3.0, d2[n], -d1[n]);
    double t0[N], t1[N], t2[N];
    FORN t0[n] = r * E0[n];
    FORN t1[n] = r * E1[n];
    FORN t2[n] = r * E2[n];
    FORN t0[n] = fma(t0[n], E0[n], c0[n]);
    FORN t1[n] = fma(t1[n], E1[n], c1[n]);
    FORN t2[n] = fma(t2[n], E2[n], c2[n]);
    FORN c0[n] = fma(D0[n], D0[n], t0[n]);
    FORN c1[n] = fma(D1[n], D1[n], t1[n]);
    FORN c2[n] = fma(D2[n], D2[n], t2[n]);
}
template <int N>
__device__ __forceinline__ void corr_n(const double (&d1)[N], const double (&d2)[N], const double (&d3)[N], const double (&d4)[N],
                                       const double (&c0)[N], const double (&c1)[N], const double (&c2)[N],
                                       double (&num)[N], double (&den)[N]) {
    double tau[N], s0[N], s1[N], s2[N], q0[N], q1[N], q2[N], S[N], Y0[N], Y1[N], Y2[N];
    FORN tau[n] = c2[n] - c0[n];
    FORN s0[n] = c0[n] * c0[n];
    FORN s1[n] = c1[n] * c1[n];
    FORN s2[n] = c2[n] * c2[n];
    FORN Y0[n] = -0.05 * d4[n];
    FORN Y1[n] = 0.1 * d2[n];
    FORN Y2[n] = (-1.0 / 30.0) * d1[n];
    FORN tau[n] = tau[n] * tau[n];
    FORN q2[n] = s0[n] * s1[n];
    FORN q0[n] = s1[n] * s2[n];
    FORN q1[n] = s0[n] * s2[n];
    FORN Y0[n] = fma(0.2, d3[n], Y0[n]);                         // 0.3/6 (4 d3 - d4)
    FORN Y1[n] = fma(0.2, d3[n], Y1[n]);                         // 0.6/6 (d2 + 2 d3)
    FORN Y2[n] = fma(1.0 / 12.0, d2[n], Y2[n]);                  // 0.1/6 (5 d2 - 2 d1)
    FORN S[n] = q2[n] * s2[n];
    FORN q0[n] = fma(tau[n], q0[n], S[n]);                       // a0
    FORN q1[n] = fma(tau[n], q1[n], S[n]);                       // a1
    FORN q2[n] = fma(tau[n], q2[n], S[n]);                       // a2
    FORN num[n] = q2[n] * Y2[n];
    FORN den[n] = 0.1 * q2[n];
    FORN num[n] = fma(q1[n], Y1[n], num[n]);
    FORN den[n] = fma(0.6, q1[n], den[n]);
    FORN num[n] = fma(q0[n], Y0[n], num[n]);
    FORN den[n] = fma(0.3, q0[n], den[n]);
}
template <int N>
__device__ __forceinline__ void rcp_n(const double (&x)[N], double (&r)[N]) {
    double e[N];
    FORN asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[n]) : "d"(x[n]));
    FORN e[n] = fma(-x[n], r[n], 1.0);
    FORN r[n] = fma(r[n], e[n], r[n]);
    FORN e[n] = fma(-x[n], r[n], 1.0);
    FORN r[n] = fma(r[n], e[n], r[n]);
}

// ---- round 2: cheaper forms of the same arithmetic (FAST mode only; differences are round-off) ----------
// corr10_n: numerator and denominator of the WENO correction scaled by 10, so that the optimal weights
// (3, 6, 1) and all but two of the candidate coefficients are FP64 immediates (no constant registers) and
// two operations per reconstruction disappear:
//     sum(w_k p_k) = c + [a0 (2 d3 - d4/2) + a1 (d2 + 2 d3) + a2 (5/6 d2 - 1/3 d1)] / [3 a0 + 6 a1 + a2]
template <int N>
__device__ __forceinline__ void corr10_n(const double (&d1)[N], const double (&d2)[N], const double (&d3)[N], const double (&d4)[N],
                                         const double (&c0)[N], const double (&c1)[N], const double (&c2)[N],
                                         double (&num)[N], double (&den)[N]) {
    double tau[N], s0[N], s1[N], s2[N], q0[N], q1[N], q2[N], S[N], Z0[N], Z1[N], Z2[N];
    FORN tau[n] = c2[n] - c0[n];
    FORN s0[n] = c0[n] * c0[n];
    FORN s1[n] = c1[n] * c1[n];
    FORN s2[n] = c2[n] * c2[n];
    FORN Z0[n] = -0.5 * d4[n];
    FORN Z1[n] = fma(2.0, d3[n], d2[n]);                         // X1     = d2 + 2 d3
    FORN Z2[n] = (-1.0 / 3.0) * d1[n];
    FORN tau[n] = tau[n] * tau[n];
    FORN q2[n] = s0[n] * s1[n];
    FORN q0[n] = s1[n] * s2[n];
    FORN q1[n] = s0[n] * s2[n];
    FORN Z0[n] = fma(2.0, d3[n], Z0[n]);                         // X0 / 2 = 2 d3 - d4 / 2
    FORN Z2[n] = fma(5.0 / 6.0, d2[n], Z2[n]);                   // X2 / 6 = 5/6 d2 - 1/3 d1
    FORN S[n] = q2[n] * s2[n];
    FORN q0[n] = fma(tau[n], q0[n], S[n]);                       // a0
    FORN q1[n] = fma(tau[n], q1[n], S[n]);                       // a1
    FORN q2[n] = fma(tau[n], q2[n], S[n]);                       // a2
    FORN num[n] = q2[n] * Z2[n];
    FORN den[n] = fma(6.0, q1[n], q2[n]);
    FORN num[n] = fma(q1[n], Z1[n], num[n]);
    FORN den[n] = fma(3.0, q0[n], den[n]);
    FORN num[n] = fma(q0[n], Z0[n], num[n]);
}
// Reciprocals of mixed accuracy: entries [0, N1) get ONE Newton step after rcp.approx (relative error
// ~1e-12: they divide a WENO correction that is itself O(dx^2)..O(1e-1) of the reconstructed value, the
// result is within round-off of the two-step form), entries [N1, N) two steps (full precision: 1/h).
template <int N, int N1>
__device__ __forceinline__ void rcp_mix_n(const double (&x)[N], double (&r)[N]) {
    double e[N];
    FORN asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[n]) : "d"(x[n]));
    FORN e[n] = fma(-x[n], r[n], 1.0);
    FORN r[n] = fma(r[n], e[n], r[n]);
#pragma unroll
    for (int n = N1; n < N; n++) e[n] = fma(-x[n], r[n], 1.0);
#pragma unroll
    for (int n = N1; n < N; n++) r[n] = fma(r[n], e[n], r[n]);
}
// x > 0 from the sign and exponent word: an integer compare instead of an FP64-pipe DSETP.  Identical for
// every normal x (positive denormals below 2^-1042 count as zero; they select the other upwind side of a
// flux that is then multiplied by that denormal).
__device__ __forceinline__ bool gt0(double x) { return __double2hiint(x) > 0; }

// ---- warp reduction of the fused diagnostics -------------------------------------------------------------
// Extrema travel as order-preserving 64-bit keys (key(x) < key(y) <=> x < y for all non-NaN doubles; NaN sorts
// above +inf), so that the warp maximum is two REDUX instructions (high word, then the low words of the lanes that
// hold the maximal high word) instead of a five-step shuffle tree of emulated FP64 maxima.
__device__ __forceinline__ unsigned long long ord_key(double x) {
    const long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double ord_val(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k ^ 0x8000000000000000ull) : ~k));
}
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long k) {
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}
// Four sums at once: a reduce-scatter over the lanes (each step halves the quantities a lane carries), 6 shuffles
// instead of 20.  The totals of q[0..3] end up in lanes 0, 8, 16, 24 (returned in every lane of those groups' leaders);
// the summation order is fixed by the lane numbers: bit-reproducible.
__device__ __forceinline__ double warp_sum4(const double (&q)[4], int lane) {
    const bool up16 = lane & 16, up8 = lane & 8;
    const double k0 = up16 ? q[2] : q[0], k1 = up16 ? q[3] : q[1];
    const double s0 = up16 ? q[0] : q[2], s1 = up16 ? q[1] : q[3];
    const double r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
    const double r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
    double r = (up8 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, up8 ? r0 : r1, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;           // lanes 0..7: q[0], 8..15: q[1], 16..23: q[2], 24..31: q[3]
}
} // namespace swmhd
