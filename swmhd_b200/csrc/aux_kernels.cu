// aux_kernels.cu — halo filling (fill_halo_regions!) and the diagnostic
// reductions of the reference scripts (SWMHD_example.jl:47-63,67-77;
// divergence_sw_mhd.jl:42-59,63-75).
#include "kparams.h"
#include <cmath>

namespace swmhd {
namespace {

// One thread per halo cell.  Every halo cell reads its *interior* source directly
// (x wrap and y wrap/mirror composed), so one launch fills edges and corners with
// no ordering constraint between the x and y passes.
__global__ void halo_kernel(const HaloParams p) {
    const int k = blockIdx.y;
    double *a = p.U[k];
    const int Nx = p.Nx, Ny = p.Ny, P = p.P;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrow = p.j_hi - p.j_lo + 1;
    const int nxw = nrow > 0 ? nrow * 6 : 0;
    if (t < nxw) {                       // periodic x wrap of parent row r
        int r = p.j_lo + t / 6, c = t % 6;
        if (r >= p.rows[k] || r < 3 || r > Ny + 2) return;     // interior rows only: halo rows belong to the y part
        if (k == 1 && p.by && p.y_mode == 2 && p.first && r == 3) return;   // wall row of v: zeroed below
        int dst = c < 3 ? c : Nx + c;    // 0,1,2 | Nx+3,Nx+4,Nx+5
        int src = c < 3 ? Nx + c : c;    // Nx..Nx+2 | 3,4,5
        a[(size_t)r * P + dst] = a[(size_t)r * P + src];
        return;
    }
    t -= nxw;
    if (t >= 6 * P) return;
    const int slot = t / P, pi = t % P;
    const int si = pi < 3 ? pi + Nx : (pi >= Nx + 3 ? pi - Nx : pi);   // x-wrapped source column
    const bool south = slot < 3;
    const int kk = south ? slot + 1 : slot - 2;                          // 1..3
    if (p.y_mode == 0) return;
    if (p.y_mode == 1) {
        // Periodic: c[1-k] = c[Ny+1-k], c[Ny+k] = c[k]
        int dj = south ? 1 - kk : Ny + kk, sj = south ? Ny + 1 - kk : kk;
        a[(size_t)(dj + 2) * P + pi] = a[(size_t)(sj + 2) * P + si];
        return;
    }
    if (south ? !p.first : !p.last) return;
    if (k == 1) {                        // v|vh: impenetrable wall rows j = 1 and Ny+1
        if (kk == 1) a[(size_t)((south ? 1 : Ny + 1) + 2) * P + pi] = 0.0;
        return;
    }
    // centre-located in y: mirror (no-flux), with the gradient offset on A
    int dj = south ? 1 - kk : Ny + kk, sj = south ? kk : Ny + 1 - kk;
    double val = a[(size_t)(sj + 2) * P + si];
    if (k == 3 && p.grad) {
        // explicit roundings (no FMA contraction): the halo must be bit-identical to the oracle's
        if (south) val = __dsub_rn(val, __dmul_rn(__dmul_rn(p.gs, (double)(2 * kk - 1)), p.dy));
        else       val = __dadd_rn(val, __dmul_rn(__dmul_rn(p.gn, (double)(2 * kk - 1)), p.dy));
    }
    a[(size_t)(dj + 2) * P + pi] = val;
}

// ---------------------------------------------------------------------------
struct DiagAcc { double v[NDIAG]; };

__device__ __forceinline__ double warp_sum(double x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_max(double x) {
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_down_sync(0xffffffffu, x, o));
    return x;
}

constexpr int DT = 256;

__global__ void __launch_bounds__(DT) diag_kernel(const DiagParams p) {
    const int Nx = p.Nx, Ny = p.Ny, P = p.P;
    const double *u = p.U[0], *v = p.U[1], *h = p.U[2], *A = p.U[3];
    auto ix = [&](int i, int j) { return (size_t)(i + 2) + (size_t)P * (size_t)(j + 2); };
    auto sq = [](double x) { return x * x; };
    auto ixf_h = [&](int i, int j) { return 0.5 * (h[ix(i - 1, j)] + h[ix(i, j)]); };
    auto iyf_h = [&](int i, int j) { return 0.5 * (h[ix(i, j - 1)] + h[ix(i, j)]); };
    auto dxA = [&](int i, int j) { return (A[ix(i, j)] - A[ix(i - 1, j)]) / p.dx; };
    auto dyA = [&](int i, int j) { return (A[ix(i, j)] - A[ix(i, j - 1)]) / p.dy; };
    // KE bracket u^2 + ℑxyᶠᶜᵃ(v^2) at fcc; ME bracket Bx^2 + ℑxyᶜᶠᵃ(By^2) at cfc
    auto keb = [&](int i, int j) {
        return sq(u[ix(i, j)]) + 0.5 * (0.5 * (sq(v[ix(i - 1, j)]) + sq(v[ix(i, j)])) +
                                        0.5 * (sq(v[ix(i - 1, j + 1)]) + sq(v[ix(i, j + 1)])));
    };
    auto sqBx = [&](int i, int j) { return sq(-dyA(i, j) / iyf_h(i, j)); };
    auto sqBy = [&](int i, int j) { return sq(dxA(i, j) / ixf_h(i, j)); };
    auto meb = [&](int i, int j) {
        return sqBx(i, j) + 0.5 * (0.5 * (sqBy(i, j - 1) + sqBy(i + 1, j - 1)) + 0.5 * (sqBy(i, j) + sqBy(i + 1, j)));
    };
    auto hBx = [&](int i, int j) { return -(0.5 * (0.5 * (dyA(i - 1, j) + dyA(i, j)) + 0.5 * (dyA(i - 1, j + 1) + dyA(i, j + 1)))); };
    auto hBy = [&](int i, int j) { return 0.5 * (0.5 * (dxA(i, j - 1) + dxA(i + 1, j - 1)) + 0.5 * (dxA(i, j) + dxA(i + 1, j))); };

    double ke = 0, me = 0, pe = 0, sh = 0, mu = 0, mA = 0, mh = -INFINITY, md = 0, nf = 0;
    const long long ncell = (long long)Nx * Ny;
    for (long long c = (long long)blockIdx.x * DT + threadIdx.x; c < ncell; c += (long long)gridDim.x * DT) {
        int i = (int)(c % Nx) + 1, j = (int)(c / Nx) + 1;
        double hh = h[ix(i, j)], aa = A[ix(i, j)], uu = u[ix(i, j)], vv = v[ix(i, j)];
        double wgt = (p.form == 0) ? 0.5 * hh : 0.5 * (1.0 / hh);
        ke += wgt * (0.5 * (keb(i, j) + keb(i + 1, j)));
        me += (0.5 * hh) * (0.5 * (meb(i, j) + meb(i, j + 1)));
        double dh = hh - p.h_ref;
        pe += (0.5 * p.g) * (dh * dh);
        sh += hh;
        double speed = (p.form == 0) ? fabs(uu) : fabs(uu / ixf_h(i, j));
        mu = fmax(mu, speed);
        mA = fmax(mA, fabs(aa));
        mh = fmax(mh, -hh);
        md = fmax(md, fabs((hBx(i + 1, j) - hBx(i, j)) / p.dx + (hBy(i, j + 1) - hBy(i, j)) / p.dy));
        if (!(isfinite(hh) && isfinite(aa) && isfinite(uu) && isfinite(vv))) nf += 1.0;
    }
    __shared__ double red[NDIAG][DT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double r[NDIAG] = {warp_sum(ke), warp_sum(me), warp_sum(pe), warp_sum(sh), warp_max(mu),
                       warp_max(mA), warp_max(mh), warp_max(md), warp_sum(nf)};
    if (lane == 0)
        for (int q = 0; q < NDIAG; q++) red[q][wid] = r[q];
    __syncthreads();
    if (threadIdx.x < NDIAG) {
        int q = threadIdx.x;
        double acc = red[q][0];
        for (int w = 1; w < DT / 32; w++) {
            if (q >= 4 && q <= 7) acc = fmax(acc, red[q][w]);
            else acc += red[q][w];
        }
        p.partials[(size_t)blockIdx.x * NDIAG + q] = acc;
    }
}

// Fixed-order two-level final reduction (deterministic for a given launch geometry): block b folds
// the contiguous chunk b of the per-tile partials (thread-strided, then a fixed binary tree) into
// stage[b]; the last block to finish folds stage[0..FB) the same way.  Slots 4..7 are maxima.
// The ticket counter of the "last block folds" pattern is owned by the CONTEXT (one word behind its
// stage[] scratch, zeroed at create): two contexts reducing concurrently on one GPU never share it.
constexpr int FT = 256, FB = 64;

__device__ __forceinline__ void block_fold(double (&v)[NDIAG], double (*sh)[FT]) {
#pragma unroll
    for (int q = 0; q < NDIAG; q++) sh[q][threadIdx.x] = v[q];
    __syncthreads();
    for (int s = FT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
#pragma unroll
            for (int q = 0; q < NDIAG; q++) {
                const double x = sh[q][threadIdx.x + s];
                sh[q][threadIdx.x] = (q >= 4 && q <= 7) ? fmax(sh[q][threadIdx.x], x) : sh[q][threadIdx.x] + x;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FT) diag_final_kernel(const double *partials, int nblocks, double *stage, double *out, unsigned int *ticket) {
    __shared__ double sh[NDIAG][FT];
    __shared__ bool last;
    const int per = (nblocks + FB - 1) / FB;
    const int lo = blockIdx.x * per, hi = min(nblocks, lo + per);
    double v[NDIAG];
#pragma unroll
    for (int q = 0; q < NDIAG; q++) v[q] = (q >= 4 && q <= 7) ? -INFINITY : 0.0;
    for (int b = lo + threadIdx.x; b < hi; b += FT) {
#pragma unroll
        for (int q = 0; q < NDIAG; q++) {
            const double x = partials[(size_t)b * NDIAG + q];
            v[q] = (q >= 4 && q <= 7) ? fmax(v[q], x) : v[q] + x;
        }
    }
    block_fold(v, sh);
    if (threadIdx.x < NDIAG) stage[(size_t)blockIdx.x * NDIAG + threadIdx.x] = sh[threadIdx.x][0];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
#pragma unroll
    for (int q = 0; q < NDIAG; q++) {
        const bool is_max = (q >= 4 && q <= 7);
        v[q] = is_max ? -INFINITY : 0.0;
        if (threadIdx.x < FB) v[q] = stage[(size_t)threadIdx.x * NDIAG + q];
    }
    block_fold(v, sh);
    if (threadIdx.x < NDIAG) out[threadIdx.x] = sh[threadIdx.x][0];
    if (threadIdx.x == 0) *ticket = 0;
}

// ---------------------------------------------------------------------------
// Output quantities of the reference's field writer (SWMHD_example.jl:67-69,81-84;
// divergence_sw_mhd.jl:63-66,77-82): u, v (velocities: uh/ℑx h, vh/ℑy h for the conservative form)
// and the speed s = sqrt(u^2 + ℑxyᶠᶜᵃ(v^2)) at (Face, Center) — AbstractOperations interpolates the
// second operand to the location of the first.  One thread per cell of the region
// [0, Nx+2] x [0, Ny+2]; the caller fills the remaining halo cells with the halo kernel.
__global__ void output_kernel(OutputParams p) {
    const int P = p.P;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // logical i in [0, Nx+2]
    const int j = blockIdx.y;                                     // logical j in [0, Ny+2]
    if (i > p.Nx + 2 || j > p.Ny + 2) return;
    auto ix = [&](int a, int b) { return (size_t)(a + 2) + (size_t)P * (size_t)(b + 2); };
    const double *U = p.U[0], *V = p.U[1], *H = p.U[2];
    auto uvel = [&](int a, int b) { return p.form == 0 ? U[ix(a, b)] : U[ix(a, b)] / (0.5 * (H[ix(a - 1, b)] + H[ix(a, b)])); };
    auto vvel = [&](int a, int b) { return p.form == 0 ? V[ix(a, b)] : V[ix(a, b)] / (0.5 * (H[ix(a, b - 1)] + H[ix(a, b)])); };
    const double u = uvel(i, j), v = vvel(i, j);
    auto sq = [](double x) { return x * x; };
    const double v2 = 0.5 * (0.5 * (sq(vvel(i - 1, j)) + sq(v)) + 0.5 * (sq(vvel(i - 1, j + 1)) + sq(vvel(i, j + 1))));
    p.out_u[ix(i, j)] = u;
    p.out_v[ix(i, j)] = v;
    p.out_s[ix(i, j)] = sqrt(sq(u) + v2);
}

} // namespace

cudaError_t launch_output(const OutputParams &p, cudaStream_t st) {
    dim3 grid((p.Nx + 3 + 127) / 128, p.Ny + 3);
    output_kernel<<<grid, 128, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_halo(const HaloParams &p, cudaStream_t st) {
    int nrow = p.j_hi - p.j_lo + 1;
    int n = (nrow > 0 ? nrow * 6 : 0) + 6 * p.P;
    dim3 grid((n + 255) / 256, 4);
    halo_kernel<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

int diag_blocks(int Nx, int Ny) {
    long long n = ((long long)Nx * Ny + DT - 1) / DT;
    return (int)(n < 148 * 8 ? n : 148 * 8);
}

// stage[] = FB x NDIAG partials + one word for the ticket counter (must be zero before the first launch)
int diag_stage_doubles() { return FB * NDIAG + 1; }
static unsigned int *diag_ticket(double *stage) { return reinterpret_cast<unsigned int *>(stage + FB * NDIAG); }

cudaError_t launch_diag(const DiagParams &p, double *out9, cudaStream_t st) {
    diag_kernel<<<p.nblocks, DT, 0, st>>>(p);
    diag_final_kernel<<<FB, FT, 0, st>>>(p.partials, p.nblocks, p.stage, out9, diag_ticket(p.stage));
    return cudaGetLastError();
}


cudaError_t launch_diag_final(const double *partials, int nblocks, double *stage, double *out9, cudaStream_t st) {
    diag_final_kernel<<<FB, FT, 0, st>>>(partials, nblocks, stage, out9, diag_ticket(stage));
    return cudaGetLastError();
}

} // namespace swmhd
