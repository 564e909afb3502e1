// fp64_ops.cu — issue cost of DFMA / DADD / DMUL / mixes on the FP64 pipe (ILP 4, 12 warps per SM)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double *out, int iters, double a, double b) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (OP == 0) x[i] = fma(x[i], a, b);
                if (OP == 1) x[i] = __dadd_rn(x[i], b);
                if (OP == 2) x[i] = __dmul_rn(x[i], a);
                if (OP == 3) { if (u & 1) x[i] = __dadd_rn(x[i], b); else x[i] = fma(x[i], a, b); }
                if (OP == 4) { if (u & 1) x[i] = __dmul_rn(x[i], a); else x[i] = fma(x[i], a, b); }
                if (OP == 5) { if (u & 1) x[i] = __dmul_rn(x[i], a); else x[i] = __dadd_rn(x[i], b); }
                if (OP == 6) { x[i] = fma(x[i], a, b); asm volatile("" ::: "memory"); }
                if (OP == 7) { x[i] = (x[i] > 0.5) ? fma(x[i], a, b) : -x[i]; }   // DFMA + DSETP + 2 FSEL
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}
template <int OP>
void run(const char *name, double *d) {
    const int warps = 12, sms = 148, iters = 4096;
    k<OP><<<sms, warps * 32>>>(d, 16, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    k<OP><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    double cyc; cudaMemcpy(&cyc, d + sms * warps * 32, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.3f cycles per (source-level) op per SMSP\n", name, cyc / ((double)iters * 32 * warps / 4.0));
}
int main() {
    double *d; cudaMalloc(&d, 148 * 1024 * 8 + 64);
    run<0>("DFMA", d); run<1>("DADD", d); run<2>("DMUL", d); run<3>("DFMA/DADD alternating", d);
    run<4>("DFMA/DMUL alternating", d); run<5>("DADD/DMUL alternating", d); run<7>("DFMA + DSETP + select", d);
    return 0;
}
