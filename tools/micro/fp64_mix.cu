// fp64_mix.cu — do non-FP64 instructions issue in the shadow of the FP64 pipe's 2-cycle cadence?
// Per loop trip and thread: 32 DFMA (4 independent chains) + 32*K independent integer / FP32 / LDS instructions.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int KIND>
__global__ void k(double *out, int iters, double a, double b, int ia) {
    __shared__ double sm[1024];
    double x[4];
    int y[4] = {1, 2, 3, 4};
    float f[4] = {1.f, 2.f, 3.f, 4.f};
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double ld = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                x[i] = fma(x[i], a, b);
#pragma unroll
                for (int q = 0; q < K; q++) {
                    if (KIND == 0) y[i] = y[i] * ia + q;                 // IMAD
                    if (KIND == 1) f[i] = fmaf(f[i], 1.0001f, 0.5f);     // FFMA
                    if (KIND == 2) ld += sm[(threadIdx.x + y[i] + q + u) & 1023];   // LDS + DADD + address
                    if (KIND == 3) y[i] = (y[i] ^ ia) + q;               // LOP3 + IADD (ALU)
                }
            }
    }
    double s = ld;
#pragma unroll
    for (int i = 0; i < 4; i++) s += x[i] + y[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int K, int KIND> void run(const char *name, double *d) {
    const int sms = 148, warps = 16, iters = 2048;
    k<K, KIND><<<sms, warps * 32>>>(d, 8, 1.0000001, 1e-9, 3); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<K, KIND><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_smsp = 32.0 * iters * warps / 4;
    printf("%-8s K=%d: %.2f cycles per DFMA(+K others) per SMSP at 1.95 GHz\n", name, K, ms * 1e-3 * 1.95e9 / dfma_per_smsp);
}
int main() {
    double *d; cudaMalloc(&d, 148 * 1024 * 8);
    run<0, 0>("none", d);
    run<1, 0>("IMAD", d); run<2, 0>("IMAD", d); run<3, 0>("IMAD", d);
    run<1, 1>("FFMA", d); run<2, 1>("FFMA", d);
    run<1, 3>("ALU", d); run<2, 3>("ALU", d);
    run<1, 2>("LDS", d);
    return 0;
}
