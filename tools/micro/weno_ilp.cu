// weno_ilp.cu — FP64 pipe efficiency of the WENO5-Z arithmetic: four reconstructions per loop trip,
// written chain after chain vs interleaved statement by statement (device_prims.cuh batched forms).
#include <cstdio>
#include "../../swmhd_b200/csrc/device_prims.cuh"
using namespace swmhd;
template <int MODE>
__global__ void __launch_bounds__(128, 3) k(double *out, int iters, double a) {
    double d[4][4], acc[4] = {0, 0, 0, 0};
    for (int n = 0; n < 4; n++) for (int j = 0; j < 4; j++) d[n][j] = 1e-3 * (threadIdx.x + 1) * (n + 1) + 1e-4 * j;
    const double es = 1e-6;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int n = 0; n < 4; n++) {
                double c0 = es, c1 = es, c2 = es, num, den;
                weno_beta_acc4(d[n][0], d[n][1], d[n][2], d[n][3], c0, c1, c2);
                weno_corr4(d[n][0], d[n][1], d[n][2], d[n][3], c0, c1, c2, num, den);
                acc[n] = fma(num, frcp(den), acc[n]);
            }
        } else {
            double d1[4], d2[4], d3[4], d4[4], c0[4] = {es, es, es, es}, c1[4] = {es, es, es, es}, c2[4] = {es, es, es, es}, num[4], den[4], rc[4];
#pragma unroll
            for (int n = 0; n < 4; n++) { d1[n] = d[n][0]; d2[n] = d[n][1]; d3[n] = d[n][2]; d4[n] = d[n][3]; }
            beta_acc_n<4>(d1, d2, d3, d4, c0, c1, c2);
            corr_n<4>(d1, d2, d3, d4, c0, c1, c2, num, den);
            rcp_n<4>(den, rc);
#pragma unroll
            for (int n = 0; n < 4; n++) acc[n] = fma(num[n], rc[n], acc[n]);
        }
#pragma unroll
        for (int n = 0; n < 4; n++) d[n][it & 3] = fma(acc[n], a, d[n][it & 3]);   // keep the inputs live and changing
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}
template <int MODE> void run(const char *name, double *d, int ctas_per_sm) {
    const int sms = 148, iters = 2000;
    k<MODE><<<sms * ctas_per_sm, 128>>>(d, 8, 1e-9); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas_per_sm, 128>>>(d, iters, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d + sms * ctas_per_sm * 128, 8, cudaMemcpyDeviceToHost);
    // 4 reconstructions of ~49 FP64 instructions per trip and warp; warps per SMSP = ctas_per_sm
    const double fp64_per_smsp = 200.0 * iters * ctas_per_sm;        // warp instructions per SMSP (200 per trip and warp)
    printf("%-12s %d warps/SMSP: block-0 cycles %.0f, event %.3f ms -> %.2f cycles per FP64 instr per SMSP (clock64), %.2f ns-based at 1.95 GHz\n",
           name, ctas_per_sm, cyc, ms, cyc / fp64_per_smsp, ms * 1e-3 * 1.95e9 / fp64_per_smsp);
}
int main() {
    double *d; cudaMalloc(&d, 148 * 8 * 128 * 8 + 64);
    for (int c : {1, 2, 3, 6}) { run<0>("sequential", d, c); run<1>("interleaved", d, c); }
    return 0;
}
