// fp64_pipe.cu — what the FP64 pipe of one SM sustains: DFMA issue rate vs warps/SM and ILP,
// and the dependent-issue latency.  nvcc -arch=sm_100a -O3 fp64_pipe.cu -o fp64_pipe
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}
template <int ILP>
void run(int warps, double *d) {
    int iters = 4096;
    int sms = 148;
    k<ILP><<<sms, warps * 32>>>(d, 16, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d + sms * warps * 32, 8, cudaMemcpyDeviceToHost);
    double ninst_per_smsp = (double)iters * 8 * ILP * warps / 4.0;  // warp-instr per SMSP
    printf("warps/SM %2d ILP %d: %.3f cycles per warp-DFMA per SMSP (pipe), per-warp interval %.2f cyc, %.2f TFLOP/s\n", warps, ILP,
           cyc / ninst_per_smsp, cyc / ((double)iters * 8 * ILP), 2.0 * iters * 8 * ILP * warps * 32 * sms / (ms * 1e-3) / 1e12);
}
int main() {
    double *d; cudaMalloc(&d, 148 * 1024 * 8 + 64);
    for (int w : {4, 8, 12, 16, 24, 32}) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); }
    return 0;
}
