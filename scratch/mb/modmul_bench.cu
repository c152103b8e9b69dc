// Microbenchmark: sustained Montgomery-product throughput of ff.cuh under different occupancies / ILP.
#include "../../za_b200/csrc/ff.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace za;
template <int ILP>
__global__ void __launch_bounds__(128) k(Fq* out, int iters) {
    Fq x[ILP], y[ILP];
    for (int j = 0; j < ILP; j++) { for (int i = 0; i < 8; i++) { x[j].v[i] = threadIdx.x * 7 + i + j; y[j].v[i] = blockIdx.x + 3 * i + j; } x[j].v[7] &= 0x0fffffff; y[j].v[7] &= 0x0fffffff; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = x[j] * y[j];
    }
    Fq acc = x[0];
    for (int j = 1; j < ILP; j++) acc = acc + x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int ILP>
void run(int blocks_per_sm, int iters) {
    int sms = 148;
    Fq* out; cudaMalloc(&out, (size_t)sms * blocks_per_sm * 128 * sizeof(Fq));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<ILP><<<sms * blocks_per_sm, 128>>>(out, 10);
    cudaEventRecord(a);
    k<ILP><<<sms * blocks_per_sm, 128>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double mm = (double)sms * blocks_per_sm * 128 * iters * ILP;
    printf("ILP=%d warps/SM=%d: %.2f G modmul/s = %.2f T IMAD-class/s (x264)  err=%s\n", ILP, blocks_per_sm * 4, mm / ms / 1e6, mm * 264 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    for (int bps : {1, 2, 4, 6, 8}) run<1>(bps, 2000);
    for (int bps : {1, 2, 4}) run<2>(bps, 1000);
    for (int bps : {1, 2, 4}) run<4>(bps, 500);
    return 0;
}
