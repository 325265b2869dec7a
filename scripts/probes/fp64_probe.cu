// Measures FP64 (DFMA/DADD), 64-bit integer add and F2F throughput: lanes per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(double* out, int iters, long long* cycles) {
    double v[16]; long long q[16]; float f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k] = 0.001 * (threadIdx.x + k); q[k] = threadIdx.x + k; f[k] = 0.5f + k; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) v[k] = fma(v[k], 1.0000001, 0.5);
            if (MODE == 1) v[k] = v[k] + 1.5;
            if (MODE == 2) q[k] = q[k] + (q[(k + 1) & 15] | 1);
            if (MODE == 3) v[k] = (double)(float)v[k] + 1.0;      // F2F f64->f32->f64 + DADD
            if (MODE == 4) f[k] = fmaf(f[k], 1.0001f, 0.5f);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k] + (double)q[k] + f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, threads = 32 * warps_per_sm, iters = 1000;
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * threads); cudaMalloc(&cyc, sizeof(long long) * sms);
    probe<MODE><<<sms, threads>>>(out, 10, cyc);
    probe<MODE><<<sms, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int b = 0; b < sms; ++b) avg += h[b]; avg /= sms;
    double steps = (double)iters * 16 * warps_per_sm;
    printf("%-22s warps/SM %2d: %.2f clk per warp-op per SM, %.2f lanes/clk/SM\n", name, warps_per_sm, avg / steps, 32.0 * steps / avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 16, 32}) {
        run<0>("dfma", w); run<1>("dadd", w); run<2>("int64 add", w); run<3>("f2f+f2f+dadd", w); run<4>("ffma", w);
    }
    return 0;
}
