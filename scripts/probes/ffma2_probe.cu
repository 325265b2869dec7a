// Packed fp32 (FFMA2: fma.rn.f32x2) against scalar FFMA on the device: warp-instructions per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}

template <int MODE>
__global__ void probe(float* out, int iters, long long* cycles) {
    float v[16];
    uint64_t p[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        v[k] = 0.001f * (threadIdx.x + k);
        float2 f = make_float2(v[k], v[k] + 1.f);
        p[k] = *reinterpret_cast<uint64_t*>(&f);
    }
    const float2 cf = make_float2(1.0001f, 0.9999f), df = make_float2(0.5f, 0.25f);
    const uint64_t c2 = *reinterpret_cast<const uint64_t*>(&cf), d2 = *reinterpret_cast<const uint64_t*>(&df);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) v[k] = fmaf(v[k], 1.0001f, 0.5f);     // 1 FFMA
            if (MODE == 1) p[k] = fma2(p[k], c2, d2);             // 1 FFMA2 (2 flop-lanes per thread)
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { float2 f = *reinterpret_cast<float2*>(&p[k]); s += v[k] + f.x + f.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, threads = 32 * warps_per_sm, blocks = sms, iters = 4000;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&cyc, sizeof(long long) * blocks);
    probe<MODE><<<blocks, threads>>>(out, 10, cyc);
    probe<MODE><<<blocks, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int b = 0; b < blocks; ++b) avg += h[b]; avg /= blocks;
    double winstr = (double)iters * 16 * warps_per_sm;
    printf("%-8s warps/SM %2d: %.0f clk, %.3f warp-instr/clk/SM, %.1f fp32 fma lanes/clk/SM\n", name, warps_per_sm, avg,
           winstr / avg, winstr / avg * 32 * (MODE == 1 ? 2 : 1));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 32}) { run<0>("FFMA", w); run<1>("FFMA2", w); }
    return 0;
}
