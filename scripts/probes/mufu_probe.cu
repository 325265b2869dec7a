// Measures MUFU (ex2 / rcp) and tcgen05.ld throughput on the device: warp-instructions per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void probe(float* out, int iters, long long* cycles) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.001f * (threadIdx.x + k);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) v[k] = ex2a(v[k]);                      // 1 MUFU.EX2
            if (MODE == 1) v[k] = rcpa(v[k]);                      // 1 MUFU.RCP
            if (MODE == 2) v[k] = rcpa(1.f + ex2a(v[k]));          // EX2 + FADD + RCP (the sigmoid chain)
            if (MODE == 3) v[k] = lg2a(v[k]);
            if (MODE == 4) v[k] = fmaf(v[k], 1.0001f, 0.5f);       // FFMA reference
            if (MODE == 5) v[k] = rcpa(1.f + ex2a(fminf(v[k], 60.f)));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int mufu_per_elem, int warps_per_sm) {
    int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    int sms = prop.multiProcessorCount;
    int threads = 32 * warps_per_sm, blocks = sms, iters = 2000;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&cyc, sizeof(long long) * blocks);
    probe<MODE><<<blocks, threads>>>(out, 10, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int b = 0; b < blocks; ++b) avg += h[b]; avg /= blocks;
    double warp_instr = (double)iters * 16 * warps_per_sm;               // element-steps per SM
    printf("%-28s warps/SM %2d: %.3f ms, %.0f clk/SM, %.2f clk per warp-step/SM, %.2f lanes/clk/SM per MUFU-op (x%d ops)\n", name,
           warps_per_sm, ms, avg, avg / warp_instr, 32.0 * warp_instr * mufu_per_elem / avg, mufu_per_elem);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("ex2", 1, w);
        run<1>("rcp", 1, w);
        run<3>("lg2", 1, w);
        run<2>("rcp(1+ex2(x))", 2, w);
        run<5>("rcp(1+ex2(min(x,60)))", 2, w);
        run<4>("ffma", 1, w);
    }
    return 0;
}
