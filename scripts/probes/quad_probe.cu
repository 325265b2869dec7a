// What the pipes allow for the shared-plan inner loop: per quad of background columns 7 packed fp32 ops (FMUL2, 2 FADD2,
// 4 FFMA2) + 2 MUFU.RCP, data in registers (no tensor-memory loads), at several warps per SM.  Variants:
//   0: the loop as shipped (2 reciprocals per quad)            1: one reciprocal per quad (+3 scalar multiplies)
//   2: packed ops only (no MUFU)                               3: MUFU only
// Prints clocks per quad per warp-scheduler slot (SM sub-partition): the MUFU floor is 16, the FMA-pipe floor 14.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o quad_probe quad_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__device__ __forceinline__ void quad(f32x2 A2, f32x2 AA2, f32x2 AA2x2, f32x2 ds, f32x2 dq, f32x2 one2, f32x2 two2, f32x2& a1, f32x2& a0) {
    const f32x2 sm = mul2(A2, ds);
    const f32x2 t1 = add2(sm, one2);
    const f32x2 den = fma2(AA2, dq, t1);
    const f32x2 w = fma2(AA2x2, dq, sm);
    float dlo, dhi;
    upk(den, dlo, dhi);
    f32x2 r;
    if (MODE == 0 || MODE == 3) r = pk(rcp(dlo), rcp(dhi));
    else if (MODE == 1) { const float rr = rcp(dlo * dhi); r = pk(rr * dhi, rr * dlo); }
    else r = den;
    if (MODE == 3) { a1 = add2(a1, r); return; }
    a1 = fma2(r, add2(sm, two2), a1);
    a0 = fma2(r, w, a0);
}

template <int MODE>
__global__ void probe(float* out, int iters, long long* cycles) {
    f32x2 ds[8], dq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ds[k] = pk(0.5f + 0.01f * (threadIdx.x & 7) + 0.1f * k, 0.7f + 0.05f * k); dq[k] = pk(0.2f + 0.03f * k, 0.1f + 0.02f * k); }
    const f32x2 one2 = pk(1.f, 1.f), two2 = pk(2.f, 2.f);
    f32x2 a1[2] = {pk(0.f, 0.f), pk(0.f, 0.f)}, a0[2] = {pk(0.f, 0.f), pk(0.f, 0.f)};
    float A = 1.0f + 0.001f * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        A = A * 1.0001f;
        const float AA = A * A;
        const f32x2 A2 = pk(A, A), AA2 = pk(AA, AA), AAx = pk(2.f * AA, 2.f * AA);
#pragma unroll
        for (int rep = 0; rep < 3; ++rep)              // 24 quads per "instance" (96 columns)
#pragma unroll
            for (int k = 0; k < 8; ++k) quad<MODE>(A2, AA2, AAx, ds[k], dq[k], one2, two2, a1[k & 1], a0[k & 1]);
    }
    long long t1 = clock64();
    float x0, x1, y0, y1;
    upk(add2(a1[0], a1[1]), x0, x1); upk(add2(a0[0], a0[1]), y0, y1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + y0 + y1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, threads = 32 * warps_per_sm, blocks = sms, iters = 2000;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&cyc, sizeof(long long) * blocks);
    probe<MODE><<<blocks, threads>>>(out, 10, cyc);
    probe<MODE><<<blocks, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int b = 0; b < blocks; ++b) avg += h[b]; avg /= blocks;
    const double quads_per_smsp = (double)iters * 24 * warps_per_sm / 4.0;
    printf("%-28s warps/SM %2d: %9.0f clk, %6.2f clk per quad per sub-partition\n", name, warps_per_sm, avg, avg / quads_per_smsp);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 20, 24, 32}) {
        run<0>("7 packed + 2 rcp (shipped)", w);
        run<1>("7 packed + 3 mul + 1 rcp", w);
        run<2>("7 packed only", w);
        run<3>("4 packed + 2 rcp", w);
    }
    return 0;
}
