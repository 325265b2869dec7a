// Shared-plan fast path of the fused coalition evaluation (binary-logistic head).
//
// When every instance of a bucket evaluates the SAME coalition plan and all G groups vary (the common case: on
// Adult-shaped data every instance has M = G), the masked score separates:
//     t(i, s, j) = a(i, s) + d(s, j),   a(i, s) = scale * sum_k z_sk XW_i[k],   d(s, j) = scale * (score_j - sum_k z_sk BW[j][k])
// so  2^t = A(i, s) * Dm(s, j)  with Dm = 2^d independent of the instance.  Dm (S x N floats, 0.8 MB for Adult) is
// computed once per plan, row-normalised; a warp owns 32 coalition rows and streams the instances through them.  Two
// elements share a reciprocal,  p1a + p1b = (2 + sm) / (1 + sm + q),  sm = A (Dma + Dmb),  q = A^2 (Dma Dmb),  so what the
// kernel keeps per row are the pair sums and pair products of Dm -- in TENSOR MEMORY (explain_shared_tmem_kernel, the
// default) or, in the first version kept for comparisons, the raw row in registers (explain_shared_kernel,
// DKS_SHARED_DM=regs).  No GEMM, no EX2 per element: 3.5 packed-fp32 lane-ops + 0.5 MUFU.  Output: (sum p1, sum p0) per
// (instance, coalition); wls_pmat_kernel / wls_shared_kernel apply the link and solve with what the plan precomputed.
// Instances with a partial varying set, per-instance plans and other heads go through the general kernels.
#pragma once

#include "dks_kernels.cuh"
#include "dks_tc.cuh"

namespace dks {
namespace shared_path {

constexpr int MAXN = 128;            // background rows per launch (a warp's slice of tensor memory / a lane's registers)
constexpr int WARPS_PER_CTA = 12;
constexpr float U_CLAMP = 1.152921504606846976e18f;   // 2^60: (1 + ua)(1 + ub) stays finite in fp32

// d(s, j) = scale * (score_j - sum_k z_sk BW[j][k])  for the full varying set (k = group index), in log2 units.
// Rows are normalised: dme[s] = rint(max_j d(s, j)), DmT[j][s] = 2^(d(s, j) - dme[s]) <= sqrt(2), float32, transposed so that
// consecutive coalitions are contiguous (coalesced row loads by lanes).  The exponent goes back in through a(i, s), so
// products of two entries of a row (the kernel stores pair sums and pair products) cannot overflow or vanish together.
__device__ __forceinline__ double plan_d(const uint64_t* __restrict__ zz, const double* __restrict__ BW,
                                         const double* __restrict__ scores, int j, int G, double scale) {
    double c = 0.0;
    for (int k = 0; k < G; ++k)
        if ((zz[k >> 6] >> (k & 63)) & 1ull) c += BW[(size_t)j * G + k];
    return scale * (scores[j] - c);
}
__global__ void plan_dme_kernel(const uint64_t* __restrict__ z, int W, int S, int S_pad, const double* __restrict__ BW,
                                const double* __restrict__ scores, int N, int G, double scale, double* __restrict__ dme) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S_pad) return;
    double mx = 0.0;
    if (s < S) {
        mx = -1.0e300;
        for (int j = 0; j < N; ++j) mx = fmax(mx, plan_d(z + (size_t)s * W, BW, scores, j, G, scale));
        mx = rint(mx);
    }
    dme[s] = mx;
}
__global__ void plan_dm_kernel(const uint64_t* __restrict__ z, int W, int S, int S_pad, const double* __restrict__ BW,
                               const double* __restrict__ scores, int N, int G, double scale, const double* __restrict__ dme,
                               float* __restrict__ DmT) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * S_pad) return;
    const int j = idx / S_pad, s = idx - j * S_pad;
    float out = 0.f;
    if (s < S) out = (float)exp2(plan_d(z + (size_t)s * W, BW, scores, j, G, scale) - dme[s]);
    DmT[idx] = out;
}

struct SharedParams {
    int n, N, G, S, S_pad;
    double scale;
    const float* DmT;        // [N][S_pad] rows normalised by 2^-dme[s]
    const double* dme;       // [S_pad]
    const uint64_t* z;       // [S][W]
    const double* XT;        // [n][ceil(G/4)][16] nibble tables: scale * sum of the contributions a nibble selects
    const int* list;         // instances on this path
    const int* count;        // their number (device)
    float2* sums;            // [n][S_pad] (sum p1, sum p0)
    int accumulate;          // add to sums instead of overwriting them (second and later background chunks)
    float* acache;           // [n][S_pad] A(i, s) of sixteen-word rows, kept between the launches of the background chunks
    int acache_mode;         // 0: off, 1: this launch writes it (first chunk), 2: this launch reads it
};

// Two sigmoids with one reciprocal.  With ua = 2^ta, ub = 2^tb:  (1+ua)(1+ub) = 1 + sm + q,  sm = ua + ub, q = ua*ub
//   p1a + p1b = (2 + sm) / (1 + sm + q)          p0a + p0b = (sm + 2q) / (1 + sm + q)
// 12 FP32-pipe ops + 1 MUFU per pair.  CLAMP: bound u at 2^60 so q cannot overflow (only needed for extreme scores).
template <bool CLAMP>
__device__ __forceinline__ void pair_acc(float A, float dma, float dmb, float& a1, float& a0) {
    float ua = A * dma, ub = A * dmb;
    if (CLAMP) { ua = fminf(ua, U_CLAMP); ub = fminf(ub, U_CLAMP); }
    const float q = ua * ub, sm = ua + ub;
    const float t1 = 1.f + sm;
    const float r = rcp_approx(fmaf(ua, ub, t1));
    a1 = fmaf(r, t1 + 1.f, a1);
    a0 = fmaf(r, fmaf(2.f, q, sm), a0);
}
__device__ __forceinline__ void single_acc(float A, float dma, float& a1, float& a0) {
    const float ua = fminf(A * dma, U_CLAMP);
    const float r = rcp_approx(1.f + ua);
    a1 += r;
    a0 = fmaf(ua, r, a0);
}

// Four sigmoids (two pairs, each sharing one reciprocal) in packed arithmetic: da = (dm0, dm1), db = (dm2, dm3) pair up as
// (0,2) and (1,3).  10 packed ops + 2 MUFU per four elements (pair_acc: 11 + 1 per two).
__device__ __forceinline__ void quad_acc(f32x2 A2, f32x2 da, f32x2 db, f32x2 one2, f32x2 two2, f32x2& a1, f32x2& a0) {
    const f32x2 u = f2_mul(A2, da), v = f2_mul(A2, db);
    const f32x2 q = f2_mul(u, v), sm = f2_add(u, v);
    const f32x2 t1 = f2_add(sm, one2);
    const f32x2 den = f2_add(q, t1);
    float dlo, dhi;
    f2_unpack(den, dlo, dhi);
    const f32x2 r = f2_pack(rcp_approx(dlo), rcp_approx(dhi));
    a1 = f2_fma(r, f2_add(t1, one2), a1);
    a0 = f2_fma(r, f2_fma(two2, q, sm), a0);
}

// NTAIL = N % 16 (compile time): the last, partial chunk is straight-line code
template <bool CLAMP, int NTAIL>
__device__ __forceinline__ void row_sums(const float (&dm)[MAXN], float A, int nfull, float& s1, float& s0) {
    float acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc0[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < MAXN / 16; ++c) {
        if (c < nfull) {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2)
                pair_acc<CLAMP>(A, dm[c * 16 + jj], dm[c * 16 + jj + 1], acc1[(jj >> 1) & 3], acc0[(jj >> 1) & 3]);
        } else if (NTAIL > 0 && c == nfull) {
#pragma unroll
            for (int jj = 0; jj + 1 < NTAIL; jj += 2)
                pair_acc<CLAMP>(A, dm[c * 16 + jj], dm[c * 16 + jj + 1], acc1[(jj >> 1) & 3], acc0[(jj >> 1) & 3]);
            if (NTAIL & 1) single_acc(A, dm[c * 16 + NTAIL - 1], acc1[3], acc0[3]);   // odd number of background rows
        }
    }
    s1 = (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]);
    s0 = (acc0[0] + acc0[1]) + (acc0[2] + acc0[3]);
}

// Same sums with packed arithmetic (the common case: no clamping needed).  dm[2j], dm[2j+1] travel as one 64-bit operand.
template <int NTAIL>
__device__ __forceinline__ void row_sums_packed(const float (&dm)[MAXN], float A, int nfull, float& s1, float& s0) {
    const f32x2 A2 = f2_pack(A, A), one2 = f2_pack(1.f, 1.f), two2 = f2_pack(2.f, 2.f);
    f32x2 acc1[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)}, acc0[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
    float t1s = 0.f, t0s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXN / 16; ++c) {
        if (c < nfull) {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4)
                quad_acc(A2, f2_pack(dm[c * 16 + jj], dm[c * 16 + jj + 1]), f2_pack(dm[c * 16 + jj + 2], dm[c * 16 + jj + 3]),
                         one2, two2, acc1[(jj >> 2) & 1], acc0[(jj >> 2) & 1]);
        } else if (NTAIL > 0 && c == nfull) {
#pragma unroll
            for (int jj = 0; jj + 3 < NTAIL; jj += 4)
                quad_acc(A2, f2_pack(dm[c * 16 + jj], dm[c * 16 + jj + 1]), f2_pack(dm[c * 16 + jj + 2], dm[c * 16 + jj + 3]),
                         one2, two2, acc1[(jj >> 2) & 1], acc0[(jj >> 2) & 1]);
            constexpr int Q = NTAIL & ~3;                 // what the quads covered
            if ((NTAIL & 3) >= 2) pair_acc<false>(A, dm[c * 16 + Q], dm[c * 16 + Q + 1], t1s, t0s);
            if (NTAIL & 1) single_acc(A, dm[c * 16 + NTAIL - 1], t1s, t0s);
        }
    }
    float a, b, c2, d;
    f2_unpack(f2_add(acc1[0], acc1[1]), a, b);
    f2_unpack(f2_add(acc0[0], acc0[1]), c2, d);
    s1 = (a + b) + t1s;
    s0 = (c2 + d) + t0s;
}

// one warp = 32 coalition rows (one per lane) x a strided subset of the instances
// W = 64-bit words per coalition row (1: up to 64 groups, 2: up to 128; the tensor-memory version below also 16: up to 1024)
template <int NTAIL, int W>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA, 1) explain_shared_kernel(SharedParams p) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int n_rg = p.S_pad / 32;                       // row groups
    const int total_warps = gridDim.x * WARPS_PER_CTA;
    const int nparts = total_warps / n_rg;               // replicas of every row group
    if (nparts == 0 || gw >= nparts * n_rg) return;
    const int rg = gw % n_rg, part = gw / n_rg;
    const int s = rg * 32 + lane;
    const int cnt = *p.count;
    const int N = p.N, G = p.G;

    // this lane's row of Dm, in registers (chunks of 16 columns; unused chunks are never touched)
    float dm[MAXN];
    float dmax = 0.f;
#pragma unroll
    for (int c = 0; c < MAXN / 16; ++c) {
        if (c * 16 < N) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int j = c * 16 + jj;
                dm[j] = j < N ? p.DmT[(size_t)j * p.S_pad + s] : 0.f;
                dmax = fmaxf(dmax, dm[j]);
            }
        }
    }
    uint64_t zz[W];
#pragma unroll
    for (int w = 0; w < W; ++w) zz[w] = s < p.S ? p.z[(size_t)s * W + w] : 0ull;
    const int nfull = N / 16;
    const double es = p.dme[s];                          // exponent the row of Dm was normalised by

    const int ntab = (G + 3) / 4;
    for (int m = part; m < cnt; m += nparts) {
        const int i = p.list[m];
        // a = scale * sum_k z_k XW_i[k] in float64, one table entry per nibble of the row (prep_kernel built the tables);
        // the loads are independent, two partial sums keep the add chain short.
        // A = 2^a = 2^n * 2^f with n = rint(a), |f| <= 1/2 (f exact in fp32 to 3e-8, ex2.approx to ~1e-7 relative)
        const double* xt = p.XT + (size_t)i * ntab * 16;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
#pragma unroll 4
            for (int t = 0; t < 16 && 16 * w + t < ntab; t += 2) {
                a0 += __ldg(xt + (16 * w + t) * 16 + (int)((zz[w] >> (4 * t)) & 15ull));
                if (16 * w + t + 1 < ntab) a1 += __ldg(xt + (16 * w + t + 1) * 16 + (int)((zz[w] >> (4 * t + 4)) & 15ull));
            }
        }
        double a = (a0 + a1) + es;
        a = fmin(fmax(a, -120.0), 120.0);
        const double an = rint(a);
        const float A = ex2_approx((float)(a - an)) * __int_as_float((127 + (int)an) << 23);
        float s1, s0;
        // with u <= 1e18 the product q = ua*ub and the reciprocal of (1+ua)(1+ub) stay normal fp32 numbers: no clamps needed
        const bool risky = __any_sync(0xffffffffu, A * dmax > 1.0e18f);
        if (risky) row_sums<true, NTAIL>(dm, A, nfull, s1, s0);
        else row_sums_packed<NTAIL>(dm, A, nfull, s1, s0);
        if (s < p.S) {
            float2* dst = p.sums + (size_t)i * p.S_pad + s;
            if (p.accumulate) { const float2 o = *dst; s1 += o.x; s0 += o.y; }
            *dst = make_float2(s1, s0);
        }
    }
}

// ---- the same kernel with the Dm rows parked in TENSOR MEMORY ------------------------------------------------------
// The register version above keeps a lane's row of Dm (up to 128 floats) in registers: 168 registers per thread, 12
// warps per SM, and the kernel is latency bound.  TMEM (512 columns x 128 lanes x 32 bit per SM) is otherwise idle on this
// path, so it holds the rows instead: warp w owns TMEM lanes 32*(w%4) .. +31 (the hardware's lane quarter of that warp)
// and the column range (w/4)*cstride .. +N; it writes its 32 rows once (tcgen05.st) and re-reads them 16 columns at a time
// (tcgen05.ld, next chunk in flight while the current one is consumed).  That frees ~100 registers per thread: 16 or 20
// warps per SM instead of 12.
constexpr int TM_MAX_WARPS = 20;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float v0, float v1, float v2, float v3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v0)),
                 "r"(__float_as_uint(v1)), "r"(__float_as_uint(v2)), "r"(__float_as_uint(v3))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The same four sigmoids from what tensor memory holds for a quad of columns (0,2) (1,3): ds = (dm0 + dm2, dm1 + dm3) and
// dq = (dm0 dm2, dm1 dm3), both independent of the instance:  sm = A ds,  q = A^2 dq.  7 packed ops + 2 MUFU per four
// elements (A2 = (A, A), AA2 = (A^2, A^2), AA2x2 = 2 AA2).
__device__ __forceinline__ void quad_acc_sq(f32x2 A2, f32x2 AA2, f32x2 AA2x2, f32x2 ds, f32x2 dq, f32x2 one2, f32x2 two2,
                                            f32x2& a1, f32x2& a0) {
    const f32x2 sm = f2_mul(A2, ds);
    const f32x2 t1 = f2_add(sm, one2);
    const f32x2 den = f2_fma(AA2, dq, t1);
    const f32x2 w = f2_fma(AA2x2, dq, sm);
    float dlo, dhi;
    f2_unpack(den, dlo, dhi);
    const f32x2 r = f2_pack(rcp_approx(dlo), rcp_approx(dhi));
    a1 = f2_fma(r, f2_add(sm, two2), a1);
    a0 = f2_fma(r, w, a0);
}

// one 16-column chunk of a row: four packed quads (or the clamped scalar pairs), NV = valid columns of this chunk
// v holds, per full quad of columns, (ds.lo, ds.hi, dq.lo, dq.hi); columns past the last full quad of the tail chunk are
// raw Dm values.
template <int NV>
__device__ __forceinline__ void chunk_sums(const float (&v)[16], float A, f32x2 A2, f32x2 AA2, f32x2 AA2x2, f32x2 one2,
                                           f32x2 two2, f32x2 (&acc1)[2], f32x2 (&acc0)[2], float& t1s, float& t0s) {
#pragma unroll
    for (int jj = 0; jj + 3 < NV; jj += 4)
        quad_acc_sq(A2, AA2, AA2x2, f2_pack(v[jj], v[jj + 1]), f2_pack(v[jj + 2], v[jj + 3]), one2, two2, acc1[(jj >> 2) & 1],
                    acc0[(jj >> 2) & 1]);
    constexpr int Q = NV & ~3;
    if ((NV & 3) >= 2) pair_acc<false>(A, v[Q], v[Q + 1], t1s, t0s);
    if (NV & 1) single_acc(A, v[NV - 1], t1s, t0s);
}

template <int NTAIL, int W>
__global__ void __launch_bounds__(32 * TM_MAX_WARPS, 1) explain_shared_tmem_kernel(SharedParams p, int warps_used, int cstride) {
    __shared__ uint32_t s_tmem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tbase = s_tmem;

    const int n_rg = p.S_pad / 32;                       // row groups
    const int total_warps = gridDim.x * warps_used;
    const int nparts = total_warps / n_rg;               // replicas of every row group
    const int gw = blockIdx.x * warps_used + warp;
    const bool active = warp < warps_used && nparts > 0 && gw < nparts * n_rg;
    if (active) {
        const int rg = gw % n_rg, part = gw / n_rg;
        const int s = rg * 32 + lane;
        const int cnt = *p.count;
        const int N = p.N, G = p.G;
        const int nfull = N / 16;
        // this warp's slice of tensor memory: its lane quarter, column range warp / 4
        const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * cstride);
        const double es = p.dme[s];                      // exponent the row of Dm was normalised by (entries <= sqrt 2)
        for (int c = 0; c * 16 < N; ++c) {
            float v[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int j = c * 16 + jj;
                v[jj] = j < N ? p.DmT[(size_t)j * p.S_pad + s] : 0.f;
            }
            // every full quad of columns (0,2) (1,3) is stored as pair sums and pair products
            const int nv = c < nfull ? 16 : NTAIL;
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
                if (jj + 3 < nv) {
                    const float d0 = v[jj], d1 = v[jj + 1], d2 = v[jj + 2], d3 = v[jj + 3];
                    v[jj] = d0 + d2; v[jj + 1] = d1 + d3; v[jj + 2] = d0 * d2; v[jj + 3] = d1 * d3;
                }
            }
            if (c < nfull) {
                tmem_st16(taddr + c * 16, v);
            } else {
                // the tail is written four columns at a time: the slice stride is N rounded up to 4, and a wider store
                // would run into the next warp's slice
#pragma unroll
                for (int q = 0; q < (NTAIL + 3) / 4; ++q) tmem_st4(taddr + c * 16 + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
        tmem_st_wait();
        // rows of one or two words stay in registers; sixteen-word rows (more than 128 groups) are re-read per instance
        constexpr int WR = W <= 2 ? W : 1;
        uint64_t zz[WR];
#pragma unroll
        for (int w = 0; w < WR; ++w) zz[w] = s < p.S ? p.z[(size_t)s * W + w] : 0ull;
        const int ntab = (G + 3) / 4;
        const f32x2 one2 = f2_pack(1.f, 1.f), two2 = f2_pack(2.f, 2.f);

        // Up to 16 groups (four nibbles): the table entries of the NEXT instance are loaded one iteration ahead (the row's
        // nibbles, hence the offsets, do not depend on the instance), so neither the index load nor the table load sits
        // in front of A.  Wider problems load in place: their per-instance arithmetic is long enough to hide it.
        const bool ahead = W <= 2 && ntab <= 4;
        int off[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) off[t] = t * 16 + (int)((zz[0] >> (4 * t)) & 15ull);
        int i_cur = part < cnt ? p.list[part] : 0;
        int i_nx = part + nparts < cnt ? p.list[part + nparts] : 0;
        double nx[4] = {0.0, 0.0, 0.0, 0.0};
        if (ahead && part < cnt) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (t < ntab) nx[t] = __ldg(p.XT + (size_t)i_cur * ntab * 16 + off[t]);
        }
        for (int m = part; m < cnt; m += nparts) {
            int i;
            double a;
            bool cached = false;                             // only ever set for sixteen-word rows
            if (ahead) {
                i = i_cur;
                a = (nx[0] + nx[1]) + (nx[2] + nx[3]);
                i_cur = i_nx;
                if (m + nparts < cnt) {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (t < ntab) nx[t] = __ldg(p.XT + (size_t)i_cur * ntab * 16 + off[t]);
                }
                if (m + 2 * nparts < cnt) i_nx = p.list[m + 2 * nparts];
            } else if (W <= 2) {
                i = p.list[m];
                const double* xt = p.XT + (size_t)i * ntab * 16;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int w = 0; w < WR; ++w) {
#pragma unroll 4
                    for (int t = 0; t < 16 && 16 * w + t < ntab; t += 2) {
                        a0 += __ldg(xt + (16 * w + t) * 16 + (int)((zz[w] >> (4 * t)) & 15ull));
                        if (16 * w + t + 1 < ntab) a1 += __ldg(xt + (16 * w + t + 1) * 16 + (int)((zz[w] >> (4 * t + 4)) & 15ull));
                    }
                }
                a = a0 + a1;
            } else if (p.acache_mode == 2) {
                // sixteen-word rows, second and later background chunks: A(i, s) does not depend on the chunk (the row
                // exponent dme[s] covers the whole background) -- the first chunk's launch left it in acache
                i = p.list[m];
                a = 0.0;
                cached = true;
            } else {
                // sixteen-word rows: one word (sixteen nibble tables) at a time, the word re-read from the plan (L1/L2
                // resident: 128 B per row); four partial sums keep the float64 add chains short
                i = p.list[m];
                const double* xt = p.XT + (size_t)i * ntab * 16;
                const uint64_t* zrow = p.z + (size_t)s * W;
                const int nwords = (ntab + 15) >> 4;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int w = 0; w < nwords; ++w) {
                    const uint64_t zw = s < p.S ? __ldg(zrow + w) : 0ull;
                    const double* xw = xt + (size_t)w * 256;
                    const int nt = ntab - 16 * w < 16 ? ntab - 16 * w : 16;      // tables this word addresses
                    if (nt == 16) {
#pragma unroll
                        for (int t = 0; t < 16; t += 4) {
                            a0 += __ldg(xw + t * 16 + (int)((zw >> (4 * t)) & 15ull));
                            a1 += __ldg(xw + (t + 1) * 16 + (int)((zw >> (4 * t + 4)) & 15ull));
                            a2 += __ldg(xw + (t + 2) * 16 + (int)((zw >> (4 * t + 8)) & 15ull));
                            a3 += __ldg(xw + (t + 3) * 16 + (int)((zw >> (4 * t + 12)) & 15ull));
                        }
                    } else {
                        for (int t = 0; t < nt; ++t) a0 += __ldg(xw + t * 16 + (int)((zw >> (4 * t)) & 15ull));
                    }
                }
                a = (a0 + a1) + (a2 + a3);
            }
            float A;
            if (W > 2 && cached) {
                A = p.acache[(size_t)i * p.S_pad + s];
            } else {
                a += es;
                a = fmin(fmax(a, -120.0), 120.0);
                const double an = rint(a);
                A = ex2_approx((float)(a - an)) * __int_as_float((127 + (int)an) << 23);
                if (W > 2 && p.acache_mode == 1) p.acache[(size_t)i * p.S_pad + s] = A;
            }
            // A^2 must stay finite in fp32 (the normalised entries are <= sqrt 2, so A bounds every u): rows beyond that
            // take the clamped scalar path on the raw row from global memory (rare: saturated scores)
            const bool risky = __any_sync(0xffffffffu, A > 1.0e18f);
            if (risky) {
                float r1 = 0.f, r0 = 0.f;
                for (int j = 0; j + 1 < N; j += 2)
                    pair_acc<true>(A, p.DmT[(size_t)j * p.S_pad + s], p.DmT[(size_t)(j + 1) * p.S_pad + s], r1, r0);
                if (N & 1) single_acc(A, p.DmT[(size_t)(N - 1) * p.S_pad + s], r1, r0);
                if (s < p.S) {
                    float2* dst = p.sums + (size_t)i * p.S_pad + s;
                    if (p.accumulate) { const float2 o = *dst; r1 += o.x; r0 += o.y; }
                    *dst = make_float2(r1, r0);
                }
                continue;
            }
            const f32x2 A2 = f2_pack(A, A);
            const float AA = A * A;
            const f32x2 AA2 = f2_pack(AA, AA), AA2x2 = f2_pack(2.f * AA, 2.f * AA);
            f32x2 acc1[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)}, acc0[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
            float t1s = 0.f, t0s = 0.f;
            // chunks of 16 columns, the next one in flight while this one is consumed
            float va[16], vb[16];
            const int nch = nfull + (NTAIL > 0 ? 1 : 0);
            tc::tmem_ld16(taddr, va);
            for (int c = 0; c < nch; c += 2) {
                tc::tmem_ld_wait(va);
                if (c + 1 < nch) tc::tmem_ld16(taddr + (c + 1) * 16, vb);
                if (c < nfull) chunk_sums<16>(va, A, A2, AA2, AA2x2, one2, two2, acc1, acc0, t1s, t0s);
                else if (NTAIL > 0) chunk_sums<NTAIL>(va, A, A2, AA2, AA2x2, one2, two2, acc1, acc0, t1s, t0s);
                if (c + 1 < nch) {
                    tc::tmem_ld_wait(vb);
                    if (c + 2 < nch) tc::tmem_ld16(taddr + (c + 2) * 16, va);
                    if (c + 1 < nfull) chunk_sums<16>(vb, A, A2, AA2, AA2x2, one2, two2, acc1, acc0, t1s, t0s);
                    else if (NTAIL > 0) chunk_sums<NTAIL>(vb, A, A2, AA2, AA2x2, one2, two2, acc1, acc0, t1s, t0s);
                }
            }
            float q0, q1, q2, q3;
            f2_unpack(f2_add(acc1[0], acc1[1]), q0, q1);
            f2_unpack(f2_add(acc0[0], acc0[1]), q2, q3);
            float s1 = (q0 + q1) + t1s, s0 = (q2 + q3) + t0s;
            if (s < p.S) {
                float2* dst = p.sums + (size_t)i * p.S_pad + s;
                if (p.accumulate) { const float2 o = *dst; s1 += o.x; s0 += o.y; }
                *dst = make_float2(s1, s0);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// 0: registers, 1: tensor memory (default); DKS_SHARED_DM=regs selects the register version for comparisons
inline bool shared_dm_in_tmem() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("DKS_SHARED_DM");
        mode = (e && e[0] == 'r') ? 0 : 1;
    }
    return mode == 1;
}

inline void launch_explain_shared_chunk(const SharedParams& p, int words, int grid, cudaStream_t stream) {
    if (shared_dm_in_tmem() || words > 2) {                  // sixteen-word rows exist for the tensor-memory kernel only
        // column stride of a warp's slice: N rounded up to 4.  Reads are whole 16-column chunks (the last one may look into
        // the next slice, which is harmless), so the last slice must leave room for a full chunk: 5 slices up to N = 100,
        // 4 up to N = 128
        const int cstride = (p.N + 3) / 4 * 4;
        const int reach = (p.N + 15) / 16 * 16;              // columns a slice's reads can touch
        int slices = 5;
        while (slices > 1 && (slices - 1) * cstride + reach > 512) --slices;
        const int warps_used = 4 * slices;
        switch (p.N % 16) {
#define DKS_CASE(T)                                                                                          \
    case T:                                                                                                  \
        if (words == 1) explain_shared_tmem_kernel<T, 1><<<grid, 32 * TM_MAX_WARPS, 0, stream>>>(p, warps_used, cstride); \
        else if (words == 2) explain_shared_tmem_kernel<T, 2><<<grid, 32 * TM_MAX_WARPS, 0, stream>>>(p, warps_used, cstride); \
        else explain_shared_tmem_kernel<T, 16><<<grid, 32 * TM_MAX_WARPS, 0, stream>>>(p, warps_used, cstride);           \
        break;
            DKS_CASE(0) DKS_CASE(1) DKS_CASE(2) DKS_CASE(3) DKS_CASE(4) DKS_CASE(5) DKS_CASE(6) DKS_CASE(7)
            DKS_CASE(8) DKS_CASE(9) DKS_CASE(10) DKS_CASE(11) DKS_CASE(12) DKS_CASE(13) DKS_CASE(14) DKS_CASE(15)
#undef DKS_CASE
        }
        return;
    }

    const int threads = 32 * WARPS_PER_CTA;
    switch (p.N % 16) {
#define DKS_CASE(T)                                                             \
    case T:                                                                     \
        if (words == 1) explain_shared_kernel<T, 1><<<grid, threads, 0, stream>>>(p); \
        else explain_shared_kernel<T, 2><<<grid, threads, 0, stream>>>(p);      \
        break;
        DKS_CASE(0) DKS_CASE(1) DKS_CASE(2) DKS_CASE(3) DKS_CASE(4) DKS_CASE(5) DKS_CASE(6) DKS_CASE(7)
        DKS_CASE(8) DKS_CASE(9) DKS_CASE(10) DKS_CASE(11) DKS_CASE(12) DKS_CASE(13) DKS_CASE(14) DKS_CASE(15)
#undef DKS_CASE
    }
}

// Backgrounds larger than MAXN rows go through in chunks of MAXN columns of Dm (one launch each, sums accumulated)
inline int launch_explain_shared(SharedParams p, int words, int grid, cudaStream_t stream) {
    const int N = p.N;
    const float* dm = p.DmT;
    int launches = 0;
    const bool use_cache = words > 2 && p.acache != nullptr && N > MAXN;
    for (int j0 = 0; j0 < N; j0 += MAXN, ++launches) {
        p.N = N - j0 < MAXN ? N - j0 : MAXN;
        p.DmT = dm + (size_t)j0 * p.S_pad;
        p.accumulate = j0 > 0;
        p.acache_mode = use_cache ? (j0 == 0 ? 1 : 2) : 0;
        launch_explain_shared_chunk(p, words, grid, stream);
    }
    return launches;
}

struct WlsSharedParams {
    int n, N, G, C, S, S_pad, link, uniform_w;
    const float2* sums;      // [n][S_pad]
    const uint64_t* z;
    const double* w;
    const double* ainv;      // [(G-1) x (G-1)]
    const double* dlink;     // [n][C]
    const double* linkfnull;
    const double* fnull;
    const int* list;
    const int* count;
    double* phi;             // [C][n][G]
};

// ---- projection form of the solve for a shared plan ----------------------------------------------------------------
// beta = inv(A) E^T W (y - z_L delta) = P y - delta d,  P[k][s] = w_s sum_l inv(A)[k][l] (z_sl - z_sL),  d = P z_L.
// P depends only on the plan: computed once (float32 copy for the kernel, float64 for d).
__global__ void plan_pmat_kernel(const uint64_t* __restrict__ z, const double* __restrict__ w,
                                 const double* __restrict__ ainv, int S, int S_pad, int M, float* __restrict__ pmat) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nA = M - 1, L = M - 1;
    if (idx >= nA * S_pad) return;
    const int k = idx / S_pad, s = idx - k * S_pad;
    double acc = 0.0;
    if (s < S) {
        const uint64_t zz = z[s];
        const int zl = (int)((zz >> L) & 1ull);
        for (int l = 0; l < nA; ++l) {
            const int e = (int)((zz >> l) & 1ull) - zl;
            if (e) acc += ainv[k * nA + l] * (double)e;
        }
        acc *= w[s];
    }
    pmat[idx] = (float)acc;
}
__global__ void plan_dvec_kernel(const uint64_t* __restrict__ z, const float* __restrict__ pmat, int S, int S_pad, int M,
                                 double* __restrict__ dvec) {
    // d = P z_L with the SAME float32 P the kernel multiplies y by, so the delta term is removed consistently
    const int k = blockIdx.x, nA = M - 1, L = M - 1;
    if (k >= nA) return;
    double acc = 0.0;
    for (int s = threadIdx.x; s < S; s += 32)
        if ((z[s] >> L) & 1ull) acc += (double)pmat[(size_t)k * S_pad + s];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) dvec[k] = acc;
}

struct WlsPmatParams {
    int n, N, G, C, S, S_pad, link, uniform_w;
    const float2* sums;
    const float* pmat;       // [(G-1)][S_pad]
    const double* dvec;      // [(G-1)]
    const double* dlink;
    const double* linkfnull;
    const double* fnull;
    const int* list;
    const int* count;
    double* phi;
};
constexpr int PMAT_MAXK = 24;        // coefficients held in registers per thread
inline int wls_pmat_kpad(int G) { return (G - 1 + 3) / 4 * 4; }       // coefficient rows padded to a multiple of four
inline size_t wls_pmat_smem(int G, int S_pad, bool as_double) {
    return (size_t)wls_pmat_kpad(G) * S_pad * (as_double ? sizeof(double) : sizeof(float));
}

// Persistent CTAs; P resident in shared memory.  Per coalition row: y, then KPAD multiply-adds in float64.
// KPAD (compile time) = coefficients rounded up to a multiple of four, the padding rows of P are zero: the inner loop has
// no bounds checks (the checks were a fifth of the instructions).  PT = float by default; a float64 copy of the table
// (no float -> double conversions in the loop, but half the CTAs per SM) measured slower.
template <int KPAD, typename PT, int THREADS>
__global__ void __launch_bounds__(THREADS) wls_pmat_kernel(WlsPmatParams p) {
    extern __shared__ __align__(16) unsigned char s_praw[];
    PT* s_P = reinterpret_cast<PT*>(s_praw);                        // [KPAD][S_pad]
    __shared__ double s_part[THREADS / 32][KPAD];
    __shared__ LogTabEntry s_logtab[DKS_LOGTAB_SIZE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int G = p.G, nA = G - 1;
    const int cnt = *p.count;
    if ((int)blockIdx.x >= cnt) return;
    if (threadIdx.x < DKS_LOGTAB_SIZE) logtab_fill(s_logtab, threadIdx.x);
    for (int idx = threadIdx.x; idx < KPAD * p.S_pad; idx += THREADS) s_P[idx] = idx < nA * p.S_pad ? (PT)p.pmat[idx] : (PT)0;
    __syncthreads();
    const double lf1 = p.linkfnull[1], f1 = p.fnull[1], inv_n = 1.0 / (double)p.N;
    const size_t slab = (size_t)p.n * G;
    int i_next = p.list[blockIdx.x];
    double delta_next = p.dlink[(size_t)i_next * p.C + 1];
    constexpr int INFLIGHT = THREADS >= 512 ? 4 : 8;                  // independent loads in flight per thread
    for (int m = blockIdx.x; m < cnt; m += gridDim.x) {
        const int i = i_next;
        const double delta = delta_next;
        if (m + (int)gridDim.x < cnt) {            // next instance's index and delta: off the critical path
            i_next = p.list[m + gridDim.x];
            delta_next = p.dlink[(size_t)i_next * p.C + 1];
        }
        const float2* sums = p.sums + (size_t)i * p.S_pad;
        double Tk[KPAD];
#pragma unroll
        for (int k = 0; k < KPAD; ++k) Tk[k] = 0.0;
        for (int s0 = 0; s0 < p.S; s0 += INFLIGHT * THREADS) {
            float2 a[INFLIGHT];
#pragma unroll
            for (int r = 0; r < INFLIGHT; ++r) {
                const int s = s0 + r * THREADS + threadIdx.x;
                a[r] = s < p.S ? sums[s] : make_float2(1.f, 1.f);
            }
#pragma unroll
            for (int r = 0; r < INFLIGHT; ++r) {
                const int s = s0 + r * THREADS + threadIdx.x;
                if (s < p.S) {
                    double y;
                    if (p.link == DKS_LINK_LOGIT) y = fast_log_ratio(a[r].x, a[r].y, s_logtab) - lf1;
                    else y = (p.uniform_w ? (double)a[r].x * inv_n : (double)a[r].x) - f1;
#pragma unroll
                    for (int k = 0; k < KPAD; ++k) Tk[k] = fma((double)s_P[(size_t)k * p.S_pad + s], y, Tk[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KPAD; ++k) {
            const double r = warp_sum(Tk[k]);
            if (lane == 0) s_part[wib][k] = r;
        }
        __syncthreads();
        if (wib == 0) {
            double beta = 0.0;
            if (lane < nA) {
#pragma unroll
                for (int wq = 0; wq < THREADS / 32; ++wq) beta += s_part[wq][lane];      // fixed order: reproducible
                beta -= delta * p.dvec[lane];
            }
            const double sum = warp_sum(beta);
            if (lane < G) {
                double val = lane < nA ? beta : delta - sum;       // the eliminated (last) group takes the remainder
                if (fabs(val) < 1e-10) val = 0.0;
                p.phi[slab + (size_t)i * G + lane] = val;
                p.phi[(size_t)i * G + lane] = (val == 0.0) ? 0.0 : -val;
            }
        }
        __syncthreads();
    }
}

// picks the instantiation; returns false when no variant fits shared memory
inline bool launch_wls_pmat(const WlsPmatParams& p, int n, int sm_count, int max_smem, cudaStream_t stream, cudaError_t* err) {
    const int kpad = wls_pmat_kpad(p.G);
    const size_t sm_d = wls_pmat_smem(p.G, p.S_pad, true), sm_f = wls_pmat_smem(p.G, p.S_pad, false);
    // measured on B200 (Adult shape): float32 table, 2 CTAs of 256 threads per SM: 57 us; float64 table, 1 CTA of 512
    // threads: 90 us.  DKS_PMAT=double selects the float64 variant for comparisons.
    static int want_double = -1;
    if (want_double < 0) { const char* e = getenv("DKS_PMAT"); want_double = (e && e[0] == 'd') ? 1 : 0; }
    const bool as_double = want_double && sm_d + 8192 <= (size_t)max_smem;
    if (!as_double && sm_f + 8192 > (size_t)max_smem) return false;
    const size_t smem = as_double ? sm_d : sm_f;
    int per_sm = (int)((size_t)max_smem / (smem + 8192));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const int grid = n < sm_count * per_sm ? n : sm_count * per_sm;
    *err = cudaSuccess;
#define DKS_PM(K)                                                                                                         \
    case K:                                                                                                               \
        if (as_double) {                                                                                                  \
            *err = cudaFuncSetAttribute(wls_pmat_kernel<K, double, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (*err == cudaSuccess) wls_pmat_kernel<K, double, 512><<<grid, 512, smem, stream>>>(p);                     \
        } else {                                                                                                          \
            *err = cudaFuncSetAttribute(wls_pmat_kernel<K, float, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (*err == cudaSuccess) wls_pmat_kernel<K, float, 256><<<grid, 256, smem, stream>>>(p);                      \
        }                                                                                                                 \
        break;
    switch (kpad) {
        DKS_PM(4) DKS_PM(8) DKS_PM(12) DKS_PM(16) DKS_PM(20) DKS_PM(24)
        default: return false;
    }
#undef DKS_PM
    return true;
}

// Persistent CTAs of 8 warps, each looping over instances: y = link(ey) - link(fnull) per coalition, E^T W y in 2^-40
// fixed point (integer adds: exact, order-independent), beta = inv(E^T W E) (E^T W y), phi.
constexpr int WLS_THREADS = 256;
inline size_t wls_shared_smem(int G) { return sizeof(double) * (size_t)(G - 1) * (G - 1); }
template <int W>
__global__ void __launch_bounds__(WLS_THREADS) wls_shared_kernel(WlsSharedParams p) {
    extern __shared__ double s_ainv[];                   // [(G-1)][(G-1)]
    __shared__ long long s_part[WLS_THREADS / 32][64 * W];
    __shared__ double s_rhs[64 * W];
    __shared__ LogTabEntry s_logtab[DKS_LOGTAB_SIZE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int G = p.G, nA = G - 1, L = G - 1;
    const int cnt = *p.count;
    if ((int)blockIdx.x >= cnt) return;
    if (threadIdx.x < DKS_LOGTAB_SIZE) logtab_fill(s_logtab, threadIdx.x);
    for (int idx = threadIdx.x; idx < nA * nA; idx += blockDim.x) s_ainv[idx] = p.ainv[idx];
    __syncthreads();
    const double lf1 = p.linkfnull[1], f1 = p.fnull[1], inv_n = 1.0 / (double)p.N;
    const size_t slab = (size_t)p.n * G;
    for (int m = blockIdx.x; m < cnt; m += gridDim.x) {
        const int i = p.list[m];
        const double delta = p.dlink[(size_t)i * p.C + 1];
        const float2* sums = p.sums + (size_t)i * p.S_pad;
        // thread handles coalitions tid, tid+256, ...; sixteen coefficients of E^T W y per pass over the rows
        for (int k0 = 0; k0 < nA; k0 += 16) {
            long long Tk[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) Tk[k] = 0;
#pragma unroll 2
            for (int s = threadIdx.x; s < p.S; s += WLS_THREADS) {
                const float2 a = sums[s];
                double y;
                if (p.link == DKS_LINK_LOGIT) y = fast_log_ratio(a.x, a.y, s_logtab) - lf1;
                else y = (p.uniform_w ? (double)a.x * inv_n : (double)a.x) - f1;
                const uint64_t* zrow = p.z + (size_t)s * W;
                const bool zl = (zrow[L >> 6] >> (L & 63)) & 1ull;
                const double v = p.w[s] * (y - (zl ? delta : 0.0));
                const uint64_t zw = zrow[k0 >> 6];               // a 16-bit window never straddles two words
                const uint32_t zb = (uint32_t)((zl ? ~zw : zw) >> (k0 & 63));
                const long long vi = zl ? -to_fix(v) : to_fix(v);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k0 + k < nA && ((zb >> k) & 1u)) Tk[k] += vi;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (k0 + k < nA) {
                    const long long r = warp_sum_ll(Tk[k]);
                    if (lane == 0) s_part[wib][k0 + k] = r;
                }
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < nA) {
            long long acc = 0;
#pragma unroll
            for (int wq = 0; wq < WLS_THREADS / 32; ++wq) acc += s_part[wq][threadIdx.x];
            s_rhs[threadIdx.x] = from_fix(acc);
        }
        __syncthreads();
        if (wib == 0) {
            double sum = 0.0;
            double beta[2 * W];
#pragma unroll
            for (int h = 0; h < 2 * W; ++h) beta[h] = 0.0;
#pragma unroll
            for (int h = 0; h < 2 * W; ++h) {
                const int k = lane + 32 * h;
                if (k < nA) {
                    double b0 = 0.0, b1 = 0.0;      // two chains: the dot product is latency bound otherwise
                    int l = 0;
                    for (; l + 1 < nA; l += 2) {
                        b0 = fma(s_ainv[k * nA + l], s_rhs[l], b0);
                        b1 = fma(s_ainv[k * nA + l + 1], s_rhs[l + 1], b1);
                    }
                    if (l < nA) b0 = fma(s_ainv[k * nA + l], s_rhs[l], b0);
                    beta[h] = b0 + b1;
                    sum += beta[h];
                }
            }
            sum = warp_sum(sum);
#pragma unroll
            for (int h = 0; h < 2 * W; ++h) {
                const int k = lane + 32 * h;
                if (k < G) {
                    double val = k < nA ? beta[h] : delta - sum;       // the eliminated (last) group takes the remainder
                    if (fabs(val) < 1e-10) val = 0.0;
                    p.phi[slab + (size_t)i * G + k] = val;
                    p.phi[(size_t)i * G + k] = (val == 0.0) ? 0.0 : -val;
                }
            }
        }
        // s_part / s_rhs are rewritten only after the next instance's row loop, which ends with __syncthreads
    }
}

}  // namespace shared_path
}  // namespace dks
