// Fused coalition kernel on tcgen05 / TMEM (sm_100a), binary-logistic head.
//
// Per instance i and 128-coalition tile the masked-batch scores are one small dense contraction
//     T[s][j] = sum_k Z[s][k] * Delta_i[j][k],      Delta_i[j][k] = scale*(XW_i[v_k] - BW[j][v_k]),  k < M
//                                                   Delta_i[j][M] = scale*score_j   (Z[s][M] = 1)
// (scale = -kappa*log2 e, so exp(-kappa*score) = 2^T).  Z is 0/1 and exact in bf16; Delta is split into three bf16
// terms (hi/mid/lo, ~fp32-exact) and accumulated in fp32 in TMEM by three tcgen05.mma (M=128, N=Npad, K=16).
// Warp roles of the persistent CTA (one per SM):
//   warp 0      producer: builds the instance's B operand (Delta splits) and each tile's A operand (Z bits expanded
//               to bf16 in registers, never read from HBM as a matrix) in shared memory, issues the MMAs;
//   warps 4-11  epilogue, two groups of four warps = two TMEM accumulator buffers: tcgen05.ld the 128 x N scores,
//               p1 = 1/(1+2^T), background-weighted sums, link -> y[s] (float64) in shared memory;
//   warps 1-3   constrained WLS of the previous instance (float64) while the next one is being evaluated.
// All hand-offs are mbarriers (tcgen05.commit for MMA completion); no __syncthreads in the steady state.
#pragma once

#include <cuda_bf16.h>

#include "dks_kernels.cuh"

namespace dks {
namespace tc {

constexpr int TILE_S = 128;      // coalitions per MMA tile (UMMA M)
constexpr int KP = 16;           // K per split: up to 15 varying groups + the constant column
constexpr int NSPLIT = 3;        // bf16 hi/mid/lo
constexpr int MAX_NPAD = 128;    // background rows per accumulator buffer (TMEM columns)
constexpr int N_EPI_WARPS = 8, N_WLS_WARPS = 3;
constexpr int NTHREADS = 32 * (1 + N_WLS_WARPS + N_EPI_WARPS);
constexpr int TMEM_COLS = 256;   // two accumulator buffers of 128 fp32 columns
constexpr uint32_t SPIN_LIMIT = 1u << 26;

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must not hang the GPU -- flag the status word and trap instead
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* status) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) {
            status[0] = DKS_ERR_CUDA;
            status[1] = -77;
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, single CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (canonical layout ((8,m),(T,2)):((1T,SBO),(1,LBO)):
// 8x16-byte core matrices; LBO = bytes between the two K-adjacent core matrices of one K=16 step, SBO = bytes
// between core matrices adjacent along M/N).  cute::UMMA::SmemDescriptor bit layout, version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: fp32 accumulate, bf16 A/B, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_S >> 4) << 24);
}

// ---- shared memory carve-up ---------------------------------------------------------------------------------
struct Smem {
    uint64_t* bars;      // [8]: tmem_full[2], tmem_empty[2], ys_full[2], ys_empty[2]
    uint32_t* tmem_ptr;  // [1]
    int* vi;             // [16] varying position -> group (producer warp only)
    double* chol;        // [15*15]
    double* rhs;         // [16]
    double* ys;          // [2][S_cap]
    float* wb;           // [MAX_NPAD] background weights
    unsigned char* A;    // [2][128*KP*2]
    unsigned char* B;    // [2][NSPLIT][Npad*KP*2]
};
__host__ __device__ inline size_t smem_bytes(int S_cap, int Npad) {
    return 128 /*bars + tmem ptr*/ + 2 * 16 * sizeof(int) + (15 * 15 + 16) * sizeof(double) + 2 * (size_t)S_cap * sizeof(double) +
           MAX_NPAD * sizeof(float) + 2 * (size_t)TILE_S * KP * 2 + 2 * NSPLIT * (size_t)Npad * KP * 2 + 64;
}
__device__ inline Smem carve(unsigned char* base, int S_cap, int Npad) {
    Smem s;
    s.bars = reinterpret_cast<uint64_t*>(base);
    s.tmem_ptr = reinterpret_cast<uint32_t*>(base + 64);
    s.vi = reinterpret_cast<int*>(base + 128);
    s.chol = reinterpret_cast<double*>(base + 128 + 2 * 16 * sizeof(int));
    s.rhs = s.chol + 15 * 15;
    s.ys = s.rhs + 16;
    s.wb = reinterpret_cast<float*>(s.ys + 2 * (size_t)S_cap);
    unsigned char* p = reinterpret_cast<unsigned char*>(s.wb + MAX_NPAD);
    p = reinterpret_cast<unsigned char*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
    s.A = p;
    s.B = p + 2 * (size_t)TILE_S * KP * 2;
    return s;
}

struct TcParams {
    ExplainParams p;
    const double* BW;      // [N][G] grouped background contributions, float64 (R == 1)
    const double* scores;  // [N]
    int Npad;
    int uniform_w;
    float* dbg_T;          // optional [S_cap][Npad] dump of the scores of instance dbg_i
    int dbg_i;
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// instances this CTA handles: i = blockIdx.x + q*gridDim.x; every role walks the same list
__device__ __forceinline__ int tiles_of(const ExplainParams& p, int i, int& M, int& S) {
    M = p.Mcnt[i];
    if (M < 2) { S = 0; return 0; }
    S = dks_effective_S(M, p.S_req);
    if (p.ext_z == nullptr) {
        PlanDev pd = p.plans[M];
        if (pd.z == nullptr || pd.S != S) { S = 0; return 0; }   // reported by the WLS warps
    }
    if (S > p.S_cap) { S = 0; return 0; }
    return (S + TILE_S - 1) / TILE_S;
}

__global__ void __launch_bounds__(NTHREADS, 1) explain_tcgen05_kernel(TcParams tp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ExplainParams& p = tp.p;
    const int Npad = tp.Npad, N = p.N, G = p.G;
    Smem sm = carve(smem_raw, p.S_cap, Npad);
    uint64_t* tmem_full = sm.bars;
    uint64_t* tmem_empty = sm.bars + 2;
    uint64_t* ys_full = sm.bars + 4;
    uint64_t* ys_empty = sm.bars + 6;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t slab = (size_t)p.n * G;

    if (threadIdx.x == 0) {
        mbar_init(&tmem_full[0], 1);  mbar_init(&tmem_full[1], 1);
        mbar_init(&tmem_empty[0], 128); mbar_init(&tmem_empty[1], 128);
        mbar_init(&ys_full[0], 32 * N_EPI_WARPS); mbar_init(&ys_full[1], 32 * N_EPI_WARPS);
        mbar_init(&ys_empty[0], 32 * N_WLS_WARPS); mbar_init(&ys_empty[1], 32 * N_WLS_WARPS);
        fence_barrier_init();
    }
    for (int j = threadIdx.x; j < MAX_NPAD; j += blockDim.x) sm.wb[j] = j < N ? p.wbf[j] : 0.f;
    if (warp == 0) tmem_alloc(sm.tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr;

    const uint32_t a_bytes = TILE_S * KP * 2, b_split_bytes = (uint32_t)Npad * KP * 2;

    if (warp == 0) {
        // =================================== producer / MMA issuer ===================================
        const uint32_t idesc = make_idesc(Npad);
        uint32_t g = 0;  // global tile counter of this CTA
        int q = 0;       // instance ordinal of this CTA
        for (int i = blockIdx.x; i < p.n; i += gridDim.x, ++q) {
            int M, S;
            const int T = tiles_of(p, i, M, S);
            if (T == 0) continue;
            const uint64_t vm = p.vmask[i];
            const uint64_t* zp = p.ext_z ? p.ext_z + (size_t)i * p.ext_stride : p.plans[M].z;
            // before touching B[q&1] / A buffers the MMAs that read them two tiles ago must be done
            unsigned char* Bq = sm.B + (size_t)(q & 1) * NSPLIT * b_split_bytes;
            if (g >= 1) {  // every MMA issued so far has completed => the A/B buffers it read are free
                uint32_t u = (g - 1) >> 1;
                mbar_wait(&tmem_full[(g - 1) & 1], u & 1, p.status);
            }
            // ---- B operand of this instance: Delta splits, K-major core matrices [kc][j][8] ----
            {
                {   // lane k finds the k-th varying group
                    int cnt = 0, gsel = 0;
                    for (int gI = 0; gI < G; ++gI)
                        if ((vm >> gI) & 1ull) { if (cnt == lane) gsel = gI; ++cnt; }
                    __syncwarp();
                    if (lane < KP) sm.vi[lane] = lane < M ? gsel : 0;
                    __syncwarp();
                }
                int vi[KP];
#pragma unroll
                for (int kk = 0; kk < KP; ++kk) vi[kk] = sm.vi[kk];
                double xw[KP];
#pragma unroll
                for (int kk = 0; kk < KP; ++kk) xw[kk] = kk < M ? p.XW[(size_t)i * G + vi[kk]] : 0.0;
                for (int j = lane; j < Npad; j += 32) {
                    float hi[KP], mid[KP], lo[KP];
#pragma unroll
                    for (int kk = 0; kk < KP; ++kk) {
                        double v = 0.0;
                        if (j < N) {
                            if (kk < M) v = p.scale * (xw[kk] - tp.BW[(size_t)j * G + vi[kk]]);
                            else if (kk == M) v = p.scale * tp.scores[j];
                        }
                        float vf = (float)v;
                        float h = __bfloat162float(__float2bfloat16_rn(vf));
                        float r1 = vf - h;
                        float m = __bfloat162float(__float2bfloat16_rn(r1));
                        float r2 = (r1 - m) + (float)(v - (double)vf);
                        hi[kk] = h; mid[kk] = m; lo[kk] = r2;
                    }
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc) {
                        uint4 wh, wm, wl;
                        wh.x = pack_bf16(hi[kc * 8 + 0], hi[kc * 8 + 1]); wh.y = pack_bf16(hi[kc * 8 + 2], hi[kc * 8 + 3]);
                        wh.z = pack_bf16(hi[kc * 8 + 4], hi[kc * 8 + 5]); wh.w = pack_bf16(hi[kc * 8 + 6], hi[kc * 8 + 7]);
                        wm.x = pack_bf16(mid[kc * 8 + 0], mid[kc * 8 + 1]); wm.y = pack_bf16(mid[kc * 8 + 2], mid[kc * 8 + 3]);
                        wm.z = pack_bf16(mid[kc * 8 + 4], mid[kc * 8 + 5]); wm.w = pack_bf16(mid[kc * 8 + 6], mid[kc * 8 + 7]);
                        wl.x = pack_bf16(lo[kc * 8 + 0], lo[kc * 8 + 1]); wl.y = pack_bf16(lo[kc * 8 + 2], lo[kc * 8 + 3]);
                        wl.z = pack_bf16(lo[kc * 8 + 4], lo[kc * 8 + 5]); wl.w = pack_bf16(lo[kc * 8 + 6], lo[kc * 8 + 7]);
                        const size_t off = (size_t)kc * Npad * 16 + (size_t)j * 16;
                        *reinterpret_cast<uint4*>(Bq + 0 * b_split_bytes + off) = wh;
                        *reinterpret_cast<uint4*>(Bq + 1 * b_split_bytes + off) = wm;
                        *reinterpret_cast<uint4*>(Bq + 2 * b_split_bytes + off) = wl;
                    }
                }
            }
            // ---- tiles ----
            for (int t = 0; t < T; ++t, ++g) {
                const uint32_t buf = g & 1, u = g >> 1;
                if (t > 0 && g >= 2) mbar_wait(&tmem_full[buf], (u - 1) & 1, p.status);  // A[buf] free (MMA g-2 done)
                unsigned char* Ab = sm.A + (size_t)buf * a_bytes;
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    const int sl = lane + 32 * r4;
                    const int s = t * TILE_S + sl;
                    uint32_t zz = 0;
                    if (s < S) zz = (uint32_t)(zp[s] & 0xFFFFull) | (1u << M);   // constant column carries score_j
                    uint32_t w[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        w[c] = (((zz >> (2 * c)) & 1u) ? 0x00003F80u : 0u) | (((zz >> (2 * c + 1)) & 1u) ? 0x3F800000u : 0u);
                    *reinterpret_cast<uint4*>(Ab + 0 * (TILE_S * 16) + sl * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(Ab + 1 * (TILE_S * 16) + sl * 16) = make_uint4(w[4], w[5], w[6], w[7]);
                }
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    mbar_wait(&tmem_empty[buf], (u & 1) ^ 1, p.status);   // epilogue drained this accumulator
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(Ab), TILE_S * 16, 128);
                    const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll
                    for (int sp = 0; sp < NSPLIT; ++sp) {
                        const uint64_t bdesc = make_smem_desc(smem_u32(Bq + sp * b_split_bytes), (uint32_t)Npad * 16, 128);
                        umma_bf16(d_tmem, adesc, bdesc, idesc, sp > 0 ? 1u : 0u);
                    }
                    umma_commit(&tmem_full[buf]);   // arrives when the three MMAs have completed
                }
                __syncwarp();
            }
        }
    } else if (warp >= 1 + N_WLS_WARPS) {
        // =================================== epilogue ===================================
        const int ew = warp - (1 + N_WLS_WARPS);      // 0..7
        const int grp = ew >> 2;                      // accumulator buffer this group drains
        const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31 (hardware rule: warp id % 4)
        const int row_in_tile = quarter * 32 + lane;
        const double lf1 = p.linkfnull[1], f1 = p.fnull[1];
        const float inv_n = 1.0f / (float)N;
        uint32_t g = 0;
        int q = 0;
        for (int i = blockIdx.x; i < p.n; i += gridDim.x, ++q) {
            int M, S;
            const int T = tiles_of(p, i, M, S);
            double* ys = sm.ys + (size_t)(q & 1) * p.S_cap;
            // y buffer of ordinal q-2 must have been consumed by the WLS warps
            mbar_wait(&ys_empty[q & 1], ((q >> 1) & 1) ^ 1, p.status);
            for (int t = 0; t < T; ++t, ++g) {
                if ((int)(g & 1) != grp) continue;
                const uint32_t buf = g & 1, u = g >> 1;
                mbar_wait(&tmem_full[buf], u & 1, p.status);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 128;
                float acc1 = 0.f, acc0 = 0.f;
                for (int c0 = 0; c0 < Npad; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (tp.dbg_T != nullptr && i == tp.dbg_i) {
                        const int s = t * TILE_S + row_in_tile;
                        if (s < p.S_cap)
                            for (int jj = 0; jj < 16; ++jj) tp.dbg_T[(size_t)s * Npad + c0 + jj] = v[jj];
                    }
                    if (c0 + 16 <= N) {
                        if (tp.uniform_w) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                float tt = fminf(v[jj], 120.f);
                                float uu = ex2_approx(tt);
                                float rr = rcp_approx(1.f + uu);
                                acc1 += rr;
                                acc0 = fmaf(uu, rr, acc0);
                            }
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                float tt = fminf(v[jj], 120.f);
                                float uu = ex2_approx(tt);
                                float rr = rcp_approx(1.f + uu);
                                float wj = sm.wb[c0 + jj];
                                acc1 = fmaf(wj, rr, acc1);
                                acc0 = fmaf(wj, uu * rr, acc0);
                            }
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            if (c0 + jj < N) {
                                float tt = fminf(v[jj], 120.f);
                                float uu = ex2_approx(tt);
                                float rr = rcp_approx(1.f + uu);
                                float wj = tp.uniform_w ? 1.f : sm.wb[c0 + jj];
                                acc1 = fmaf(wj, rr, acc1);
                                acc0 = fmaf(wj, uu * rr, acc0);
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty[buf]);       // accumulator buffer may be overwritten
                const int s = t * TILE_S + row_in_tile;
                if (s < S) {
                    double y;
                    if (p.link == DKS_LINK_LOGIT) y = log((double)acc1 / (double)acc0) - lf1;
                    else y = (double)(tp.uniform_w ? acc1 * inv_n : acc1) - f1;
                    ys[s] = y;
                }
            }
            mbar_arrive(&ys_full[q & 1]);            // release-arrive: this thread's y values are published
        }
    } else {
        // =================================== WLS warps (float64) ===================================
        const int ww = warp - 1;                      // 0..2
        const int wtid = threadIdx.x - 32;            // 0..95
        int cachedM = -1;
        int q = 0;
        for (int i = blockIdx.x; i < p.n; i += gridDim.x, ++q) {
            int M, S;
            const int T = tiles_of(p, i, M, S);
            const int C = p.C;
            for (int idx = wtid; idx < C * G; idx += 32 * N_WLS_WARPS)
                p.phi[(size_t)(idx / G) * slab + (size_t)i * G + idx % G] = 0.0;
            mbar_wait(&ys_full[q & 1], (q >> 1) & 1, p.status);
            const double* ys = sm.ys + (size_t)(q & 1) * p.S_cap;
            if (T == 0) {
                if (M == 1) {
                    if (wtid < C) {
                        int gI = __ffsll((long long)p.vmask[i]) - 1;
                        p.phi[(size_t)wtid * slab + (size_t)i * G + gI] = p.dlink[(size_t)i * C + wtid];
                    }
                } else if (M >= 2 && wtid == 0) {
                    int S0 = dks_effective_S(M, p.S_req);
                    bool missing = p.ext_z == nullptr && (p.plans[M].z == nullptr || p.plans[M].S != S0);
                    if (missing) { atomicCAS(&p.status[0], 0, DKS_ERR_PLAN_MISSING); p.status[1] = M; }
                    else { atomicCAS(&p.status[0], 0, DKS_ERR_INVALID); p.status[1] = i; }
                }
                mbar_arrive(&ys_empty[q & 1]);
                continue;
            }
            const uint64_t* zp;
            const double* wp;
            const double* chol = nullptr;
            if (p.ext_z) { zp = p.ext_z + (size_t)i * p.ext_stride; wp = p.ext_w + (size_t)i * p.ext_stride; }
            else { zp = p.plans[M].z; wp = p.plans[M].w; chol = p.plans[M].chol; }
            const int nA = M - 1;
            if (chol != nullptr) {
                if (M != cachedM) {
                    named_bar_sync(1, 32 * N_WLS_WARPS);
                    for (int idx = wtid; idx < nA * nA; idx += 32 * N_WLS_WARPS) sm.chol[idx] = chol[idx];
                    cachedM = M;
                }
            } else {
                cachedM = -1;
                named_bar_sync(1, 32 * N_WLS_WARPS);
                wls_build_normal(zp, wp, S, M, sm.chol, ww, N_WLS_WARPS);
                named_bar_sync(1, 32 * N_WLS_WARPS);
                if (ww == 0) {
                    bool ok = wls_cholesky_warp(sm.chol, nA);
                    if (!ok && lane == 0) { atomicCAS(&p.status[0], 0, DKS_ERR_NUMERIC); p.status[1] = i; }
                }
            }
            const double delta = p.dlink[(size_t)i * C + 1];
            wls_build_rhs(zp, wp, ys, S, M, delta, sm.rhs, ww, N_WLS_WARPS);
            named_bar_sync(1, 32 * N_WLS_WARPS);
            if (wtid == 0) {
                int vi[KP];
                {
                    const uint64_t vm = p.vmask[i];
                    int k = 0;
                    for (int gI = 0; gI < G; ++gI) if (((vm >> gI) & 1ull) && k < KP) vi[k++] = gI;
                }
                wls_solve_write(sm.chol, sm.rhs, M, delta, vi, p.phi + slab + (size_t)i * G, 1.0);
                double* phi0 = p.phi + (size_t)i * G;
                const double* phi1 = p.phi + slab + (size_t)i * G;
                for (int k = 0; k < M; ++k) { double v = phi1[vi[k]]; phi0[vi[k]] = (v == 0.0) ? 0.0 : -v; }
            }
            named_bar_sync(1, 32 * N_WLS_WARPS);
            mbar_arrive(&ys_empty[q & 1]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace tc

// ---- host glue --------------------------------------------------------------------------------------------------
inline int tc_npad(int N) { return (N + 15) / 16 * 16; }

inline bool tc_supported(const dks_ctx* ctx, const ExplainParams& p) {
    if (ctx->act != DKS_ACT_BINARY_LOGISTIC || ctx->R != 1) return false;
    if (ctx->G > tc::KP - 1) return false;            // M + constant column must fit one K = 16 step
    if (ctx->N > tc::MAX_NPAD) return false;          // one accumulator buffer holds the whole background
    if ((long long)tc::smem_bytes(p.S_cap, tc_npad(ctx->N)) > (long long)ctx->max_smem_optin) return false;
    return true;
}

inline int tc_launch(dks_ctx* ctx, const ExplainParams& p) {
    tc::TcParams tp;
    tp.p = p;
    tp.BW = ctx->d_BW;
    tp.scores = ctx->d_scores;
    tp.Npad = tc_npad(ctx->N);
    tp.uniform_w = ctx->uniform_w ? 1 : 0;
    tp.dbg_T = ctx->dbg_T;
    tp.dbg_i = ctx->dbg_i;
    size_t smem = tc::smem_bytes(p.S_cap, tp.Npad);
    cudaError_t e = cudaFuncSetAttribute(tc::explain_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return DKS_ERR_CUDA;
    int grid = ctx->sm_count < p.n ? ctx->sm_count : p.n;
    tc::explain_tcgen05_kernel<<<grid, tc::NTHREADS, smem, ctx->stream>>>(tp);
    ctx->launches += 1;
    return DKS_OK;
}

inline int tc_fit(dks_ctx*) { return DKS_OK; }
inline int tc_plan_changed(dks_ctx*, int) { return DKS_OK; }
inline void tc_release(dks_ctx*) {}

}  // namespace dks
