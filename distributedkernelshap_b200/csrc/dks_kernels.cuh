// Device code of the B200 KernelSHAP engine: fit kernels (K0), per-instance preparation, and the fused
// coalition kernel (mask/impute + predict + background reduction + link + constrained WLS).
//
// Algebra used throughout: the model head sees linear scores, so a masked row's score is
//     score(s, j) = base_j + sum_{k in varying} z_sk * (XW_i[k] - BW[j][k])
// with XW_i[k] = sum_{col in group k} x_i[col] W[col]  and  BW[j][k] likewise for background row j.  The
// masked batch (S*N x D, KernelExplainer.allocate/addsample) is therefore never materialised.
#pragma once

#include "dks_common.cuh"
#include "dks_linkmath.cuh"

namespace dks {

// ------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- cheap link arithmetic (per-row log around a 64-entry table): dks_linkmath.cuh

// fixed-point accumulation of E^T W y: v -> round(v * 2^40) as int64 (|v| = w |y| < 8e6 fits), integer adds are exact and
// order-independent (bit-reproducible whatever the reduction order); resolution 2^-40 ~ 9e-13 per row.  (64-bit integer
// adds run at ~30 lanes/clk/SM, half the DADD rate: used where order-independence matters, not for speed.)
#define DKS_FIX_SCALE 1099511627776.0          /* 2^40 */
#define DKS_FIX_INV 9.094947017729282379150390625e-13   /* 2^-40 */
__device__ __forceinline__ long long to_fix(double v) { return __double2ll_rn(v * DKS_FIX_SCALE); }
__device__ __forceinline__ double from_fix(long long t) { return (double)t * DKS_FIX_INV; }
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// np.isclose(a, b, rtol=1e-5, atol=1e-8, equal_nan=True) as used by KernelExplainer.not_equal
__device__ __forceinline__ bool np_isclose(double a, double b) {
    if (a == b) return true;
    if (isnan(a) && isnan(b)) return true;
    if (!isfinite(a) || !isfinite(b)) return false;
    return fabs(a - b) <= 1e-8 + 1e-5 * fabs(b);
}

// link(p): shap.common.LogitLink.f / IdentityLink.f
__device__ __forceinline__ double link_f(double p, int link) {
    return link == DKS_LINK_LOGIT ? log(p / (1.0 - p)) : p;
}

// model head in float64 on R scores -> C outputs (C = 2 for the binary head, else R)
__device__ inline void head_f64(const double* z, int R, int act, double kappa, double* out) {
    if (act == DKS_ACT_BINARY_LOGISTIC) {
        // softmax([-kz/2, kz/2]) evaluated the numerically stable way sklearn does
        double t = kappa * z[0];
        double e = exp(-fabs(t));
        double big = 1.0 / (1.0 + e), small = e / (1.0 + e);
        out[1] = t >= 0 ? big : small;
        out[0] = t >= 0 ? small : big;
    } else if (act == DKS_ACT_SOFTMAX) {
        double m = z[0];
        for (int r = 1; r < R; ++r) m = fmax(m, z[r]);
        double sum = 0;
        for (int r = 0; r < R; ++r) { out[r] = exp(z[r] - m); sum += out[r]; }
        for (int r = 0; r < R; ++r) out[r] /= sum;
    } else {
        for (int r = 0; r < R; ++r) out[r] = z[r];
    }
}

// ------------------------------------------------------------------------------------------------------
// K0: fit (DenseData + KernelExplainer.__init__)
// ------------------------------------------------------------------------------------------------------
// BW[j][g][r] = sum_{col in g} bg[j][col] * W[r][col]
__global__ void fit_bw_kernel(const double* __restrict__ bg, const double* __restrict__ W,
                              const int32_t* __restrict__ goff, const int32_t* __restrict__ gcols, int N, int D,
                              int G, int R, double* __restrict__ BW) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * G * R) return;
    int r = idx % R, g = (idx / R) % G, j = idx / (R * G);
    double acc = 0;
    for (int c = goff[g]; c < goff[g + 1]; ++c) {
        int col = gcols[c];
        acc += bg[(size_t)j * D + col] * W[(size_t)r * D + col];
    }
    BW[idx] = acc;
}

__global__ void fit_scores_kernel(const double* __restrict__ BW, const double* __restrict__ b, int N, int G, int R,
                                  double* __restrict__ scores) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * R) return;
    int r = idx % R, j = idx / R;
    double acc = b[r];
    for (int g = 0; g < G; ++g) acc += BW[((size_t)j * G + g) * R + r];
    scores[idx] = acc;
}

__global__ void fit_colstats_kernel(const double* __restrict__ bg, int N, int D, double* __restrict__ colmin,
                                    double* __restrict__ colmax, int* __restrict__ colnan) {
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= D) return;
    double mn = INFINITY, mx = -INFINITY;
    int nan = 0;
    for (int j = 0; j < N; ++j) {
        double v = bg[(size_t)j * D + col];
        if (isnan(v)) nan = 1;
        else { mn = fmin(mn, v); mx = fmax(mx, v); }
    }
    colmin[col] = mn; colmax[col] = mx; colnan[col] = nan;
}

// fnull[c] = sum_j w_j f(bg_j)[c]; linkfnull = link(fnull); Bbar[g][r] = sum_j w_j BW[j][g][r].  One block.
__global__ void fit_fnull_kernel(const double* __restrict__ scores, const double* __restrict__ BW,
                                 const double* __restrict__ wbg, int N, int G, int R, int C, int act, double kappa,
                                 int link, double* __restrict__ fnull, double* __restrict__ linkfnull,
                                 double* __restrict__ Bbar) {
    int t = threadIdx.x;
    if (t < C) {
        double acc = 0;
        for (int j = 0; j < N; ++j) {
            double out[DKS_MAX_OUT];
            head_f64(scores + (size_t)j * R, R, act, kappa, out);
            acc += out[t] * wbg[j];
        }
        fnull[t] = acc;
        linkfnull[t] = link_f(acc, link);
    }
    for (int idx = t; idx < G * R; idx += blockDim.x) {
        double acc = 0;
        for (int j = 0; j < N; ++j) acc += wbg[j] * BW[(size_t)j * G * R + idx];
        Bbar[idx] = acc;
    }
}

// scaled float copies consumed by the fused kernel: BWs[r][g][j] = scale*BW[j][g][r], bases[r][j] = scale*score
__global__ void fit_scale_kernel(const double* __restrict__ BW, const double* __restrict__ scores,
                                 const double* __restrict__ wbg, int N, int G, int R, double scale,
                                 float* __restrict__ BWs, float* __restrict__ bases, float* __restrict__ wbf) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N * G * R) {
        int j = idx % N, g = (idx / N) % G, r = idx / (N * G);
        BWs[idx] = (float)(scale * BW[((size_t)j * G + g) * R + r]);
    }
    if (idx < N * R) {
        int j = idx % N, r = idx / N;
        bases[idx] = (float)(scale * scores[(size_t)j * R + r]);
    }
    if (idx < N) wbf[idx] = (float)wbg[idx];
}

// f(X) for n rows, float64 (model check against the Python callable)
__global__ void predict_kernel(const double* __restrict__ X, const double* __restrict__ W,
                               const double* __restrict__ b, int n, int D, int R, int C, int act, double kappa,
                               double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double z[DKS_MAX_OUT], o[DKS_MAX_OUT];
    for (int r = 0; r < R; ++r) {
        double acc = b[r];
        for (int c = 0; c < D; ++c) acc += X[(size_t)i * D + c] * W[(size_t)r * D + c];
        z[r] = acc;
    }
    head_f64(z, R, act, kappa, o);
    for (int c = 0; c < C; ++c) out[(size_t)i * C + c] = o[c];
}

// Packed fp32 (sm_100 FFMA2/FMUL2/FADD2: two fp32 lanes per thread in one 64-bit register pair, one issue slot).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}


// ------------------------------------------------------------------------------------------------------
// Preparation: grouped instance contributions, varying_groups(), f(x), link deltas
// ------------------------------------------------------------------------------------------------------
// Fused preparation: a block handles `ipb` instances.  Phase 1, one thread per (instance, group): grouped
// contribution XW[i][g][r] and the "group varies" flag (KernelExplainer.varying_groups); phase 2, one thread per
// instance: varying bit-mask, M, histogram of M, f(x), link(f(x)) - link(fnull).
// STAGE: the block first copies what phase 1 reads -- its instances' rows of X, W, the column statistics and the group
// tables -- into shared memory with coalesced loads, so the per-(instance, group) loop runs without dependent global loads.
// (Measured on the Adult shape: 13.3 us per launch under ncu against 13.8 us unstaged -- the kernel is bound by launch +
// one cold DRAM round trip + the serial per-instance tail, not by the column loop; kept because it is never slower.)
// The host picks STAGE when the tables fit shared memory.
template <bool STAGE>
__global__ void prep_kernel(const double* __restrict__ X, const double* __restrict__ W, const double* __restrict__ b,
                            const double* __restrict__ bg, const int32_t* __restrict__ goff,
                            const int32_t* __restrict__ gcols, const double* __restrict__ colmin,
                            const double* __restrict__ colmax, const int* __restrict__ colnan,
                            const double* __restrict__ linkfnull, int n, int N, int D, int G, int R, int C, int act,
                            double kappa, int link, int ipb, double* __restrict__ XW, uint64_t* __restrict__ vmask,
                            int* __restrict__ Mcnt, double* __restrict__ dlink, int* __restrict__ hist,
                            int* __restrict__ counts, int* __restrict__ idx_full, int* __restrict__ idx_other,
                            double* __restrict__ XT, double xt_scale) {
    extern __shared__ __align__(16) unsigned char prep_smem[];
    double* sXW = reinterpret_cast<double*>(prep_smem);                        // [ipb][G][R]
    double* sX = sXW + (size_t)ipb * G * R;                                    // STAGE: [ipb][D]
    double* sW = sX + (STAGE ? (size_t)ipb * D : 0);                           //        [R][D]
    double* sMin = sW + (STAGE ? (size_t)R * D : 0);                           //        [D]
    double* sMax = sMin + (STAGE ? D : 0);                                     //        [D]
    int* sNan = reinterpret_cast<int*>(sMax + (STAGE ? D : 0));                //        [D]
    int* sCols = sNan + (STAGE ? D : 0);                                       //        [D]
    int* sOff = sCols + (STAGE ? D : 0);                                       //        [G + 1]
    unsigned char* sflag = reinterpret_cast<unsigned char*>(sOff + (STAGE ? G + 1 : 0));   // [ipb][G]
    const int i0 = blockIdx.x * ipb;
    if (STAGE) {
        const int rows = min(ipb, n - i0);
        const double* Xb = X + (size_t)i0 * D;
        for (int idx = threadIdx.x; idx < rows * D; idx += blockDim.x) sX[idx] = Xb[idx];
        for (int idx = threadIdx.x; idx < R * D; idx += blockDim.x) sW[idx] = W[idx];
        for (int idx = threadIdx.x; idx < D; idx += blockDim.x) {
            sMin[idx] = colmin[idx]; sMax[idx] = colmax[idx]; sNan[idx] = colnan[idx]; sCols[idx] = gcols[idx];
        }
        for (int idx = threadIdx.x; idx <= G; idx += blockDim.x) sOff[idx] = goff[idx];
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < ipb * G; idx += blockDim.x) {
        const int li = idx / G, g = idx - li * G, i = i0 + li;
        if (i >= n) continue;
        bool varies = false;
        double acc[8];
        for (int r = 0; r < R; ++r) acc[r] = 0;
        const int c0 = STAGE ? sOff[g] : goff[g], c1 = STAGE ? sOff[g + 1] : goff[g + 1];
        for (int c = c0; c < c1; ++c) {
            const int col = STAGE ? sCols[c] : gcols[c];
            const double xv = STAGE ? sX[(size_t)li * D + col] : X[(size_t)i * D + col];
            for (int r = 0; r < R; ++r) acc[r] += xv * (STAGE ? sW[(size_t)r * D + col] : W[(size_t)r * D + col]);
            if ((STAGE ? sNan[col] : colnan[col]) || isnan(xv)) {
                if (!varies)
                    for (int j = 0; j < N && !varies; ++j) varies = !np_isclose(xv, bg[(size_t)j * D + col]);
            } else {
                // |x-b| - rtol|b| is decreasing for b <= x and increasing for b >= x: the extremes decide
                const double mn = STAGE ? sMin[col] : colmin[col], mx = STAGE ? sMax[col] : colmax[col];
                varies = varies || !np_isclose(xv, mn) || !np_isclose(xv, mx);
            }
        }
        for (int r = 0; r < R; ++r) {
            sXW[(size_t)idx * R + r] = acc[r];
            XW[((size_t)i * G + g) * R + r] = acc[r];
        }
        sflag[idx] = varies ? 1 : 0;
    }
    __syncthreads();
    if (XT != nullptr) {
        // XT[i][t][x] = xt_scale * sum_{b<4} bit_b(x) XW[i][4t+b]: the shared-plan kernel adds one table entry per nibble
        // of a coalition row instead of one term per group (R == 1)
        const int ntab = (G + 3) / 4;
        for (int idx = threadIdx.x; idx < ipb * ntab * 16; idx += blockDim.x) {
            const int li = idx / (ntab * 16), rem = idx - li * ntab * 16, t = rem >> 4, x = rem & 15;
            const int i = i0 + li;
            if (i >= n) continue;
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (((x >> b) & 1) && 4 * t + b < G) acc += sXW[(size_t)li * G + 4 * t + b];
            XT[((size_t)i * ntab + t) * 16 + x] = xt_scale * acc;
        }
    }
    for (int li = threadIdx.x; li < ipb; li += blockDim.x) {
        const int i = i0 + li;
        if (i >= n) continue;
        uint64_t m = 0;                       // varying bit-mask (groups 0..63; wider problems only use the count)
        int M = 0;
        double z[8], o[DKS_MAX_OUT];
        for (int r = 0; r < R; ++r) z[r] = b[r];
        for (int g = 0; g < G; ++g) {
            if (sflag[li * G + g]) { if (g < 64) m |= (1ull << g); ++M; }
            for (int r = 0; r < R; ++r) z[r] += sXW[((size_t)li * G + g) * R + r];
        }
        vmask[i] = m;
        Mcnt[i] = M;
        atomicAdd(&hist[M], 1);
        // bucket: all groups vary (candidates for the shared-plan fast path) / everything else
        if (M == G && M >= 2) idx_full[atomicAdd(&counts[0], 1)] = i;
        else idx_other[atomicAdd(&counts[1], 1)] = i;
        head_f64(z, R, act, kappa, o);
        for (int c = 0; c < C; ++c) dlink[(size_t)i * C + c] = link_f(o[c], link) - linkfnull[c];
    }
}
inline size_t prep_smem_bytes(bool stage, int ipb, int G, int R, int D) {
    size_t b = sizeof(double) * (size_t)ipb * G * R + (size_t)ipb * G + 16;
    if (stage) b += sizeof(double) * ((size_t)ipb * D + (size_t)R * D + 2 * (size_t)D) + sizeof(int) * (2 * (size_t)D + G + 1);
    return b;
}

// ------------------------------------------------------------------------------------------------------
// Constrained WLS (KernelExplainer.solve without the l1 branch), float64, one CTA
// ------------------------------------------------------------------------------------------------------
// With L the last varying position and z' = (z_L ? ~z : z):  e_k e_l = z'_k & z'_l  and  e_k = (z_L ? -1 : 1) z'_k,
// where e_k = z_k - z_L is a column of upstream's `etmp`.

// A = E^T diag(w) E for k,l < M-1 (symmetric, row-major nA x nA), built warp-per-entry
__device__ inline void wls_build_normal(const uint64_t* __restrict__ zp, const double* __restrict__ wp, int S, int M,
                                        double* A, int warp, int nwarps) {
    const int nA = M - 1, L = M - 1;
    const int lane = threadIdx.x & 31;
    const int npairs = nA * (nA + 1) / 2;
    for (int pr = warp; pr < npairs; pr += nwarps) {
        int k = 0, rem = pr;
        while (rem > k) { rem -= (k + 1); ++k; }  // pr = k(k+1)/2 + l, l <= k
        int l = rem;
        double acc = 0;
        for (int s = lane; s < S; s += 32) {
            uint64_t z = zp[s];
            if ((z >> L) & 1ull) z = ~z;
            if (((z >> k) & (z >> l)) & 1ull) acc += wp[s];
        }
        acc = warp_sum(acc);
        if (lane == 0) { A[k * nA + l] = acc; A[l * nA + k] = acc; }
    }
}

// in-place lower Cholesky of the nA x nA matrix A (row-major) by warp 0; returns false if not positive definite
__device__ inline bool wls_cholesky_warp(double* A, int nA) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
    for (int c = 0; c < nA; ++c) {
        double d = A[c * nA + c];
        if (!(d > 0.0)) ok = false;
        d = sqrt(d);
        __syncwarp();
        if (lane == 0) A[c * nA + c] = d;
        for (int r = c + 1 + lane; r < nA; r += 32) A[r * nA + c] /= d;
        __syncwarp();
        for (int r = c + 1 + lane; r < nA; r += 32) {
            double lrc = A[r * nA + c];
            for (int c2 = c + 1; c2 <= r; ++c2) A[r * nA + c2] -= lrc * A[c2 * nA + c];
        }
        __syncwarp();
    }
    return ok;
}

// rhs[k] = sum_s w_s e_sk (y_s - z_sL * delta), warp-per-k
__device__ inline void wls_build_rhs(const uint64_t* __restrict__ zp, const double* __restrict__ wp,
                                     const double* ys, int S, int M, double delta, double* rhs, int warp, int nwarps) {
    const int nA = M - 1, L = M - 1;
    const int lane = threadIdx.x & 31;
    for (int k = warp; k < nA; k += nwarps) {
        double acc = 0;
        for (int s = lane; s < S; s += 32) {
            uint64_t z = zp[s];
            int zl = (int)((z >> L) & 1ull), zk = (int)((z >> k) & 1ull);
            int e = zk - zl;
            if (e != 0) {
                double yy = ys[s] - (zl ? delta : 0.0);
                acc += wp[s] * (double)e * yy;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) rhs[k] = acc;
    }
}

// solve (L L^T) beta = rhs in place (thread 0), then write phi for output dim c of instance i
__device__ inline void wls_solve_write(const double* Lf, double* rhs, int M, double delta, const int* vi,
                                       double* __restrict__ phi_row, double sign) {
    const int nA = M - 1;
    for (int r = 0; r < nA; ++r) {
        double v = rhs[r];
        for (int c = 0; c < r; ++c) v -= Lf[r * nA + c] * rhs[c];
        rhs[r] = v / Lf[r * nA + r];
    }
    for (int r = nA - 1; r >= 0; --r) {
        double v = rhs[r];
        for (int c = r + 1; c < nA; ++c) v -= Lf[c * nA + r] * rhs[c];
        rhs[r] = v / Lf[r * nA + r];
    }
    double sum = 0;
    for (int k = 0; k < nA; ++k) {
        double v = rhs[k];
        sum += v;
        if (fabs(v) < 1e-10) v = 0;
        phi_row[vi[k]] = sign * v;
    }
    double last = delta - sum;
    if (fabs(last) < 1e-10) last = 0;
    phi_row[vi[nA]] = sign * last;
}

// Block-level: Cholesky factor and inverse of the nA x nA matrix in A (shared memory, row-major; A must be followed by
// nA*nA doubles of scratch).  Needs blockDim.x >= max(32, nA).  Returns (in every thread) whether A was positive definite.
__device__ inline bool wls_factor_invert(double* A, int nA, double* __restrict__ chol, double* __restrict__ ainv) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < 32) {
        bool ok = wls_cholesky_warp(A, nA);
        if (!ok) s_ok = 0;
    }
    __syncthreads();
    if (chol != nullptr)
        for (int idx = threadIdx.x; idx < nA * nA; idx += blockDim.x) chol[idx] = A[idx];
    // inverse of E^T W E (what upstream's np.linalg.inv computes): column c of the inverse solves L L^T x = e_c
    if ((int)threadIdx.x < nA) {
        const int c = threadIdx.x;
        double* x = A + nA * nA + c * nA;   // scratch column in shared memory
        for (int r = 0; r < nA; ++r) {
            double v = (r == c) ? 1.0 : 0.0;
            for (int k = 0; k < r; ++k) v -= A[r * nA + k] * x[k];
            x[r] = v / A[r * nA + r];
        }
        for (int r = nA - 1; r >= 0; --r) {
            double v = x[r];
            for (int k = r + 1; k < nA; ++k) v -= A[k * nA + r] * x[k];
            x[r] = v / A[r * nA + r];
        }
        if (ainv != nullptr)
            for (int r = 0; r < nA; ++r) ainv[r * nA + c] = x[r];
    }
    __syncthreads();
    return s_ok != 0;
}

// factor the normal matrix of a shared plan once (dks_set_shared_plan): one CTA
__global__ void plan_factor_kernel(const uint64_t* __restrict__ z, const double* __restrict__ w, int S, int M,
                                   double* __restrict__ chol, double* __restrict__ ainv, int* __restrict__ status) {
    extern __shared__ double sm_d[];
    double* A = sm_d;
    const int nA = M - 1;
    wls_build_normal(z, w, S, M, A, threadIdx.x >> 5, blockDim.x >> 5);
    __syncthreads();
    const bool ok = wls_factor_invert(A, nA, chol, ainv);
    if (!ok && threadIdx.x == 0) { status[0] = DKS_ERR_NUMERIC; status[1] = M; }
}

// ---- plans of 65..128 groups (two words per row) -----------------------------------------------------------------
__device__ __forceinline__ bool zbit2(const uint64_t* __restrict__ row, int k) { return (row[k >> 6] >> (k & 63)) & 1ull; }

// Same factorisation for two-word plans: the matrix (up to 127 x 127) lives in shared memory, the columns of the inverse
// are solved in a global scratch buffer [nA][nA].
__global__ void plan_factor_wide_kernel(const uint64_t* __restrict__ z, const double* __restrict__ w, int S, int M,
                                        double* __restrict__ chol, double* __restrict__ ainv, double* __restrict__ scratch,
                                        int* __restrict__ status) {
    extern __shared__ double sm_d[];
    double* A = sm_d;
    const int nA = M - 1, L = M - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int npairs = nA * (nA + 1) / 2;
    for (int pr = warp; pr < npairs; pr += nwarps) {
        int k = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
        while (k * (k + 1) / 2 > pr) --k;
        while ((k + 1) * (k + 2) / 2 <= pr) ++k;
        const int l = pr - k * (k + 1) / 2;
        double acc = 0;
        for (int s = lane; s < S; s += 32) {
            const uint64_t* row = z + (size_t)s * 2;
            const bool zl = zbit2(row, L);
            if ((zbit2(row, k) != zl) && (zbit2(row, l) != zl)) acc += w[s];
        }
        acc = warp_sum(acc);
        if (lane == 0) { A[k * nA + l] = acc; A[l * nA + k] = acc; }
    }
    __syncthreads();
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < 32) { if (!wls_cholesky_warp(A, nA)) s_ok = 0; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nA * nA; idx += blockDim.x) chol[idx] = A[idx];
    for (int c = threadIdx.x; c < nA; c += blockDim.x) {
        double* x = scratch + (size_t)c * nA;
        for (int r = 0; r < nA; ++r) {
            double v = (r == c) ? 1.0 : 0.0;
            for (int k = 0; k < r; ++k) v -= A[r * nA + k] * x[k];
            x[r] = v / A[r * nA + r];
        }
        for (int r = nA - 1; r >= 0; --r) {
            double v = x[r];
            for (int k = r + 1; k < nA; ++k) v -= A[k * nA + r] * x[k];
            x[r] = v / A[r * nA + r];
        }
        for (int r = 0; r < nA; ++r) ainv[r * nA + c] = x[r];
    }
    if (threadIdx.x == 0 && !s_ok) { status[0] = DKS_ERR_NUMERIC; status[1] = M; }
}

// Push all-gather: every rank stores its block of phi into slab `rank` of each peer's gathered buffer through NVLink
// peer memory (128-bit stores; blockIdx.y = peer slot).  The caller follows up with a cross-GPU barrier.
struct PeerPush { double* dst[16]; int npeers; };
__global__ void push_phi_kernel(const double* __restrict__ src, PeerPush pp, long long n_doubles) {
    double* dst = pp.dst[blockIdx.y];
    const long long nvec = n_doubles >> 1;
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(dst);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) d2[i] = s2[i];
    if ((n_doubles & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n_doubles - 1] = src[n_doubles - 1];
}

// Same, for the rows of an instance list only (the fused shared-plan kernel has already stored its instances into the
// peers' buffers; what the general kernels computed still has to travel).  phi is [C][n][G].
__global__ void push_rows_kernel(const double* __restrict__ src, PeerPush pp, const int* __restrict__ list,
                                 const int* __restrict__ count, int n, int G, int C) {
    const int cnt = *count;
    const long long per = (long long)C * G;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < cnt * per; idx += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(idx / per), rem = (int)(idx - q * per), c = rem / G, g = rem - c * G;
        const size_t off = ((size_t)c * n + list[q]) * G + g;
        const double v = src[off];
        for (int r = 0; r < pp.npeers; ++r) pp.dst[r][off] = v;
    }
}

// ---- KernelShap.build_explanation post-processing off the resident phi (kernel_shap.py:36-109, :112-207, :952-956) ----------
// Segment sums over consecutive groups (sum_categories), |.| accumulated per (output, segment) in 2^-40 fixed point
// (shared-memory atomics per block, one global atomic per block and cell: order-independent), argmax of the raw prediction.
__global__ void phi_summary_kernel(const double* __restrict__ phi, int C, int n, int G, const int* __restrict__ seg, int Gp,
                                   double* __restrict__ phi_sum, unsigned long long* __restrict__ absacc,
                                   const double* __restrict__ dlink, const double* __restrict__ linkfnull,
                                   int* __restrict__ argmax) {
    extern __shared__ unsigned long long s_abs[];           // [C * Gp]
    for (int idx = threadIdx.x; idx < C * Gp; idx += blockDim.x) s_abs[idx] = 0ull;
    __syncthreads();
    const long long total = (long long)C * n * Gp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int gp = (int)(idx % Gp);
        const long long ci = idx / Gp;
        const int i = (int)(ci % n), c = (int)(ci / n);
        const int g0 = seg ? seg[gp] : gp, g1 = seg ? seg[gp + 1] : gp + 1;
        double v = 0.0;
        for (int g = g0; g < g1; ++g) v += phi[((size_t)c * n + i) * G + g];
        if (phi_sum) phi_sum[idx] = v;
        atomicAdd(&s_abs[c * Gp + gp], (unsigned long long)to_fix(fabs(v)));
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < C * Gp; idx += blockDim.x)
        if (s_abs[idx]) atomicAdd(&absacc[idx], s_abs[idx]);
    if (argmax != nullptr) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            int best = 0;
            double bv = dlink[(size_t)i * C] + linkfnull[0];
            for (int c = 1; c < C; ++c) {
                const double v = dlink[(size_t)i * C + c] + linkfnull[c];
                if (v > bv) { bv = v; best = c; }
            }
            argmax[i] = best;
        }
    }
}
// mean |phi| per output and aggregated over outputs ([C + 1][Gp]) and their descending order (ties: higher index first,
// what reversing a stable ascending argsort gives).  One block.
__global__ void phi_rank_kernel(const unsigned long long* __restrict__ absacc, int C, int n, int Gp, double* __restrict__ mean_abs,
                                int* __restrict__ order) {
    for (int idx = threadIdx.x; idx < (C + 1) * Gp; idx += blockDim.x) {
        const int r = idx / Gp, g = idx - r * Gp;
        double v = 0.0;
        if (r < C) v = from_fix((long long)absacc[r * Gp + g]) / (double)n;
        else for (int c = 0; c < C; ++c) v += from_fix((long long)absacc[c * Gp + g]) / (double)n;
        mean_abs[idx] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < (C + 1) * Gp; idx += blockDim.x) {
        const int r = idx / Gp, g = idx - r * Gp;
        const double v = mean_abs[idx];
        int rank = 0;
        for (int l = 0; l < Gp; ++l) {
            const double o = mean_abs[r * Gp + l];
            if (o > v || (o == v && l > g)) ++rank;
        }
        order[r * Gp + rank] = g;
    }
}

// Cross-GPU completion of the push all-gather without a library barrier: every rank keeps a flag word per peer in
// peer-mapped memory.  After the solve kernels (whose epilogues stored phi into the peers' buffers) one thread per peer
// publishes this rank's step count into the peer's flag array (system-scope release) and waits until the peer's count has
// reached the same step (acquire).  The step counter lives on the device, so a replayed CUDA graph keeps counting.
struct PeerFlags {
    unsigned long long* mine;            // [world] flags written by the peers
    unsigned long long* peer[16];        // peer[r]: rank r's flag array, mapped here
    unsigned long long* step;            // this rank's step counter (device memory)
    int world, rank;
};
__global__ void peer_sync_kernel(PeerFlags f, int* __restrict__ status) {
    __shared__ unsigned long long s_step;
    if (threadIdx.x == 0) { s_step = *f.step + 1ull; *f.step = s_step; }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= f.world || t == f.rank) return;
    const unsigned long long e = s_step;
    __threadfence_system();              // everything this GPU stored before (previous kernels included) is ordered first
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.peer[t] + f.rank), "l"(e) : "memory");
    unsigned long long seen = 0ull;
    for (long long spins = 0; spins < (1ll << 31); ++spins) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f.mine + t) : "memory");
        if (seen >= e) return;
    }
    if (atomicCAS(&status[0], 0, DKS_ERR_CUDA) == 0) status[1] = -78;     // a peer never arrived
}

// an instance list that must be empty (shapes no kernel covers): report instead of computing
__global__ void flag_unsupported_kernel(const int* __restrict__ count, int detail, int* __restrict__ status) {
    if (*count > 0 && atomicCAS(&status[0], 0, DKS_ERR_UNSUPPORTED) == 0) status[1] = detail;
}

// normal matrix of the first `rows` rows of a plan (its enumerated prefix), unfactored: the per-instance sampler adds
// the sampled rows' part to it
__global__ void plan_prefix_normal_kernel(const uint64_t* __restrict__ z, const double* __restrict__ w, int rows, int M,
                                          double* __restrict__ afix) {
    extern __shared__ double sm_d[];
    const int nA = M - 1;
    wls_build_normal(z, w, rows, M, sm_d, threadIdx.x >> 5, blockDim.x >> 5);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nA * nA; idx += blockDim.x) afix[idx] = sm_d[idx];
}

// ------------------------------------------------------------------------------------------------------
// Fused coalition kernel, CUDA-core (SIMT) version.  One CTA per instance (grid-stride), one thread per
// coalition row.  Handles the binary-logistic and identity heads; any plan source.
// ------------------------------------------------------------------------------------------------------
struct SimtSmem {
    double* ys;     // [S_cap] link(ey) - link(fnull) per coalition
    double* A;      // [63*63] normal matrix / its Cholesky factor
    double* rhs;    // [64]
    double* xw;     // [64] scaled grouped contributions of the instance (varying positions)
    int* vi;        // [64] varying position -> group index
    float* Bs;      // [M][N] scaled grouped background contributions of the varying groups, then bases[N], wb[N]
};

__device__ inline SimtSmem simt_carve(unsigned char* base, int S_cap, int R, int ny) {
    SimtSmem s;
    s.ys = reinterpret_cast<double*>(base);
    s.A = s.ys + (size_t)S_cap * ny;
    s.rhs = s.A + 63 * 63;
    s.xw = s.rhs + 64;
    s.vi = reinterpret_cast<int*>(s.xw + 64 * (size_t)R);
    s.Bs = reinterpret_cast<float*>(s.vi + 64);
    return s;
}

// ny = number of y buffers (1, or C for the softmax head), R = score rows staged per instance
__host__ __device__ inline size_t simt_smem_bytes(int S_cap, int N, int Mmax, int R = 1, int ny = 1) {
    return sizeof(double) * ((size_t)S_cap * ny + 63 * 63 + 64 + 64 * (size_t)R) + sizeof(int) * 64 +
           sizeof(float) * ((size_t)Mmax * N * R + (size_t)N * R + (size_t)N);
}

__global__ void __launch_bounds__(256) explain_simt_kernel(ExplainParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bool softmax = p.act == DKS_ACT_SOFTMAX;
    SimtSmem sm = simt_carve(smem_raw, p.S_cap, softmax ? p.R : 1, softmax ? p.C : 1);
    const int tid = threadIdx.x;
    const int N = p.N, G = p.G, C = p.C;
    const size_t slab = (size_t)p.n * G;

    const int ninst = dks_inst_count(p);
    for (int qi = blockIdx.x; qi < ninst; qi += gridDim.x) {
        const int i = dks_inst_at(p, qi);
        const int M = p.Mcnt[i];
        const uint64_t vm = p.vmask[i];
        __syncthreads();  // previous instance done with shared memory
        for (int idx = tid; idx < C * G; idx += blockDim.x) p.phi[(size_t)(idx / G) * slab + (size_t)i * G + idx % G] = 0.0;
        if (M == 0) continue;
        if (M == 1) {
            if (tid < C) {
                int g = __ffsll((long long)vm) - 1;
                p.phi[(size_t)tid * slab + (size_t)i * G + g] = p.dlink[(size_t)i * C + tid];
            }
            continue;
        }
        const int S = dks_effective_S(M, p.S_req);
        const uint64_t* zp;
        const double* wp;
        const double* chol = nullptr;
        if (p.ext_z != nullptr) {
            zp = p.ext_z + (size_t)i * p.ext_stride;
            wp = p.ext_w + (size_t)i * p.ext_stride;
            if (p.ext_chol != nullptr) chol = p.ext_chol + (size_t)i * p.ext_fstride;   // factored with the plan
        } else {
            PlanDev pd = p.plans[M];
            if (pd.z == nullptr || pd.S != S) {
                if (tid == 0) { if (atomicCAS(&p.status[0], 0, DKS_ERR_PLAN_MISSING) == 0) p.status[1] = M; }
                continue;
            }
            zp = pd.z; wp = pd.w; chol = pd.chol;
        }
        if (S > p.S_cap) {
            if (tid == 0) { if (atomicCAS(&p.status[0], 0, DKS_ERR_INVALID) == 0) p.status[1] = i; }
            continue;
        }

        if (tid == 0) {
            int k = 0;
            for (int g = 0; g < G; ++g) if ((vm >> g) & 1ull) sm.vi[k++] = g;
        }
        __syncthreads();

        float* Bs = sm.Bs;
        float* bases = Bs + (size_t)M * N;
        float* wb = bases + N;

        if (p.act == DKS_ACT_BINARY_LOGISTIC) {
            // stage this instance's varying columns of the background table
            for (int idx = tid; idx < M * N; idx += blockDim.x) {
                int k = idx / N, j = idx - k * N;
                Bs[idx] = p.BWs[(size_t)sm.vi[k] * N + j];
            }
            for (int j = tid; j < N; j += blockDim.x) { bases[j] = p.bases[j]; wb[j] = p.wbf[j]; }
            if (tid < M) sm.xw[tid] = p.scale * p.XW[(size_t)i * G + sm.vi[tid]];
            __syncthreads();

            const double lf1 = p.linkfnull[1], f1 = p.fnull[1];
            for (int s = tid; s < S; s += blockDim.x) {
                const uint64_t z = zp[s];
                double a = 0;
                for (int k = 0; k < M; ++k) if ((z >> k) & 1ull) a += sm.xw[k];
                const float af = (float)a;
                float acc1 = 0.f, acc0 = 0.f;
                for (int j = 0; j < N; ++j) {
                    float c = 0.f;
                    for (int k = 0; k < M; ++k) if ((z >> k) & 1ull) c += Bs[k * N + j];
                    float t = (bases[j] - c) + af;     // = -kappa*log2(e) * masked score
                    t = fminf(fmaxf(t, -120.f), 120.f);
                    float u = ex2_approx(t);           // exp(-kappa*score)
                    float r = rcp_approx(1.f + u);     // p1 = sigmoid(kappa*score)
                    acc1 = fmaf(wb[j], r, acc1);
                    acc0 = fmaf(wb[j], u * r, acc0);   // p0 = 1 - p1, accumulated without cancellation
                }
                double y;
                if (p.link == DKS_LINK_LOGIT) y = log((double)acc1 / (double)acc0) - lf1;
                else y = (double)acc1 - f1;
                sm.ys[s] = y;
            }
            __syncthreads();

            // WLS for output 1; output 0 is its exact negation (p0 = 1 - p1 row-wise)
            const double* Lf;
            if (chol != nullptr) {
                for (int idx = tid; idx < (M - 1) * (M - 1); idx += blockDim.x) sm.A[idx] = chol[idx];
            } else {
                wls_build_normal(zp, wp, S, M, sm.A, threadIdx.x >> 5, blockDim.x >> 5);
                __syncthreads();
                if (tid < 32) {
                    bool ok = wls_cholesky_warp(sm.A, M - 1);
                    if (!ok && tid == 0) { if (atomicCAS(&p.status[0], 0, DKS_ERR_NUMERIC) == 0) p.status[1] = i; }
                }
            }
            Lf = sm.A;
            const double delta = p.dlink[(size_t)i * C + 1];
            wls_build_rhs(zp, wp, sm.ys, S, M, delta, sm.rhs, threadIdx.x >> 5, blockDim.x >> 5);
            __syncthreads();
            if (tid == 0) {
                wls_solve_write(Lf, sm.rhs, M, delta, sm.vi, p.phi + slab + (size_t)i * G, 1.0);
                double* phi0 = p.phi + (size_t)i * G;
                const double* phi1 = p.phi + slab + (size_t)i * G;
                for (int k = 0; k < M; ++k) { double v = phi1[sm.vi[k]]; phi0[sm.vi[k]] = (v == 0.0) ? 0.0 : -v; }
            }
        } else if (p.act == DKS_ACT_SOFTMAX) {
            // ---- general softmax head: R = C score rows, outputs softmax(scores); scale = log2(e) ----
            const int R = p.R;
            float* basesR = Bs + (size_t)R * M * N;      // [R][N]
            float* wbR = basesR + (size_t)R * N;         // [N]
            for (int idx = tid; idx < R * M * N; idx += blockDim.x) {
                const int r = idx / (M * N), rem = idx - r * (M * N), k = rem / N, j = rem - k * N;
                Bs[idx] = p.BWs[((size_t)r * G + sm.vi[k]) * N + j];
            }
            for (int idx = tid; idx < R * N; idx += blockDim.x) basesR[idx] = p.bases[idx];
            for (int j = tid; j < N; j += blockDim.x) wbR[j] = p.wbf[j];
            for (int idx = tid; idx < R * M; idx += blockDim.x) {
                const int r = idx / M, k = idx - r * M;
                sm.xw[r * 64 + k] = p.scale * p.XW[((size_t)i * G + sm.vi[k]) * R + r];
            }
            __syncthreads();
            for (int s = tid; s < S; s += blockDim.x) {
                const uint64_t z = zp[s];
                float af[8], acc[8];
                for (int r = 0; r < R; ++r) {
                    double a = 0;
                    for (int k = 0; k < M; ++k) if ((z >> k) & 1ull) a += sm.xw[r * 64 + k];
                    af[r] = (float)a;
                    acc[r] = 0.f;
                }
                for (int j = 0; j < N; ++j) {
                    float t[8], mx = -3.0e38f;
                    for (int r = 0; r < R; ++r) {
                        float c = 0.f;
                        const float* Br = Bs + (size_t)r * M * N;
                        for (int k = 0; k < M; ++k) if ((z >> k) & 1ull) c += Br[k * N + j];
                        t[r] = (basesR[r * N + j] - c) + af[r];
                        mx = fmaxf(mx, t[r]);
                    }
                    float den = 0.f;
                    for (int r = 0; r < R; ++r) { t[r] = ex2_approx(t[r] - mx); den += t[r]; }
                    const float inv = wbR[j] * rcp_approx(den);
                    for (int r = 0; r < R; ++r) acc[r] = fmaf(t[r], inv, acc[r]);
                }
                for (int c = 0; c < C; ++c) {
                    double y;
                    if (p.link == DKS_LINK_LOGIT) {
                        float rest = 0.f;                 // 1 - ey_c as the sum of the other classes: no cancellation
                        for (int c2 = 0; c2 < C; ++c2) if (c2 != c) rest += acc[c2];
                        y = log((double)acc[c] / (double)rest) - p.linkfnull[c];
                    } else {
                        y = (double)acc[c] - p.fnull[c];
                    }
                    sm.ys[(size_t)c * p.S_cap + s] = y;
                }
            }
            __syncthreads();
            if (chol != nullptr) {
                for (int idx = tid; idx < (M - 1) * (M - 1); idx += blockDim.x) sm.A[idx] = chol[idx];
            } else {
                wls_build_normal(zp, wp, S, M, sm.A, threadIdx.x >> 5, blockDim.x >> 5);
                __syncthreads();
                if (tid < 32) {
                    bool ok = wls_cholesky_warp(sm.A, M - 1);
                    if (!ok && tid == 0) { if (atomicCAS(&p.status[0], 0, DKS_ERR_NUMERIC) == 0) p.status[1] = i; }
                }
            }
            for (int c = 0; c < C; ++c) {
                __syncthreads();
                const double delta = p.dlink[(size_t)i * C + c];
                wls_build_rhs(zp, wp, sm.ys + (size_t)c * p.S_cap, S, M, delta, sm.rhs, threadIdx.x >> 5, blockDim.x >> 5);
                __syncthreads();
                if (tid == 0) wls_solve_write(sm.A, sm.rhs, M, delta, sm.vi, p.phi + (size_t)c * slab + (size_t)i * G, 1.0);
            }
        } else if (p.act == DKS_ACT_IDENTITY) {
            // identity head: the background average commutes with the head, so
            // ey_r(s) = fnull_r + sum_k z_sk (XW_i[k][r] - Bbar[k][r])   -- float64 throughout
            const double* Lf;
            if (chol != nullptr) {
                for (int idx = tid; idx < (M - 1) * (M - 1); idx += blockDim.x) sm.A[idx] = chol[idx];
            } else {
                wls_build_normal(zp, wp, S, M, sm.A, threadIdx.x >> 5, blockDim.x >> 5);
                __syncthreads();
                if (tid < 32) {
                    bool ok = wls_cholesky_warp(sm.A, M - 1);
                    if (!ok && tid == 0) { if (atomicCAS(&p.status[0], 0, DKS_ERR_NUMERIC) == 0) p.status[1] = i; }
                }
            }
            Lf = sm.A;
            for (int r = 0; r < p.R; ++r) {
                __syncthreads();
                if (tid < M) {
                    int g = sm.vi[tid];
                    sm.xw[tid] = p.XW[((size_t)i * G + g) * p.R + r] - p.Bbar[(size_t)g * p.R + r];
                }
                __syncthreads();
                const double fn = p.fnull[r], lfn = p.linkfnull[r];
                for (int s = tid; s < S; s += blockDim.x) {
                    const uint64_t z = zp[s];
                    double a = fn;
                    for (int k = 0; k < M; ++k) if ((z >> k) & 1ull) a += sm.xw[k];
                    sm.ys[s] = link_f(a, p.link) - lfn;
                }
                __syncthreads();
                const double delta = p.dlink[(size_t)i * C + r];
                wls_build_rhs(zp, wp, sm.ys, S, M, delta, sm.rhs, threadIdx.x >> 5, blockDim.x >> 5);
                __syncthreads();
                if (tid == 0) wls_solve_write(Lf, sm.rhs, M, delta, sm.vi, p.phi + (size_t)r * slab + (size_t)i * G, 1.0);
            }
        }
    }
}

}  // namespace dks
