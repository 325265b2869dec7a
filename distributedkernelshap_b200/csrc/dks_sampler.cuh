// Device-side per-instance coalition plans (the sampling part of KernelExplainer.explain).
//
// Upstream draws a fresh plan for every instance from an advancing MT19937 stream.  Here every instance gets its own
// plan from a counter-based generator, Philox4x32-10 keyed by the seed with counter (draw t, global row, block), so the
// plan of a row does not depend on batching, sharding or the number of GPUs.  The sequential semantics are upstream's:
//   draw t: subset size ~ p (sizes not fully enumerated), then a uniform subset of that size;
//   a mask seen before adds 1 to the weight of its first occurrence (and of its complement row);
//   a new mask takes the next row, followed by its complement when the size is "paired" and a row is left;
//   stop when the budget S is filled; sampled weights are rescaled to the mass the enumerated sizes left over.
// The enumerated prefix (deterministic per M) is copied from the shared plan of that M.  One CTA per instance: draws are
// produced in batches of 1024, first occurrences resolved with a shared-memory hash table (atomicMin on the draw index),
// row positions with a block prefix sum -- the result is independent of thread scheduling.
// tests/sampler_twin.py is the NumPy twin; tests compare the plans bit for bit.
#pragma once

#include "dks_common.cuh"

namespace dks {
namespace sampler {

constexpr int THREADS = 256;
constexpr int DRAWS_PER_THREAD = 4;
constexpr int BATCH = THREADS * DRAWS_PER_THREAD;
constexpr int MAX_SAMPLED = 4096;          // rows the sampled part of a plan may have
constexpr int TABLE_CAP = 2 * MAX_SAMPLED; // hash table slots (power of two)

using SamplingInfo = ::DksSamplingInfo;

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// position (from the LSB) of the r-th (0-based) set bit of x
__device__ __forceinline__ int nth_set_bit(uint64_t x, int r) {
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    const int clo = __popc(lo);
    if (r < clo) return (int)__fns(lo, 0, r + 1);
    return 32 + (int)__fns(hi, 0, r - clo + 1);
}

__device__ __forceinline__ uint32_t hash_mask(uint64_t m) {
    m ^= m >> 33; m *= 0xff51afd7ed558ccdull; m ^= m >> 33; m *= 0xc4ceb9fe1a85ec53ull; m ^= m >> 33;
    return (uint32_t)m;
}

struct SamplerParams {
    int n, G, S_req, stride;
    uint64_t seed;
    long long row_offset;           // global index of row 0 of this call
    const int* Mcnt;
    const PlanDev* plans;           // shared plans: source of the enumerated prefix
    const SamplingInfo* info;       // [DKS_MAX_GROUPS + 1]
    uint64_t* out_z;                // [n][stride]
    double* out_w;                  // [n][stride]
    int* status;
};

__global__ void __launch_bounds__(THREADS) sample_plans_kernel(SamplerParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t* tkey = reinterpret_cast<uint64_t*>(smraw);                 // [TABLE_CAP]
    uint32_t* tfirst = reinterpret_cast<uint32_t*>(tkey + TABLE_CAP);     // [TABLE_CAP] first draw index
    uint32_t* tcount = tfirst + TABLE_CAP;                                // [TABLE_CAP] multiplicity among counted draws
    uint32_t* rowslot = tcount + TABLE_CAP;                               // [MAX_SAMPLED] table slot of each sampled row
    __shared__ uint32_t s_warp[THREADS / 32];
    __shared__ uint32_t s_carry;
    __shared__ double s_red[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;

    for (int i = blockIdx.x; i < p.n; i += gridDim.x) {
        const int M = p.Mcnt[i];
        uint64_t* oz = p.out_z + (size_t)i * p.stride;
        double* ow = p.out_w + (size_t)i * p.stride;
        if (M < 2) continue;
        const int S = dks_effective_S(M, p.S_req);
        const PlanDev pd = p.plans[M];
        if (pd.z == nullptr || pd.S != S || S > p.stride) {
            if (tid == 0) { if (atomicCAS(&p.status[0], 0, DKS_ERR_PLAN_MISSING) == 0) p.status[1] = M; }
            for (int s = tid; s < p.stride; s += THREADS) { oz[s] = 0ull; ow[s] = 0.0; }
            continue;
        }
        const SamplingInfo& inf = p.info[M];
        const int nfixed = inf.nfixed, left0 = S - nfixed;
        // enumerated prefix (deterministic per M); the rest starts empty
        for (int s = tid; s < S; s += THREADS) {
            oz[s] = s < nfixed ? pd.z[s] : 0ull;
            ow[s] = s < nfixed ? pd.w[s] : 0.0;
        }
        if (left0 <= 0 || inf.ncdf <= 0) continue;
        if (left0 > MAX_SAMPLED) {
            if (tid == 0) { if (atomicCAS(&p.status[0], 0, DKS_ERR_UNSUPPORTED) == 0) p.status[1] = M; }
            continue;
        }
        __syncthreads();
        for (int h = tid; h < TABLE_CAP; h += THREADS) { tkey[h] = 0ull; tfirst[h] = 0xFFFFFFFFu; tcount[h] = 0u; }
        if (tid == 0) s_carry = 0u;
        __syncthreads();

        const uint64_t fullmask = M >= 64 ? ~0ull : ((1ull << M) - 1ull);
        const uint64_t grow = (uint64_t)(p.row_offset + i);
        const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
        const uint32_t ndraws = 4u * (uint32_t)left0;            // upstream draws 4 * samples_left size picks at most

        for (uint32_t t0 = 0; t0 < ndraws; t0 += BATCH) {
            uint64_t mask[DRAWS_PER_THREAD];
            uint32_t slot[DRAWS_PER_THREAD];
            bool paired[DRAWS_PER_THREAD], valid[DRAWS_PER_THREAD];
            // ---- 1. generate this thread's draws (consecutive t) and register first occurrences
#pragma unroll
            for (int j = 0; j < DRAWS_PER_THREAD; ++j) {
                const uint32_t t = t0 + (uint32_t)tid * DRAWS_PER_THREAD + j;
                valid[j] = t < ndraws;
                mask[j] = 0ull; slot[j] = 0u; paired[j] = false;
                if (!valid[j]) continue;
                uint32_t rnd[4];
                philox4x32_10(k0, k1, t, (uint32_t)grow, (uint32_t)(grow >> 32), 0u, rnd);
                const double u = ((double)rnd[0] + 0.5) * 2.3283064365386963e-10;     // (r + 1/2) / 2^32
                int idx = 0;
                while (idx < inf.ncdf - 1 && u >= inf.cdf[idx]) ++idx;
                const int size = idx + inf.n_full + 1;
                paired[j] = size <= inf.n_paired;
                uint64_t avail = fullmask, mk = 0ull;
                int have = 1;                                     // rnd[1..3] are still unused
                uint32_t blockno = 0u;
                for (int c = 0; c < size; ++c) {
                    if (have == 4) { ++blockno; philox4x32_10(k0, k1, t, (uint32_t)grow, (uint32_t)(grow >> 32), blockno, rnd); have = 0; }
                    const uint32_t r32 = rnd[have++];
                    const int pick = (int)(((uint64_t)r32 * (uint64_t)(M - c)) >> 32);
                    const int bit = nth_set_bit(avail, pick);
                    mk |= 1ull << bit;
                    avail &= ~(1ull << bit);
                }
                mask[j] = mk;
                uint32_t h = hash_mask(mk) & (TABLE_CAP - 1);
                while (true) {
                    const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&tkey[h]), 0ull, (unsigned long long)mk);
                    if (old == 0ull || old == (unsigned long long)mk) break;
                    h = (h + 1) & (TABLE_CAP - 1);
                }
                slot[j] = h;
                atomicMin(&tfirst[h], t);
            }
            __syncthreads();
            // ---- 2. rows each draw would create, exclusive prefix over the batch in draw order
            uint32_t rows[DRAWS_PER_THREAD], local = 0u;
#pragma unroll
            for (int j = 0; j < DRAWS_PER_THREAD; ++j) {
                const uint32_t t = t0 + (uint32_t)tid * DRAWS_PER_THREAD + j;
                const bool fresh = valid[j] && tfirst[slot[j]] == t;
                rows[j] = fresh ? (paired[j] ? 2u : 1u) : 0u;
                local += rows[j];
            }
            uint32_t incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_warp[wib] = incl;
            __syncthreads();
            uint32_t warp_off = 0u, total = 0u;
#pragma unroll
            for (int wq = 0; wq < THREADS / 32; ++wq) { if (wq < wib) warp_off += s_warp[wq]; total += s_warp[wq]; }
            const uint32_t carry = s_carry;
            uint32_t before = carry + warp_off + incl - local;    // rows created by earlier draws
            // ---- 3. draws made while the budget was not yet full count; new masks among them take rows
#pragma unroll
            for (int j = 0; j < DRAWS_PER_THREAD; ++j) {
                if (valid[j] && before < (uint32_t)left0) {
                    atomicAdd(&tcount[slot[j]], 1u);
                    if (rows[j] > 0u) {
                        oz[nfixed + before] = mask[j];
                        rowslot[before] = slot[j];
                        if (rows[j] == 2u && before + 1u < (uint32_t)left0) {
                            oz[nfixed + before + 1u] = mask[j] ^ fullmask;
                            rowslot[before + 1u] = slot[j];
                        }
                    }
                }
                before += rows[j];
            }
            __syncthreads();
            if (tid == 0) s_carry = carry + total;
            __syncthreads();
            if (carry + total >= (uint32_t)left0) break;
        }
        // ---- 4. weights: multiplicity of the row's mask, rescaled to the mass left for the sampled sizes
        const uint32_t filled = min(s_carry, (uint32_t)left0);
        double part = 0.0;
        for (uint32_t r = tid; r < filled; r += THREADS) part += (double)tcount[rowslot[r]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_red[wib] = part;
        __syncthreads();
        double totalw = 0.0;
#pragma unroll
        for (int wq = 0; wq < THREADS / 32; ++wq) totalw += s_red[wq];
        const double scale = totalw > 0.0 ? inf.weight_left / totalw : 0.0;
        for (uint32_t r = tid; r < filled; r += THREADS) ow[nfixed + r] = (double)tcount[rowslot[r]] * scale;
        __syncthreads();
    }
}

inline size_t smem_bytes() { return (size_t)TABLE_CAP * (8 + 4 + 4) + (size_t)MAX_SAMPLED * 4; }

}  // namespace sampler
}  // namespace dks
