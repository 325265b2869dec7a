// Host-side emulation of the CUDA execution model for the SIMPLE kernels of the engine (no PTX, no tensor memory): one
// std::thread per CUDA thread of a block, a std::barrier for __syncthreads, `__shared__` arrays as function statics (blocks
// run one after the other).  Test infrastructure: lets the CPU suite execute the very kernel source that nvcc compiles
// (csrc/dks_wide.cuh) and check its indexing against a plain reference, here where there is no GPU.
#pragma once

#include <barrier>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() = default;
    dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace emu {
inline std::unique_ptr<std::barrier<>> block_barrier;
inline double shuffle_slots[1024];
}
inline thread_local dim3 threadIdx;
inline dim3 blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::block_barrier->arrive_and_wait(); }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

namespace dks {
// butterfly reduction of a warp's 32 values in the association order of the shfl.xor loop; must be called by every thread
// of the block (true for the kernels emulated here)
inline double warp_sum(double v) {
    const unsigned tid = threadIdx.x;
    emu::shuffle_slots[tid] = v;
    __syncthreads();
    double a[32], b[32];
    const unsigned base = tid & ~31u;
    for (int l = 0; l < 32; ++l) a[l] = base + l < blockDim.x ? emu::shuffle_slots[base + l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
        for (int l = 0; l < 32; ++l) b[l] = a[l] + a[l ^ o];
        std::memcpy(a, b, sizeof(a));
    }
    __syncthreads();
    return a[tid & 31];
}
}  // namespace dks

namespace emu {
template <typename Kernel, typename Params>
void launch(dim3 grid, dim3 block, Kernel kernel, const Params& p) {
    gridDim = grid;
    blockDim = block;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = dim3(bx, by, 0);
            block_barrier = std::make_unique<std::barrier<>>((std::ptrdiff_t)block.x);
            std::vector<std::thread> threads;
            threads.reserve(block.x);
            for (unsigned t = 0; t < block.x; ++t)
                threads.emplace_back([&, t] {
                    threadIdx = dim3(t, 0, 0);
                    kernel(p);
                });
            for (auto& th : threads) th.join();
        }
}
}  // namespace emu
