// Walk-step micro-benchmark (measurement tooling, not product code):  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o walk_mb scripts/walk_microbench.cu
// An idealised tree walk with nothing around it: every lane runs N chains of [feature LDS -> compare -> child select -> node LDG]
// through complete synthetic trees (all lanes of a warp inside the same tree), 1024 lanes per SM.  It answers two questions:
//   * what bounds a walk step -- variants without the feature LDS / without the compare / with 4..16 chains;
//   * whether 4-byte (rank-coded) nodes would beat the 8-byte slots -- they do not: the step rate is the same.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
constexpr int kRows = 16;
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void step8(uint2 &n, uint32_t fcol, uint32_t wl, uint32_t wh) {
    const float fv = lds_f32(fcol + (n.y >> 20));
    uint32_t a = (n.y & 0xFFFF8u) | wl;
    asm("{.reg .pred p; setp.gtu.f32 p, %1, %2; @p add.u32 %0, %0, 8;}" : "+r"(a) : "f"(fv), "f"(__uint_as_float(n.x)));
    asm("{.reg .u64 ad; mov.b64 ad, {%2, %3}; ld.global.nc.v2.u32 {%0, %1}, [ad];}" : "=r"(n.x), "=r"(n.y) : "r"(a), "r"(wh));
}
// VAR 3: rank-coded features in ONE register (DESIGN.md "what is left"): no feature LDS; the node carries a rotate
// amount (top 5 bits of hi32) and a top-aligned integer threshold (lo32); fcol stands in for the lane's rank word
template <int VAR>
__device__ __forceinline__ void step8v(uint2 &n, uint32_t fcol, uint32_t wl, uint32_t wh) {
    if (VAR == 3 || VAR == 4) {      // VAR 4: the rotate amount in the LOW 5 bits of hi32 (shf.wrap reads them directly: no extract)
        const uint32_t v = __funnelshift_l(fcol, fcol, VAR == 3 ? (n.y >> 27) : n.y);
        uint32_t a = (n.y & 0xFFFF8u) | wl;
        asm("{.reg .pred p; setp.gt.u32 p, %1, %2; @p add.u32 %0, %0, 8;}" : "+r"(a) : "r"(v), "r"(n.x));
        asm("{.reg .u64 ad; mov.b64 ad, {%2, %3}; ld.global.nc.v2.u32 {%0, %1}, [ad];}" : "=r"(n.x), "=r"(n.y) : "r"(a), "r"(wh));
        return;
    }
    float fv;
    if (VAR == 1) fv = __uint_as_float(fcol); else fv = lds_f32(fcol + (n.y >> 20));
    uint32_t a = (n.y & 0xFFFF8u) | wl;
    if (VAR != 2) asm("{.reg .pred p; setp.gtu.f32 p, %1, %2; @p add.u32 %0, %0, 8;}" : "+r"(a) : "f"(fv), "f"(__uint_as_float(n.x)));
    else a += (__float_as_uint(fv) & 8u);
    asm("{.reg .u64 ad; mov.b64 ad, {%2, %3}; ld.global.nc.v2.u32 {%0, %1}, [ad];}" : "=r"(n.x), "=r"(n.y) : "r"(a), "r"(wh));
}
template <int VAR, int CH>
__global__ void __launch_bounds__(1024, 1) kv(const void *tbl, int n_trees, int depth, int iters, unsigned *sink, int tree_stride) {
    extern __shared__ __align__(4096) uint32_t rows[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int r = 0; r < kRows; ++r) {
        uint32_t h = (tid * 2654435761u) ^ (r * 40503u) ^ (blockIdx.x * 97u);
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        rows[(warp * kRows + r) * 32 + lane] = __float_as_uint((float)(h & 1023));
    }
    __syncthreads();
    const uint32_t fcol = (uint32_t)__cvta_generic_to_shared(rows + warp * kRows * 32 + lane);
    const uint64_t base = (uint64_t)tbl;
    const uint32_t wl = (uint32_t)base, wh = (uint32_t)(base >> 32);
    uint32_t acc = 0;
    int t = (blockIdx.x * 32 + warp) * 8 % n_trees;
    for (int it = 0; it < iters; ++it) {
        uint2 n[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) n[c] = __ldg(reinterpret_cast<const uint2 *>(tbl) + (size_t)((t + c) % n_trees) * tree_stride);
        for (int d = 0; d < depth; ++d) {
#pragma unroll
            for (int c = 0; c < CH; ++c) step8v<VAR>(n[c], fcol, wl, wh);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) acc ^= n[c].x;
        t = (t + CH) % n_trees;
    }
    if (acc == 0x1234567u) atomicAdd(sink, 1u);
}
__device__ __forceinline__ void step4(uint32_t &n, uint32_t fcol, uint32_t wl, uint32_t wh) {
    const uint32_t key = lds_u32(((n >> 21) & 0x780u) | fcol);
    uint32_t a = (n & 0xFFFCu) | wl;
    asm("{.reg .pred p; setp.gt.u32 p, %1, %2; @p add.u32 %0, %0, 4;}" : "+r"(a) : "r"(key), "r"(n));
    asm("{.reg .u64 ad; mov.b64 ad, {%1, %2}; ld.global.nc.u32 %0, [ad];}" : "=r"(n) : "r"(a), "r"(wh));
}
// trees: depth D complete trees; tree t occupies nodes [t*(2^(D+1)) ...]; children of node i (heap order within tree, BFS) adjacent
// COHERENCE experiment (round 2): production lanes of a warp are not random -- requests of one (family, orientation)
// share most of their feature values (down 1..4, flags, similar clocks), so the lanes of a warp sit on the same few
// nodes of a tree level.  g_coh = percentage of lanes whose feature values are the warp's common ones.
__device__ int g_coh = 0;
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const void *tbl, int n_trees, int depth, int iters, unsigned *sink, int tree_stride) {
    extern __shared__ __align__(4096) uint32_t rows[];   // 32 warps x 16 rows x 32 lanes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int r = 0; r < kRows; ++r) {
        uint32_t h = (tid * 2654435761u) ^ (r * 40503u) ^ (blockIdx.x * 97u);
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        uint32_t hw = ((uint32_t)(warp + 1) * 2654435761u) ^ (r * 40503u) ^ (blockIdx.x * 97u);     // the warp's common value
        hw ^= hw >> 15; hw *= 2246822519u; hw ^= hw >> 13;
        uint32_t hs = (tid * 374761393u) ^ (blockIdx.x * 668265263u);                               // does this lane follow it?
        hs ^= hs >> 13; hs *= 1274126177u; hs ^= hs >> 16;
        if ((int)(hs % 100u) < g_coh) h = hw;
        if (MODE == 8) rows[(warp * kRows + r) * 32 + lane] = __float_as_uint((float)(h & 1023));
        else rows[(warp * kRows + r) * 32 + lane] = ((uint32_t)r << 28) | ((h & 1023) << 16);
    }
    __syncthreads();
    const uint32_t fcol = (uint32_t)__cvta_generic_to_shared(rows + warp * kRows * 32 + lane);
    const uint64_t base = (uint64_t)tbl;
    const uint32_t wl = (uint32_t)base, wh = (uint32_t)(base >> 32);
    uint32_t acc = 0;
    int t = (blockIdx.x * 32 + warp) * 8 % n_trees;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 8) {
            uint2 n[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) n[c] = __ldg(reinterpret_cast<const uint2 *>(tbl) + (size_t)((t + c) % n_trees) * tree_stride);
            for (int d = 0; d < depth; ++d) {
#pragma unroll
                for (int c = 0; c < 8; ++c) step8(n[c], fcol, wl, wh);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) acc ^= n[c].x;
        } else {
            uint32_t n[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) n[c] = __ldg(reinterpret_cast<const uint32_t *>(tbl) + (size_t)((t + c) % n_trees) * tree_stride);
            for (int d = 0; d < depth; ++d) {
#pragma unroll
                for (int c = 0; c < 8; ++c) step4(n[c], fcol, wl, wh);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) acc ^= n[c];
        }
        t = (t + 8) % n_trees;
    }
    if (acc == 0x1234567u) atomicAdd(sink, 1u);
}
int main() {
    const int D = 3, n_trees = 1200, stride = 16;   // 16 slots per tree (15 used), complete depth-3 trees + leaf level
    // 8-byte table
    std::vector<uint64_t> t8((size_t)n_trees * stride);
    std::vector<uint32_t> t4((size_t)n_trees * stride);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
    for (int t = 0; t < n_trees; ++t) {
        // heap nodes 0..14: node i children 2i+1, 2i+2 (adjacent), leaves 7..14 self-loop
        for (int i = 0; i < stride; ++i) {
            const uint32_t row = rnd() % 11, thr = rnd() % 1024;
            uint32_t child = (i < 7) ? (2 * i + 1) : i;          // leaves point at themselves (never go right: thr huge)
            const uint32_t slot8 = (uint32_t)(t * stride + child) * 8u, slot4 = (uint32_t)(t * stride + child) * 4u;
            float thrf = (i < 7) ? (float)thr : 1e30f;
            uint32_t lo; memcpy(&lo, &thrf, 4);
            t8[(size_t)t * stride + i] = ((uint64_t)(((row * 128u) << 20) | slot8) << 32) | lo;
            const uint32_t j = (i < 7) ? thr : 4095u;
            t4[(size_t)t * stride + i] = (row << 28) | (j << 16) | (slot4 & 0xFFFFu);
        }
    }
    // note: 4-byte table 1200*16*4 = 76.8 KB > 64 KB window: limit trees so offsets fit 16 bits
    const int n4 = 1000;
    void *d8, *d4; unsigned *sink;
    cudaMalloc(&d8, (1 << 20) * 2); cudaMalloc(&d4, (1 << 20) * 2); cudaMalloc(&sink, 4);
    char *a8 = (char *)(((uintptr_t)d8 + (1 << 20) - 1) >> 20 << 20), *a4 = (char *)(((uintptr_t)d4 + (1 << 20) - 1) >> 20 << 20);
    cudaMemcpy(a8, t8.data(), t8.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(a4, t4.data(), t4.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int coh : {0, 75, 95, 100}) {
        cudaMemcpyToSymbol(g_coh, &coh, sizeof(int));
        for (int depth : {3, 6})
            for (int mode : {8, 4, 8, 4}) {
                const int iters = 2000;
                cudaEventRecord(e0);
                if (mode == 8) k<8><<<148, 1024, 65536>>>(a8, n_trees, depth, iters, sink, stride);
                else k<4><<<148, 1024, 65536>>>(a4, n4, depth, iters, sink, stride);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("coherent lanes %3d%% depth %d mode %dB: %.2f ms  %.3e lane-steps/s  err=%s\n", coh, depth, mode, ms,
                       148.0 * 1024 * iters * 8.0 * depth / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
            }
    }
    { const int zero = 0; cudaMemcpyToSymbol(g_coh, &zero, sizeof(int)); }
    for (int depth : {3, 6}) {
        for (int mode : {8, 4, 8, 4}) {
            const int iters = 2000;
            cudaEventRecord(e0);
            if (mode == 8) k<8><<<148, 1024, 65536>>>(a8, n_trees, depth, iters, sink, stride);
            else k<4><<<148, 1024, 65536>>>(a4, n4, depth, iters, sink, stride);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double steps = 148.0 * 1024 * iters * 8.0 * depth;
            printf("depth %d mode %dB: %.2f ms  %.3e lane-steps/s  err=%s\n", depth, mode, ms, steps / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        }
    }
    {
        auto run = [&](const char *name, auto kern, int ch, int depth) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            const int iters = 2000 * 8 / ch;
            kern<<<148, 1024, 65536>>>(a8, n_trees, depth, iters, sink, stride);
            cudaEventRecord(e0);
            kern<<<148, 1024, 65536>>>(a8, n_trees, depth, iters, sink, stride);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%-28s depth %d: %.2f ms %.3e lane-steps/s %s\n", name, depth, ms, 148.0 * 1024 * iters * ch * depth / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        };
        for (int depth : {3, 6}) {
            run("8B normal 8 chains", kv<0, 8>, 8, depth);
            run("8B no LDS 8 chains", kv<1, 8>, 8, depth);
            run("8B no setp 8 chains", kv<2, 8>, 8, depth);
            run("8B rank register 8 chains", kv<3, 8>, 8, depth);
            run("8B rank register 12 chains", kv<3, 12>, 12, depth);
            run("8B rank reg, no extract 8 ch", kv<4, 8>, 8, depth);
            run("8B normal 4 chains", kv<0, 4>, 4, depth);
            run("8B normal 12 chains", kv<0, 12>, 12, depth);
            run("8B normal 16 chains", kv<0, 16>, 16, depth);
        }
    }
    return 0;
}
