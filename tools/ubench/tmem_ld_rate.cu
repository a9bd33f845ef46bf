// tmem_ld_rate.cu -- microbenchmark: how fast can the epilogue warps of one CTA pull a 128-lane x
// 256-column fp32 accumulator (128 KB) out of tensor memory?  Variants: instruction shape
// (32x32b.x32 / 16x256b.x8), number of warps (4 / 8 / 16), wait::ld after every load or after a pair.
// Decides whether the per-layer drain of the fused MRF kernel (about 2.4k cycles) is a hardware floor
// (DESIGN.md 5.4).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define R32(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), \
    "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),         \
    "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define L32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"

__device__ __forceinline__ void ld_32x32b_x32(uint32_t a, uint32_t (&r)[32]) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 : R32(r) : "r"(a)); }
__device__ __forceinline__ void ld_16x256b_x8(uint32_t a, uint32_t (&r)[32]) { asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " L32 : R32(r) : "r"(a)); }
#define S32(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), \
    "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),         \
    "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
#define T32 "[%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
__device__ __forceinline__ void st_32x32b_x32(uint32_t a, const uint32_t (&r)[32]) { asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 " T32 ::"r"(a), S32(r) : "memory"); }
__device__ __forceinline__ void st_16x256b_x8(uint32_t a, const uint32_t (&r)[32]) { asm volatile("tcgen05.st.sync.aligned.16x256b.x8.b32 " T32 ::"r"(a), S32(r) : "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// SHAPE 0: 32x32b.x32 (warp = lane quarter x 32 columns per load), 1: 16x256b.x8 (16 lanes x 64 columns per load)
template <int SHAPE, int PAIR>
__global__ void __launch_bounds__(512) bench(int nwarps, int iters, long long *out, uint32_t *sink)
{
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    __syncthreads();
    if (warp < nwarps) {
        const int quarter = warp & 3;
        const int cshare = 256 / (nwarps / 4);                   // columns per warp
        const int c0 = (warp >> 2) * cshare;
        const uint32_t lane_base = tm + ((uint32_t)(quarter * 32) << 16);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // one pass = this warp's share of a 128 x 256 accumulator
            if (SHAPE == 0) {
                for (int c = 0; c < cshare; c += 32 * (PAIR ? 2 : 1)) {
                    uint32_t r[32], q[32];
                    ld_32x32b_x32(lane_base + (uint32_t)(c0 + c), r);
                    if (PAIR) ld_32x32b_x32(lane_base + (uint32_t)(c0 + c + 32), q);
                    wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc ^= r[i];
                    if (PAIR) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc ^= q[i];
                    }
                }
            } else {
                for (int c = 0; c < cshare; c += 64) {
                    uint32_t r[32], q[32];
                    ld_16x256b_x8(lane_base + (uint32_t)(c0 + c), r);
                    if (PAIR) ld_16x256b_x8(lane_base + (16u << 16) + (uint32_t)(c0 + c), q);
                    wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc ^= r[i];
                    if (!PAIR) {
                        ld_16x256b_x8(lane_base + (16u << 16) + (uint32_t)(c0 + c), q);
                        wait_ld();
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc ^= q[i];
                }
            }
        }
        t1 = clock64();
    }
    if (acc == 0x12345678u) sink[tid] = acc;
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// stores: SHAPE as above; one wait::st per pass
template <int SHAPE>
__global__ void __launch_bounds__(512) bench_st(int nwarps, int iters, long long *out)
{
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = tid * 32 + i;
    __syncthreads();
    if (warp < nwarps) {
        const int quarter = warp & 3;
        const int cshare = 256 / (nwarps / 4);
        const int c0 = (warp >> 2) * cshare;
        const uint32_t lane_base = tm + ((uint32_t)(quarter * 32) << 16);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (SHAPE == 0) {
                for (int c = 0; c < cshare; c += 32) st_32x32b_x32(lane_base + (uint32_t)(c0 + c), r);
            } else {
                for (int c = 0; c < cshare; c += 64) {
                    st_16x256b_x8(lane_base + (uint32_t)(c0 + c), r);
                    st_16x256b_x8(lane_base + (16u << 16) + (uint32_t)(c0 + c), r);
                }
            }
            wait_st();
        }
        t1 = clock64();
    }
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int SHAPE>
static void run_st(const char *name, int nwarps)
{
    long long *d_out, h[148];
    cudaMalloc(&d_out, sizeof(h));
    const int iters = 200;
    bench_st<SHAPE><<<148, 512>>>(nwarps, 20, d_out);
    bench_st<SHAPE><<<148, 512>>>(nwarps, iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / iters;
    printf("STORE %-12s warps=%2d  cycles per 128x256 fp32 accumulator = %8.1f  (%.1f B/clk/SM)\n", name, nwarps, cyc, 131072.0 / cyc);
    cudaFree(d_out);
}

template <int SHAPE, int PAIR>
static void run(const char *name, int nwarps)
{
    long long *d_out, h[148];
    uint32_t *sink;
    cudaMalloc(&d_out, sizeof(h));
    cudaMalloc(&sink, 512 * 4);
    const int iters = 200;
    bench<SHAPE, PAIR><<<148, 512>>>(nwarps, 20, d_out, sink);
    bench<SHAPE, PAIR><<<148, 512>>>(nwarps, iters, d_out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / iters;
    printf("%-16s wait_after_%s warps=%2d  cycles per 128x256 fp32 accumulator = %8.1f  (%.1f B/clk/SM)\n", name, PAIR ? "pair" : "each", nwarps, cyc,
           131072.0 / cyc);
    cudaFree(d_out);
    cudaFree(sink);
}

int main()
{
    for (int nw : {4, 8, 16}) {
        run<0, 0>("32x32b.x32", nw);
        run<0, 1>("32x32b.x32", nw);
        run<1, 0>("16x256b.x8", nw);
        run<1, 1>("16x256b.x8", nw);
    }
    for (int nw : {4, 8, 16}) {
        run_st<0>("32x32b.x32", nw);
        run_st<1>("16x256b.x8", nw);
    }
    return 0;
}
