// store_rate.cu -- how fast can one SM push a 128 KB tile to global memory while all 148 SMs do the same?
//   (a) STG.128 from 512 threads, a warp instruction = 4 rows x one full 128-byte line (the fused MRF kernel's final phase)
//   (b) cp.async.bulk shared -> global (TMA bulk store) of the same bytes, issued by one thread in 16 KB pieces
// Prints bytes per clock per SM and the aggregate rate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, int tiles_per_cta, long long *cyc)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 1024 / 16; i += 512) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(i, i, i, i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < tiles_per_cta; ++t) {
        float *tile = out + ((size_t)(blockIdx.x * tiles_per_cta + t)) * (32 * 1024);     // 128 KB = 1024 rows x 32 floats
        if (MODE == 0) {
            // thread: 4 adjacent channels (16 B) of rows r; warp w covers 64 "columns" like the kernel: 16 stores per thread
            const float4 v = make_float4((float)tid, 1.f, 2.f, 3.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int row = 4 * ((warp >> 2) * 64 + 8 * (i >> 1) + 2 * (lane & 3) + (i & 1)) + (warp & 3);
                *reinterpret_cast<float4 *>(tile + (size_t)row * 32 + 4 * (lane >> 2)) = v;
            }
        } else {
            if (tid == 0) {
                for (int c = 0; c < 8; ++c)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(tile + c * 4096), "r"(smem_u32(smem + c * 16384)), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
        }
    }
    if (MODE == 1 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    const int tiles = 64;
    float *out; long long *cyc, h;
    cudaMalloc(&out, (size_t)148 * tiles * 128 * 1024);
    cudaMalloc(&cyc, 8);
    for (int grid : {148, 37, 8, 1})
    for (int mode = 0; mode < 2; ++mode) {
        auto fn = mode ? k<1> : k<0>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            fn<<<grid, 512, 128 * 1024>>>(out, tiles, cyc);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)tiles * 128 * 1024;
        printf("grid %3d %s: %.1f B/clk/SM (CTA 0: %lld cycles for %d tiles of 128 KB), aggregate %.2f TB/s\n", grid, mode ? "cp.async.bulk smem->global" : "STG.128 (final-phase pattern)",
               bytes / (double)h, h, tiles, (double)grid * bytes / (ms * 1e-3) / 1e12);
    }
    return 0;
}
