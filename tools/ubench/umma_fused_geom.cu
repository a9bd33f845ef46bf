// umma_fused_geom.cu -- tensor-pipe rate of the fused MRF kernel's exact operand geometry, nothing else running:
// A = 128-row window of the tap-reversed weight array (K-major, no swizzle, LBO_A = tap_blocks * CH * 16),
// B = polyphase activation sub-buffers (rows 16 B apart, LBO_B = (NCOL + 17) * 16), N = NCOL, one K-step of
// k + S - 1 shifted steps per "chunk".  Prints cycles per MMA for (CH, k) and for row-aligned B starts.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred;
}

template <int CH, int K, int NCOL, int ALIGNED>
__global__ void __launch_bounds__(128) bench(int iters, long long *out)
{
    constexpr int S = 128 / CH, GROUPS = CH / 8, NROWS = NCOL + 17, LBO_B = NROWS * 16, SUB = GROUPS * LBO_B;
    constexpr int TAPB = K + 2 * S - 2, LBO_A = TAPB * CH * 16, NJ = K + S - 1, C = (K - 1) / 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x2c002c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem);                 // one weight chunk: 2 groups x TAPB blocks x CH rows x 16 B
    const uint32_t b_base = a_base + 48 * 1024;             // activation buffer: S sub-buffers
    constexpr uint64_t HI = ((uint64_t)(128u >> 4) | ((uint64_t)1 << 14)) << 32;
    long long t0 = 0, t1 = 0;
    if (tid < 32) {
        const uint32_t leader = elect_one();
        int q0 = ((-C) % S + S) % S, ro0 = (-C - q0) / S;
        const uint32_t a_fix = ((LBO_A >> 4) & 0x3FFFu) << 16, b_fix = ((LBO_B >> 4) & 0x3FFFu) << 16;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
                uint32_t a_lo = (((a_base + (uint32_t)(K + S - 2) * (CH * 16u)) & 0x3FFFFu) >> 4) | a_fix;
                uint32_t b_lo = (((b_base + (uint32_t)q0 * SUB + (uint32_t)(ro0 * 16) + 128u) & 0x3FFFFu) >> 4) | b_fix;
                int q = q0;
#pragma unroll 1
                for (int j = 0; j < NJ; ++j) {
                    mma_ss(tm, HI | a_lo, HI | (ALIGNED ? (b_lo & ~7u) : b_lo), idesc, (it | j) ? 1u : 0u);
                    a_lo -= (uint32_t)CH;
                    b_lo += (uint32_t)(SUB >> 4);
                    if (++q == S) { q = 0; b_lo -= (uint32_t)(S * (SUB >> 4) - 1); }
                }
            }
            __syncwarp();
        }
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        t1 = clock64();
        if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int CH, int K, int NCOL, int ALIGNED>
void run(long long *d_out)
{
    long long h;
    cudaFuncSetAttribute(bench<CH, K, NCOL, ALIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 256;
    for (int rep = 0; rep < 2; ++rep) bench<CH, K, NCOL, ALIGNED><<<148, 128, 200 * 1024>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("CH=%3d k=%2d NCOL=%3d %s: cycles/mma=%7.2f (ideal %d)\n", CH, K, NCOL, ALIGNED ? "B rows 128-B aligned" : "B rows as in the kernel", (double)h / (iters * (K + 128 / CH - 1)), NCOL / 2);
}

int main()
{
    long long *d_out;
    cudaMalloc(&d_out, 8);
    run<32, 11, 256, 0>(d_out); run<32, 11, 256, 1>(d_out);
    run<32, 3, 256, 0>(d_out);  run<32, 3, 128, 0>(d_out); run<32, 3, 128, 1>(d_out);
    run<64, 7, 256, 0>(d_out);  run<64, 7, 256, 1>(d_out);
    run<128, 3, 256, 0>(d_out); run<128, 11, 256, 0>(d_out); run<128, 11, 256, 1>(d_out);
    return 0;
}
