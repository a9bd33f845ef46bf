// umma_issue.cu -- how cheaply can one thread issue tcgen05.mma?  Variants of the issue loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
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

// VAR 0: if (tid == 0) loop, descriptors rebuilt per MMA with adds
// VAR 1: warp-convergent loop, elect.sync leader, descriptor lo-word incremented by constants, unroll 8
template <int VAR, int N, int TAPS>
__global__ void __launch_bounds__(128) bench(int iters, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
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
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem);
    constexpr uint32_t A_ROWS = 256 + 64;
    constexpr uint32_t lbo_a = A_ROWS * 16u;          // A: 2 K-groups of A_ROWS rows
    const uint32_t b_base = a_base + 16 * 1024;       // B: TAPS blocks of [2][N][8] halfs
    constexpr uint32_t lbo_b = N * 16u;
    long long t0 = 0, t1 = 0;
    if (tid < 32) {
        const uint32_t leader = elect_one();
        t0 = clock64();
        if (VAR == 0) {
            if (tid == 0) {
                for (int it = 0; it < iters; ++it) {
                    for (int a = 0; a < TAPS; ++a) {
                        const uint64_t adesc = make_desc(a_base + a * 16u, lbo_a, 128u);
                        const uint64_t bdesc = make_desc(b_base + a * (2u * lbo_b), lbo_b, 128u);
                        mma_ss(tm, adesc, bdesc, idesc, (it | a) ? 1u : 0u);
                    }
                }
            }
        } else {
            const uint64_t adesc0 = make_desc(a_base, lbo_a, 128u);
            const uint64_t bdesc0 = make_desc(b_base, lbo_b, 128u);
            for (int it = 0; it < iters; ++it) {
                if (leader) {
#pragma unroll
                    for (int a = 0; a < TAPS; ++a) {
                        mma_ss(tm, adesc0 + (uint64_t)a, bdesc0 + (uint64_t)(a * ((2u * lbo_b) >> 4)), idesc, (it | a) ? 1u : 0u);
                    }
                }
                __syncwarp();
            }
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

template <int VAR, int N, int TAPS>
void run(long long *d_out)
{
    long long h;
    cudaFuncSetAttribute(bench<VAR, N, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int iters = 512;
    for (int rep = 0; rep < 2; ++rep) bench<VAR, N, TAPS><<<148, 128, 96 * 1024>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    const double cyc = (double)h / (iters * TAPS);
    printf("var=%d N=%3d taps=%2d cycles/mma=%7.2f floor=%5.1f util=%5.1f%%\n", VAR, N, TAPS, cyc, N / 2.0, 100.0 * (N / 2.0) / cyc);
}

int main()
{
    long long *d_out;
    cudaMalloc(&d_out, 8);
    run<0, 32, 8>(d_out); run<1, 32, 8>(d_out);
    run<0, 64, 8>(d_out); run<1, 64, 8>(d_out);
    run<1, 16, 8>(d_out);
    run<1, 32, 16>(d_out);
    run<1, 128, 8>(d_out); run<1, 256, 8>(d_out);
    return 0;
}
