// umma_rate.cu -- microbenchmark: cycles per tcgen05.mma (kind::f16, fp16 x fp16 -> fp32) as a
// function of M, N and the A-operand source (shared memory vs tensor memory), one issuing
// thread per CTA, one CTA per SM.  Decides the tile orientation of the MRF kernels.
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: A from smem, 1: A from tmem
template <int MODE>
__global__ void __launch_bounds__(128) bench(int M, int N, int iters, int a_rows, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // fp16 1.0
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
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_base = smem_u32(smem);              // A: up to a_rows rows, 2 K-groups
        const uint32_t lbo_a = (uint32_t)a_rows * 16u;
        const uint32_t b_base = a_base + 24 * 1024;          // B: N rows, 2 K-groups, several taps
        const uint32_t lbo_b = (uint32_t)N * 16u;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t shift = (uint32_t)(it & 7) * 16u;       // emulate tap shifts
            const uint32_t dcol = tm + (uint32_t)((it >> 4) & 1) * (uint32_t)N;
            const uint64_t bdesc = make_desc(b_base + ((it & 3) * 2u) * lbo_b % 8192u, lbo_b, 128u);
            if (MODE == 0) {
                const uint64_t adesc = make_desc(a_base + shift, lbo_a, 128u);
                mma_ss(dcol, adesc, bdesc, idesc, (it & 15) ? 1u : 0u);
            } else {
                mma_ts(dcol, tm + 256u + (uint32_t)(it & 7) * 8u, bdesc, idesc, (it & 15) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main()
{
    long long *d_out, h;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4096;
    const int Ms[] = {128, 64};
    const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
    for (int mode = 0; mode < 2; ++mode)
        for (int M : Ms)
            for (int N : Ns) {
                if (M == 128 && N % 16) continue;
                for (int grid : {1, 148}) {
                    for (int rep = 0; rep < 2; ++rep) {
                        if (mode == 0) bench<0><<<grid, 128, 64 * 1024>>>(M, N, iters, 160, d_out);
                        else bench<1><<<grid, 128, 64 * 1024>>>(M, N, iters, 160, d_out);
                    }
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
                    const double cyc = (double)h / iters;
                    const double ideal = (M == 128 ? 128.0 : 128.0) * N / 256.0;   // doc floor
                    printf("A=%s M=%3d N=%3d grid=%3d  cycles/mma=%7.2f  doc_floor=%6.1f  MAC/clk=%7.1f  util=%5.1f%%\n", mode ? "tmem" : "smem", M, N, grid,
                           cyc, ideal, (double)M * N * 16 / cyc, 100.0 * (double)M * N * 16 / cyc / 4096.0);
                }
            }
    return 0;
}
