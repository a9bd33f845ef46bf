// conv_ref.cu -- validation convolution on plain CUDA cores.
//
// NOT the product path: it exists so that GPU tests can check the tcgen05 implicit-GEMM
// kernel (conv_umma.cu) layer by layer on the device, and it is only reachable through the
// explicit debug switch of the C-ABI (zvx_set_debug_kernels).  It consumes the ORIGINAL
// (OC, IC, K) fp16 weight tensor, so it also validates the weight packing.  Arithmetic
// per element is the reference's: fp16-rounded input x fp16 weight, fp32 accumulate
// (/root/reference/ggml/src/ggml.c:3769-3786; ggml-cpu/ggml-cpu.c:9952, :1463-1503).
#include "zvx_common.cuh"
#include "zvx_internal.h"

namespace zvx {

__global__ void __launch_bounds__(128) conv_ref_kernel(const ConvParams p)
{
    const int tile    = blockIdx.x;
    const int u       = find_segment(p.tile_start, p.B, tile);
    const int t0      = (tile - __ldg(p.tile_start + u)) * 128;
    const int seg_f0  = __ldg(p.seg_start + u);
    const int seg_len = (__ldg(p.seg_start + u + 1) - seg_f0) * p.rate_in;
    const size_t seg_row0 = (size_t)seg_f0 * p.rate_in;
    const int oc = blockIdx.y;
    const int t  = t0 + threadIdx.x;
    if (t >= seg_len) return;

    float acc = 0.f;
    for (int a = 0; a < p.ntaps; ++a) {
        const int t_in = t + p.tap_off0 + a * p.tap_step;
        if (t_in < 0 || t_in >= seg_len) continue;
        const size_t e0 = (seg_row0 + (size_t)t_in) * (size_t)p.ldx + p.x_ch_off;
        const int wt = p.w_tap0 + a * p.w_tap_stride;
        for (int ic = 0; ic < p.Cin; ++ic) {
            float xq;
            if (p.pro_mode == PRO_F16) {
                xq = __half2float(reinterpret_cast<const __half *>(p.x)[e0 + ic]);
            } else {
                ProCh pc = {0.f, 1.f, 1.f, 0.f};
                if (p.pro_mode == PRO_NORM || p.pro_mode == PRO_MEL) {
                    pc.mu   = p.p_mu[(size_t)u * p.p_stat_stride + ic];
                    pc.rstd = p.p_rstd[(size_t)u * p.p_stat_stride + ic];
                }
                if (p.pro_mode == PRO_NORM) {
                    pc.g = p.p_g[(size_t)u * p.p_gb_stride + ic];
                    pc.b = p.p_b[(size_t)u * p.p_gb_stride + ic];
                }
                float x = reinterpret_cast<const float *>(p.x)[e0 + ic];
                if (p.pro_mode == PRO_SUM3) {
                    x = __fmul_rn(__fadd_rn(__fadd_rn(x, p.x2[e0 + ic]), p.x3[e0 + ic]), p.sum_scale);
                    xq = __half2float(__float2half_rn(lrelu_f(x, p.pro_slope)));
                } else
                xq = __half2float(__float2half_rn(prologue_apply(p.pro_mode, x, p.pro_slope, pc)));
            }
            const float w = __half2float(p.w_raw[((size_t)oc * p.Cin + ic) * p.w_taps_total + wt]);
            acc = fmaf(xq, w, acc);
        }
    }
    const size_t orow = (seg_row0 + (size_t)t) * (size_t)p.out_mul + p.out_add;
    epilogue_store(p, acc, orow, oc);
}

cudaError_t conv_ref_launch(const ConvParams &p, int total_tiles, cudaStream_t st)
{
    dim3 grid(total_tiles, p.Cout, 1);
    conv_ref_kernel<<<grid, 128, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace zvx
