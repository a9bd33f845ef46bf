// mrf_fused_host.h -- host-side preparation for the fused MRF residual-block kernel:
// weight packing (tap-reversed, zero-padded, per 16-channel K-step chunks), epilogue scatter
// tables and the launch plan (how many conv pairs are chained per launch).  Plain C++ (no
// CUDA types: fp16 values travel as uint16_t) so that tests/cpu/mrf_fused_emul.cpp can run
// exactly this code against a direct convolution.
#pragma once

#include <stdint.h>

#include <vector>

#include "mrf_fused.cuh"

namespace zvx {
namespace mrf {

// raw: (OC = CH, IC = CH, K = k) fp16, K fastest (ggml ne [K, IC, OC]).
// packed: [K-step c][group g2 of 2][tap block tb][oc row][8 ic rows]; tap block tb holds tap
// t = (k - 1 + S - 1) - tb, zeros when t is outside [0, k).  Output rows and K slots are in the
// kernel's row order (row r <-> channel row_to_chan(r), mrf_fused.cuh).
inline std::vector<uint16_t> pack_weights(const uint16_t *raw, int CH, int k)
{
    const int S = 128 / CH, TB = tap_blocks(k, S);
    std::vector<uint16_t> pk((size_t)(CH / 16) * 2 * TB * CH * 8, 0);
    size_t o = 0;
    for (int c = 0; c < CH / 16; ++c)
        for (int g2 = 0; g2 < 2; ++g2)
            for (int tb = 0; tb < TB; ++tb) {
                const int t = (k - 1 + S - 1) - tb;
                for (int oc = 0; oc < CH; ++oc)
                    for (int e = 0; e < 8; ++e, ++o)
                        if (t >= 0 && t < k) pk[o] = raw[((size_t)row_to_chan(oc) * CH + row_to_chan(c * 16 + g2 * 8 + e)) * k + t];
            }
    return pk;
}

// per-channel vector (bias) in the kernel's row order
inline std::vector<float> to_row_order(const std::vector<float> &v)
{
    std::vector<float> r(v.size());
    for (size_t i = 0; i < v.size(); ++i) r[i] = v[(size_t)row_to_chan((int)i)];
    return r;
}

// Scatter table from a layer whose output positions are laid out for dilation d_cur into a
// buffer laid out for dilation d_next: entry [s][n'] for output position p = S n' + s.
inline std::vector<uint32_t> make_table(int CH, int ncol, int d_cur, int d_next)
{
    const int S = 128 / CH, groups = CH / 8;
    const int Wp = wp_of(CH, ncol);
    std::vector<uint32_t> t((size_t)S * ncol, tbl_trash(ncol));
    for (int s = 0; s < S; ++s)
        for (int n = 0; n < ncol; ++n) {
            const int p = S * n + s;
            if (p >= Wp) continue;
            const int tau = pos_to_tau(p, d_cur, Wp);
            t[(size_t)s * ncol + n] = tbl_pack(dest_unit(tau, d_next, Wp, S, groups, ncol), tau);
        }
    return t;
}

// One launch = a chain of conv pairs [p0, p1) of a residual block.
struct ChainPlan { int p0, p1, halo, valid; };

// Greedy split of the block's P pairs so that every launch keeps at least `min_eff` of its
// window as valid output (halo recompute is the price of fusing; a split costs one fp32 round
// trip of y through HBM).
inline std::vector<ChainPlan> plan_chains(int CH, int ncol, int k, const int *dil, int P, double min_eff)
{
    const int S = 128 / CH;
    const int Wp = wp_of(CH, ncol);
    std::vector<ChainPlan> out;
    int p = 0;
    while (p < P) {
        int halo = 0, q = p;
        while (q < P) {
            const int h = (k - 1) / 2 * dil[q] + (k - 1) / 2;
            if (q > p && (double)(Wp - 2 * (halo + h)) / (S * ncol) < min_eff) break;
            halo += h;
            ++q;
        }
        out.push_back({p, q, halo, Wp - 2 * halo});
        p = q;
    }
    return out;
}

}  // namespace mrf
}  // namespace zvx
