// zvx_api.cu -- context, weight upload/repack, layer schedule and the C ABI (include/zvx.h).
//
// The schedule below is the B200 restatement of the two reference graphs:
//   StyleTTSDecoder   /root/reference/src/stylettsdec.cpp:371-441
//   HiFiGAN           /root/reference/src/hifigan.cpp:242-345
// Every conv is one launch of the tcgen05 implicit-GEMM kernel (conv_umma.cu) with the
// neighbouring elementwise ops fused into its prologue / epilogue.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/zvx.h"
#include "zvx_internal.h"
#include "mrf_fused_host.h"

using namespace zvx;

namespace {

thread_local std::string g_create_error;

struct DevTensor {
    void   *d = nullptr;
    int     dtype = 0;
    int     nd = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    size_t  nbytes = 0;
};

struct ConvVariant {
    __half *packed = nullptr;
    __half *packed_pair = nullptr;   // the same blocks with each CTA's half of the output channels contiguous (CTA-pair kernel), or null
    int ntaps = 0, w_tap0 = 0, w_tap_stride = 1, tap_off0 = 0, tap_step = 1, out_add = 0;
};

struct ConvLayer {
    __half *packed_fold = nullptr;   // decoder conv2 with its block's learned 1x1 shortcut folded in: per N-chunk this conv's blocks, then the
    int fold_ic = 0;                 // shortcut's (fold_ic input channels, one tap); see ConvParams::xb
    int OC = 0, IC = 0, K = 0, NC = 0;
    const __half *raw = nullptr;
    const float  *bias = nullptr;
    std::vector<ConvVariant> var;
};

// one launch of the fused MRF kernel: conv pairs [p0, p1) of a residual block
struct FusedChain {
    int p0 = 0, p1 = 0, halo = 0, valid = 0, wincfg = 0, nlayers = 0;
    mrf::Layer layers[mrf::MAX_LAYERS];
    const uint32_t *tbl0 = nullptr;
    double flops_per_row = 0.0;
};
struct FusedBlock { int CH = 0, k = 0, ncol = 256; std::vector<FusedChain> chains; };
struct WinCfg { int rate_idx, valid; };

struct ResBlkW  { ConvLayer conv1, conv2, conv1x1; bool learned_sc = false; const float *n1w, *n1b, *n2w, *n2b; int cin, cout; };
struct AdaBlkW  { ConvLayer conv1, conv2, conv1x1; bool learned_sc = false; int ada1, ada2; int cin, cout; };

}  // namespace

struct zvx_ctx {
    zvx_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    zvx_ctx *lane = nullptr;             // second stream + workspace sharing this context's weights (zvx_synth_batch pipelining)
    bool is_lane = false;
    int e2e_chunks = 2;
    int use_ref_kernels = 0;
    int debug_stop = -1;

    std::map<std::string, DevTensor> W;
    std::vector<void *> owned;           // every cudaMalloc to free

    // ---- decoder ----
    ResBlkW enc[2];
    ConvLayer asr0; const float *asr1w = nullptr, *asr1b = nullptr;
    AdaBlkW dec[5];
    ConvLayer to_out;
    AdainTable adain;
    // ---- vocoder ----
    const float *mel_mean = nullptr, *mel_scale = nullptr;
    ConvLayer input_conv;
    std::vector<ConvLayer> up;                    // per stage, var = phases
    std::vector<ConvLayer> upf;                   // per stage: all phases fused into one conv (OC == 0: not used)
    int use_fused_upconv = 1;
    std::vector<ConvLayer> mrf1, mrf2;            // [stage*nb*nd + j*nd + d]
    ConvLayer output_conv;
    std::vector<float> out_w_kc;                  // host fp32 copy [K][C] of the output conv (constant-bank path)
    float out_b_host = 0.f;
    std::vector<FusedBlock> fused;                // [stage*nb + j]; CH == 0: not fused
    std::vector<WinCfg> wincfg;                   // window tilings used by the fused chains
    std::map<std::vector<int>, const uint32_t *> tbl_cache;
    int use_fused = 1;
    int fused_prefetch = -1;                      // -1: automatic (measured: pays only at CH = 32), 0 / 1 / 2 (bulk request): ZVX_FUSED_PREFETCH
    int fused_persistent = 1;
    int fused_ncol128_ctas = 2;                   // CTAs per SM of a 128-column launch (1: leaves room for another launch's CTA)
    int fused_flags = 0;
    int conv_persistent = 1;
    float *feat = nullptr; int2 *feat_tab = nullptr; size_t feat_cap = 0, feat_tab_cap = 0;   // length regulator staging
    // single-utterance calls (the reference-facing eval() path: ~100 small launches) are replayed from CUDA graphs,
    // one per (stage, length, switches); dropped whenever a workspace buffer they point into is reallocated
    // the three independent residual blocks of an MRF stage run on three streams (inside a graph capture: three parallel
    // branches of the graph): a single short utterance leaves most launches under-filled, and for large batches the tail
    // of one block's persistent kernel overlaps the start of the next
    int h2d_chain = 1;          // zvx_synth_batch: a sub-batch's input copies wait for the previous sub-batch's
    int device_split = 1;       // zvx_synth_batch_device: two half batches, on the context and on its lane
    cudaEvent_t split_ev[2] = {nullptr, nullptr};
    int fork_branches = 1;
    cudaStream_t fork_stream[2] = {nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr, join_ev[2] = {nullptr, nullptr};
    float *fkY1[2] = {nullptr, nullptr}, *fkT2[2] = {nullptr, nullptr};
    __half *fkH16[2] = {nullptr, nullptr};
    struct GraphEntry { int kind, L, flags; cudaGraphExec_t exec; int64_t launches; };
    std::vector<GraphEntry> graphs;
    std::map<std::pair<int, int>, int> graph_seen;   // (kind, L) -> calls so far: a length is captured the second time it is seen
    int use_graphs = 1;
    int chunk_group_max = 8;   // zvx_vocode_chunked: at most this many chunks per vocoder pass
    int conv_smem_kb = 100;   // shared-memory budget of a one-tile conv CTA (two CTAs per SM)
    int conv_tma = 1;       // PRO_F16 operands of the one-tile conv kernel staged by TMA (cp.async.bulk.tensor) instead of cp.async
    int conv_cluster = 1;   // CTAs per cluster of the one-tile conv kernel sharing every weight stage by multicast; measured on
                            // B200 (profiles/r02_conv_cluster_ab.txt): 2 -> +7 %, 4 -> +19 % time on the decoder convs, so off
    int branch_f16 = 0; // fused MRF blocks write their outputs as fp16, the consumer sums three fp16 tensors (see run_vocoder)
    int conv_fold = 1;  // decoder: the learned 1x1 shortcut of a block is computed by the block's conv2 (extra K-chunks), ZVX_CONV_FOLD
    int conv_epi8 = 1;  // see run_conv
    int conv_pair = 0;  // PRO_F16 convs of the one-tile kernel as tcgen05 CTA pairs (M = 256, half a weight stage per SM)
    int upconv_mt2 = 1; // two M-tiles per CTA for up-convs with NC <= 96 (ZVX_UPCONV_MT2)
    int conv_mt2 = 0;   // two M-tiles per CTA: measured slower on B200 while the A producer is the limit (profiles/)
    std::vector<int> tile256_cfg;                 // per rate index: wincfg entry of the 256-row tiling
    int num_sms = 148;
    double fused_min_eff = 0.8;
    int *d_wins = nullptr;                        // [nwincfg][B+1] window prefixes
    std::vector<int> total_wins;
    std::vector<int> rates;                       // rows per frame after stage i (index 0 = 1)
    std::vector<int> chans;                       // channels after stage i (index 0 = input conv out)

    // ---- workspace ----
    int64_t cap_frames = 0;
    int cap_batch = 0;
    float *enc_in = nullptr, *sc = nullptr, *h528 = nullptr, *e0 = nullptr, *h1056 = nullptr, *catA = nullptr,
          *catB = nullptr, *asr = nullptr, *d1 = nullptr, *d2 = nullptr, *mel = nullptr, *style = nullptr;
    float *mu = nullptr, *rstd = nullptr, *adain_gb = nullptr;
    // InstanceNorm statistics fused into the producing conv's epilogue: per-tile partial sums (stat_part, [tiles][C]) +
    // a finalize launch; asr_mu / asr_rstd: statistics of the asr_res part of the concatenated decoder input (computed once)
    double2 *stat_part = nullptr; size_t stat_part_cap = 0;
    float *asr_mu = nullptr, *asr_rstd = nullptr;
    int fused_stats = 1;
    float *v0 = nullptr, *U = nullptr, *CS = nullptr, *Y1 = nullptr, *VA = nullptr, *VB = nullptr, *T2 = nullptr, *wav = nullptr;
    int branch_sum_in_consumer = 1;               // fused stages: write the 3 branch outputs, the next kernel sums them
    int stage_is_split[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // set per run: stage i's output lives in CS/VA/VB (3 buffers)
    // per-conv MRF path (stage 0, 256 channels): fp16 copies lrelu(U) (shared by the three blocks) and lrelu(y) (per
    // branch) written by the producing conv's epilogue, so that every conv1 reads a ready-made fp16 operand (cp.async)
    // instead of converting fp32 rows in its producer warps; sized for the stages that are not fused (chain_elems per frame)
    __half *U16 = nullptr, *Y16[3] = {nullptr, nullptr, nullptr};
    int64_t chain_elems = 0;
    int mrf_f16_chain = 1;
    // stage hand-off of the fused MRF stages: the last residual block's final phase adds the other two blocks' outputs,
    // averages, applies the consumer's leaky-ReLU and writes ONE fp16 operand (S16, ping-pong per stage) instead of three
    // fp32 tensors that the next up-conv / the output conv would each read (12 -> 2 bytes per element on the consumer side)
    // Measured on B200 (profiles/r02_stage_handoff_ab.txt): consumers -0.63 ms (out_conv 0.80 -> 0.46, up-convs -0.30), but
    // the final phase of the last block, which sits between two windows of the persistent kernel, pays the two extra
    // reads: stages 1-3 +1.6 ms.  Off by default; the code stays as the measured alternative (ZVX_STAGE_HANDOFF=1).
    __half *S16[2] = {nullptr, nullptr};
    int stage_handoff = 0;
    __half *H16 = nullptr, *X16 = nullptr, *R16 = nullptr;   // X16: decoder conv operand (normalised, activated, fp16); R16: raw input as fp16
    int dec_prepass = 1;
    int *d_seg = nullptr;                         // [B+1] frames prefix
    int *d_tiles = nullptr;                       // [nrates][B+1] tile prefixes
    int *d_err = nullptr;
    std::vector<int> h_seg, h_tiles;              // host mirrors
    std::vector<int> total_tiles;                 // per rate
    int *pin_tables = nullptr;                    // pinned staging for seg/tile tables
    cudaEvent_t tables_event = nullptr;
    bool tables_pending = false;
    float *pin_in = nullptr, *pin_out = nullptr;  // pinned staging for host-pointer API
    size_t pin_in_cap = 0, pin_out_cap = 0;
    int last_B = 0;
    int last_max_len = 0;
    int64_t last_frames = 0;
    // ---- per-launch profiling (CUDA events on the launch stream) ----
    bool prof = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct ProfEntry { int kind, stage; double flops, bytes; size_t ev; };
    std::vector<ProfEntry> prof_entries;
};

namespace {

int fail(zvx_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return 1;
}

#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <typename T>
int dev_alloc(zvx_ctx *ctx, T **p, size_t n)
{
    void *d = nullptr;
    CK(ctx, cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    ctx->owned.push_back(d);
    *p = reinterpret_cast<T *>(d);
    return 0;
}

void dev_free(zvx_ctx *ctx, void *p)
{
    if (!p) return;
    auto it = std::find(ctx->owned.begin(), ctx->owned.end(), p);
    if (it != ctx->owned.end()) ctx->owned.erase(it);
    cudaFree(p);
}

const DevTensor *find_w(zvx_ctx *ctx, const std::string &name)
{
    auto it = ctx->W.find(name);
    return it == ctx->W.end() ? nullptr : &it->second;
}

// checked_get_tensor (utils.cpp:9-17): a missing tensor is an error
int get_w(zvx_ctx *ctx, const std::string &name, int dtype, const DevTensor **out)
{
    const DevTensor *t = find_w(ctx, name);
    if (!t) return fail(ctx, "tensor '%s' not found", name.c_str());
    if (t->dtype != dtype)
        return fail(ctx, "tensor '%s' has dtype %d, expected %d (conv kernels must be F16, everything else F32)",
                    name.c_str(), t->dtype, dtype);
    *out = t;
    return 0;
}

int pick_nc(int oc)
{
    if (oc <= 256) return oc;
    for (int nc = 256; nc >= 16; nc -= 16)
        if (oc % nc == 0) return nc;
    return 0;
}

// Pack raw (OC, IC, K) fp16 (K fastest) into UMMA-ready blocks, ordered
// [n-chunk][k-chunk of 64 channels][tap] with each block laid out [kc/8][NC][8]:
// the K-major no-swizzle canonical layout (8-row groups contiguous, LBO = NC*16 bytes).
int pack_variant(zvx_ctx *ctx, const std::vector<__half> &raw, int OC, int IC, int K, int NC, ConvVariant &v)
{
    std::vector<__half> pk((size_t)OC * IC * v.ntaps);
    size_t o = 0;
    for (int n = 0; n < OC / NC; ++n)
        for (int c0 = 0; c0 < IC; c0 += 64) {
            const int kc = std::min(64, IC - c0);
            for (int a = 0; a < v.ntaps; ++a) {
                const int tap = v.w_tap0 + a * v.w_tap_stride;
                for (int g = 0; g < kc / 8; ++g)
                    for (int nn = 0; nn < NC; ++nn)
                        for (int e = 0; e < 8; ++e)
                            pk[o++] = raw[((size_t)(n * NC + nn) * IC + (c0 + g * 8 + e)) * K + tap];
            }
        }
    if (dev_alloc(ctx, &v.packed, pk.size())) return 1;
    CK(ctx, cudaMemcpy(v.packed, pk.data(), pk.size() * sizeof(__half), cudaMemcpyHostToDevice));
    // CTA pairs (conv_umma.cu, PAIR): CTA r of a pair holds output channels [r NC/2, (r+1) NC/2) of every block; laid out
    // [half][kc/8][NC/2][8] its share of a weight stage is ONE bulk copy (8 copies of 1.4 KB per stage made the loader the limit)
    if (ctx->conv_pair && NC % 16 == 0 && NC >= 32 && (int64_t)IC * v.ntaps >= 512) {
        std::vector<__half> pp(pk.size());
        size_t o2 = 0, blk = 0;
        for (int n = 0; n < OC / NC; ++n)
            for (int c0 = 0; c0 < IC; c0 += 64) {
                const int kc = std::min(64, IC - c0);
                for (int a = 0; a < v.ntaps; ++a) {
                    for (int r = 0; r < 2; ++r)
                        for (int g = 0; g < kc / 8; ++g)
                            for (int nn = 0; nn < NC / 2; ++nn)
                                for (int e = 0; e < 8; ++e)
                                    pp[o2++] = pk[blk + ((size_t)g * NC + (size_t)r * (NC / 2) + nn) * 8 + e];
                    blk += (size_t)kc * NC;
                }
            }
        if (dev_alloc(ctx, &v.packed_pair, pp.size())) return 1;
        CK(ctx, cudaMemcpy(v.packed_pair, pp.data(), pp.size() * sizeof(__half), cudaMemcpyHostToDevice));
    }
    return 0;
}

// host copies of raw conv weights are needed for packing
struct HostW;
int pack_fold(zvx_ctx *ctx, const std::vector<__half> &raw, const std::vector<__half> &raw_sc, ConvLayer &L, int IC_sc);
struct HostW { std::map<std::string, std::vector<__half>> h; std::map<std::string, std::vector<float>> f; };

// conv2 (OC, IC, K) followed, per N-chunk, by the 1x1 shortcut (OC, IC_sc, 1): the order the kernel's weight loader walks
int pack_fold(zvx_ctx *ctx, const std::vector<__half> &raw, const std::vector<__half> &raw_sc, ConvLayer &L, int IC_sc)
{
    const int OC = L.OC, IC = L.IC, K = L.K, NC = L.NC;
    if ((int)raw.size() != OC * IC * K || (int)raw_sc.size() != OC * IC_sc) return fail(ctx, "fold: unexpected weight sizes");
    std::vector<__half> pk((size_t)OC * ((size_t)IC * K + IC_sc));
    size_t o = 0;
    for (int n = 0; n < OC / NC; ++n) {
        for (int c0 = 0; c0 < IC; c0 += 64) {
            const int kc = std::min(64, IC - c0);
            for (int tap = 0; tap < K; ++tap)
                for (int g = 0; g < kc / 8; ++g)
                    for (int nn = 0; nn < NC; ++nn)
                        for (int e = 0; e < 8; ++e) pk[o++] = raw[((size_t)(n * NC + nn) * IC + (c0 + g * 8 + e)) * K + tap];
        }
        for (int c0 = 0; c0 < IC_sc; c0 += 64) {
            const int kc = std::min(64, IC_sc - c0);
            for (int g = 0; g < kc / 8; ++g)
                for (int nn = 0; nn < NC; ++nn)
                    for (int e = 0; e < 8; ++e) pk[o++] = raw_sc[(size_t)(n * NC + nn) * IC_sc + (c0 + g * 8 + e)];
        }
    }
    if (dev_alloc(ctx, &L.packed_fold, pk.size())) return 1;
    CK(ctx, cudaMemcpy(L.packed_fold, pk.data(), pk.size() * sizeof(__half), cudaMemcpyHostToDevice));
    L.fold_ic = IC_sc;
    return 0;
}

int make_conv(zvx_ctx *ctx, HostW &hw, const std::string &prefix, bool with_bias, int dilation, ConvLayer &L,
              int force_pad = -1)
{
    const DevTensor *w, *b = nullptr;
    if (get_w(ctx, prefix + ".w", ZVX_F16, &w)) return 1;
    if (with_bias && get_w(ctx, prefix + ".b", ZVX_F32, &b)) return 1;
    L.K = (int)w->ne[0];
    L.IC = (int)w->ne[1];
    L.OC = (int)w->ne[2];
    L.raw = reinterpret_cast<const __half *>(w->d);
    L.bias = b ? reinterpret_cast<const float *>(b->d) : nullptr;
    if (b && b->ne[0] != L.OC) return fail(ctx, "%s.b has %lld elements, expected %d", prefix.c_str(), (long long)b->ne[0], L.OC);
    if (L.IC % 16 != 0) return fail(ctx, "%s: input channels %d not a multiple of 16", prefix.c_str(), L.IC);
    if (L.OC == 1) return 0;   // single-channel output conv runs on CUDA cores, no packing
    L.NC = pick_nc(L.OC);
    if (L.NC == 0 || L.OC % 16 != 0) return fail(ctx, "%s: unsupported output channels %d", prefix.c_str(), L.OC);
    ConvVariant v;
    v.ntaps = L.K;
    v.w_tap0 = 0;
    v.w_tap_stride = 1;
    v.tap_step = dilation;
    v.tap_off0 = -(force_pad >= 0 ? force_pad : (L.K - 1) / 2 * dilation);
    v.out_add = 0;
    if (pack_variant(ctx, hw.h[prefix + ".w"], L.OC, L.IC, L.K, L.NC, v)) return 1;
    L.var.push_back(v);
    return 0;
}

// ConvTranspose1d(stride s, padding p = s/2 + s%2, output_padding s%2) emulated by the
// reference as zero-stuffing + stride-1 conv with the pre-flipped kernel
// (hifigan.cpp:44-65).  Polyphase form: output o = q*s + phi receives
// exactly the taps k' = k0 + a*s with k0 = (off - phi) mod s, reading x[q + d0 + a].
int make_upconv(zvx_ctx *ctx, HostW &hw, const std::string &prefix, int s, ConvLayer &L)
{
    const DevTensor *w, *b;
    if (get_w(ctx, prefix + ".w", ZVX_F16, &w)) return 1;
    if (get_w(ctx, prefix + ".b", ZVX_F32, &b)) return 1;
    L.K = (int)w->ne[0];
    L.IC = (int)w->ne[1];
    L.OC = (int)w->ne[2];
    L.raw = reinterpret_cast<const __half *>(w->d);
    L.bias = reinterpret_cast<const float *>(b->d);
    L.NC = pick_nc(L.OC);
    if (L.NC == 0 || L.OC % 16 != 0 || L.IC % 16 != 0) return fail(ctx, "%s: unsupported shape", prefix.c_str());
    const int p = s / 2 + s % 2, op = s % 2;
    const int off = L.K - 1 - p;
    if ((L.K - 2 * p + op) != s) return fail(ctx, "%s: kernel %d / stride %d does not give L_out = s*L_in", prefix.c_str(), L.K, s);
    for (int phi = 0; phi < s; ++phi) {
        ConvVariant v;
        const int k0 = (((off - phi) % s) + s) % s;
        v.w_tap0 = k0;
        v.w_tap_stride = s;
        v.ntaps = (L.K - k0 + s - 1) / s;
        v.tap_off0 = (phi + k0 - off) / s;    // exact: phi + k0 - off is a multiple of s
        if ((phi + k0 - off) % s != 0) return fail(ctx, "%s: polyphase derivation broken", prefix.c_str());
        v.tap_step = 1;
        v.out_add = phi;
        if (pack_variant(ctx, hw.h[prefix + ".w"], L.OC, L.IC, L.K, L.NC, v)) return 1;
        L.var.push_back(v);
    }
    return 0;
}

template <typename T>
int upload_vec(zvx_ctx *ctx, const std::vector<T> &v, const T **out)
{
    T *d = nullptr;
    if (dev_alloc(ctx, &d, v.size())) return 1;
    CK(ctx, cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = d;
    return 0;
}

// All s output phases of a ConvTranspose1d in ONE implicit GEMM: "output channel" (phi, oc), taps =
// the union of the phases' input offsets (3 for K = 2s), zero weights where a phase does not use an
// offset.  Row t of the fused output [rows_in][s*OC] is exactly rows t*s .. t*s+s-1 of the true
// output [rows_in*s][OC], so no scatter is needed.  Used where s*OC <= 256 (one N-chunk): there
// the per-phase launches are dominated by per-CTA overhead and re-read the input s times.
int make_upconv_fused(zvx_ctx *ctx, HostW &hw, const std::string &prefix, int s, const ConvLayer &L, ConvLayer &F)
{
    if (s * L.OC > 256) return 0;
    int dmin = 1 << 30, dmax = -(1 << 30);
    for (const ConvVariant &v : L.var) { dmin = std::min(dmin, v.tap_off0); dmax = std::max(dmax, v.tap_off0 + v.ntaps - 1); }
    const int KF = dmax - dmin + 1, OCF = s * L.OC;
    const std::vector<__half> &raw = hw.h[prefix + ".w"];
    const std::vector<float> &bias = hw.f[prefix + ".b"];
    if ((int)bias.size() != L.OC) return fail(ctx, "%s: bias not available for the fused up-conv", prefix.c_str());
    std::vector<__half> rf((size_t)OCF * L.IC * KF, __float2half(0.f));
    std::vector<float> bf(OCF);
    for (int phi = 0; phi < s; ++phi) {
        const ConvVariant &v = L.var[phi];
        for (int oc = 0; oc < L.OC; ++oc) {
            bf[phi * L.OC + oc] = bias[oc];
            for (int ic = 0; ic < L.IC; ++ic)
                for (int t = 0; t < v.ntaps; ++t)
                    rf[((size_t)(phi * L.OC + oc) * L.IC + ic) * KF + (v.tap_off0 - dmin + t)] =
                        raw[((size_t)oc * L.IC + ic) * L.K + v.w_tap0 + t * v.w_tap_stride];
        }
    }
    F.OC = OCF; F.IC = L.IC; F.K = KF; F.NC = pick_nc(OCF);
    if (F.NC == 0) return 0;
    const __half *d_raw = nullptr; const float *d_bias = nullptr;
    if (upload_vec(ctx, rf, &d_raw) || upload_vec(ctx, bf, &d_bias)) return 1;
    F.raw = d_raw; F.bias = d_bias;
    ConvVariant fv;
    fv.ntaps = KF; fv.w_tap0 = 0; fv.w_tap_stride = 1; fv.tap_off0 = dmin; fv.tap_step = 1; fv.out_add = 0;
    if (pack_variant(ctx, rf, OCF, L.IC, KF, F.NC, fv)) return 1;
    F.var.push_back(fv);
    return 0;
}

int f32_ptr(zvx_ctx *ctx, const std::string &name, int64_t n, const float **out)
{
    const DevTensor *t;
    if (get_w(ctx, name, ZVX_F32, &t)) return 1;
    if (n > 0 && t->ne[0] * t->ne[1] != n) return fail(ctx, "tensor '%s' has %lld elements, expected %lld", name.c_str(), (long long)(t->ne[0] * t->ne[1]), (long long)n);
    *out = reinterpret_cast<const float *>(t->d);
    return 0;
}

int build_decoder(zvx_ctx *ctx, HostW &hw)
{
    const int D = ctx->cfg.dim_in, BN = 2 * D, R = ctx->cfg.residual_dim, S = ctx->cfg.style_dim;
    char nm[128];
    const int enc_dims[2][2] = {{D, BN}, {BN, BN}};
    for (int i = 0; i < 2; ++i) {
        ResBlkW &b = ctx->enc[i];
        b.cin = enc_dims[i][0];
        b.cout = enc_dims[i][1];
        b.learned_sc = b.cin != b.cout;
        snprintf(nm, sizeof nm, "_mel_decoder.encode.%d", i);
        const std::string p = nm;
        if (make_conv(ctx, hw, p + ".conv1", true, 1, b.conv1)) return 1;
        if (make_conv(ctx, hw, p + ".conv2", true, 1, b.conv2)) return 1;
        if (b.learned_sc && make_conv(ctx, hw, p + ".conv1x1", false, 1, b.conv1x1)) return 1;
        if (b.learned_sc && ctx->conv_fold && b.cin >= 64 && b.cin % 8 == 0 && pack_fold(ctx, hw.h[p + ".conv2.w"], hw.h[p + ".conv1x1.w"], b.conv2, b.cin)) return 1;
        if (f32_ptr(ctx, p + ".norm1.w", b.cin, &b.n1w) || f32_ptr(ctx, p + ".norm1.b", b.cin, &b.n1b) ||
            f32_ptr(ctx, p + ".norm2.w", b.cin, &b.n2w) || f32_ptr(ctx, p + ".norm2.b", b.cin, &b.n2b))
            return 1;
        if (b.conv1.IC != b.cin || b.conv1.OC != b.cin || b.conv2.IC != b.cin || b.conv2.OC != b.cout)
            return fail(ctx, "%s: conv shapes do not match ResBlk1d(%d,%d)", nm, b.cin, b.cout);
    }
    if (make_conv(ctx, hw, "_mel_decoder.asr_res.0", true, 1, ctx->asr0)) return 1;
    if (ctx->asr0.OC != R || ctx->asr0.IC != D) return fail(ctx, "asr_res.0 shape mismatch");
    if (f32_ptr(ctx, "_mel_decoder.asr_res.1.w", R, &ctx->asr1w) || f32_ptr(ctx, "_mel_decoder.asr_res.1.b", R, &ctx->asr1b)) return 1;

    const int dec_dims[5][2] = {{BN + R, BN}, {BN + R, BN}, {BN + R, D}, {D, D}, {D, D}};
    ctx->adain.n = 0;
    ctx->adain.total = 0;
    ctx->adain.style_dim = S;
    for (int i = 0; i < 5; ++i) {
        AdaBlkW &b = ctx->dec[i];
        b.cin = dec_dims[i][0];
        b.cout = dec_dims[i][1];
        b.learned_sc = b.cin != b.cout;
        snprintf(nm, sizeof nm, "_mel_decoder.decode.%d", i);
        const std::string p = nm;
        if (make_conv(ctx, hw, p + ".conv1", true, 1, b.conv1)) return 1;
        if (make_conv(ctx, hw, p + ".conv2", true, 1, b.conv2)) return 1;
        if (b.learned_sc && make_conv(ctx, hw, p + ".conv1x1", false, 1, b.conv1x1)) return 1;
        if (b.learned_sc && ctx->conv_fold && b.cin >= 64 && b.cin % 8 == 0 && pack_fold(ctx, hw.h[p + ".conv2.w"], hw.h[p + ".conv1x1.w"], b.conv2, b.cin)) return 1;
        if (b.conv1.IC != b.cin || b.conv1.OC != b.cout || b.conv2.IC != b.cout || b.conv2.OC != b.cout)
            return fail(ctx, "%s: conv shapes do not match AdainResBlk1d(%d,%d)", nm, b.cin, b.cout);
        for (int k = 1; k <= 2; ++k) {
            const int C = k == 1 ? b.cin : b.cout;
            snprintf(nm, sizeof nm, "_mel_decoder.decode.%d.norm%d.fc", i, k);
            const float *fw = nullptr, *fb = nullptr;
            if (f32_ptr(ctx, std::string(nm) + ".w", (int64_t)2 * C * S, &fw) || f32_ptr(ctx, std::string(nm) + ".b", 2 * C, &fb)) return 1;
            AdainDesc &d = ctx->adain.d[ctx->adain.n];
            d.fc_w = fw;
            d.fc_b = fb;
            d.C = C;
            d.out_off = ctx->adain.total;
            (k == 1 ? b.ada1 : b.ada2) = ctx->adain.n;
            ctx->adain.total += 2 * C;
            ctx->adain.n++;
        }
    }
    if (make_conv(ctx, hw, "_mel_decoder.to_out.0", true, 1, ctx->to_out)) return 1;
    if (ctx->to_out.OC != ctx->cfg.num_mels) return fail(ctx, "to_out.0 has %d output channels, expected %d", ctx->to_out.OC, ctx->cfg.num_mels);
    return 0;
}

int get_table(zvx_ctx *ctx, int CH, int ncol, int d_cur, int d_next, const uint32_t **out)
{
    const std::vector<int> key = {CH, ncol, d_cur, d_next};
    auto it = ctx->tbl_cache.find(key);
    if (it != ctx->tbl_cache.end()) { *out = it->second; return 0; }
    if (upload_vec(ctx, mrf::make_table(CH, ncol, d_cur, d_next), out)) return 1;
    ctx->tbl_cache[key] = *out;
    return 0;
}

// Prepare the fused residual-block launches (mrf_fused.cu) for every stage whose channel count
// divides 128: packed weights, cumulative conv2 biases, scatter tables, window tilings.
int build_fused(zvx_ctx *ctx, HostW &hw)
{
    const zvx_config &c = ctx->cfg;
    const int nb = c.num_resblocks, nd = c.num_resblock_dilations;
    ctx->fused.assign((size_t)c.num_upsamples * nb, FusedBlock());
    char nm[128];
    for (int i = 0; i < c.num_upsamples; ++i) {
        const int CH = ctx->chans[i + 1];
        if (CH != 32 && CH != 64 && CH != 128) continue;
        if (2 * nd > mrf::MAX_LAYERS) continue;
        // columns per window: 128 -> two CTAs per SM (one's epilogue overlaps the other's MMAs) at the
        // price of a larger halo fraction; worth it where the MMA phase per layer is short (CH = 32)
        // columns per window, chosen per residual block below: 128 -> two CTAs per SM (one's epilogue
        // overlaps the other's MMAs) at the price of a larger halo fraction; pays off only for short kernels
        int ncol128_maxk = 0;                     // measured on B200 (profiles/r02_ab_small_experiments.txt): no longer pays anywhere
        {
            char key[40];
            snprintf(key, sizeof key, "ZVX_NCOL128_MAXK_%d", CH);
            if (const char *e = getenv(key)) ncol128_maxk = atoi(e);
        }

        for (int j = 0; j < nb; ++j) {
            const size_t idx0 = ((size_t)i * nb + j) * nd;
            const int k = ctx->mrf1[idx0].K;
            bool ok = k <= mrf::MAX_K && (k & 1);
            for (int d = 0; d < nd; ++d) ok = ok && ctx->mrf1[idx0 + d].K == k && ctx->mrf2[idx0 + d].K == k;
            int dil[8];
            for (int d = 0; d < nd; ++d) {
                dil[d] = c.resblock_dilations[j * nd + d];
                ok = ok && (dil[d] == 1 || dil[d] == 3 || dil[d] == 5);
            }
            if (!ok) continue;
            FusedBlock &fb = ctx->fused[(size_t)i * nb + j];
            const int ncol = (k <= ncol128_maxk && mrf_fused_supported(CH, 128)) ? 128 : 256;
            fb.k = k;
            fb.ncol = ncol;
            std::vector<float> cum(CH, 0.f);
            for (const mrf::ChainPlan &cp : mrf::plan_chains(CH, ncol, k, dil, nd, ctx->fused_min_eff)) {
                if (cp.valid <= 0) return fail(ctx, "fused MRF: window too small for kernel %d", k);
                FusedChain fc;
                fc.p0 = cp.p0; fc.p1 = cp.p1; fc.halo = cp.halo; fc.valid = cp.valid;
                fc.nlayers = 2 * (cp.p1 - cp.p0);
                fc.wincfg = -1;
                for (size_t w = 0; w < ctx->wincfg.size(); ++w)
                    if (ctx->wincfg[w].rate_idx == i + 1 && ctx->wincfg[w].valid == cp.valid) fc.wincfg = (int)w;
                if (fc.wincfg < 0) { ctx->wincfg.push_back({i + 1, cp.valid}); fc.wincfg = (int)ctx->wincfg.size() - 1; }
                if (get_table(ctx, CH, ncol, 1, dil[cp.p0], &fc.tbl0)) return 1;
                for (int l = 0; l < fc.nlayers; ++l) {
                    const int p = cp.p0 + l / 2;
                    const bool second = l & 1;
                    snprintf(nm, sizeof nm, "_meldec.blocks.%d.convs%d.%d.1", i * nb + j, second ? 2 : 1, p);
                    const std::string name = nm;
                    mrf::Layer &L = fc.layers[l];
                    L.k = k;
                    L.d = second ? 1 : dil[p];
                    L.accumulate = second ? 1 : 0;
                    L.out_slope = 0.1f;
                    const std::vector<__half> &raw = hw.h[name + ".w"];
                    const std::vector<float> &bias = hw.f[name + ".b"];
                    if ((int)raw.size() != CH * CH * k || (int)bias.size() != CH) return fail(ctx, "%s: unexpected size for the fused path", nm);
                    if (upload_vec(ctx, mrf::pack_weights(reinterpret_cast<const uint16_t *>(raw.data()), CH, k), &L.w)) return 1;
                    if (second) {
                        for (int q = 0; q < CH; ++q) cum[q] += bias[q];   // y_true = accumulator + sum of conv2 biases so far
                        if (upload_vec(ctx, mrf::to_row_order(cum), &L.bias)) return 1;
                    } else {
                        if (upload_vec(ctx, mrf::to_row_order(bias), &L.bias)) return 1;
                    }
                    L.tbl = nullptr;
                    if (l + 1 < fc.nlayers) {
                        const int d_next = (l & 1) ? dil[p + 1] : 1;
                        if (get_table(ctx, CH, ncol, L.d, d_next, &L.tbl)) return 1;
                    }
                    fc.flops_per_row += 2.0 * CH * CH * k;
                }
                // the cumulative bias restarts with every launch: y is re-read in fp32 with the
                // previous launches' biases already folded in
                std::fill(cum.begin(), cum.end(), 0.f);
                fb.chains.push_back(fc);
            }
            fb.CH = CH;
        }
    }
    return 0;
}

int build_vocoder(zvx_ctx *ctx, HostW &hw)
{
    const zvx_config &c = ctx->cfg;
    if (f32_ptr(ctx, "hifigan.mean", c.num_mels, &ctx->mel_mean) || f32_ptr(ctx, "hifigan.scale", c.num_mels, &ctx->mel_scale)) return 1;
    if (make_conv(ctx, hw, "_meldec.input_conv", true, 1, ctx->input_conv, (c.kernel_size - 1) / 2)) return 1;
    if (ctx->input_conv.IC != c.num_mels) return fail(ctx, "input_conv expects %d channels", ctx->input_conv.IC);
    ctx->rates.assign(1, 1);
    ctx->chans.assign(1, ctx->input_conv.OC);
    char nm[128];
    const int nb = c.num_resblocks, nd = c.num_resblock_dilations;
    ctx->up.resize(c.num_upsamples);
    ctx->upf.assign(c.num_upsamples, ConvLayer());
    ctx->mrf1.resize((size_t)c.num_upsamples * nb * nd);
    ctx->mrf2.resize((size_t)c.num_upsamples * nb * nd);
    for (int i = 0; i < c.num_upsamples; ++i) {
        snprintf(nm, sizeof nm, "_meldec.upsamples.%d.1", i);
        if (make_upconv(ctx, hw, nm, c.upsample_scales[i], ctx->up[i])) return 1;
        if (make_upconv_fused(ctx, hw, nm, c.upsample_scales[i], ctx->up[i], ctx->upf[i])) return 1;
        if (ctx->up[i].IC != ctx->chans.back()) return fail(ctx, "%s: expects %d input channels, have %d", nm, ctx->up[i].IC, ctx->chans.back());
        ctx->rates.push_back(ctx->rates.back() * c.upsample_scales[i]);
        ctx->chans.push_back(ctx->up[i].OC);
        for (int j = 0; j < nb; ++j)
            for (int d = 0; d < nd; ++d) {
                const size_t idx = ((size_t)i * nb + j) * nd + d;
                const int dil = c.resblock_dilations[j * nd + d];
                snprintf(nm, sizeof nm, "_meldec.blocks.%d.convs1.%d.1", i * nb + j, d);
                if (make_conv(ctx, hw, nm, true, dil, ctx->mrf1[idx])) return 1;
                snprintf(nm, sizeof nm, "_meldec.blocks.%d.convs2.%d.1", i * nb + j, d);
                if (make_conv(ctx, hw, nm, true, 1, ctx->mrf2[idx])) return 1;
                if (ctx->mrf1[idx].IC != ctx->chans.back() || ctx->mrf1[idx].OC != ctx->chans.back() ||
                    ctx->mrf2[idx].IC != ctx->chans.back() || ctx->mrf2[idx].OC != ctx->chans.back())
                    return fail(ctx, "%s: channel mismatch", nm);
            }
    }
    if (build_fused(ctx, hw)) return 1;
    if (ctx->rates.back() != c.hop_size) return fail(ctx, "product of upsample_scales (%d) != hop_size (%d)", ctx->rates.back(), c.hop_size);
    if (make_conv(ctx, hw, "_meldec.output_conv.1", true, 1, ctx->output_conv, (c.kernel_size - 1) / 2)) return 1;
    if (ctx->output_conv.OC != 1 || ctx->output_conv.IC != ctx->chans.back()) return fail(ctx, "output_conv shape mismatch");
    {
        const std::vector<__half> &ow = hw.h["_meldec.output_conv.1.w"];
        const std::vector<float> &ob = hw.f["_meldec.output_conv.1.b"];
        const int C = ctx->output_conv.IC, K = ctx->output_conv.K;
        if ((int)ow.size() == C * K && ob.size() == 1) {
            ctx->out_w_kc.resize((size_t)K * C);
            for (int cc = 0; cc < C; ++cc)
                for (int k = 0; k < K; ++k) ctx->out_w_kc[(size_t)k * C + cc] = __half2float(ow[(size_t)cc * K + k]);
            ctx->out_b_host = ob[0];
        }
    }
    return 0;
}

// ---------------------------------------------------------------- workspace
int64_t max_stage_elems(const zvx_ctx *ctx)
{
    int64_t m = 0;
    for (size_t i = 1; i < ctx->rates.size(); ++i) m = std::max<int64_t>(m, (int64_t)ctx->rates[i] * ctx->chans[i]);
    return m;
}

void drop_graphs(zvx_ctx *ctx)
{
    for (auto &g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
}

int reserve(zvx_ctx *ctx, int64_t frames, int batch)
{
    if (batch > ctx->cap_batch || frames > ctx->cap_frames) drop_graphs(ctx);
    const zvx_config &c = ctx->cfg;
    // Growth frees the old buffers first (they are large), so a failed allocation must leave the context in a state a
    // smaller retry can recover from: pointers nulled and the capacity zeroed BEFORE anything is freed, the new capacity
    // published only after every allocation has succeeded.
    if (batch > ctx->cap_batch) {
        const int nb = std::max(std::max(batch, 64), 2 * ctx->cap_batch);   // grow geometrically: every growth reallocates
        const int nr = (int)std::max<size_t>(ctx->rates.size(), 1);
        const int nw = (int)std::max<size_t>(ctx->wincfg.size(), 1);
        ctx->cap_batch = 0;
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        {
            int **ib[] = {&ctx->d_seg, &ctx->d_tiles, &ctx->d_wins};
            for (int **b : ib) { dev_free(ctx, *b); *b = nullptr; }
            float **fb[] = {&ctx->mu, &ctx->rstd, &ctx->adain_gb, &ctx->style, &ctx->asr_mu, &ctx->asr_rstd};
            for (float **b : fb) { dev_free(ctx, *b); *b = nullptr; }
        }
        if (ctx->pin_tables) cudaFreeHost(ctx->pin_tables);
        ctx->pin_tables = nullptr;
        ctx->tables_pending = false;
        if (dev_alloc(ctx, &ctx->d_seg, nb + 1) || dev_alloc(ctx, &ctx->d_tiles, (size_t)nr * (nb + 1)) ||
            dev_alloc(ctx, &ctx->d_wins, (size_t)nw * (nb + 1)))
            return 1;
        const int maxc = 2 * c.dim_in + c.residual_dim;
        if (dev_alloc(ctx, &ctx->mu, (size_t)nb * maxc) || dev_alloc(ctx, &ctx->rstd, (size_t)nb * maxc)) return 1;
        if (dev_alloc(ctx, &ctx->adain_gb, (size_t)nb * std::max(ctx->adain.total, 1))) return 1;
        if (dev_alloc(ctx, &ctx->style, (size_t)nb * c.style_dim)) return 1;
        if (dev_alloc(ctx, &ctx->asr_mu, (size_t)nb * std::max(c.residual_dim, 1)) || dev_alloc(ctx, &ctx->asr_rstd, (size_t)nb * std::max(c.residual_dim, 1))) return 1;
        CK(ctx, cudaMallocHost(&ctx->pin_tables, sizeof(int) * (size_t)(nr + 1 + nw) * (nb + 1)));
        ctx->cap_batch = nb;
    }
    if (frames > ctx->cap_frames) {
        const int64_t F = std::max<int64_t>(frames, 256);
        ctx->cap_frames = 0;
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        float **bufs[] = {&ctx->enc_in, &ctx->sc, &ctx->h528, &ctx->e0, &ctx->h1056, &ctx->catA, &ctx->catB, &ctx->asr,
                          &ctx->d1, &ctx->d2, &ctx->mel, &ctx->v0, &ctx->U, &ctx->CS, &ctx->Y1, &ctx->VA, &ctx->VB, &ctx->T2, &ctx->wav};
        for (float **b : bufs) { dev_free(ctx, *b); *b = nullptr; }
        dev_free(ctx, ctx->H16); ctx->H16 = nullptr;
        for (int j = 0; j < 2; ++j) {
            dev_free(ctx, ctx->fkY1[j]); dev_free(ctx, ctx->fkT2[j]); dev_free(ctx, ctx->fkH16[j]);
            ctx->fkY1[j] = ctx->fkT2[j] = nullptr; ctx->fkH16[j] = nullptr;
        }
        dev_free(ctx, ctx->X16); ctx->X16 = nullptr;
        dev_free(ctx, ctx->R16); ctx->R16 = nullptr;
        dev_free(ctx, ctx->U16); ctx->U16 = nullptr;
        for (int j = 0; j < 2; ++j) { dev_free(ctx, ctx->S16[j]); ctx->S16[j] = nullptr; }
        for (int j = 0; j < 3; ++j) { dev_free(ctx, ctx->Y16[j]); ctx->Y16[j] = nullptr; }
        const int D = c.dim_in, BN = 2 * D, R = c.residual_dim;
        if (c.with_decoder) {
            if (dev_alloc(ctx, &ctx->X16, F * (BN + R)) || dev_alloc(ctx, &ctx->R16, F * (BN + R))) return 1;
            if (dev_alloc(ctx, &ctx->enc_in, F * D) || dev_alloc(ctx, &ctx->sc, F * BN) || dev_alloc(ctx, &ctx->h528, F * D) ||
                dev_alloc(ctx, &ctx->e0, F * BN) || dev_alloc(ctx, &ctx->h1056, F * BN) || dev_alloc(ctx, &ctx->catA, F * (BN + R)) ||
                dev_alloc(ctx, &ctx->catB, F * (BN + R)) || dev_alloc(ctx, &ctx->asr, F * R) || dev_alloc(ctx, &ctx->d1, F * D) ||
                dev_alloc(ctx, &ctx->d2, F * D))
                return 1;
        }
        if (dev_alloc(ctx, &ctx->mel, F * c.num_mels)) return 1;
        if (c.with_vocoder) {
            const int64_t S = max_stage_elems(ctx);
            if (dev_alloc(ctx, &ctx->v0, F * ctx->chans[0]) || dev_alloc(ctx, &ctx->U, F * S) || dev_alloc(ctx, &ctx->CS, F * S) ||
                dev_alloc(ctx, &ctx->Y1, F * S) || dev_alloc(ctx, &ctx->VA, F * S) || dev_alloc(ctx, &ctx->VB, F * S) ||
                dev_alloc(ctx, &ctx->T2, F * S) ||
                dev_alloc(ctx, &ctx->H16, F * S) || dev_alloc(ctx, &ctx->wav, F * c.hop_size))
                return 1;
            if (ctx->stage_handoff)
                for (int j = 0; j < 2; ++j)
                    if (dev_alloc(ctx, &ctx->S16[j], F * S)) return 1;
            // fp16 operand chain of the stages that run conv by conv
            ctx->chain_elems = 0;
            if (ctx->mrf_f16_chain)
                for (int i = 0; i < c.num_upsamples; ++i) {
                    bool all_fused = ctx->use_fused != 0;
                    for (int j = 0; j < c.num_resblocks; ++j) all_fused = all_fused && ctx->fused[(size_t)i * c.num_resblocks + j].CH != 0;
                    if (!all_fused) ctx->chain_elems = std::max<int64_t>(ctx->chain_elems, (int64_t)ctx->rates[i + 1] * ctx->chans[i + 1]);
                }
            if (ctx->chain_elems > 0) {
                if (dev_alloc(ctx, &ctx->U16, F * ctx->chain_elems)) return 1;
                for (int j = 0; j < 3; ++j)
                    if (dev_alloc(ctx, &ctx->Y16[j], F * ctx->chain_elems)) return 1;
            }
            // per-branch temporaries of the forked MRF stages (run_vocoder): + 2 x (2 fp32 + 1 fp16) stage buffers
            if (ctx->fork_branches)
                for (int j = 0; j < 2; ++j)
                    if (dev_alloc(ctx, &ctx->fkY1[j], F * S) || dev_alloc(ctx, &ctx->fkT2[j], F * S) || dev_alloc(ctx, &ctx->fkH16[j], F * S)) return 1;
        }
        ctx->cap_frames = F;
    }
    return 0;
}

int ensure_fork(zvx_ctx *ctx);

// upload utterance segmentation: frames prefix + per-rate 128-row tile prefixes
int set_batch(zvx_ctx *ctx, int B, const int32_t *L, bool reserve_workspace = true)
{
    if (B <= 0) return fail(ctx, "empty batch");
    int64_t frames = 0;
    for (int b = 0; b < B; ++b) {
        if (L[b] <= 0) return fail(ctx, "utterance %d has non-positive length %d", b, L[b]);
        frames += L[b];
    }
    if (frames * (int64_t)ctx->cfg.hop_size * 64 > (int64_t)1 << 40) return fail(ctx, "batch too large");
    if (reserve(ctx, reserve_workspace ? frames : 0, B)) return 1;
    // the pinned table buffer is reused: wait until the previous batch's table upload has been consumed
    if (ctx->tables_pending) CK(ctx, cudaEventSynchronize(ctx->tables_event));
    const int nr = (int)ctx->rates.size();
    const int stride = ctx->cap_batch + 1;
    int *tab = ctx->pin_tables;
    ctx->h_seg.assign(B + 1, 0);
    ctx->total_tiles.assign(std::max(nr, 1), 0);
    tab[0] = 0;
    for (int b = 0; b < B; ++b) tab[b + 1] = tab[b] + L[b];
    for (int b = 0; b <= B; ++b) ctx->h_seg[b] = tab[b];
    for (int r = 0; r < std::max(nr, 1); ++r) {
        int *t = tab + (size_t)(r + 1) * stride;
        const int rate = nr ? ctx->rates[r] : 1;
        t[0] = 0;
        for (int b = 0; b < B; ++b) t[b + 1] = t[b] + (int)(((int64_t)L[b] * rate + 127) / 128);
        ctx->total_tiles[r] = t[B];
    }
    CK(ctx, cudaMemcpyAsync(ctx->d_seg, tab, sizeof(int) * (B + 1), cudaMemcpyHostToDevice, ctx->stream));
    for (int r = 0; r < std::max(nr, 1); ++r)
        CK(ctx, cudaMemcpyAsync(ctx->d_tiles + (size_t)r * stride, tab + (size_t)(r + 1) * stride, sizeof(int) * (B + 1),
                                cudaMemcpyHostToDevice, ctx->stream));
    // window prefixes of the fused MRF chains: ceil(rows / valid) windows per utterance
    ctx->total_wins.assign(ctx->wincfg.size(), 0);
    for (size_t w = 0; w < ctx->wincfg.size(); ++w) {
        int *t = tab + (size_t)(std::max(nr, 1) + 1 + w) * stride;
        const int rate = ctx->rates[ctx->wincfg[w].rate_idx], valid = ctx->wincfg[w].valid;
        t[0] = 0;
        for (int b = 0; b < B; ++b) t[b + 1] = t[b] + (int)(((int64_t)L[b] * rate + valid - 1) / valid);
        ctx->total_wins[w] = t[B];
        CK(ctx, cudaMemcpyAsync(ctx->d_wins + w * stride, t, sizeof(int) * (B + 1), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(ctx, cudaEventRecord(ctx->tables_event, ctx->stream));
    ctx->tables_pending = true;
    if (ctx->cfg.with_decoder && ctx->fused_stats) {
        // per-tile statistics partials of the widest decoder conv output
        const size_t need = (size_t)ctx->total_tiles[0] * (size_t)(2 * ctx->cfg.dim_in);
        if (need > ctx->stat_part_cap) {
            CK(ctx, cudaStreamSynchronize(ctx->stream));
            const size_t cap = std::max(need, ctx->stat_part_cap * 2);
            dev_free(ctx, ctx->stat_part); ctx->stat_part = nullptr; ctx->stat_part_cap = 0;
            if (dev_alloc(ctx, &ctx->stat_part, cap)) return 1;
            ctx->stat_part_cap = cap;
            drop_graphs(ctx);
        }
    }
    if (ctx->fork_branches && ctx->cfg.with_vocoder && ensure_fork(ctx)) return 1;
    ctx->last_B = B;
    ctx->last_frames = frames;
    ctx->last_max_len = 0;
    for (int b = 0; b < B; ++b) ctx->last_max_len = std::max(ctx->last_max_len, (int)L[b]);
    return 0;
}

// ---------------------------------------------------------------- per-launch profiling
int prof_begin(zvx_ctx *ctx, int kind, int stage, double flops, double bytes)
{
    if (!ctx->prof) return 0;
    while (ctx->ev_pool.size() < ctx->ev_used + 2) {
        cudaEvent_t e;
        CK(ctx, cudaEventCreate(&e));
        ctx->ev_pool.push_back(e);
    }
    zvx_ctx::ProfEntry pe = {kind, stage, flops, bytes, ctx->ev_used};
    ctx->prof_entries.push_back(pe);
    CK(ctx, cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream));
    ctx->ev_used += 2;
    return 0;
}
int prof_end(zvx_ctx *ctx)
{
    if (!ctx->prof) return 0;
    CK(ctx, cudaEventRecord(ctx->ev_pool[ctx->prof_entries.back().ev + 1], ctx->stream));
    return 0;
}

// ---------------------------------------------------------------- conv launch helper
struct ConvCall {
    int kind = ZVX_K_DEC_CONV, stage = 0;
    const ConvLayer *L = nullptr;
    int variant = 0;
    const void *x = nullptr; int ldx = 0, x_ch_off = 0;
    const float *x2 = nullptr, *x3 = nullptr; float sum_scale = 0.f; bool sum3_half = false;   // PRO_SUM3 (sum3_half: fp16 sources)
    int rate_idx = 0;          // index into ctx->rates of the INPUT rate
    int pro_mode = PRO_CVT; float pro_slope = 0.f;
    const float *mu = nullptr, *rstd = nullptr; int stat_stride = 0;
    const float *g = nullptr, *b = nullptr; int gb_stride = 0;
    bool use_bias = true;
    const float *res = nullptr; int ldres = 0, res_ch_off = 0;
    const float *acc_in = nullptr;
    float scale = 0.f;
    float *out32 = nullptr; int ldo32 = 0, o32_ch_off = 0;
    __half *out16 = nullptr; int ldo16 = 0, o16_ch_off = 0; float out16_slope = 0.f;
    int out_mul = 1;
    double flops = 0.0;        // algorithmic FLOPs of the launch when they differ from 2*rows*OC*IC*taps
    bool stats = false;        // the epilogue also emits the per-tile statistics partials of the output (ctx->stat_part)
    __half *raw16_out = nullptr;              // run_norm_conv: the norm pass also writes the plain fp16 copy of its input here
    const void *xb = nullptr; int ldxb = 0;   // fold: fp16 source of the layer's folded 1x1 shortcut (ConvLayer::packed_fold)
};

int run_conv(zvx_ctx *ctx, const ConvCall &cc)
{
    const ConvLayer &L = *cc.L;
    const ConvVariant &v = L.var[cc.variant];
    ConvParams p;
    memset(&p, 0, sizeof p);
    p.x = cc.x; p.ldx = cc.ldx; p.x_ch_off = cc.x_ch_off; p.Cin = L.IC;
    p.x2 = cc.x2; p.x3 = cc.x3; p.sum_scale = cc.sum_scale;
    p.seg_start = ctx->d_seg;
    p.tile_start = ctx->d_tiles + (size_t)cc.rate_idx * (ctx->cap_batch + 1);
    p.B = ctx->last_B;
    p.rate_in = ctx->rates.empty() ? 1 : ctx->rates[cc.rate_idx];
    p.ntaps = v.ntaps; p.tap_off0 = v.tap_off0; p.tap_step = v.tap_step;
    p.w_packed = v.packed; p.w_raw = L.raw; p.w_taps_total = L.K; p.w_tap0 = v.w_tap0; p.w_tap_stride = v.w_tap_stride;
    p.Cout = L.OC; p.NC = L.NC;
    p.pro_mode = (cc.pro_mode == PRO_SUM3 && cc.sum3_half) ? PRO_SUM3H : cc.pro_mode; p.pro_slope = cc.pro_slope;
    p.p_mu = cc.mu; p.p_rstd = cc.rstd; p.p_stat_stride = cc.stat_stride;
    p.p_g = cc.g; p.p_b = cc.b; p.p_gb_stride = cc.gb_stride;
    p.bias = cc.use_bias ? L.bias : nullptr;
    p.res = cc.res; p.ldres = cc.ldres; p.res_ch_off = cc.res_ch_off;
    p.acc_in = cc.acc_in;
    p.has_scale = cc.scale != 0.f; p.scale = cc.scale;
    p.out32 = cc.out32; p.ldo32 = cc.ldo32; p.o32_ch_off = cc.o32_ch_off;
    p.out16 = cc.out16; p.ldo16 = cc.ldo16; p.o16_ch_off = cc.o16_ch_off; p.out16_slope = cc.out16_slope;
    p.out_mul = cc.out_mul; p.out_add = v.out_add;
    p.stats_out = cc.stats ? ctx->stat_part : nullptr;
    p.err_flag = ctx->d_err;
    const bool fold = cc.xb != nullptr;
    if (fold) {
        if (!L.packed_fold || cc.pro_mode != PRO_F16) return fail(ctx, "conv: folded shortcut without packed weights / fp16 operand");
        p.xb = cc.xb; p.ldxb = cc.ldxb; p.Cin_b = L.fold_ic; p.w_packed = L.packed_fold;
    }
    int tiles = ctx->total_tiles[cc.rate_idx];
    // two M-tiles per CTA (each weight stage feeds 256 rows) whenever that still fills the GPU
    p.mt = 1;
    // (measured per launch kind, profiles/r02_ab_conv_epilogue_and_pairs.txt: it pays only for the narrow up-conv of the last stage --
    //  NC = 96, two taps, a streaming kernel whose per-CTA set-up is a third of a CTA's life -- and costs 17-36 % elsewhere)
    const bool mt2_here = ctx->conv_mt2 || (ctx->upconv_mt2 && cc.kind == ZVX_K_UPCONV && L.NC <= 96);
    if (!ctx->use_ref_kernels && mt2_here && cc.rate_idx < (int)ctx->tile256_cfg.size() &&
        (int64_t)tiles * (L.OC / L.NC) >= (int64_t)3 * ctx->num_sms) {
        const int w = ctx->tile256_cfg[cc.rate_idx];
        p.mt = 2;
        p.tile_start = ctx->d_wins + (size_t)w * (ctx->cap_batch + 1);
        tiles = ctx->total_wins[w];
    }
    ctx->launches++;
    const double rows = (double)ctx->last_frames * p.rate_in;
    if (prof_begin(ctx, cc.kind, cc.stage, cc.flops > 0.0 ? cc.flops : 2.0 * rows * L.OC * ((double)L.IC * v.ntaps + (fold ? L.fold_ic : 0)), 0.0)) return 1;
    if (ctx->use_ref_kernels) {
        CK(ctx, conv_ref_launch(p, tiles, ctx->stream));
    } else {
        // persistent kernel whenever there are at least ~2 work items per SM, else one tile per CTA
        // (measured per launch kind, profiles/: it wins on the MRF convs -- long K loops, residual-stream
        // epilogue -- and loses on the short decoder / up-conv launches)
        if (ctx->conv_persistent && cc.kind == ZVX_K_MRF_CONV && p.mt == 1 && (int64_t)tiles * (L.OC / L.NC) >= (int64_t)2 * ctx->num_sms) {
            const size_t smem = conv_umma_pk_plan(p, 226 * 1024);
            if (smem > 227 * 1024) return fail(ctx, "conv needs %zu bytes of shared memory", smem);
            CK(ctx, conv_umma_pk_launch(p, tiles, ctx->num_sms, smem, ctx->stream));
        } else {
            p.use_tma = ctx->conv_tma && p.pro_mode == PRO_F16;
            // a second epilogue warp per lane quarter where the epilogue is a large share of a CTA's life and the launch
            // is not bound by the number of co-resident CTAs (measured per launch kind, profiles/r02_ab_conv_epilogue.txt)
            p.epi8 = ctx->conv_epi8 && p.mt == 1 && (cc.kind == ZVX_K_DEC_CONV || (cc.kind == ZVX_K_UPCONV && L.NC >= 128));
            p.pair = ctx->conv_pair && cc.kind == ZVX_K_DEC_CONV && v.packed_pair && p.use_tma && p.mt == 1 && tiles >= 2;
            p.tma_row0 = 0;
            p.tma_rows = (long long)ctx->last_frames * p.rate_in;
            const size_t smem = conv_umma_plan(p, p.mt == 2 ? 226 * 1024 : (size_t)ctx->conv_smem_kb * 1024);
            if (smem > 227 * 1024) return fail(ctx, "conv needs %zu bytes of shared memory", smem);
            if (fold && p.Cin_b <= 0) return fail(ctx, "conv: the launch cannot fold its shortcut (needs the TMA-staged fp16 operand)");
            if (p.pair && fold) p.pair = 0;                   // (the pair layout of the folded weights is not built)
            if (p.pair) p.w_packed = v.packed_pair;          // (conv_umma_plan clears the flag when the launch does not qualify)
            p.cluster = 1;
            if (ctx->conv_cluster > 1 && tiles >= 2 * ctx->conv_cluster && (int64_t)L.IC * v.ntaps >= 512) {
                p.cluster = ctx->conv_cluster;
                const int kc_last = L.IC % 64 ? L.IC % 64 : 64;
                if (((size_t)kc_last * L.NC * 2) % (16 * (size_t)p.cluster) != 0) p.cluster = 1;   // shares must be 16-byte multiples
            }
            CK(ctx, conv_umma_launch(p, tiles, smem, ctx->stream));
        }
    }
    return prof_end(ctx);
}

int run_stats(zvx_ctx *ctx, const float *x, int ld, int ch_off, int C)
{
    ctx->launches++;
    if (prof_begin(ctx, ZVX_K_STATS, 0, 0.0, 2.0 * (double)ctx->last_frames * C * sizeof(float))) return 1;
    CK(ctx, stats_launch(x, ld, ch_off, C, ctx->d_seg, ctx->last_B, 1, ctx->mu, ctx->rstd, ctx->stream));
    return prof_end(ctx);
}

// mean / rstd of the tensor the previous conv launch wrote (its epilogue left per-tile partials in ctx->stat_part);
// C_tail > 0: the asr_res channels of the concatenated decoder input follow, copied from ctx->asr_mu / asr_rstd
int run_stats_finalize(zvx_ctx *ctx, int C, int C_tail)
{
    ctx->launches++;
    if (prof_begin(ctx, ZVX_K_STATS, 0, 0.0, 16.0 * (double)ctx->total_tiles[0] * C)) return 1;
    CK(ctx, stats_finalize_launch(ctx->stat_part, C, ctx->d_tiles, ctx->d_seg, ctx->last_B, 1, ctx->asr_mu, ctx->asr_rstd, C_tail,
                                  ctx->mu, ctx->rstd, ctx->stream));
    return prof_end(ctx);
}

// A decoder conv whose input is InstanceNorm/AdaIN-affine + leaky-ReLU of x (stats already in ctx->mu/rstd):
// either fused into the conv's prologue (PRO_NORM) or as a stand-alone fp16 pass + PRO_F16 conv.
int run_norm_conv(zvx_ctx *ctx, ConvCall cc)
{
    if (!ctx->dec_prepass || ctx->use_ref_kernels) return run_conv(ctx, cc);
    const int C = cc.L->IC;
    ctx->launches++;
    if (prof_begin(ctx, ZVX_K_NORM_AFFINE, 0, 0.0, 6.0 * (double)ctx->last_frames * C)) return 1;
    CK(ctx, norm_act_f16_launch(reinterpret_cast<const float *>(cc.x), cc.ldx, cc.x_ch_off, C, ctx->d_seg, ctx->last_B, ctx->last_max_len,
                                cc.mu, cc.rstd, cc.g, cc.b, cc.gb_stride, cc.pro_slope, ctx->X16, cc.raw16_out, ctx->stream));
    if (prof_end(ctx)) return 1;
    cc.x = ctx->X16; cc.ldx = C; cc.x_ch_off = 0; cc.pro_mode = PRO_F16;
    cc.mu = cc.rstd = cc.g = cc.b = nullptr;
    return run_conv(ctx, cc);
}

// fp32 -> fp16 copy of a conv input that is used raw (1x1 shortcuts, asr_res, to_out); returns the
// ConvCall rewritten to read it (PRO_F16).  `fresh` = false reuses the copy made by the previous call.
int use_raw_f16(zvx_ctx *ctx, ConvCall &cc, bool fresh)
{
    if (!ctx->dec_prepass || ctx->use_ref_kernels || cc.pro_mode != PRO_CVT) return 0;
    const int C = cc.L->IC;
    if (fresh) {
        ctx->launches++;
        if (prof_begin(ctx, ZVX_K_NORM_AFFINE, 0, 0.0, 6.0 * (double)ctx->last_frames * C)) return 1;
        CK(ctx, cvt_f16_launch(reinterpret_cast<const float *>(cc.x), cc.ldx, cc.x_ch_off, C, (size_t)ctx->last_frames, ctx->R16, ctx->stream));
        if (prof_end(ctx)) return 1;
    }
    cc.x = ctx->R16; cc.ldx = C; cc.x_ch_off = 0; cc.pro_mode = PRO_F16;
    return 0;
}

// ---------------------------------------------------------------- decoder schedule
// enc_in [F][D] and style [B][S] are already on the device.
int run_decoder(zvx_ctx *ctx, float *mel_out)
{
    const zvx_config &c = ctx->cfg;
    const int D = c.dim_in, BN = 2 * D, R = c.residual_dim, CAT = BN + R;
    const float inv_sqrt2 = (float)(1.0 / std::sqrt(2.0));

    ctx->launches++;
    if (prof_begin(ctx, ZVX_K_ADAIN_FC, 0, 2.0 * ctx->last_B * (double)ctx->adain.total * c.style_dim,
                   (double)ctx->adain.total * c.style_dim * sizeof(float)))
        return 1;
    CK(ctx, adain_fc_launch(ctx->adain, ctx->style, ctx->last_B, ctx->adain_gb, ctx->stream));
    if (prof_end(ctx)) return 1;

    // The learned shortcut (1x1 conv of the block input, plus its fp16 copy) does not depend on the block's
    // norm -> conv1 chain: it runs on a forked stream and is joined before conv2 reads it.
    const bool fork_ok = ctx->fork_branches && ctx->fork_stream[0] && !ctx->prof && ctx->debug_stop < 0 && !ctx->use_ref_kernels;
    auto forked = [&](auto body) -> int {
        if (!fork_ok) return body();
        cudaStream_t main_stream = ctx->stream;
        cudaError_t e = cudaEventRecord(ctx->fork_ev, main_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->fork_stream[0], ctx->fork_ev, 0);
        if (e != cudaSuccess) return fail(ctx, "fork: %s", cudaGetErrorString(e));
        ctx->stream = ctx->fork_stream[0];
        const int rc = body();
        e = cudaEventRecord(ctx->join_ev[0], ctx->stream);
        ctx->stream = main_stream;
        if (rc) return rc;
        if (e != cudaSuccess) return fail(ctx, "fork: %s", cudaGetErrorString(e));
        return 0;
    };
    auto join = [&]() -> int {
        if (!fork_ok) return 0;
        CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->join_ev[0], 0));
        return 0;
    };

    // The learned 1x1 shortcut folded into the block's conv2 (ConvParams::xb): its weights follow conv2's in ConvLayer::packed_fold,
    // its operand is the fp16 copy of the block input; no shortcut launch, no `sc` round trip.
    auto fold_ok = [&](const ConvLayer &conv2) {
        return ctx->conv_fold && conv2.packed_fold && ctx->dec_prepass && !ctx->use_ref_kernels && ctx->conv_tma && !ctx->conv_mt2;
    };

    // InstanceNorm statistics (ggml_norm, ggml-cpu.c:6880-6929; call sites stylettsdec.cpp:94-98,119-123,191): the conv that
    // WRITES a tensor leaves per-tile sums in its epilogue and a small finalize launch turns them into mean / rstd right
    // before the consumer; only tensors no conv of this schedule produced (enc_in, the asr_res branch) take the stand-alone pass
    const bool fs = ctx->fused_stats && !ctx->use_ref_kernels && !ctx->conv_mt2 && ctx->stat_part;

    // ---- encode.0 / encode.1 : ResBlk1d (stylettsdec.cpp:69-149) ----
    const float *x = ctx->enc_in; int ldx = D;
    float *enc_out[2] = {ctx->e0, ctx->catA};
    int enc_ld[2] = {BN, CAT};
    float *enc_h[2] = {ctx->h528, ctx->h1056};
    for (int i = 0; i < 2; ++i) {
        ResBlkW &b = ctx->enc[i];
        const float *sc = x; int ldsc = ldx;
        const void *fold_x = nullptr; int fold_ld = 0;
        if (b.learned_sc && fold_ok(b.conv2)) {
            // the fp16 copy of the block input is written by conv1's norm pass below (i == 0: x is enc_in; the copy is reused by asr_res)
            fold_x = ctx->R16; fold_ld = b.cin; sc = nullptr; ldsc = 0;
        } else if (b.learned_sc) {
            ConvCall s; s.L = &b.conv1x1; s.x = x; s.ldx = ldx; s.pro_mode = PRO_CVT; s.use_bias = false;
            s.out32 = ctx->sc; s.ldo32 = b.cout;
            if (forked([&]() -> int {
                    if (use_raw_f16(ctx, s, true)) return 1;     // i == 0: x is enc_in; the copy is reused by asr_res below
                    return run_conv(ctx, s);
                }))
                return 1;
            sc = ctx->sc; ldsc = b.cout;
        }
        if (fs && i > 0 ? run_stats_finalize(ctx, b.cin, 0) : run_stats(ctx, x, ldx, 0, b.cin)) return 1;
        ConvCall c1; c1.L = &b.conv1; c1.x = x; c1.ldx = ldx; c1.pro_mode = PRO_NORM; c1.pro_slope = 0.2f;
        c1.mu = ctx->mu; c1.rstd = ctx->rstd; c1.stat_stride = b.cin; c1.g = b.n1w; c1.b = b.n1b; c1.gb_stride = 0;
        c1.out32 = enc_h[i]; c1.ldo32 = b.cin; c1.stats = fs;
        if (fold_x) c1.raw16_out = ctx->R16;
        if (run_norm_conv(ctx, c1)) return 1;
        if (fs ? run_stats_finalize(ctx, b.cin, 0) : run_stats(ctx, enc_h[i], b.cin, 0, b.cin)) return 1;
        ConvCall c2; c2.L = &b.conv2; c2.x = enc_h[i]; c2.ldx = b.cin; c2.pro_mode = PRO_NORM; c2.pro_slope = 0.2f;
        c2.mu = ctx->mu; c2.rstd = ctx->rstd; c2.stat_stride = b.cin; c2.g = b.n2w; c2.b = b.n2b; c2.gb_stride = 0;
        c2.res = sc; c2.ldres = ldsc; c2.scale = inv_sqrt2;
        c2.xb = fold_x; c2.ldxb = fold_ld;
        c2.out32 = enc_out[i]; c2.ldo32 = enc_ld[i]; c2.stats = fs;
        if (b.learned_sc && !fold_x && join()) return 1;
        if (run_norm_conv(ctx, c2)) return 1;
        x = enc_out[i]; ldx = enc_ld[i];
    }
    // ---- asr_res = IN_affine(conv1x1(enc_seq) + b)  (stylettsdec.cpp:382-396), into both concat buffers ----
    {
        ConvCall a; a.L = &ctx->asr0; a.x = ctx->enc_in; a.ldx = D; a.pro_mode = PRO_CVT; a.out32 = ctx->asr; a.ldo32 = R;
        if (use_raw_f16(ctx, a, !ctx->enc[0].learned_sc)) return 1;   // enc_in was converted for encode.0's shortcut
        if (run_conv(ctx, a)) return 1;
        if (run_stats(ctx, ctx->asr, R, 0, R)) return 1;
        ctx->launches++;
        if (prof_begin(ctx, ZVX_K_NORM_AFFINE, 0, 0.0, 3.0 * (double)ctx->last_frames * R * sizeof(float))) return 1;
        CK(ctx, norm_affine_launch(ctx->asr, R, R, ctx->d_seg, ctx->last_B, ctx->mu, ctx->rstd, ctx->asr1w, ctx->asr1b,
                                   ctx->catA, ctx->catB, CAT, BN, ctx->stream));
        if (prof_end(ctx)) return 1;
        if (fs) {
            // statistics of the asr_res channels of the concatenated input: the same for decode.0-2, computed once
            ctx->launches++;
            if (prof_begin(ctx, ZVX_K_STATS, 0, 0.0, 2.0 * (double)ctx->last_frames * R * sizeof(float))) return 1;
            CK(ctx, stats_launch(ctx->catA, CAT, BN, R, ctx->d_seg, ctx->last_B, 1, ctx->asr_mu, ctx->asr_rstd, ctx->stream));
            if (prof_end(ctx)) return 1;
        }
    }
    // ---- decode.0-4 : AdainResBlk1d (stylettsdec.cpp:242-304) ----
    const float *din[5]  = {ctx->catA, ctx->catB, ctx->catA, ctx->d1, ctx->d2};
    const int    dinl[5] = {CAT, CAT, CAT, D, D};
    float       *dout[5] = {ctx->catB, ctx->catA, ctx->d1, ctx->d2, ctx->d1};
    const int    doutl[5] = {CAT, CAT, D, D, D};
    for (int i = 0; i < 5; ++i) {
        AdaBlkW &b = ctx->dec[i];
        const AdainDesc &a1 = ctx->adain.d[b.ada1], &a2 = ctx->adain.d[b.ada2];
        float *h = b.cout == BN ? ctx->h1056 : ctx->h528;
        const float *sc = din[i]; int ldsc = dinl[i];
        const void *fold_x = nullptr; int fold_ld = 0;
        if (b.learned_sc && fold_ok(b.conv2)) {
            fold_x = ctx->R16; fold_ld = b.cin; sc = nullptr; ldsc = 0;      // written by conv1's norm pass below
        } else if (b.learned_sc) {
            ConvCall s; s.L = &b.conv1x1; s.x = din[i]; s.ldx = dinl[i]; s.pro_mode = PRO_CVT; s.use_bias = false;
            s.out32 = ctx->sc; s.ldo32 = b.cout;
            if (forked([&]() -> int {
                    if (use_raw_f16(ctx, s, true)) return 1;
                    return run_conv(ctx, s);
                }))
                return 1;
            sc = ctx->sc; ldsc = b.cout;
        }
        if (fs ? run_stats_finalize(ctx, dinl[i] == CAT ? BN : b.cin, dinl[i] == CAT ? R : 0) : run_stats(ctx, din[i], dinl[i], 0, b.cin)) return 1;
        ConvCall c1; c1.L = &b.conv1; c1.x = din[i]; c1.ldx = dinl[i]; c1.pro_mode = PRO_NORM; c1.pro_slope = 0.2f;
        c1.mu = ctx->mu; c1.rstd = ctx->rstd; c1.stat_stride = b.cin;
        c1.g = ctx->adain_gb + a1.out_off; c1.b = ctx->adain_gb + a1.out_off + a1.C; c1.gb_stride = ctx->adain.total;
        c1.out32 = h; c1.ldo32 = b.cout; c1.stats = fs;
        if (fold_x) c1.raw16_out = ctx->R16;
        if (run_norm_conv(ctx, c1)) return 1;
        if (fs ? run_stats_finalize(ctx, b.cout, 0) : run_stats(ctx, h, b.cout, 0, b.cout)) return 1;
        ConvCall c2; c2.L = &b.conv2; c2.x = h; c2.ldx = b.cout; c2.pro_mode = PRO_NORM; c2.pro_slope = 0.2f;
        c2.mu = ctx->mu; c2.rstd = ctx->rstd; c2.stat_stride = b.cout;
        c2.g = ctx->adain_gb + a2.out_off; c2.b = ctx->adain_gb + a2.out_off + a2.C; c2.gb_stride = ctx->adain.total;
        c2.res = sc; c2.ldres = ldsc; c2.scale = inv_sqrt2;
        c2.xb = fold_x; c2.ldxb = fold_ld;
        c2.out32 = dout[i]; c2.ldo32 = doutl[i]; c2.stats = fs && i < 4;
        if (b.learned_sc && !fold_x && join()) return 1;
        if (run_norm_conv(ctx, c2)) return 1;
    }
    // ---- to_out (stylettsdec.cpp:432-441) ----
    ConvCall o; o.L = &ctx->to_out; o.x = dout[4]; o.ldx = D; o.pro_mode = PRO_CVT; o.out32 = mel_out; o.ldo32 = c.num_mels;
    return run_conv(ctx, o);
}

// ---------------------------------------------------------------- vocoder schedule
// streams and events of the forked MRF stages (their per-branch temporaries are part of the workspace, reserve())
int ensure_fork(zvx_ctx *ctx)
{
    if (ctx->fork_stream[0]) return 0;
    for (int j = 0; j < 2; ++j) CK(ctx, cudaEventCreateWithFlags(&ctx->join_ev[j], cudaEventDisableTiming));
    CK(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    CK(ctx, cudaStreamCreateWithFlags(&ctx->fork_stream[1], cudaStreamNonBlocking));
    CK(ctx, cudaStreamCreateWithFlags(&ctx->fork_stream[0], cudaStreamNonBlocking));
    return 0;
}

int run_vocoder(zvx_ctx *ctx, const float *mel_in, float *wav_out, int16_t *pcm_out = nullptr)
{
    const zvx_config &c = ctx->cfg;
    const int nb = c.num_resblocks, nd = c.num_resblock_dilations;
    // (mel - mean) / scale -> input_conv (hifigan.cpp:242-265)
    {
        ConvCall ic; ic.kind = ZVX_K_VOC_INPUT_CONV; ic.L = &ctx->input_conv; ic.x = mel_in; ic.ldx = c.num_mels; ic.pro_mode = PRO_MEL;
        ic.mu = ctx->mel_mean; ic.rstd = ctx->mel_scale; ic.stat_stride = 0;
        ic.out32 = ctx->v0; ic.ldo32 = ctx->chans[0];
        if (run_conv(ctx, ic)) return 1;
    }
    const float *vin = ctx->v0, *vin2 = nullptr, *vin3 = nullptr;   // vin2/vin3: stage output still split in 3 branches
    const __half *vin16 = nullptr;                                  // stage output handed over as ONE ready-made fp16 operand
    bool vin_half = false;                                          // vin / vin2 / vin3 are fp16 tensors (ZVX_BRANCH_F16)
    float *vout[2] = {ctx->VA, ctx->VB};
    const float third = (float)(1.0 / (float)nb);
    const bool fork_ok = ctx->fork_branches && ctx->fork_stream[0] && ctx->fkY1[0] && !ctx->prof && ctx->debug_stop < 0 && !ctx->use_ref_kernels;
    for (int i = 0; i < 8; ++i) ctx->stage_is_split[i] = 0;
    for (int i = 0; i < c.num_upsamples; ++i) {
        if (ctx->debug_stop >= 0 && i >= ctx->debug_stop) return 0;
        const int cin = ctx->chans[i], ch = ctx->chans[i + 1];
        const int s = c.upsample_scales[i];
        // does a block of this stage run conv by conv, and does the fp16 chain fit the workspace?
        bool any_unfused = false;
        for (int j = 0; j < nb; ++j) any_unfused = any_unfused || !(ctx->use_fused && ctx->fused[(size_t)i * nb + j].CH);
        const bool chain = any_unfused && !ctx->use_ref_kernels && ctx->U16 && nb <= 3 &&
                           (int64_t)ctx->rates[i + 1] * ch <= ctx->chain_elems;
        // leaky_relu(0.1) -> ConvTranspose1d, one launch per output phase (hifigan.cpp:281-297, :22-71)
        if (ctx->use_fused_upconv && ctx->upf[i].OC) {
            ConvCall u; u.kind = ZVX_K_UPCONV; u.stage = i; u.L = &ctx->upf[i]; u.x = vin; u.ldx = cin; u.rate_idx = i;
            u.pro_mode = PRO_LRELU; u.pro_slope = 0.1f; u.out32 = ctx->U; u.ldo32 = s * ch; u.out_mul = 1;
            if (vin16) { u.x = vin16; u.pro_mode = PRO_F16; }
            else if (vin2) { u.pro_mode = PRO_SUM3; u.x2 = vin2; u.x3 = vin3; u.sum_scale = third; u.sum3_half = vin_half; }
            u.flops = 2.0 * (double)ctx->last_frames * ctx->rates[i] * s * ch * cin * (ctx->up[i].K / s);
            if (chain) { u.out16 = ctx->U16; u.ldo16 = s * ch; u.out16_slope = 0.1f; }
            if (run_conv(ctx, u)) return 1;
        } else {
            const bool pre = vin2 && !vin16 && !ctx->use_ref_kernels;
            if (pre) {
                // one pass makes the fp16 operand lrelu(mean of the 3 branches) for all s phase launches
                const size_t n = (size_t)ctx->last_frames * ctx->rates[i] * cin;
                ctx->launches++;
                if (prof_begin(ctx, ZVX_K_NORM_AFFINE, i, 0.0, 14.0 * (double)n)) return 1;
                if (vin_half)
                    CK(ctx, sum3h_act_f16_launch(reinterpret_cast<const __half *>(vin), reinterpret_cast<const __half *>(vin2),
                                                 reinterpret_cast<const __half *>(vin3), third, 0.1f, n, ctx->H16, ctx->stream));
                else
                    CK(ctx, sum3_act_f16_launch(vin, vin2, vin3, third, 0.1f, n, ctx->H16, ctx->stream));
                if (prof_end(ctx)) return 1;
            }
            // the s output phases are independent launches (disjoint rows of U): dealt round-robin to the three streams
            cudaStream_t main_stream = ctx->stream;
            if (fork_ok) CK(ctx, cudaEventRecord(ctx->fork_ev, main_stream));
            bool used[2] = {false, false};
            int rc = 0;
            for (int phi = 0; phi < s && !rc; ++phi) {
                ConvCall u; u.kind = ZVX_K_UPCONV; u.stage = i; u.L = &ctx->up[i]; u.variant = phi; u.x = vin; u.ldx = cin; u.rate_idx = i;
                u.pro_mode = PRO_LRELU; u.pro_slope = 0.1f; u.out32 = ctx->U; u.ldo32 = ch; u.out_mul = s;
                if (vin16) { u.x = vin16; u.pro_mode = PRO_F16; }
                else if (pre) { u.x = ctx->H16; u.pro_mode = PRO_F16; }
                else if (vin2) { u.pro_mode = PRO_SUM3; u.x2 = vin2; u.x3 = vin3; u.sum_scale = third; u.sum3_half = vin_half; }
                if (chain) { u.out16 = ctx->U16; u.ldo16 = ch; u.out16_slope = 0.1f; }
                const int lane_id = fork_ok ? phi % 3 : 0;
                if (lane_id > 0) {
                    ctx->stream = ctx->fork_stream[lane_id - 1];
                    if (!used[lane_id - 1] && cudaStreamWaitEvent(ctx->stream, ctx->fork_ev, 0) != cudaSuccess) rc = 1;
                    used[lane_id - 1] = true;
                }
                if (!rc) rc = run_conv(ctx, u);
                ctx->stream = main_stream;
            }
            for (int j = 0; j < 2; ++j)
                if (used[j] && (cudaEventRecord(ctx->join_ev[j], ctx->fork_stream[j]) != cudaSuccess ||
                                cudaStreamWaitEvent(main_stream, ctx->join_ev[j], 0) != cudaSuccess))
                    rc = 1;
            if (rc) return ctx->err.empty() ? fail(ctx, "up-conv phase launch failed") : 1;
        }
        // MRF: three residual blocks on U, averaged (hifigan.cpp:300-315, :97-183).  When all three run as
        // fused chains, each writes its own output buffer and the branch sum / average is applied by the
        // consumer (next up-conv or the output conv, PRO_SUM3): the fused kernel's final phase is then pure
        // stores instead of a read-modify-write of the running sum.
        // The per-conv path (stage 0, 256 channels) does the same: block j keeps its residual stream in
        // branch buffer j, so no conv epilogue reads a running sum.
        const bool split = ctx->branch_sum_in_consumer && !ctx->use_ref_kernels && nb == 3;
        float *branch_out[3] = {ctx->CS, ctx->VA, ctx->VB};
        // stage hand-off: every block of the stage is a fused chain -> the last chain of the last block runs after the
        // other blocks, sums the three outputs in the reference's order and emits the consumer's fp16 operand
        bool handoff = split && ctx->stage_handoff && ctx->S16[0] && ctx->use_fused && ctx->debug_stop < 0;
        for (int j = 0; j < nb; ++j) handoff = handoff && ctx->fused[(size_t)i * nb + j].CH != 0;
        // fp16 branch outputs: every block of the stage is a fused chain -> its final phase writes y_j as fp16 and the consumer
        // sums three fp16 tensors (6 instead of 12 bytes per element; the sum is rounded to fp16 right after anyway)
        bool half_out = split && ctx->branch_f16 && !handoff && ctx->use_fused && ctx->debug_stop < 0;
        for (int j = 0; j < nb; ++j) half_out = half_out && ctx->fused[(size_t)i * nb + j].CH != 0;
        const FusedChain *deferred = nullptr;
        const float *deferred_yin = nullptr;
        int deferred_CH = 0, deferred_ncol = 0;
        // blocks 1 and 2 go to their own streams (own temporaries), joined before the consumer
        const bool fork = split && fork_ok;
        cudaStream_t main_stream = ctx->stream;
        if (fork) CK(ctx, cudaEventRecord(ctx->fork_ev, main_stream));
        for (int j = 0; j < nb; ++j) {
            const FusedBlock &fb = ctx->fused[(size_t)i * nb + j];
            struct StreamSwap {          // launches of this block go to ctx->stream: point it at the branch's stream
                zvx_ctx *c; cudaStream_t saved;
                ~StreamSwap() { c->stream = saved; }
            } swap_back = {ctx, main_stream};
            float *bY1 = ctx->Y1, *bT2 = ctx->T2;
            __half *bH16 = ctx->H16;
            if (fork && j > 0) {
                ctx->stream = ctx->fork_stream[j - 1];
                CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->fork_ev, 0));
                bY1 = ctx->fkY1[j - 1]; bT2 = ctx->fkT2[j - 1]; bH16 = ctx->fkH16[j - 1];
            }
            struct Join {                // ... and the main stream waits for it at the end of the iteration
                zvx_ctx *c; cudaStream_t main; cudaEvent_t ev; bool on;
                ~Join() { if (on) { cudaEventRecord(ev, c->stream); cudaStreamWaitEvent(main, ev, 0); } }
            } join = {ctx, main_stream, (fork && j > 0) ? ctx->join_ev[j - 1] : nullptr, fork && j > 0};
            if (ctx->use_fused && !ctx->use_ref_kernels && fb.CH) {
                // whole residual block (or chains of its conv pairs) on chip: mrf_fused.cu
                float *tmp[2] = {bY1, split ? bT2 : vout[(i + 1) & 1]};   // (the stage-input buffer is free after the up-conv)
                const float *yin = ctx->U;
                for (size_t q = 0; q < fb.chains.size(); ++q) {
                    const FusedChain &fc = fb.chains[q];
                    const bool lastc = q + 1 == fb.chains.size();
                    if (handoff && lastc && j == nb - 1) {        // launched after the join, below
                        deferred = &fc; deferred_yin = yin; deferred_CH = fb.CH; deferred_ncol = fb.ncol;
                        break;
                    }
                    mrf::Params fp;
                    memset(&fp, 0, sizeof fp);
                    fp.y_in = yin;
                    fp.in_slope = 0.1f;
                    fp.tbl0 = fc.tbl0;
                    fp.nlayers = fc.nlayers;
                    for (int l = 0; l < fc.nlayers; ++l) fp.L[l] = fc.layers[l];
                    fp.seg_start = ctx->d_seg;
                    fp.win_start = ctx->d_wins + (size_t)fc.wincfg * (ctx->cap_batch + 1);
                    fp.B = ctx->last_B;
                    fp.ncol = fb.ncol;
                    fp.prefetch = ctx->fused_prefetch >= 0 ? ctx->fused_prefetch : (fb.CH == 32 ? 1 : 0);
                    fp.flags = ctx->fused_flags;
                    fp.resident_ctas = ctx->fused_persistent ? ctx->num_sms * (fb.ncol == 128 ? ctx->fused_ncol128_ctas : 1) : 0;
                    fp.rate = ctx->rates[i + 1];
                    fp.halo = fc.halo;
                    fp.valid = fc.valid;
                    fp.err_flag = ctx->d_err;
                    if (!lastc) {
                        fp.out = tmp[q & 1];
                    } else if (split && half_out) {
                        fp.out16 = reinterpret_cast<uint16_t *>(branch_out[j]);     // y_j as fp16; summed by the consumer
                        fp.out16_slope = 1.0f;                                       // max(x, 1 * x) = x
                    } else if (split) {
                        fp.out = branch_out[j];                        // y_j; summed by the consumer
                    } else if (j == 0 && nb > 1) {
                        fp.out = ctx->CS;                              // cs = y_0
                    } else {
                        fp.acc_in = j > 0 ? ctx->CS : nullptr;         // cs = cs + y_j
                        if (j == nb - 1) { fp.out = vout[i & 1]; if (nb > 1) { fp.has_scale = 1; fp.scale = third; } }
                        else fp.out = ctx->CS;
                    }
                    ctx->launches++;
                    const double rows = (double)ctx->last_frames * fp.rate;
                    if (prof_begin(ctx, ZVX_K_MRF_CONV, i, rows * fc.flops_per_row, 0.0)) return 1;
                    CK(ctx, mrf_fused_launch(fb.CH, fp, ctx->total_wins[fc.wincfg], ctx->stream));
                    if (prof_end(ctx)) return 1;
                    yin = fp.out;                                      // (only chains that are not the last feed another one)
                }
                continue;
            }
            float *Y = split ? branch_out[j] : (j == 0) ? ctx->CS : bY1;
            for (int d = 0; d < nd; ++d) {
                const size_t idx = ((size_t)i * nb + j) * nd + d;
                const float *yin = d == 0 ? ctx->U : Y;
                ConvCall c1; c1.kind = ZVX_K_MRF_CONV; c1.stage = i; c1.L = &ctx->mrf1[idx]; c1.x = yin; c1.ldx = ch; c1.rate_idx = i + 1;
                c1.pro_mode = PRO_LRELU; c1.pro_slope = 0.1f;
                if (chain) { c1.x = d == 0 ? ctx->U16 : ctx->Y16[j]; c1.pro_mode = PRO_F16; }   // lrelu(.) already applied by the writer
                c1.out16 = bH16; c1.ldo16 = ch; c1.out16_slope = 0.1f;
                if (run_conv(ctx, c1)) return 1;
                ConvCall c2; c2.kind = ZVX_K_MRF_CONV; c2.stage = i; c2.L = &ctx->mrf2[idx]; c2.x = bH16; c2.ldx = ch; c2.rate_idx = i + 1; c2.pro_mode = PRO_F16;
                c2.res = yin; c2.ldres = ch;
                const bool last = d == nd - 1;
                if (last && j > 0 && !split) {
                    c2.acc_in = ctx->CS;                       // cs = cs + y_j
                    if (j == nb - 1) { c2.scale = third; c2.out32 = vout[i & 1]; }   // c = cs / num_blocks
                    else c2.out32 = ctx->CS;
                } else {
                    c2.out32 = Y;
                }
                c2.ldo32 = ch;
                if (chain && !last) { c2.out16 = ctx->Y16[j]; c2.ldo16 = ch; c2.out16_slope = 0.1f; }
                if (run_conv(ctx, c2)) return 1;
            }
        }
        if (deferred) {
            const FusedChain &fc = *deferred;
            mrf::Params fp;
            memset(&fp, 0, sizeof fp);
            fp.y_in = deferred_yin;
            fp.in_slope = 0.1f;
            fp.tbl0 = fc.tbl0;
            fp.nlayers = fc.nlayers;
            for (int l = 0; l < fc.nlayers; ++l) fp.L[l] = fc.layers[l];
            fp.seg_start = ctx->d_seg;
            fp.win_start = ctx->d_wins + (size_t)fc.wincfg * (ctx->cap_batch + 1);
            fp.B = ctx->last_B;
            fp.ncol = deferred_ncol;
            fp.prefetch = ctx->fused_prefetch >= 0 ? ctx->fused_prefetch : (deferred_CH == 32 ? 1 : 0);
            fp.flags = ctx->fused_flags;
            fp.resident_ctas = ctx->fused_persistent ? ctx->num_sms * (deferred_ncol == 128 ? ctx->fused_ncol128_ctas : 1) : 0;
            fp.rate = ctx->rates[i + 1];
            fp.halo = fc.halo;
            fp.valid = fc.valid;
            fp.err_flag = ctx->d_err;
            fp.acc_in = branch_out[0];                     // c = ((y_0 + y_1) + y_2) / 3   (hifigan.cpp:300-315)
            fp.acc_in2 = branch_out[1];
            fp.has_scale = 1; fp.scale = third;
            fp.out = nullptr;
            fp.out16 = reinterpret_cast<uint16_t *>(ctx->S16[i & 1]);
            fp.out16_slope = i + 1 == c.num_upsamples ? 0.01f : 0.1f;      // leaky_relu of the consumer (hifigan.cpp:281 / :324)
            ctx->launches++;
            const double rows = (double)ctx->last_frames * fp.rate;
            if (prof_begin(ctx, ZVX_K_MRF_CONV, i, rows * fc.flops_per_row, 0.0)) return 1;
            CK(ctx, mrf_fused_launch(deferred_CH, fp, ctx->total_wins[fc.wincfg], ctx->stream));
            if (prof_end(ctx)) return 1;
            vin = vin2 = vin3 = nullptr;
            vin16 = ctx->S16[i & 1];
            vin_half = false;
        } else if (split) { vin = branch_out[0]; vin2 = branch_out[1]; vin3 = branch_out[2]; vin16 = nullptr; vin_half = half_out; ctx->stage_is_split[i] = half_out ? 2 : 1; }
        else { vin = vout[i & 1]; vin2 = vin3 = nullptr; vin16 = nullptr; vin_half = false; }
    }
    if (ctx->debug_stop >= 0) return 0;
    // leaky_relu(0.01) -> output_conv -> tanh (hifigan.cpp:324-345)
    const int last = c.num_upsamples;
    ctx->launches++;
    {
        const double rows = (double)ctx->last_frames * ctx->rates[last];
        if (prof_begin(ctx, ZVX_K_OUT_CONV, last, 2.0 * rows * ctx->chans[last] * ctx->output_conv.K,
                       rows * (ctx->chans[last] + 1) * sizeof(float)))
            return 1;
    }
    CK(ctx, out_conv_launch(vin, vin2, vin3, vin16, vin_half ? 1 : 0, third, ctx->chans[last], ctx->output_conv.K, ctx->output_conv.raw, ctx->output_conv.bias,
                            ctx->out_w_kc.empty() ? nullptr : ctx->out_w_kc.data(), ctx->out_b_host, 0.01f,
                            ctx->d_seg, ctx->d_tiles + (size_t)last * (ctx->cap_batch + 1), ctx->last_B, ctx->rates[last],
                            ctx->total_tiles[last], wav_out, pcm_out, ctx->stream));
    return prof_end(ctx);
}

int check_device_error(zvx_ctx *ctx)
{
    // A bounded barrier wait that expired traps (ptx_sm100.cuh: mbar_wait): the synchronise below then reports a
    // sticky launch failure and the primary context is gone -- say so, the caller cannot recover in this process.
    const cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (es != cudaSuccess)
        return fail(ctx, "cudaStreamSynchronize failed: %s (fatal: the CUDA context is lost, restart the process)", cudaGetErrorString(es));
    int flag = 0;
    CK(ctx, cudaMemcpy(&flag, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) return fail(ctx, "device pipeline timeout flag set");
    return 0;
}

// Run `body` (kernel launches on ctx->stream only, no synchronisation) for a single utterance of L frames: the
// first time it is captured into a CUDA graph, afterwards the graph is replayed (one launch instead of ~100).
template <class F>
int run_graphed(zvx_ctx *ctx, int kind, int L, F body)
{
    if (!ctx->use_graphs || ctx->prof || ctx->debug_stop >= 0 || ctx->use_ref_kernels) return body();
    const int flags = ctx->use_fused;
    for (size_t i = 0; i < ctx->graphs.size(); ++i) {
        const zvx_ctx::GraphEntry g = ctx->graphs[i];
        if (g.kind == kind && g.L == L && g.flags == flags) {
            // least-recently-used order: a hit moves to the back, eviction takes the front
            ctx->graphs.erase(ctx->graphs.begin() + (long)i);
            ctx->graphs.push_back(g);
            CK(ctx, cudaGraphLaunch(g.exec, ctx->stream));
            ctx->launches += g.launches;
            return 0;
        }
    }
    // a length seen for the first time runs directly: natural variable-length traffic would otherwise pay capture +
    // instantiation of ~100 nodes on almost every call and evict the graphs of the lengths that do repeat
    if (ctx->graph_seen.size() > 4096) ctx->graph_seen.clear();
    if (ctx->graph_seen[std::make_pair(kind, L)]++ == 0) return body();
    const int64_t l0 = ctx->launches;
    CK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        if (!rc) return fail(ctx, "CUDA graph capture failed: %s", cudaGetErrorString(e));
        return 1;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return fail(ctx, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
    if (ctx->graphs.size() >= 16) { cudaGraphExecDestroy(ctx->graphs.front().exec); ctx->graphs.erase(ctx->graphs.begin()); }
    ctx->graphs.push_back({kind, L, flags, exec, ctx->launches - l0});
    CK(ctx, cudaGraphLaunch(exec, ctx->stream));
    return 0;
}

}  // namespace

// A lane borrows the weights (device pointers) of its parent and owns only a stream, the utterance
// tables and a workspace: zvx_synth_batch alternates sub-batches between the context and its lane so
// that the H2D / D2H copies of one sub-batch run under the kernels of the other.
int make_lane(zvx_ctx *parent)
{
    if (parent->lane) return 0;
    zvx_ctx *l = new zvx_ctx(*parent);
    l->is_lane = true;
    l->lane = nullptr;
    l->owned.clear();
    l->stream = nullptr;
    l->tables_event = nullptr;
    l->tables_pending = false;
    l->launches = 0;
    l->prof = false;
    l->ev_pool.clear();
    l->ev_used = 0;
    l->prof_entries.clear();
    l->cap_frames = 0;
    l->cap_batch = 0;
    float **fp[] = {&l->enc_in, &l->sc, &l->h528, &l->e0, &l->h1056, &l->catA, &l->catB, &l->asr, &l->d1, &l->d2, &l->mel, &l->style,
                    &l->mu, &l->rstd, &l->adain_gb, &l->v0, &l->U, &l->CS, &l->Y1, &l->VA, &l->VB, &l->T2, &l->wav, &l->pin_in, &l->pin_out};
    for (float **q : fp) *q = nullptr;
    l->H16 = l->X16 = l->R16 = nullptr;
    l->U16 = nullptr; l->Y16[0] = l->Y16[1] = l->Y16[2] = nullptr; l->chain_elems = 0;
    l->stat_part = nullptr; l->stat_part_cap = 0; l->asr_mu = l->asr_rstd = nullptr;
    l->S16[0] = l->S16[1] = nullptr;
    l->d_seg = l->d_tiles = l->d_wins = l->d_err = l->pin_tables = nullptr;
    l->pin_in_cap = l->pin_out_cap = 0;
    l->graphs.clear();           // the copies of the parent's graph handles are not the lane's to destroy
    l->graph_seen.clear();
    for (int j = 0; j < 2; ++j) { l->fork_stream[j] = nullptr; l->join_ev[j] = nullptr; l->fkY1[j] = l->fkT2[j] = nullptr; l->fkH16[j] = nullptr; }
    l->fork_ev = nullptr;
    l->feat = nullptr; l->feat_tab = nullptr; l->feat_cap = l->feat_tab_cap = 0;
    // the lane is published only once it is complete: a half-built lane (null stream) must never be enqueued on
    auto build = [&]() -> int {
        zvx_ctx *ctx = parent;   // error reporting goes to the parent
        CK(ctx, cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
        CK(ctx, cudaEventCreateWithFlags(&l->tables_event, cudaEventDisableTiming));
        if (dev_alloc(l, &l->d_err, 1)) { ctx->err = l->err; return 1; }
        CK(ctx, cudaMemset(l->d_err, 0, sizeof(int)));
        return 0;
    };
    if (build()) { zvx_destroy(l); return 1; }
    parent->lane = l;
    return 0;
}

// ====================================================================== C ABI
extern "C" {

void zvx_default_config(zvx_config *cfg)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->device = 0;
    cfg->dim_in = 528;
    cfg->style_dim = 528;
    cfg->residual_dim = 64;
    cfg->num_mels = 80;
    cfg->hop_size = 300;
    cfg->kernel_size = 7;
    cfg->num_upsamples = 4;
    const int s[4] = {5, 5, 4, 3};
    for (int i = 0; i < 4; ++i) cfg->upsample_scales[i] = s[i];
    cfg->num_resblocks = 3;
    cfg->num_resblock_dilations = 3;
    const int d[9] = {1, 3, 5, 1, 3, 5, 1, 3, 5};
    for (int i = 0; i < 9; ++i) cfg->resblock_dilations[i] = d[i];
    cfg->with_decoder = 1;
    cfg->with_vocoder = 1;
}

const char *zvx_last_error(const zvx_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

void zvx_destroy(zvx_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->lane) { zvx_destroy(ctx->lane); ctx->lane = nullptr; }
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    for (int j = 0; j < 2; ++j) {
        if (ctx->fork_stream[j]) cudaStreamDestroy(ctx->fork_stream[j]);
        if (ctx->join_ev[j]) cudaEventDestroy(ctx->join_ev[j]);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    ctx->ev_pool.clear();
    for (int j = 0; j < 2; ++j)
        if (ctx->split_ev[j] && !ctx->is_lane) cudaEventDestroy(ctx->split_ev[j]);
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->pin_tables) cudaFreeHost(ctx->pin_tables);
    if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    if (ctx->tables_event) cudaEventDestroy(ctx->tables_event);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int zvx_create(zvx_ctx **out, const zvx_config *cfg, const zvx_tensor_desc *weights, int32_t n_weights)
{
    if (!out || !cfg) return fail(nullptr, "zvx_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, "zvx_create: no CUDA device (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "zvx_create: device %d out of range", cfg->device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, "zvx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device,
                    prop.major, prop.minor);
    if (cfg->num_upsamples < 0 || cfg->num_upsamples > 8 || cfg->num_resblocks * cfg->num_resblock_dilations > 32)
        return fail(nullptr, "zvx_create: unsupported topology");

    zvx_ctx *ctx = new zvx_ctx();
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    if (const char *e = getenv("ZVX_FUSED_MIN_EFF")) ctx->fused_min_eff = atof(e);
    if (const char *e = getenv("ZVX_FUSED_PREFETCH")) ctx->fused_prefetch = atoi(e);
    if (const char *e = getenv("ZVX_NCOL128_CTAS")) ctx->fused_ncol128_ctas = atoi(e);
    if (const char *e = getenv("ZVX_FUSED_PERSISTENT")) ctx->fused_persistent = atoi(e);
    if (const char *e = getenv("ZVX_FUSED_FLAGS")) ctx->fused_flags = atoi(e);
    if (const char *e = getenv("ZVX_E2E_CHUNKS")) ctx->e2e_chunks = std::max(1, atoi(e));
    if (const char *e = getenv("ZVX_DEC_PREPASS")) ctx->dec_prepass = atoi(e);
    if (const char *e = getenv("ZVX_BRANCH_SUM_IN_CONSUMER")) ctx->branch_sum_in_consumer = atoi(e);
    if (const char *e = getenv("ZVX_CONV_MT2")) ctx->conv_mt2 = atoi(e);
    if (const char *e = getenv("ZVX_UPCONV_MT2")) ctx->upconv_mt2 = atoi(e);
    if (const char *e = getenv("ZVX_CONV_CLUSTER")) ctx->conv_cluster = atoi(e);
    if (const char *e = getenv("ZVX_CONV_TMA")) ctx->conv_tma = atoi(e);
    if (const char *e = getenv("ZVX_CONV_PAIR")) ctx->conv_pair = atoi(e);
    if (const char *e = getenv("ZVX_CONV_EPI8")) ctx->conv_epi8 = atoi(e);
    if (const char *e = getenv("ZVX_CONV_FOLD")) ctx->conv_fold = atoi(e);
    if (const char *e = getenv("ZVX_BRANCH_F16")) ctx->branch_f16 = atoi(e);
    if (const char *e = getenv("ZVX_MRF_F16_CHAIN")) ctx->mrf_f16_chain = atoi(e);
    if (const char *e = getenv("ZVX_FUSED_STATS")) ctx->fused_stats = atoi(e);
    if (const char *e = getenv("ZVX_STAGE_HANDOFF")) ctx->stage_handoff = atoi(e);
    if (const char *e = getenv("ZVX_CONV_SMEM_KB")) ctx->conv_smem_kb = atoi(e);
    if (const char *e = getenv("ZVX_GRAPHS")) ctx->use_graphs = atoi(e);
    if (const char *e = getenv("ZVX_FORK_BRANCHES")) ctx->fork_branches = atoi(e);
    if (const char *e = getenv("ZVX_DEVICE_SPLIT")) ctx->device_split = atoi(e);
    if (const char *e = getenv("ZVX_H2D_CHAIN")) ctx->h2d_chain = atoi(e);
    if (const char *e = getenv("ZVX_CHUNK_GROUP_MAX")) ctx->chunk_group_max = std::max(1, atoi(e));
    if (const char *e = getenv("ZVX_CONV_PERSISTENT")) ctx->conv_persistent = atoi(e);
    if (const char *e = getenv("ZVX_FUSED_UPCONV")) ctx->use_fused_upconv = atoi(e);
    ctx->num_sms = prop.multiProcessorCount;
    auto bail = [&](void) { g_create_error = ctx->err; zvx_destroy(ctx); return 1; };
#define CKC(call)                                                                              \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            fail(ctx, "%s failed: %s", #call, cudaGetErrorString(e__));                        \
            return bail();                                                                     \
        }                                                                                      \
    } while (0)
    CKC(cudaSetDevice(ctx->device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&ctx->tables_event, cudaEventDisableTiming));
    CKC(conv_umma_init());
    CKC(mrf_fused_init());
    if (dev_alloc(ctx, &ctx->d_err, 1)) return bail();
    CKC(cudaMemset(ctx->d_err, 0, sizeof(int)));

    HostW hw;
    for (int i = 0; i < n_weights; ++i) {
        const zvx_tensor_desc &t = weights[i];
        if (!t.name || !t.data) { fail(ctx, "weight %d has null name/data", i); return bail(); }
        if (t.dtype != ZVX_F32 && t.dtype != ZVX_F16) { fail(ctx, "tensor '%s': unsupported dtype %d", t.name, t.dtype); return bail(); }
        DevTensor d;
        d.dtype = t.dtype;
        d.nd = t.n_dims;
        size_t n = 1;
        for (int k = 0; k < 4; ++k) { d.ne[k] = k < t.n_dims ? t.ne[k] : 1; n *= (size_t)d.ne[k]; }
        d.nbytes = n * (t.dtype == ZVX_F16 ? 2 : 4);
        void *p = nullptr;
        CKC(cudaMalloc(&p, d.nbytes));
        ctx->owned.push_back(p);
        CKC(cudaMemcpy(p, t.data, d.nbytes, cudaMemcpyHostToDevice));
        d.d = p;
        ctx->W[t.name] = d;
        if (t.dtype == ZVX_F16) {
            const __half *h = reinterpret_cast<const __half *>(t.data);
            hw.h[t.name].assign(h, h + n);
        } else if (n <= 4096) {
            const float *f = reinterpret_cast<const float *>(t.data);
            hw.f[t.name].assign(f, f + n);
        }
    }
    if (cfg->with_decoder && build_decoder(ctx, hw)) return bail();
    if (cfg->with_vocoder && build_vocoder(ctx, hw)) return bail();
    if (!cfg->with_vocoder) { ctx->rates.assign(1, 1); ctx->chans.assign(1, 0); }
    // 256-row tilings (conv_umma with two M-tiles per CTA) for every rate
    for (int r = 0; r < (int)ctx->rates.size(); ++r) {
        int w = -1;
        for (size_t q = 0; q < ctx->wincfg.size(); ++q)
            if (ctx->wincfg[q].rate_idx == r && ctx->wincfg[q].valid == 256) w = (int)q;
        if (w < 0) { ctx->wincfg.push_back({r, 256}); w = (int)ctx->wincfg.size() - 1; }
        ctx->tile256_cfg.push_back(w);
    }
    if (reserve(ctx, 512, 8)) return bail();
    CKC(cudaStreamSynchronize(ctx->stream));
#undef CKC
    *out = ctx;
    return 0;
}

void *zvx_stream(zvx_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int zvx_synchronize(zvx_ctx *ctx)
{
    if (!ctx) return 1;
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->lane && check_device_error(ctx->lane)) { ctx->err = ctx->lane->err; return 1; }
    return check_device_error(ctx);
}

int64_t zvx_kernel_launches(const zvx_ctx *ctx) { return ctx ? ctx->launches + (ctx->lane ? ctx->lane->launches : 0) : 0; }

int zvx_reserve(zvx_ctx *ctx, int64_t total_frames, int32_t max_batch)
{
    if (!ctx) return 1;
    CK(ctx, cudaSetDevice(ctx->device));
    return reserve(ctx, total_frames, max_batch);
}

int zvx_profile_begin(zvx_ctx *ctx)
{
    if (!ctx) return 1;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prof = true;
    ctx->ev_used = 0;
    ctx->prof_entries.clear();
    return 0;
}

int64_t zvx_profile_end(zvx_ctx *ctx, zvx_launch_record *recs, int64_t max_recs)
{
    if (!ctx) return -1;
    ctx->prof = false;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        fail(ctx, "zvx_profile_end: synchronize failed");
        return -1;
    }
    const int64_t n = (int64_t)ctx->prof_entries.size();
    for (int64_t i = 0; i < n && i < max_recs; ++i) {
        const zvx_ctx::ProfEntry &e = ctx->prof_entries[i];
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[e.ev], ctx->ev_pool[e.ev + 1]) != cudaSuccess) {
            fail(ctx, "cudaEventElapsedTime failed");
            return -1;
        }
        recs[i].kind = e.kind;
        recs[i].stage = e.stage;
        recs[i].flops = e.flops;
        recs[i].bytes = e.bytes;
        recs[i].ms = ms;
    }
    return n;
}

void zvx_set_debug_kernels(zvx_ctx *ctx, int32_t v) { if (ctx) ctx->use_ref_kernels = v; }
void zvx_set_debug_stop(zvx_ctx *ctx, int32_t s) { if (ctx) ctx->debug_stop = s; }
void zvx_set_fused_mrf(zvx_ctx *ctx, int32_t on) { if (ctx) ctx->use_fused = on; }

int zvx_synth_batch_device(zvx_ctx *ctx, int32_t B, const float *d_enc, const float *d_style, const int32_t *L,
                           float *d_mel, float *d_wav, int32_t sync)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_decoder || !ctx->cfg.with_vocoder) return fail(ctx, "context was built without decoder or vocoder");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->device_split && B >= 8 && !ctx->prof && ctx->debug_stop < 0) {
        // two half batches, the second on the lane (own stream + workspace): its HBM-bound small kernels can run under the
        // first half's tensor-bound persistent kernels and vice versa; the context's stream stays the ordering point
        int64_t frames = 0, acc = 0;
        for (int b = 0; b < B; ++b) frames += L[b];
        int b0 = 0;
        while (b0 < B - 1 && 2 * (acc + L[b0]) <= frames) acc += L[b0++];
        if (b0 > 0) {
            if (make_lane(ctx)) return 1;
            zvx_ctx *ln = ctx->lane;
            ln->use_ref_kernels = ctx->use_ref_kernels; ln->use_fused = ctx->use_fused;
            for (int j = 0; j < 2; ++j)
                if (!ctx->split_ev[j]) CK(ctx, cudaEventCreateWithFlags(&ctx->split_ev[j], cudaEventDisableTiming));
            const zvx_config &cfg = ctx->cfg;
            CK(ctx, cudaEventRecord(ctx->split_ev[0], ctx->stream));
            // first half on the context
            if (set_batch(ctx, b0, L)) return 1;
            CK(ctx, cudaMemcpyAsync(ctx->enc_in, d_enc, sizeof(float) * (size_t)acc * cfg.dim_in, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(ctx, cudaMemcpyAsync(ctx->style, d_style, sizeof(float) * (size_t)b0 * cfg.style_dim, cudaMemcpyDeviceToDevice, ctx->stream));
            float *melA = d_mel ? d_mel : ctx->mel;
            if (run_decoder(ctx, melA)) return 1;
            if (run_vocoder(ctx, melA, d_wav)) return 1;
            // second half on the lane, after whatever the caller had queued on the context's stream
            if (set_batch(ln, B - b0, L + b0)) { ctx->err = ln->err; return 1; }
            CK(ctx, cudaStreamWaitEvent(ln->stream, ctx->split_ev[0], 0));
            CK(ctx, cudaMemcpyAsync(ln->enc_in, d_enc + (size_t)acc * cfg.dim_in, sizeof(float) * (size_t)(frames - acc) * cfg.dim_in,
                                    cudaMemcpyDeviceToDevice, ln->stream));
            CK(ctx, cudaMemcpyAsync(ln->style, d_style + (size_t)b0 * cfg.style_dim, sizeof(float) * (size_t)(B - b0) * cfg.style_dim,
                                    cudaMemcpyDeviceToDevice, ln->stream));
            float *melB = d_mel ? d_mel + (size_t)acc * cfg.num_mels : ln->mel;
            if (run_decoder(ln, melB) || run_vocoder(ln, melB, d_wav + (size_t)acc * cfg.hop_size)) { ctx->err = ln->err; return 1; }
            CK(ctx, cudaEventRecord(ctx->split_ev[1], ln->stream));
            CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->split_ev[1], 0));
            if (!sync) return 0;
            if (check_device_error(ln)) { ctx->err = ln->err; return 1; }
            return check_device_error(ctx);
        }
    }
    if (set_batch(ctx, B, L)) return 1;
    const int64_t F = ctx->last_frames;
    // the schedule reads its inputs from the workspace: D2D copies keep the public pointers const
    CK(ctx, cudaMemcpyAsync(ctx->enc_in, d_enc, sizeof(float) * F * ctx->cfg.dim_in, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(ctx->style, d_style, sizeof(float) * (size_t)B * ctx->cfg.style_dim, cudaMemcpyDeviceToDevice, ctx->stream));
    float *mel = d_mel ? d_mel : ctx->mel;
    if (run_decoder(ctx, mel)) return 1;
    if (run_vocoder(ctx, mel, d_wav)) return 1;
    return sync ? check_device_error(ctx) : 0;
}

int zvx_vocode_batch_device(zvx_ctx *ctx, int32_t B, const float *d_mel, const int32_t *L, float *d_wav, int32_t sync)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_vocoder) return fail(ctx, "context was built without vocoder");
    CK(ctx, cudaSetDevice(ctx->device));
    if (set_batch(ctx, B, L)) return 1;
    if (run_vocoder(ctx, d_mel, d_wav)) return 1;
    return sync ? check_device_error(ctx) : 0;
}

// one sub-batch [b0, b1) of zvx_synth_batch on context / lane `c`, everything asynchronous on c->stream
// (wav: float samples; or pcm: 16-bit samples converted by the output conv itself, staged in the same device buffer)
// wait_ev / done_ev (may be null): input copies start after wait_ev and done_ev is recorded behind them, so that the input
// copies of consecutive sub-batches do not share the PCIe link (the first sub-batch can start computing earlier)
static int synth_chunk(zvx_ctx *c, int b0, int b1, const float *const *enc_seq, const float *const *style, const int32_t *L,
                       float *const *mel, float *const *wav, int16_t *const *pcm = nullptr, cudaEvent_t wait_ev = nullptr,
                       cudaEvent_t done_ev = nullptr)
{
    zvx_ctx *ctx = c;
    const zvx_config &cfg = c->cfg;
    const int n = b1 - b0;
    if (set_batch(c, n, L + b0)) return 1;
    if (wait_ev) CK(ctx, cudaStreamWaitEvent(c->stream, wait_ev, 0));
    // one copy per run of utterances whose host buffers are contiguous (the common packed layout)
    for (int b = 0; b < n;) {
        int e = b + 1;
        while (e < n && enc_seq[b0 + e] == enc_seq[b0 + e - 1] + (size_t)L[b0 + e - 1] * cfg.dim_in) ++e;
        CK(ctx, cudaMemcpyAsync(c->enc_in + (size_t)c->h_seg[b] * cfg.dim_in, enc_seq[b0 + b],
                                sizeof(float) * (size_t)(c->h_seg[e] - c->h_seg[b]) * cfg.dim_in, cudaMemcpyHostToDevice, c->stream));
        b = e;
    }
    for (int b = 0; b < n;) {
        int e = b + 1;
        while (e < n && style[b0 + e] == style[b0 + e - 1] + cfg.style_dim) ++e;
        CK(ctx, cudaMemcpyAsync(c->style + (size_t)b * cfg.style_dim, style[b0 + b], sizeof(float) * (size_t)(e - b) * cfg.style_dim,
                                cudaMemcpyHostToDevice, c->stream));
        b = e;
    }
    if (done_ev) CK(ctx, cudaEventRecord(done_ev, c->stream));
    // a single utterance (the latency path: ZeroVOXModel-style callers) replays decoder + vocoder from one CUDA graph per
    // length, like zvx_decode / zvx_vocode; batches launch kernel by kernel
    int16_t *d_pcm = reinterpret_cast<int16_t *>(c->wav);
    auto compute = [&]() -> int {
        if (run_decoder(c, c->mel)) return 1;
        return pcm ? run_vocoder(c, c->mel, nullptr, d_pcm) : run_vocoder(c, c->mel, c->wav);
    };
    if (n == 1 ? run_graphed(c, pcm ? 4 : 3, (int)L[b0], compute) : compute()) return 1;
    if (pcm) {
        for (int b = 0; b < n;) {
            int e = b + 1;
            while (e < n && pcm[b0 + e] == pcm[b0 + e - 1] + (size_t)L[b0 + e - 1] * cfg.hop_size) ++e;
            CK(ctx, cudaMemcpyAsync(pcm[b0 + b], d_pcm + (size_t)c->h_seg[b] * cfg.hop_size,
                                    sizeof(int16_t) * (size_t)(c->h_seg[e] - c->h_seg[b]) * cfg.hop_size, cudaMemcpyDeviceToHost, c->stream));
            b = e;
        }
    } else {
        for (int b = 0; b < n;) {
            int e = b + 1;
            while (e < n && wav[b0 + e] == wav[b0 + e - 1] + (size_t)L[b0 + e - 1] * cfg.hop_size) ++e;
            CK(ctx, cudaMemcpyAsync(wav[b0 + b], c->wav + (size_t)c->h_seg[b] * cfg.hop_size,
                                    sizeof(float) * (size_t)(c->h_seg[e] - c->h_seg[b]) * cfg.hop_size, cudaMemcpyDeviceToHost, c->stream));
            b = e;
        }
    }
    for (int b = 0; b < n; ++b)
        if (mel && mel[b0 + b])
            CK(ctx, cudaMemcpyAsync(mel[b0 + b], c->mel + (size_t)c->h_seg[b] * cfg.num_mels, sizeof(float) * (size_t)L[b0 + b] * cfg.num_mels,
                                    cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

static int synth_batch_impl(zvx_ctx *ctx, int32_t B, const float *const *enc_seq, const float *const *style, const int32_t *L,
                            float *const *mel, float *const *wav, int16_t *const *pcm, bool wait = true)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_decoder || !ctx->cfg.with_vocoder) return fail(ctx, "context was built without decoder or vocoder");
    if (!enc_seq || !style || !L || (!wav && !pcm)) return fail(ctx, "zvx_synth_batch: null argument");
    if (B <= 0) return fail(ctx, "empty batch");
    CK(ctx, cudaSetDevice(ctx->device));
    int64_t frames = 0;
    for (int b = 0; b < B; ++b) {
        if (L[b] <= 0) return fail(ctx, "utterance %d has non-positive length %d", b, L[b]);
        frames += L[b];
    }
    // Large batches are cut into a few sub-batches that alternate between this context and its lane
    // (own stream + workspace, shared weights): the PCIe copies of one sub-batch overlap the kernels of
    // the other.  Utterances are independent, so the result does not depend on the cut.
    int nch = ctx->e2e_chunks;
    if (ctx->prof || ctx->debug_stop >= 0 || B < 2 * nch || frames < 4096) nch = 1;
    if (nch == 1) {
        if (synth_chunk(ctx, 0, B, enc_seq, style, L, mel, wav, pcm)) return 1;
        return wait ? check_device_error(ctx) : 0;
    }
    if (make_lane(ctx)) return 1;
    ctx->lane->use_ref_kernels = ctx->use_ref_kernels;
    ctx->lane->use_fused = ctx->use_fused;
    // cut points with roughly equal frame counts
    std::vector<int> cut(1, 0);
    int64_t acc = 0;
    for (int b = 0; b < B; ++b) {
        acc += L[b];
        if ((int)cut.size() < nch && acc * nch >= frames * (int64_t)cut.size() && b + 1 < B) cut.push_back(b + 1);
    }
    cut.push_back(B);
    // size both workspaces once for the largest sub-batch (a growing reserve() would have to synchronise)
    int64_t maxf = 0; int maxb = 0;
    for (size_t q = 0; q + 1 < cut.size(); ++q) {
        int64_t f = 0;
        for (int b = cut[q]; b < cut[q + 1]; ++b) f += L[b];
        maxf = std::max(maxf, f);
        maxb = std::max(maxb, cut[q + 1] - cut[q]);
    }
    if (reserve(ctx, maxf, maxb)) return 1;
    if (reserve(ctx->lane, maxf, maxb)) { ctx->err = ctx->lane->err; return 1; }
    for (int j = 0; j < 2; ++j)
        if (!ctx->split_ev[j]) CK(ctx, cudaEventCreateWithFlags(&ctx->split_ev[j], cudaEventDisableTiming));
    const bool chain = ctx->h2d_chain != 0;
    for (size_t q = 0; q + 1 < cut.size(); ++q) {
        zvx_ctx *c = (q & 1) ? ctx->lane : ctx;
        if (synth_chunk(c, cut[q], cut[q + 1], enc_seq, style, L, mel, wav, pcm, (chain && q > 0) ? ctx->split_ev[(q - 1) & 1] : nullptr,
                        chain ? ctx->split_ev[q & 1] : nullptr)) {
            if (c != ctx) ctx->err = c->err;
            // copies into the caller's buffers may still be in flight on either stream: drain both before reporting
            cudaStreamSynchronize(ctx->stream);
            cudaStreamSynchronize(ctx->lane->stream);
            return 1;
        }
    }
    if (!wait) return 0;
    if (check_device_error(ctx->lane)) { ctx->err = ctx->lane->err; return 1; }
    return check_device_error(ctx);
}

int zvx_synth_batch(zvx_ctx *ctx, int32_t B, const float *const *enc_seq, const float *const *style, const int32_t *L,
                    float *const *mel, float *const *wav)
{
    if (ctx && !wav) return fail(ctx, "zvx_synth_batch: null argument");
    return synth_batch_impl(ctx, B, enc_seq, style, L, mel, wav, nullptr);
}

// Pipelined form: zvx_synth_batch_submit enqueues the copies and kernels of one batch and returns; the caller's buffers
// (inputs AND outputs) must stay untouched until zvx_synth_batch_wait has returned.  Consecutive submits keep both copy
// engines and the SMs busy: a batch's sub-batches alternate between the context's two streams, so the D2H of batch i
// runs under the H2D and the kernels of batch i + 1 instead of a host synchronisation in between.
int zvx_synth_batch_submit(zvx_ctx *ctx, int32_t B, const float *const *enc_seq, const float *const *style, const int32_t *L,
                           float *const *mel, float *const *wav, int16_t *const *pcm)
{
    if (ctx && (!wav == !pcm)) return fail(ctx, "zvx_synth_batch_submit: exactly one of wav / pcm");
    return synth_batch_impl(ctx, B, enc_seq, style, L, mel, wav, pcm, false);
}

int zvx_synth_batch_wait(zvx_ctx *ctx)
{
    if (!ctx) return 1;
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->lane && check_device_error(ctx->lane)) { ctx->err = ctx->lane->err; return 1; }
    return check_device_error(ctx);
}

int zvx_synth_batch_pcm16(zvx_ctx *ctx, int32_t B, const float *const *enc_seq, const float *const *style, const int32_t *L,
                          float *const *mel, int16_t *const *pcm)
{
    if (ctx && !pcm) return fail(ctx, "zvx_synth_batch_pcm16: null argument");
    return synth_batch_impl(ctx, B, enc_seq, style, L, mel, nullptr, pcm);
}

int zvx_vocode_pcm16(zvx_ctx *ctx, const float *mel, int32_t L, int16_t *pcm)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_vocoder) return fail(ctx, "context was built without the vocoder");
    if (!mel || !pcm) return fail(ctx, "zvx_vocode_pcm16: null argument");
    const zvx_config &c = ctx->cfg;
    CK(ctx, cudaSetDevice(ctx->device));
    if (set_batch(ctx, 1, &L)) return 1;
    CK(ctx, cudaMemcpyAsync(ctx->mel, mel, sizeof(float) * (size_t)L * c.num_mels, cudaMemcpyHostToDevice, ctx->stream));
    int16_t *d_pcm = reinterpret_cast<int16_t *>(ctx->wav);
    if (run_graphed(ctx, 2, L, [&]() { return run_vocoder(ctx, ctx->mel, nullptr, d_pcm); })) return 1;
    CK(ctx, cudaMemcpyAsync(pcm, d_pcm, sizeof(int16_t) * (size_t)L * c.hop_size, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

// Minimal RIFF/WAVE writer, mono 16-bit PCM: what sf_open(SFM_WRITE, SF_FORMAT_WAV | SF_FORMAT_PCM_16) +
// sf_write + sf_close produce for the reference's write_wav_file (zerovox.cpp:354-384): "RIFF" size "WAVE",
// "fmt " 16 {1, 1, rate, 2 rate, 2, 16}, "data" 2 n, little-endian samples.  Host only, no CUDA involved.
int zvx_write_wav_pcm16(const char *path, const int16_t *pcm, int64_t n_samples, int32_t sample_rate)
{
    if (!path || (!pcm && n_samples > 0) || n_samples < 0 || sample_rate <= 0) return 1;
    if (n_samples > (int64_t)0x7FFFFFD0 / 2) return 1;      // RIFF sizes are 32-bit
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    auto u32 = [](uint8_t *p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; };
    auto u16 = [](uint8_t *p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; };
    const uint32_t data_bytes = (uint32_t)(2 * n_samples);
    uint8_t h[44];
    memcpy(h, "RIFF", 4); u32(h + 4, 36u + data_bytes); memcpy(h + 8, "WAVE", 4);
    memcpy(h + 12, "fmt ", 4); u32(h + 16, 16u); u16(h + 20, 1u); u16(h + 22, 1u);
    u32(h + 24, (uint32_t)sample_rate); u32(h + 28, 2u * (uint32_t)sample_rate); u16(h + 32, 2u); u16(h + 34, 16u);
    memcpy(h + 36, "data", 4); u32(h + 40, data_bytes);
    bool ok = fwrite(h, 1, sizeof h, f) == sizeof h;
    // samples are written little-endian whatever the host order
    std::vector<uint8_t> buf;
    const int64_t CH = 1 << 16;
    for (int64_t i = 0; ok && i < n_samples; i += CH) {
        const int64_t m = std::min<int64_t>(CH, n_samples - i);
        buf.resize((size_t)(2 * m));
        for (int64_t j = 0; j < m; ++j) u16(buf.data() + 2 * j, (uint16_t)pcm[i + j]);
        ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : 1;
}

// ---------------------------------------------------------------- length regulator (SURVEY.md 8f, f2 + f1)
// Duration of one phoneme exactly as FS2Encoder::eval rounds it (fs2encoder.cpp:623-627):
//   float dur = exp(dur_data[i]) - 1.0;  int32_t duration_runded = (int32_t)(dur + 0.5);  negative -> skipped
// (exp in double on the float argument, the difference narrowed to float, + 0.5 in double, truncation).
static inline int32_t regulated_duration(float log_dur, int32_t cap)
{
    const float dur = (float)(exp((double)log_dur) - 1.0);
    const double d = (double)dur + 0.5;
    if (!(d >= 0.0)) return 0;                       // negative (or NaN): the reference's `continue`
    if (d >= (double)cap) return cap;                // cannot contribute more than the frames that are left
    return (int32_t)d;
}

int32_t zvx_regulated_frames(const float *log_dur, int32_t P, int32_t max_seq_len)
{
    if (!log_dur || P < 0 || max_seq_len <= 0) return -1;
    int64_t x = 0;
    for (int32_t i = 0; i < P && x < max_seq_len; ++i) x += regulated_duration(log_dur[i], (int32_t)(max_seq_len - x));
    return (int32_t)std::min<int64_t>(x, max_seq_len);
}

// one sub-batch [b0, b1) of zvx_synth_batch_regulated on context / lane `c`; everything asynchronous on c->stream.
// rel[b] = {first frame relative to the utterance, count} per phoneme; `tab` is the caller's staging vector for the packed
// table (it must stay alive until the stream has been synchronised).
static int regulated_chunk(zvx_ctx *c, int b0, int b1, const float *const *features, const int32_t *P, const float *const *style,
                           const std::vector<std::vector<int2>> &rel, const int32_t *L, bool pad, std::vector<int2> &tab,
                           float *const *mel, float *const *wav, int16_t *const *pcm)
{
    zvx_ctx *ctx = c;
    const zvx_config &cfg = c->cfg;
    const int n = b1 - b0;
    if (set_batch(c, n, L + b0)) return 1;
    int64_t np = 0;
    for (int b = b0; b < b1; ++b) np += P[b];
    tab.clear();
    tab.reserve((size_t)np);
    for (int b = 0; b < n; ++b)
        for (const int2 &e : rel[b0 + b]) tab.push_back(make_int2(c->h_seg[b] + e.x, e.y));
    // staging buffers (grown geometrically; growing synchronises, steady state does not)
    const size_t need_f = (size_t)np * cfg.dim_in;
    if (need_f > c->feat_cap) {
        CK(ctx, cudaStreamSynchronize(c->stream));
        const size_t cap = std::max(need_f, c->feat_cap * 2);
        dev_free(c, c->feat); c->feat = nullptr; c->feat_cap = 0;
        if (dev_alloc(c, &c->feat, cap)) return 1;
        c->feat_cap = cap;
    }
    if ((size_t)np > c->feat_tab_cap) {
        CK(ctx, cudaStreamSynchronize(c->stream));
        const size_t cap = std::max((size_t)np, std::max(c->feat_tab_cap * 2, (size_t)1024));
        dev_free(c, c->feat_tab); c->feat_tab = nullptr; c->feat_tab_cap = 0;
        if (dev_alloc(c, &c->feat_tab, cap)) return 1;
        c->feat_tab_cap = cap;
    }
    // H2D at PHONEME rate: [sum P][dim_in] instead of [sum L][dim_in] (one copy per run of contiguous host buffers)
    size_t off = 0;
    for (int b = b0; b < b1;) {
        int e = b + 1;
        size_t rows = (size_t)P[b];
        while (e < b1 && features[e] == features[e - 1] + (size_t)P[e - 1] * cfg.dim_in) rows += (size_t)P[e++];
        CK(ctx, cudaMemcpyAsync(c->feat + off * cfg.dim_in, features[b], sizeof(float) * rows * cfg.dim_in, cudaMemcpyHostToDevice, c->stream));
        off += rows;
        b = e;
    }
    CK(ctx, cudaMemcpyAsync(c->feat_tab, tab.data(), sizeof(int2) * (size_t)np, cudaMemcpyHostToDevice, c->stream));
    for (int b = 0; b < n;) {
        int e = b + 1;
        while (e < n && style[b0 + e] == style[b0 + e - 1] + cfg.style_dim) ++e;
        CK(ctx, cudaMemcpyAsync(c->style + (size_t)b * cfg.style_dim, style[b0 + b], sizeof(float) * (size_t)(e - b) * cfg.style_dim,
                                cudaMemcpyHostToDevice, c->stream));
        b = e;
    }
    // the reference clears the whole [max_seq_len][emb] buffer first (fs2encoder.cpp:614): zero tail of every utterance
    if (pad) CK(ctx, cudaMemsetAsync(c->enc_in, 0, sizeof(float) * (size_t)c->h_seg[n] * cfg.dim_in, c->stream));
    c->launches++;
    if (prof_begin(c, ZVX_K_NORM_AFFINE, 0, 0.0, 4.0 * ((double)np + (double)c->h_seg[n]) * cfg.dim_in)) return 1;
    CK(ctx, length_regulate_launch(c->feat, c->feat_tab, (int)np, cfg.dim_in, c->enc_in, c->stream));
    if (prof_end(c)) return 1;
    if (run_decoder(c, c->mel)) return 1;
    int16_t *d_pcm = reinterpret_cast<int16_t *>(c->wav);
    if (run_vocoder(c, c->mel, pcm ? nullptr : c->wav, pcm ? d_pcm : nullptr)) return 1;
    for (int b = 0; b < n; ++b) {
        const size_t s0 = (size_t)c->h_seg[b] * cfg.hop_size, ns = (size_t)L[b0 + b] * cfg.hop_size;
        if (pcm) CK(ctx, cudaMemcpyAsync(pcm[b0 + b], d_pcm + s0, sizeof(int16_t) * ns, cudaMemcpyDeviceToHost, c->stream));
        else     CK(ctx, cudaMemcpyAsync(wav[b0 + b], c->wav + s0, sizeof(float) * ns, cudaMemcpyDeviceToHost, c->stream));
        if (mel && mel[b0 + b])
            CK(ctx, cudaMemcpyAsync(mel[b0 + b], c->mel + (size_t)c->h_seg[b] * cfg.num_mels, sizeof(float) * (size_t)L[b0 + b] * cfg.num_mels,
                                    cudaMemcpyDeviceToHost, c->stream));
    }
    return 0;
}

int zvx_synth_batch_regulated(zvx_ctx *ctx, int32_t B, const float *const *features, const float *const *log_dur, const int32_t *P,
                              const float *const *style, int32_t max_seq_len, int32_t pad_to_max, int32_t *frames_out,
                              float *const *mel, float *const *wav, int16_t *const *pcm)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_decoder || !ctx->cfg.with_vocoder) return fail(ctx, "context was built without decoder or vocoder");
    if (!features || !log_dur || !P || !style || (!wav == !pcm)) return fail(ctx, "zvx_synth_batch_regulated: null argument (exactly one of wav / pcm)");
    if (B <= 0) return fail(ctx, "empty batch");
    if (max_seq_len <= 0) return fail(ctx, "zvx_synth_batch_regulated: max_seq_len must be positive");
    CK(ctx, cudaSetDevice(ctx->device));
    // host: durations -> valid frames per utterance, {first frame inside the utterance, count} per phoneme
    std::vector<int32_t> L(B);
    std::vector<std::vector<int2>> rel(B);
    int64_t frames = 0;
    for (int b = 0; b < B; ++b) {
        if (P[b] <= 0 || !features[b] || !log_dur[b] || !style[b]) return fail(ctx, "utterance %d: no phonemes / null pointer", b);
        rel[b].reserve((size_t)P[b]);
        int64_t x = 0;
        for (int32_t i = 0; i < P[b]; ++i) {
            const int32_t d = x < max_seq_len ? regulated_duration(log_dur[b][i], (int32_t)(max_seq_len - x)) : 0;
            rel[b].push_back(make_int2((int)x, d));
            x += d;
        }
        if (frames_out) frames_out[b] = (int32_t)x;
        L[b] = pad_to_max ? max_seq_len : (int32_t)x;
        if (L[b] <= 0) return fail(ctx, "utterance %d: all durations are zero", b);
        frames += L[b];
    }
    // like zvx_synth_batch: large batches alternate between the context and its lane, so that the copies of one
    // sub-batch run under the kernels of the other
    std::vector<int2> tabs[2];
    int nch = ctx->e2e_chunks > 1 ? 2 : 1;
    if (ctx->prof || ctx->debug_stop >= 0 || B < 4 || frames < 4096) nch = 1;
    if (nch == 1) {
        if (regulated_chunk(ctx, 0, B, features, P, style, rel, L.data(), pad_to_max != 0, tabs[0], mel, wav, pcm)) return 1;
        return check_device_error(ctx);
    }
    if (make_lane(ctx)) return 1;
    ctx->lane->use_ref_kernels = ctx->use_ref_kernels;
    ctx->lane->use_fused = ctx->use_fused;
    int b0 = 0;
    int64_t acc = 0;
    while (b0 < B - 1 && 2 * (acc + L[b0]) <= frames) acc += L[b0++];
    if (b0 == 0) b0 = 1;
    if (regulated_chunk(ctx, 0, b0, features, P, style, rel, L.data(), pad_to_max != 0, tabs[0], mel, wav, pcm)) {
        cudaStreamSynchronize(ctx->stream);
        return 1;
    }
    if (regulated_chunk(ctx->lane, b0, B, features, P, style, rel, L.data(), pad_to_max != 0, tabs[1], mel, wav, pcm)) {
        ctx->err = ctx->lane->err;
        cudaStreamSynchronize(ctx->stream);          // the first half's copies into the caller's buffers are still in flight
        cudaStreamSynchronize(ctx->lane->stream);
        return 1;
    }
    if (check_device_error(ctx->lane)) { ctx->err = ctx->lane->err; return 1; }
    return check_device_error(ctx);
}

int zvx_decode(zvx_ctx *ctx, const float *enc_seq, const float *style, int32_t L, float *mel)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_decoder) return fail(ctx, "context was built without decoder");
    if (!enc_seq || !style || !mel) return fail(ctx, "zvx_decode: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (set_batch(ctx, 1, &L)) return 1;
    const zvx_config &c = ctx->cfg;
    CK(ctx, cudaMemcpyAsync(ctx->enc_in, enc_seq, sizeof(float) * (size_t)L * c.dim_in, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(ctx->style, style, sizeof(float) * c.style_dim, cudaMemcpyHostToDevice, ctx->stream));
    if (run_graphed(ctx, 0, L, [&]() { return run_decoder(ctx, ctx->mel); })) return 1;
    CK(ctx, cudaMemcpyAsync(mel, ctx->mel, sizeof(float) * (size_t)L * c.num_mels, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int zvx_vocode(zvx_ctx *ctx, const float *mel, int32_t L, float *wav)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_vocoder) return fail(ctx, "context was built without vocoder");
    if (!mel || !wav) return fail(ctx, "zvx_vocode: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (set_batch(ctx, 1, &L)) return 1;
    const zvx_config &c = ctx->cfg;
    CK(ctx, cudaMemcpyAsync(ctx->mel, mel, sizeof(float) * (size_t)L * c.num_mels, cudaMemcpyHostToDevice, ctx->stream));
    if (run_graphed(ctx, 1, L, [&]() { return run_vocoder(ctx, ctx->mel, ctx->wav); })) return 1;
    if (ctx->debug_stop < 0)
        CK(ctx, cudaMemcpyAsync(wav, ctx->wav, sizeof(float) * (size_t)L * c.hop_size, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

int zvx_vocode_batch(zvx_ctx *ctx, int32_t B, const float *const *mel, const int32_t *L, float *const *wav)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_vocoder) return fail(ctx, "context was built without vocoder");
    if (!mel || !L || !wav) return fail(ctx, "zvx_vocode_batch: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (set_batch(ctx, B, L)) return 1;
    const zvx_config &c = ctx->cfg;
    for (int b = 0; b < B; ++b)
        CK(ctx, cudaMemcpyAsync(ctx->mel + (size_t)ctx->h_seg[b] * c.num_mels, mel[b], sizeof(float) * (size_t)L[b] * c.num_mels,
                                cudaMemcpyHostToDevice, ctx->stream));
    if (run_vocoder(ctx, ctx->mel, ctx->wav)) return 1;
    for (int b = 0; b < B; ++b)
        CK(ctx, cudaMemcpyAsync(wav[b], ctx->wav + (size_t)ctx->h_seg[b] * c.hop_size, sizeof(float) * (size_t)L[b] * c.hop_size,
                                cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_error(ctx);
}

// Long-form synthesis (BASELINE.json configs[2]): the vocoder's receptive field is +-19.5 mel
// frames, so a long mel can be vocoded in chunks of `chunk_frames` frames, each
// extended by `halo_frames` >= 20 real neighbouring frames on both sides; the halo part of every
// chunk's output is discarded.  True sequence ends keep the reference's per-layer zero padding.
// The result equals zvx_vocode on the whole mel; device memory is bounded by the chunk size and
// the caller receives the waveform chunk by chunk (on_chunk, may be NULL).
int zvx_vocode_chunked(zvx_ctx *ctx, const float *mel, int32_t L, int32_t chunk_frames, int32_t halo_frames, float *wav,
                       void (*on_chunk)(void *user, int64_t first_sample, int64_t n_samples), void *user)
{
    if (!ctx) return 1;
    if (!ctx->cfg.with_vocoder) return fail(ctx, "context was built without vocoder");
    if (!mel || !wav || L <= 0) return fail(ctx, "zvx_vocode_chunked: bad argument");
    if (chunk_frames <= 0 || halo_frames < 20) return fail(ctx, "zvx_vocode_chunked: chunk_frames must be > 0 and halo_frames >= 20 (receptive field 19.5 frames)");
    CK(ctx, cudaSetDevice(ctx->device));
    const zvx_config &c = ctx->cfg;
    // Chunks are independent once they carry their halo, so several of them go through the vocoder as ONE batch
    // (one pass over ~100 launches instead of one pass per chunk).  The group size ramps 1, 2, 4, 8, 8, ...: the
    // first samples leave after a single chunk's worth of work, the rest of a long utterance runs at batch throughput.
    const int32_t nchunks = (L + chunk_frames - 1) / chunk_frames;
    int32_t group = 1;
    for (int32_t c0 = 0; c0 < nchunks; c0 += group, group = std::min(2 * group, ctx->chunk_group_max)) {
        const int32_t g = std::min(group, nchunks - c0);
        std::vector<int32_t> n(g), lo(g), a(g), b(g);
        for (int32_t i = 0; i < g; ++i) {
            a[i] = (c0 + i) * chunk_frames;
            b[i] = std::min(L, a[i] + chunk_frames);
            lo[i] = std::max(0, a[i] - halo_frames);
            n[i] = std::min(L, b[i] + halo_frames) - lo[i];
        }
        if (set_batch(ctx, g, n.data())) return 1;
        for (int32_t i = 0; i < g; ++i)
            CK(ctx, cudaMemcpyAsync(ctx->mel + (size_t)ctx->h_seg[i] * c.num_mels, mel + (size_t)lo[i] * c.num_mels,
                                    sizeof(float) * (size_t)n[i] * c.num_mels, cudaMemcpyHostToDevice, ctx->stream));
        if (run_vocoder(ctx, ctx->mel, ctx->wav)) return 1;
        for (int32_t i = 0; i < g; ++i)
            CK(ctx, cudaMemcpyAsync(wav + (size_t)a[i] * c.hop_size, ctx->wav + (size_t)(ctx->h_seg[i] + a[i] - lo[i]) * c.hop_size,
                                    sizeof(float) * (size_t)(b[i] - a[i]) * c.hop_size, cudaMemcpyDeviceToHost, ctx->stream));
        if (check_device_error(ctx)) return 1;
        if (on_chunk)
            for (int32_t i = 0; i < g; ++i) on_chunk(user, (int64_t)a[i] * c.hop_size, (int64_t)(b[i] - a[i]) * c.hop_size);
    }
    return 0;
}

int zvx_debug_fetch(zvx_ctx *ctx, const char *what, float *dst, int64_t n)
{
    if (!ctx || !what || !dst) return 1;
    CK(ctx, cudaSetDevice(ctx->device));
    const std::string w = what;
    const float *src = nullptr;
    int64_t have = 0;
    const int64_t F = ctx->last_frames;
    if (w == "mel") { src = ctx->mel; have = F * ctx->cfg.num_mels; }
    else if (w == "enc_in") { src = ctx->enc_in; have = F * ctx->cfg.dim_in; }      // decoder input (length regulator output)
    else if (w == "v0") { src = ctx->v0; have = F * ctx->chans[0]; }
    else if (w == "u") { src = ctx->U; have = F * max_stage_elems(ctx); }
    else if (w.rfind("stage", 0) == 0 && w.size() == 6) {
        const int i = w[5] - '0';
        if (i < 0 || i >= ctx->cfg.num_upsamples) return fail(ctx, "bad stage");
        src = (i & 1) ? ctx->VB : ctx->VA;
        have = F * ctx->rates[i + 1] * ctx->chans[i + 1];
        if (ctx->stage_is_split[i]) {
            // the stage output is still split into its three branches: combine like the consumer does
            if (n > have) return fail(ctx, "debug tensor '%s' has %lld floats, asked for %lld", what, (long long)have, (long long)n);
            CK(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->stage_is_split[i] == 2) {           // fp16 branch tensors
                std::vector<__half> h0((size_t)n), h1((size_t)n), h2((size_t)n);
                CK(ctx, cudaMemcpy(h0.data(), ctx->CS, sizeof(__half) * n, cudaMemcpyDeviceToHost));
                CK(ctx, cudaMemcpy(h1.data(), ctx->VA, sizeof(__half) * n, cudaMemcpyDeviceToHost));
                CK(ctx, cudaMemcpy(h2.data(), ctx->VB, sizeof(__half) * n, cudaMemcpyDeviceToHost));
                const float third = (float)(1.0 / 3.0);
                for (int64_t q = 0; q < n; ++q) dst[q] = ((__half2float(h0[q]) + __half2float(h1[q])) + __half2float(h2[q])) * third;
                return 0;
            }
            std::vector<float> b1((size_t)n), b2((size_t)n);
            CK(ctx, cudaMemcpy(dst, ctx->CS, sizeof(float) * n, cudaMemcpyDeviceToHost));
            CK(ctx, cudaMemcpy(b1.data(), ctx->VA, sizeof(float) * n, cudaMemcpyDeviceToHost));
            CK(ctx, cudaMemcpy(b2.data(), ctx->VB, sizeof(float) * n, cudaMemcpyDeviceToHost));
            const float third = (float)(1.0 / 3.0);
            for (int64_t q = 0; q < n; ++q) dst[q] = ((dst[q] + b1[q]) + b2[q]) * third;
            return 0;
        }
    } else return fail(ctx, "unknown debug tensor '%s'", what);
    if (n > have) return fail(ctx, "debug tensor '%s' has %lld floats, asked for %lld", what, (long long)have, (long long)n);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(dst, src, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return 0;
}

int zvx_test_conv(zvx_ctx *ctx, const zvx_conv_test *t)
{
    if (!ctx || !t) return 1;
    CK(ctx, cudaSetDevice(ctx->device));
    // rows play the role of frames at rate index 0 (rate 1)
    int rc = 1;
    std::vector<void *> tmp;
    auto talloc = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr;
        tmp.push_back(p);
        return p;
    };
    do {
        if (set_batch(ctx, t->B, t->rows, false)) break;
        const int64_t R = ctx->last_frames;
        ConvLayer L;
        L.OC = t->Cout; L.IC = t->Cin; L.K = t->K; L.NC = pick_nc(t->Cout);
        if (L.NC == 0 || t->Cout % 16 || t->Cin % 16) { fail(ctx, "zvx_test_conv: unsupported channels"); break; }
        const size_t wn = (size_t)t->Cout * t->Cin * t->K;
        __half *d_raw = (__half *)talloc(wn * 2);
        if (!d_raw || cudaMemcpy(d_raw, t->w, wn * 2, cudaMemcpyHostToDevice) != cudaSuccess) { fail(ctx, "alloc/copy w"); break; }
        L.raw = d_raw;
        float *d_bias = nullptr;
        if (t->bias) {
            d_bias = (float *)talloc(sizeof(float) * t->Cout);
            if (!d_bias || cudaMemcpy(d_bias, t->bias, sizeof(float) * t->Cout, cudaMemcpyHostToDevice) != cudaSuccess) { fail(ctx, "alloc/copy bias"); break; }
        }
        L.bias = d_bias;
        ConvVariant v;
        v.ntaps = t->K; v.w_tap0 = 0; v.w_tap_stride = 1; v.tap_step = t->dilation; v.tap_off0 = -t->pad; v.out_add = 0;
        std::vector<__half> hraw((const __half *)t->w, (const __half *)t->w + wn);
        if (pack_variant(ctx, hraw, L.OC, L.IC, L.K, L.NC, v)) break;
        L.var.push_back(v);
        ConvCall cc;
        cc.L = &L;
        cc.pro_mode = t->pro_mode; cc.pro_slope = t->pro_slope;
        void *d_x;
        if (t->pro_mode == PRO_F16) {
            d_x = talloc((size_t)R * t->Cin * 2);
            if (!d_x || cudaMemcpy(d_x, t->x16, (size_t)R * t->Cin * 2, cudaMemcpyHostToDevice) != cudaSuccess) { fail(ctx, "alloc/copy x16"); break; }
        } else {
            d_x = talloc((size_t)R * t->Cin * 4);
            if (!d_x || cudaMemcpy(d_x, t->x, (size_t)R * t->Cin * 4, cudaMemcpyHostToDevice) != cudaSuccess) { fail(ctx, "alloc/copy x"); break; }
        }
        cc.x = d_x; cc.ldx = t->Cin;
        auto up = [&](const float *h, size_t n) -> float * {
            if (!h) return nullptr;
            float *d = (float *)talloc(n * 4);
            if (d && cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice) != cudaSuccess) d = nullptr;
            return d;
        };
        if (t->pro_mode == PRO_NORM) {
            cc.mu = up(t->mu, (size_t)t->B * t->Cin); cc.rstd = up(t->rstd, (size_t)t->B * t->Cin); cc.stat_stride = t->Cin;
            cc.g = up(t->g, (size_t)t->B * t->Cin); cc.b = up(t->b, (size_t)t->B * t->Cin); cc.gb_stride = t->Cin;
            if (!cc.mu || !cc.rstd || !cc.g || !cc.b) { fail(ctx, "norm params missing"); break; }
        } else if (t->pro_mode == PRO_MEL) {
            cc.mu = up(t->mu, t->Cin); cc.rstd = up(t->rstd, t->Cin); cc.stat_stride = 0;
            if (!cc.mu || !cc.rstd) { fail(ctx, "mel params missing"); break; }
        }
        cc.use_bias = t->bias != nullptr;
        if (t->res) { cc.res = up(t->res, (size_t)R * t->Cout); cc.ldres = t->Cout; if (!cc.res) { fail(ctx, "res"); break; } }
        cc.scale = t->scale;
        float *d_out = (float *)talloc((size_t)R * t->Cout * 4);
        __half *d_out16 = t->out16 ? (__half *)talloc((size_t)R * t->Cout * 2) : nullptr;
        if (!d_out) { fail(ctx, "alloc out"); break; }
        cudaMemsetAsync(d_out, 0xff, (size_t)R * t->Cout * 4, ctx->stream);
        cc.out32 = d_out; cc.ldo32 = t->Cout;
        cc.out16 = d_out16; cc.ldo16 = t->Cout; cc.out16_slope = t->out16_slope;
        const int saved = ctx->use_ref_kernels;
        ctx->use_ref_kernels = t->use_validation_kernel;
        const int r = run_conv(ctx, cc);
        ctx->use_ref_kernels = saved;
        if (r) break;
        if (check_device_error(ctx)) break;
        if (cudaMemcpy(t->out, d_out, (size_t)R * t->Cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { fail(ctx, "copy out"); break; }
        if (t->out16 && cudaMemcpy(t->out16, d_out16, (size_t)R * t->Cout * 2, cudaMemcpyDeviceToHost) != cudaSuccess) { fail(ctx, "copy out16"); break; }
        if (v.packed) dev_free(ctx, v.packed);
        rc = 0;
    } while (0);
    for (void *p : tmp) cudaFree(p);
    return rc;
}

}  // extern "C"
