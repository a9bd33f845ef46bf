/* zvx.h -- C ABI of the B200-native zerovox mel-decoder + vocoder hot path (libzvx.so).
 *
 * The reference has no FFI layer: its boundary for this path is the C++ class API in
 * /root/reference/src/zerovox.h plus the GGUF tensor naming of utils/zv2gguf.py.  These
 * entry points are what a reference-side shim binds (see INTEGRATION.md and
 * zerovox.cpp_b200/host/zerovox_b200.h); each cites the reference interface it replaces.
 * Plain pointers and sizes only -- no torch, no ggml types.  All functions return 0 on
 * success and a non-zero code on failure; zvx_last_error() gives the message (the C++
 * shims rethrow it as std::runtime_error, matching stylettsdec.cpp:447-448,465-466 and
 * hifigan.cpp:353-354,362-363).  One zvx_ctx per GPU; calls on one ctx are serialised by
 * the caller (the reference classes are not re-entrant either).
 */
#ifndef ZVX_H
#define ZVX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zvx_ctx zvx_ctx;

/* dtype codes equal ggml_type (ggml.h): GGML_TYPE_F32 = 0, GGML_TYPE_F16 = 1 */
enum { ZVX_F32 = 0, ZVX_F16 = 1 };

/* One weight tensor, as found by walking the GGUF / ggml weight context.
 * Replaces: checked_get_tensor(ctx_w, name)          /root/reference/src/utils.cpp:9-17
 * `ne` is ggml order (ne[0] fastest), `data` a HOST pointer (tensor->data on the CPU
 * backend); the library copies and repacks, the caller keeps ownership. */
typedef struct zvx_tensor_desc {
    const char *name;
    int32_t     dtype;
    int32_t     n_dims;
    int64_t     ne[4];
    const void *data;
} zvx_tensor_desc;

/* Topology arguments = the constructor arguments of the two reference classes.
 * Replaces: StyleTTSDecoder::StyleTTSDecoder(...)    /root/reference/src/zerovox.h:314-320
 *           HiFiGAN::HiFiGAN(...)                    /root/reference/src/zerovox.h:366-376
 * (values used by ZeroVOXModel: /root/reference/src/zerovox.cpp:117-138) */
typedef struct zvx_config {
    int32_t device;                 /* CUDA device ordinal */
    int32_t dim_in;                 /* emb_dim + punct_emb_dim (528) */
    int32_t style_dim;              /* 528 */
    int32_t residual_dim;           /* 64 */
    int32_t num_mels;               /* 80 = decoder dim_out = vocoder in_channels */
    int32_t hop_size;               /* 300 */
    int32_t kernel_size;            /* 7: vocoder input/output conv */
    int32_t num_upsamples;          /* 4 */
    int32_t upsample_scales[8];     /* 5,5,4,3 */
    int32_t num_resblocks;          /* 3 */
    int32_t num_resblock_dilations; /* 3 */
    int32_t resblock_dilations[32]; /* [num_resblocks][num_resblock_dilations] = 1,3,5 x3 */
    int32_t with_decoder;           /* build the StyleTTS decoder part (needs its tensors) */
    int32_t with_vocoder;           /* build the HiFi-GAN part */
} zvx_config;

/* Fill cfg with the values ZeroVOXModel passes (zerovox.cpp:117-138). */
void zvx_default_config(zvx_config *cfg);

/* Upload + repack weights, allocate device state.  Replaces the two constructors above
 * (graph build + ggml_gallocr_alloc_graph, stylettsdec.cpp:345-448, hifigan.cpp:223-354).
 * A missing tensor fails like checked_get_tensor (utils.cpp:12-15). */
int zvx_create(zvx_ctx **out, const zvx_config *cfg, const zvx_tensor_desc *weights, int32_t n_weights);
void zvx_destroy(zvx_ctx *ctx);

/* Message of the last failure on ctx (or of the last failed zvx_create when ctx == NULL).
 *
 * Fatal device faults.  Every barrier wait inside the kernels is bounded (about 3 s of SM clocks): a pipeline
 * that stops making progress sets a flag in device memory and executes `trap`, so a bug cannot hang the GPU.
 * Like any device-side fault the trap invalidates the CUDA primary context of the PROCESS: the failing call and
 * every later call on any zvx_ctx return non-zero with "... unspecified launch failure ... (fatal: the CUDA
 * context is lost, restart the process)", and only a new process can use the GPU again.  The reference has no
 * analogue (ggml CPU graphs cannot time out); a host that must survive this runs the library in a worker
 * process. */
const char *zvx_last_error(const zvx_ctx *ctx);

/* Replaces StyleTTSDecoder::eval            /root/reference/src/stylettsdec.cpp:457-470
 * enc_seq [L][dim_in] frame-major, style [style_dim], mel out [L][num_mels]; HOST pointers.
 * InstanceNorm/AdaIN statistics span exactly L frames. */
int zvx_decode(zvx_ctx *ctx, const float *enc_seq, const float *style, int32_t L, float *mel);

/* Replaces HiFiGAN::eval                    /root/reference/src/hifigan.cpp:358-377
 * mel [L][num_mels] -> wav [L*hop_size]; HOST pointers. */
int zvx_vocode(zvx_ctx *ctx, const float *mel, int32_t L, float *wav);

/* Batched form of decoder->eval + meldec->eval (ZeroVOXModel::eval, zerovox.cpp:330-334)
 * for B independent utterances of lengths L[b].  HOST pointers; mel may be NULL (or
 * individual entries NULL) when the caller only wants the waveform. */
int zvx_synth_batch(zvx_ctx *ctx, int32_t B, const float *const *enc_seq, const float *const *style,
                    const int32_t *L, float *const *mel, float *const *wav);

/* Pipelined form of zvx_synth_batch / zvx_synth_batch_pcm16 (exactly one of wav / pcm non-NULL): submit returns once the
 * copies and kernels of the batch are enqueued; every buffer passed to it -- inputs and outputs -- must stay untouched
 * until zvx_synth_batch_wait has returned (0 = every submitted batch completed).  Replaces a loop of ZeroVOXModel::eval
 * calls (zerovox.cpp:326-334) for callers that synthesise batch after batch: the device -> host copy of batch i runs
 * under the host -> device copy and the kernels of batch i + 1. */
int zvx_synth_batch_submit(zvx_ctx *ctx, int32_t B, const float *const *enc_seq, const float *const *style,
                           const int32_t *L, float *const *mel, float *const *wav, int16_t *const *pcm);
int zvx_synth_batch_wait(zvx_ctx *ctx);

/* Output stage: the waveform leaves the GPU as signed 16-bit PCM.
 * Replaces, for the batch, the float -> short conversion libsndfile performs inside
 * sf_write_float() when ZeroVOXModel::write_wav_file writes its SF_FORMAT_PCM_16 file
 * (/root/reference/src/zerovox.cpp:357-371; libsndfile src/pcm.c f2s_array, default
 * normalisation, clipping off: lrintf(x * 0x7FFF)).  The conversion is fused into the output
 * conv's tanh epilogue, so the device -> host copy is half the size of zvx_synth_batch's.
 * pcm[b] receives L[b]*hop_size samples; other arguments as zvx_synth_batch. */
int zvx_synth_batch_pcm16(zvx_ctx *ctx, int32_t B, const float *const *enc_seq, const float *const *style,
                          const int32_t *L, float *const *mel, int16_t *const *pcm);

/* HiFiGAN::eval (hifigan.cpp:358-377) + the same conversion: mel [L][num_mels] -> pcm [L*hop_size]. */
int zvx_vocode_pcm16(zvx_ctx *ctx, const float *mel, int32_t L, int16_t *pcm);

/* Replaces sf_open / sf_write / sf_close of ZeroVOXModel::write_wav_file (zerovox.cpp:354-384) for
 * its one format, mono SF_FORMAT_WAV | SF_FORMAT_PCM_16: a canonical 44-byte RIFF header followed by
 * little-endian samples.  Host only (no ctx, no GPU).  Returns 0 on success. */
int zvx_write_wav_pcm16(const char *path, const int16_t *pcm, int64_t n_samples, int32_t sample_rate);

/* Length regulator in front of the hot path.  Replaces the host loop at
 * the end of FS2Encoder::eval (/root/reference/src/fs2encoder.cpp:611-654: every phoneme's feature row is
 * repeated round(exp(log_duration) - 1) times into a zeroed [max_seq_len][emb] matrix; the number of valid
 * frames is returned and then ignored by ZeroVOXModel::eval, zerovox.cpp:326-334) and the host -> device copy
 * of that expanded matrix: features travel at PHONEME rate and are expanded on the GPU.
 *   features[b] [P[b]][dim_in], log_dur[b] [P[b]]  (the `features` / `log_duration_prediction` graph outputs),
 *   style[b] [style_dim]; HOST pointers.
 *   Durations are rounded on the host with the reference's own expression
 *   (float dur = exp(d) - 1.0; (int32_t)(dur + 0.5)), capped at max_seq_len frames per utterance.
 *   pad_to_max != 0: reference behaviour -- every utterance is synthesised as max_seq_len frames with a
 *   zero tail (InstanceNorm statistics span the tail); outputs hold max_seq_len frames.
 *   pad_to_max == 0: only the valid frames are synthesised (f1); outputs hold frames_out[b] frames --
 *   size them with zvx_regulated_frames first.
 *   frames_out (may be NULL) receives the valid-frame counts, i.e. FS2Encoder::eval's return value.
 *   Exactly one of wav / pcm (see zvx_synth_batch_pcm16) is non-NULL; mel as in zvx_synth_batch. */
int zvx_synth_batch_regulated(zvx_ctx *ctx, int32_t B, const float *const *features, const float *const *log_dur,
                              const int32_t *P, const float *const *style, int32_t max_seq_len, int32_t pad_to_max,
                              int32_t *frames_out, float *const *mel, float *const *wav, int16_t *const *pcm);

/* The valid-frame count FS2Encoder::eval returns for one utterance (fs2encoder.cpp:621-654); host only. */
int32_t zvx_regulated_frames(const float *log_dur, int32_t P, int32_t max_seq_len);

/* Batched HiFiGAN::eval (hifigan.cpp:358-377) for B independent mels; HOST pointers. */
int zvx_vocode_batch(zvx_ctx *ctx, int32_t B, const float *const *mel, const int32_t *L, float *const *wav);

/* Long-form synthesis (BASELINE.json configs[2]: 60 s utterances streamed in overlapping mel
 * chunks): HiFiGAN::eval on a long mel, computed chunk by chunk with `halo_frames` (>= 20: the
 * vocoder's receptive field is +-19.5 frames) real neighbouring frames on each side; on_chunk
 * (may be NULL) is called as soon as a chunk's samples are in `wav`.  Same result as zvx_vocode. */
int zvx_vocode_chunked(zvx_ctx *ctx, const float *mel, int32_t L, int32_t chunk_frames, int32_t halo_frames, float *wav,
                       void (*on_chunk)(void *user, int64_t first_sample, int64_t n_samples), void *user);

/* Same computation with inputs/outputs already resident in device memory, packed back to
 * back: d_enc [sum L][dim_in], d_style [B][style_dim], d_mel [sum L][num_mels] (may be
 * NULL), d_wav [sum L * hop].  L is a HOST array.  Asynchronous on the ctx stream unless
 * sync != 0.  Used by bench.py for the HBM-resident `value` measurement. */
int zvx_synth_batch_device(zvx_ctx *ctx, int32_t B, const float *d_enc, const float *d_style, const int32_t *L,
                           float *d_mel, float *d_wav, int32_t sync);
int zvx_vocode_batch_device(zvx_ctx *ctx, int32_t B, const float *d_mel, const int32_t *L, float *d_wav,
                            int32_t sync);

/* Stream / accounting helpers */
void *zvx_stream(zvx_ctx *ctx);                 /* cudaStream_t all work is enqueued on */
int   zvx_synchronize(zvx_ctx *ctx);
int64_t zvx_kernel_launches(const zvx_ctx *ctx); /* kernels launched by this ctx so far */
int   zvx_reserve(zvx_ctx *ctx, int64_t total_frames, int32_t max_batch);

/* Per-launch timing with CUDA events recorded on the ctx stream around every kernel this
 * library launches between zvx_profile_begin and zvx_profile_end (bench.py's roofline).
 * flops / bytes are the ALGORITHMIC work of that launch. */
enum {
    ZVX_K_DEC_CONV = 0, ZVX_K_VOC_INPUT_CONV = 1, ZVX_K_UPCONV = 2, ZVX_K_MRF_CONV = 3, ZVX_K_OUT_CONV = 4,
    ZVX_K_STATS = 5, ZVX_K_ADAIN_FC = 6, ZVX_K_NORM_AFFINE = 7
};
typedef struct zvx_launch_record {
    int32_t kind;     /* ZVX_K_* */
    int32_t stage;    /* vocoder stage for UPCONV / MRF_CONV / OUT_CONV, else 0 */
    double  flops;
    double  bytes;
    float   ms;
} zvx_launch_record;
int     zvx_profile_begin(zvx_ctx *ctx);
int64_t zvx_profile_end(zvx_ctx *ctx, zvx_launch_record *recs, int64_t max_recs); /* returns #records or -1 */

/* ---- test / debug surface (used by tests/ only) ---------------------------------- */
/* 0: tcgen05 implicit-GEMM kernels (the product path); 1: plain-CUDA validation kernels */
void zvx_set_debug_kernels(zvx_ctx *ctx, int32_t use_validation_kernels);
/* 1 (default): MRF residual blocks run as fused on-chip chains (mrf_fused.cu); 0: one
 * implicit-GEMM launch per convolution (conv_umma.cu).  Both are tcgen05 paths. */
void zvx_set_fused_mrf(zvx_ctx *ctx, int32_t on);

/* Run ONE convolution through the selected kernel on packed utterances.
 * x [sum rows][Cin] fp32 host, w (OC, IC, K) fp16 host (K fastest), bias [OC] or NULL,
 * out [sum rows * out_mul][OC] fp32 host, out16 same shape fp16 host or NULL.
 * pro_mode / slope / mu / rstd / g / b as in the fused prologue; res optional residual. */
typedef struct zvx_conv_test {
    int32_t B;
    const int32_t *rows;        /* [B] rows per utterance */
    int32_t Cin, Cout, K, dilation, pad;
    int32_t pro_mode;           /* 0 raw fp16 (x16 given), 1 cvt, 2 lrelu, 3 norm, 4 mel */
    float   pro_slope;
    const float *x;             /* fp32 input (modes 1-4) */
    const uint16_t *x16;        /* fp16 input (mode 0) */
    const uint16_t *w;
    const float *bias;
    const float *mu, *rstd, *g, *b;   /* [B][Cin] (norm) / [Cin] (mel: mu, rstd only) */
    const float *res;           /* [sum rows][Cout] or NULL */
    float   scale;              /* 0 = no scale */
    float   out16_slope;
    float  *out;                /* fp32 out */
    uint16_t *out16;            /* fp16(lrelu(out)) or NULL */
    int32_t use_validation_kernel;
} zvx_conv_test;
int zvx_test_conv(zvx_ctx *ctx, const zvx_conv_test *t);

/* Copy an internal activation of the LAST run to the host (fp32): "mel", "v0" (vocoder
 * input conv), "stage0".."stage3" need zvx_set_debug_stop(ctx, stage+1). */
int zvx_debug_fetch(zvx_ctx *ctx, const char *what, float *dst, int64_t n_floats);
void zvx_set_debug_stop(zvx_ctx *ctx, int32_t stop_after_stages); /* -1: run everything */

#ifdef __cplusplus
}
#endif
#endif /* ZVX_H */
