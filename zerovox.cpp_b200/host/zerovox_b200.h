// zerovox_b200.h -- host-side mirror of the reference's mel-decoder / vocoder classes.
//
// Drop-in replacements for the two class declarations of /root/reference/src/zerovox.h
//   ZeroVOX::StyleTTSDecoder   zerovox.h:310-360   (implemented by src/stylettsdec.cpp)
//   ZeroVOX::HiFiGAN           zerovox.h:362-402   (implemented by src/hifigan.cpp)
// with the same names, constructor arguments, eval() signatures, buffer layouts and error
// behaviour (std::runtime_error), so that ZeroVOXModel (src/zerovox.cpp:119-138,330-334)
// compiles and runs unchanged.  Only the private members differ: instead of a ggml graph the
// objects hold a handle of the C-ABI CUDA library (include/zvx.h, libzvx.so).  See
// INTEGRATION.md for the two-line change in the reference tree.
//
// Semantics kept from the reference:
//  * weights are looked up by name in `ctx_w` at construction (checked_get_tensor,
//    src/utils.cpp:9-17: a missing tensor throws); their DATA is read at the first eval(),
//    because ZeroVOXModel fills the tensors after constructing the stages
//    (src/zerovox.cpp:140-172);
//  * eval() always processes max_seq_len frames, InstanceNorm / AdaIN statistics span all of
//    them;
//  * buffers are caller-owned host memory, frame-major.
// Not kept: the reference's stdout debug prints in HiFiGAN::eval (hifigan.cpp:365-372) and its
// one-instance-per-process limitation (function-static graph buffers).
#pragma once

#include <cstdint>
#include <string>

#include "ggml.h"
#include "ggml-backend.h"

struct zvx_ctx;

namespace ZeroVOX
{
    class StyleTTSDecoder
    {
        public:

            StyleTTSDecoder(ggml_context   &ctx_w,
                            ggml_backend_t  backend,
                            uint32_t        max_seq_len,
                            uint32_t        dim_in,
                            uint32_t        style_dim,
                            uint32_t        residual_dim,
                            uint32_t        dim_out);
            ~StyleTTSDecoder();

            void eval(const float *enc_seq_data, const float *spk_emb_data, float *mel);

            // B200 extension: any number of frames <= max_seq_len (statistics over exactly n_frames)
            void eval(const float *enc_seq_data, const float *spk_emb_data, float *mel, uint32_t n_frames);

        private:

            void init();

            ggml_context *ctx_w;
            uint32_t      max_seq_len, dim_in, style_dim, residual_dim, dim_out;
            int           device;
            zvx_ctx      *zvx;
    };

    class HiFiGAN
    {
        public:

            HiFiGAN(ggml_context   &ctx_w,
                    ggml_backend_t  backend,
                    uint32_t        max_seq_len,
                    uint32_t        in_channels,
                    uint32_t        hop_size,
                    uint32_t        kernel_size,
                    int             num_upsamples,
                    const int      *upsample_scales,
                    int             num_resblocks,
                    int             num_resblock_dilations,
                    const int64_t  *resblock_dilations);
            ~HiFiGAN();

            void eval(const float *mel, float *wav);

            // B200 extension: any number of frames <= max_seq_len
            void eval(const float *mel, float *wav, uint32_t n_frames);

            // B200 extension: the waveform as signed 16-bit PCM, converted in the output
            // conv's epilogue exactly as libsndfile converts it inside ZeroVOXModel::write_wav_file
            // (zerovox.cpp:357-371): half the device -> host bytes of eval()
            void eval_pcm16(const float *mel, int16_t *pcm, uint32_t n_frames);

        private:

            void init();

            ggml_context *ctx_w;
            uint32_t      max_seq_len, in_channels, hop_size, kernel_size;
            int           num_upsamples, num_resblocks, num_resblock_dilations;
            int           upsample_scales[8];
            int           resblock_dilations[32];
            int           device;
            zvx_ctx      *zvx;
    };

    // ZeroVOXModel::write_wav_file (zerovox.cpp:337-391) for samples that are already PCM_16: mono
    // RIFF/WAVE, no libsndfile.  Returns false (and reports on stderr) when the file cannot be written.
    bool write_wav_file_pcm16(const std::string &fname, const int16_t *pcm, size_t n_samples, uint32_t sample_rate);
}
