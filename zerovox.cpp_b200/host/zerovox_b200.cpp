// zerovox_b200.cpp -- ZeroVOX::StyleTTSDecoder / ZeroVOX::HiFiGAN on top of the C-ABI CUDA
// library (include/zvx.h).  Replaces /root/reference/src/stylettsdec.cpp and src/hifigan.cpp
// behind the class API of /root/reference/src/zerovox.h:310-402.
#include "zerovox_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../../include/zvx.h"
#include "zvx_weights.h"

namespace ZeroVOX
{
    namespace
    {
        // checked_get_tensor (reference src/utils.cpp:9-17): a missing tensor is a runtime_error
        ggml_tensor *checked(ggml_context *ctx, const std::string &name)
        {
            ggml_tensor *t = ggml_get_tensor(ctx, name.c_str());
            if (!t) {
                char buf[256];
                snprintf(buf, sizeof buf, "%s: tensor '%s' not found", __func__, name.c_str());
                throw std::runtime_error(buf);
            }
            return t;
        }

        int env_device()
        {
            const char *e = getenv("ZVX_DEVICE");
            return e ? atoi(e) : 0;
        }

        [[noreturn]] void raise(zvx_ctx *z, const char *what)
        {
            throw std::runtime_error(std::string(what) + ": " + zvx_last_error(z));
        }
    }

    // ------------------------------------------------------------------ StyleTTSDecoder
    StyleTTSDecoder::StyleTTSDecoder(ggml_context &ctx_w_, ggml_backend_t /*backend*/, uint32_t max_seq_len_, uint32_t dim_in_,
                                     uint32_t style_dim_, uint32_t residual_dim_, uint32_t dim_out_)
        : ctx_w(&ctx_w_), max_seq_len(max_seq_len_), dim_in(dim_in_), style_dim(style_dim_), residual_dim(residual_dim_),
          dim_out(dim_out_), device(env_device()), zvx(nullptr)
    {
        // the same lookups the reference constructor performs (stylettsdec.cpp:33-66,163-168,220-239,334-340)
        static const char *blocks[] = {"encode.0", "encode.1", "decode.0", "decode.1", "decode.2", "decode.3", "decode.4"};
        for (const char *b : blocks) {
            const std::string p = std::string("_mel_decoder.") + b;
            checked(ctx_w, p + ".conv1.w"); checked(ctx_w, p + ".conv1.b");
            checked(ctx_w, p + ".conv2.w"); checked(ctx_w, p + ".conv2.b");
            if (b[0] == 'e') {
                checked(ctx_w, p + ".norm1.w"); checked(ctx_w, p + ".norm1.b");
                checked(ctx_w, p + ".norm2.w"); checked(ctx_w, p + ".norm2.b");
            } else {
                checked(ctx_w, p + ".norm1.fc.w"); checked(ctx_w, p + ".norm1.fc.b");
                checked(ctx_w, p + ".norm2.fc.w"); checked(ctx_w, p + ".norm2.fc.b");
            }
        }
        checked(ctx_w, "_mel_decoder.asr_res.0.w"); checked(ctx_w, "_mel_decoder.asr_res.0.b");
        checked(ctx_w, "_mel_decoder.asr_res.1.w"); checked(ctx_w, "_mel_decoder.asr_res.1.b");
        checked(ctx_w, "_mel_decoder.to_out.0.w");  checked(ctx_w, "_mel_decoder.to_out.0.b");
    }

    StyleTTSDecoder::~StyleTTSDecoder() { zvx_destroy(zvx); }

    void StyleTTSDecoder::init()
    {
        WeightSet ws;
        collect(ctx_w, {"_mel_decoder."}, ws);
        zvx_config cfg;
        zvx_default_config(&cfg);
        cfg.device = device;
        cfg.dim_in = (int32_t)dim_in;
        cfg.style_dim = (int32_t)style_dim;
        cfg.residual_dim = (int32_t)residual_dim;
        cfg.num_mels = (int32_t)dim_out;
        cfg.with_decoder = 1;
        cfg.with_vocoder = 0;
        if (zvx_create(&zvx, &cfg, ws.descs.data(), (int32_t)ws.descs.size()) != 0) raise(nullptr, "StyleTTSDecoder");
        if (zvx_reserve(zvx, max_seq_len, 1) != 0) raise(zvx, "StyleTTSDecoder");
    }

    void StyleTTSDecoder::eval(const float *enc_seq_data, const float *spk_emb_data, float *mel)
    {
        eval(enc_seq_data, spk_emb_data, mel, max_seq_len);
    }

    void StyleTTSDecoder::eval(const float *enc_seq_data, const float *spk_emb_data, float *mel, uint32_t n_frames)
    {
        if (n_frames == 0 || n_frames > max_seq_len) throw std::runtime_error("StyleTTSDecoder::eval: n_frames out of range");
        if (!zvx) init();
        if (zvx_decode(zvx, enc_seq_data, spk_emb_data, (int32_t)n_frames, mel) != 0) raise(zvx, "StyleTTSDecoder::eval");
    }

    // ------------------------------------------------------------------ HiFiGAN
    HiFiGAN::HiFiGAN(ggml_context &ctx_w_, ggml_backend_t /*backend*/, uint32_t max_seq_len_, uint32_t in_channels_, uint32_t hop_size_,
                     uint32_t kernel_size_, int num_upsamples_, const int *upsample_scales_, int num_resblocks_,
                     int num_resblock_dilations_, const int64_t *resblock_dilations_)
        : ctx_w(&ctx_w_), max_seq_len(max_seq_len_), in_channels(in_channels_), hop_size(hop_size_), kernel_size(kernel_size_),
          num_upsamples(num_upsamples_), num_resblocks(num_resblocks_), num_resblock_dilations(num_resblock_dilations_),
          device(env_device()), zvx(nullptr)
    {
        if (num_upsamples < 0 || num_upsamples > 8 || num_resblocks * num_resblock_dilations > 32)
            throw std::runtime_error("HiFiGAN: unsupported topology");
        for (int i = 0; i < num_upsamples; ++i) upsample_scales[i] = upsample_scales_[i];
        for (int i = 0; i < num_resblocks * num_resblock_dilations; ++i) resblock_dilations[i] = (int)resblock_dilations_[i];
        // the same lookups the reference constructor performs (hifigan.cpp:34-39,123-128,162-167,208-218)
        checked(ctx_w, "hifigan.mean"); checked(ctx_w, "hifigan.scale");
        checked(ctx_w, "_meldec.input_conv.w"); checked(ctx_w, "_meldec.input_conv.b");
        checked(ctx_w, "_meldec.output_conv.1.w"); checked(ctx_w, "_meldec.output_conv.1.b");
        char nm[128];
        for (int i = 0; i < num_upsamples; ++i) {
            snprintf(nm, sizeof nm, "_meldec.upsamples.%d.1", i);
            checked(ctx_w, std::string(nm) + ".w"); checked(ctx_w, std::string(nm) + ".b");
            for (int j = 0; j < num_resblocks; ++j)
                for (int d = 0; d < num_resblock_dilations; ++d)
                    for (int c = 1; c <= 2; ++c) {
                        snprintf(nm, sizeof nm, "_meldec.blocks.%d.convs%d.%d.1", i * num_resblocks + j, c, d);
                        checked(ctx_w, std::string(nm) + ".w"); checked(ctx_w, std::string(nm) + ".b");
                    }
        }
    }

    HiFiGAN::~HiFiGAN() { zvx_destroy(zvx); }

    void HiFiGAN::init()
    {
        WeightSet ws;
        collect(ctx_w, {"_meldec.", "hifigan."}, ws);
        zvx_config cfg;
        zvx_default_config(&cfg);
        cfg.device = device;
        cfg.num_mels = (int32_t)in_channels;
        cfg.hop_size = (int32_t)hop_size;
        cfg.kernel_size = (int32_t)kernel_size;
        cfg.num_upsamples = num_upsamples;
        for (int i = 0; i < num_upsamples; ++i) cfg.upsample_scales[i] = upsample_scales[i];
        cfg.num_resblocks = num_resblocks;
        cfg.num_resblock_dilations = num_resblock_dilations;
        for (int i = 0; i < num_resblocks * num_resblock_dilations; ++i) cfg.resblock_dilations[i] = resblock_dilations[i];
        cfg.with_decoder = 0;
        cfg.with_vocoder = 1;
        if (zvx_create(&zvx, &cfg, ws.descs.data(), (int32_t)ws.descs.size()) != 0) raise(nullptr, "HiFiGAN");
        if (zvx_reserve(zvx, max_seq_len, 1) != 0) raise(zvx, "HiFiGAN");
    }

    void HiFiGAN::eval(const float *mel, float *wav) { eval(mel, wav, max_seq_len); }

    void HiFiGAN::eval(const float *mel, float *wav, uint32_t n_frames)
    {
        if (n_frames == 0 || n_frames > max_seq_len) throw std::runtime_error("HiFiGAN::eval: n_frames out of range");
        if (!zvx) init();
        if (zvx_vocode(zvx, mel, (int32_t)n_frames, wav) != 0) raise(zvx, "HiFiGAN::eval");
    }

    void HiFiGAN::eval_pcm16(const float *mel, int16_t *pcm, uint32_t n_frames)
    {
        if (n_frames == 0 || n_frames > max_seq_len) throw std::runtime_error("HiFiGAN::eval_pcm16: n_frames out of range");
        if (!zvx) init();
        if (zvx_vocode_pcm16(zvx, mel, (int32_t)n_frames, pcm) != 0) raise(zvx, "HiFiGAN::eval_pcm16");
    }

    // ZeroVOXModel::write_wav_file (zerovox.cpp:337-391) without libsndfile: same file format (mono WAV,
    // PCM_16), same return convention (false + message on stderr when the file cannot be written).
    bool write_wav_file_pcm16(const std::string &fname, const int16_t *pcm, size_t n_samples, uint32_t sample_rate)
    {
        if (zvx_write_wav_pcm16(fname.c_str(), pcm, (int64_t)n_samples, (int32_t)sample_rate) != 0) {
            fprintf(stderr, "Error writing %s\n", fname.c_str());
            return false;
        }
        return true;
    }
}
