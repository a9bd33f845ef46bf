// zerovox_model_b200.cpp -- see zerovox_model_b200.h.  Caller side of the hot path:
// /root/reference/src/zerovox.cpp:21-179 (load), :198-335 (eval), :337-391 (write_wav_file).
//
// The FastSpeech2 encoder + length regulator is the reference's own code (fs2encoder.cpp, compiled unmodified against
// the reference's zerovox.h and ggml, see the Makefile); decoder + vocoder are ONE zvx context of the C-ABI CUDA library.
#include "zerovox_model_b200.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "zerovox.h"            // the reference header (-I$(REF)/src): FS2Encoder, MAX_N_PHONEMES, hyper-parameter keys
#include "zvx_weights.h"

namespace ZeroVOX
{
    namespace
    {
        double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

        // required uint32 hyper-parameter (zerovox.cpp:39-56 via GGUF_GET_KEY: a missing or mistyped key is fatal)
        uint32_t kv_u32(gguf_context *g, const char *key)
        {
            const int id = gguf_find_key(g, key);
            if (id < 0) throw std::runtime_error(std::string("key not found in model: ") + key);
            if (gguf_get_kv_type(g, id) != GGUF_TYPE_UINT32) throw std::runtime_error(std::string("key has wrong type: ") + key);
            return gguf_get_val_u32(g, id);
        }
    }

    ZeroVOXModelB200::ZeroVOXModelB200(const std::string &fname, int n_threads)
    {
        gguf_init_params params = {/*.no_alloc =*/true, /*.ctx =*/&ctx_w};
        gguf_context *g = gguf_init_from_file(fname.c_str(), params);
        if (!g) throw std::runtime_error("gguf_init_from_file() failed");
        try {
            hparams.max_seq_len            = kv_u32(g, HPARAM_MAX_SEQ_LEN);
            hparams.emb_dim                = kv_u32(g, HPARAM_EMB_DIM);
            hparams.punct_emb_dim          = kv_u32(g, HPARAM_PUNCT_EMB_DIM);
            hparams.decoder_n_head         = kv_u32(g, HPARAM_DECODER_N_HEAD);
            hparams.conv_filter_size       = kv_u32(g, HPARAM_CONV_FILTER_SIZE);
            hparams.conv_kernel_size[0]    = kv_u32(g, HPARAM_CONV_KERNEL_SIZE_0);
            hparams.conv_kernel_size[1]    = kv_u32(g, HPARAM_CONV_KERNEL_SIZE_1);
            hparams.encoder_layer          = kv_u32(g, HPARAM_ENCODER_LAYER);
            hparams.encoder_head           = kv_u32(g, HPARAM_ENCODER_HEAD);
            hparams.encoder_vp_filter_size = kv_u32(g, HPARAM_ENCODER_VP_FILTER_SIZE);
            hparams.encoder_vp_kernel_size = kv_u32(g, HPARAM_ENCODER_VP_KERNEL_SIZE);
            hparams.encoder_ve_n_bins      = kv_u32(g, HPARAM_ENCODER_VE_N_BINS);
            hparams.audio_sampling_rate    = kv_u32(g, HPARAM_AUDIO_SAMPLING_RATE);
            hparams.audio_num_mels         = kv_u32(g, HPARAM_AUDIO_NUM_MELS);
            hparams.audio_hop_size         = kv_u32(g, HPARAM_AUDIO_HOP_SIZE);

            backend = ggml_backend_cpu_init();
            if (!backend) throw std::runtime_error("ggml_backend_cpu_init() failed");
            if (n_threads > 0) ggml_backend_cpu_set_n_threads(backend, n_threads);
            buf_w = ggml_backend_alloc_ctx_tensors(ctx_w, backend);
            if (!buf_w) throw std::runtime_error("ggml_backend_alloc_ctx_tensors() failed");

            // the encoder graph is built before the tensor data is read, as in the reference (zerovox.cpp:104-115,140-172)
            encoder = new FS2Encoder(*ctx_w, backend, MAX_N_PHONEMES, hparams.emb_dim, hparams.punct_emb_dim, hparams.encoder_layer,
                                     hparams.encoder_head, hparams.conv_filter_size, hparams.conv_kernel_size,
                                     hparams.encoder_vp_kernel_size, hparams.encoder_ve_n_bins, hparams.max_seq_len);

            FILE *f = fopen(fname.c_str(), "rb");
            if (!f) throw std::runtime_error("fopen() failed");
            std::vector<uint8_t> buf;
            for (int i = 0; i < (int)gguf_get_n_tensors(g); i++) {
                ggml_tensor *t = ggml_get_tensor(ctx_w, gguf_get_tensor_name(g, i));
                const size_t offs = gguf_get_data_offset(g) + gguf_get_tensor_offset(g, i);
                buf.resize(ggml_nbytes(t));
                if (fseek(f, (long)offs, SEEK_SET) != 0 || fread(buf.data(), 1, buf.size(), f) != buf.size()) {
                    fclose(f);
                    throw std::runtime_error("reading tensor data failed");
                }
                ggml_backend_tensor_set(t, buf.data(), 0, buf.size());
            }
            fclose(f);
            init_gpu();
        } catch (...) {
            gguf_free(g);
            throw;
        }
        gguf_free(g);
    }

    ZeroVOXModelB200::~ZeroVOXModelB200()
    {
        zvx_destroy(zvx);
        delete encoder;
        if (buf_w) ggml_backend_buffer_free(buf_w);
        if (backend) ggml_backend_free(backend);
        if (ctx_w) ggml_free(ctx_w);
    }

    // decoder + vocoder in one context: the constructor arguments ZeroVOXModel passes to the two stages (zerovox.cpp:117-138)
    void ZeroVOXModelB200::init_gpu()
    {
        WeightSet ws;
        collect(ctx_w, {"_mel_decoder.", "_meldec.", "hifigan."}, ws);
        zvx_config cfg;
        zvx_default_config(&cfg);
        if (const char *e = getenv("ZVX_DEVICE")) cfg.device = atoi(e);
        const uint32_t emb = hparams.emb_dim + hparams.punct_emb_dim;
        cfg.dim_in = (int32_t)emb;
        cfg.style_dim = (int32_t)emb;
        cfg.residual_dim = 64;
        cfg.num_mels = (int32_t)hparams.audio_num_mels;
        cfg.hop_size = (int32_t)hparams.audio_hop_size;
        if (zvx_create(&zvx, &cfg, ws.descs.data(), (int32_t)ws.descs.size()) != 0)
            throw std::runtime_error(std::string("zvx_create: ") + zvx_last_error(nullptr));
        if (zvx_reserve(zvx, hparams.max_seq_len, 1) != 0) throw std::runtime_error(std::string("zvx_reserve: ") + zvx_last_error(zvx));
    }

    uint32_t ZeroVOXModelB200::eval(const int32_t *src_seq, const int32_t *puncts, const float *style, uint32_t num_phonemes,
                                    bool valid_frames_only)
    {
        uint32_t frames = 0;
        eval_batch(1, &src_seq, &puncts, &style, &num_phonemes, valid_frames_only, &frames);
        return frames;
    }

    void ZeroVOXModelB200::eval_batch(int B, const int32_t *const *src_seq, const int32_t *const *puncts, const float *const *style,
                                      const uint32_t *num_phonemes, bool valid_frames_only, uint32_t *frames_out)
    {
        if (B <= 0) throw std::runtime_error("eval_batch: empty batch");
        const uint32_t emb = hparams.emb_dim + hparams.punct_emb_dim, T = hparams.max_seq_len, hop = hparams.audio_hop_size;
        hidden_.resize((size_t)B);
        pcm_.resize((size_t)B);
        frames_.assign((size_t)B, 0);
        std::vector<const float *> pe((size_t)B);
        std::vector<int16_t *> pp((size_t)B);
        std::vector<int32_t> L((size_t)B);
        const double t0 = now_s();
        for (int b = 0; b < B; ++b) {
            hidden_[b].resize((size_t)T * emb);
            // FastSpeech2 encoder + variance adaptor + length regulator on the host (fs2encoder.cpp:594-656); the frame
            // count it returns is what the reference's caller drops (zerovox.cpp:326)
            frames_[b] = encoder->eval(src_seq[b], puncts[b], style[b], num_phonemes[b], hidden_[b].data());
            if (frames_[b] == 0 && valid_frames_only) throw std::runtime_error("eval: the length regulator produced no frames");
            L[b] = (int32_t)(valid_frames_only ? frames_[b] : T);
            pcm_[b].resize((size_t)L[b] * hop);
            pe[b] = hidden_[b].data();
            pp[b] = pcm_[b].data();
            if (frames_out) frames_out[b] = frames_[b];
        }
        const double t1 = now_s();
        if (zvx_synth_batch_pcm16(zvx, B, pe.data(), style, L.data(), nullptr, pp.data()) != 0)
            throw std::runtime_error(std::string("zvx_synth_batch_pcm16: ") + zvx_last_error(zvx));
        last_encoder_s = t1 - t0;
        last_gpu_s = now_s() - t1;
    }

    bool ZeroVOXModelB200::write_wav_file(const std::string &fname, int b)
    {
        if (b < 0 || (size_t)b >= pcm_.size()) return false;
        if (zvx_write_wav_pcm16(fname.c_str(), pcm_[b].data(), (int64_t)pcm_[b].size(), (int32_t)hparams.audio_sampling_rate) != 0) {
            fprintf(stderr, "Error writing %s\n", fname.c_str());
            return false;
        }
        return true;
    }
}
