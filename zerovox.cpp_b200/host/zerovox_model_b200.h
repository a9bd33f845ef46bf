// zerovox_model_b200.h -- ZeroVOX::ZeroVOXModel with the B200 mel-decoder + vocoder behind it.
//
// Replaces the caller of the hot path, /root/reference/src/zerovox.h:405-430 + src/zerovox.cpp:21-335:
//   * same constructor (GGUF file name; hyper-parameters from the KV block, zerovox.cpp:39-56; CPU backend for the
//     FastSpeech2 encoder, which stays the reference's own ggml graph -- fs2encoder.cpp is compiled unmodified);
//   * eval() takes the sentence as ARGUMENTS (the reference hard-codes one sentence, zerovox.cpp:204-314) and USES the
//     frame count FS2Encoder::eval returns (the reference ignores it, zerovox.cpp:326-334, and always synthesises
//     max_seq_len frames): `valid_frames_only` = true synthesises exactly the frames the length regulator produced,
//     false reproduces the reference (max_seq_len frames, zero tail, statistics over the tail -- SURVEY.md N2);
//   * eval_batch() runs several sentences: the encoder per sentence on the host, ONE zvx_synth_batch call for all
//     of them on the GPU;
//   * write_wav_file() writes what was synthesised (mono WAV / PCM_16 like zerovox.cpp:337-391, converted by the
//     output conv on the GPU; no libsndfile).
// One zvx context holds decoder and vocoder, so the mel never leaves the GPU.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

struct zvx_ctx;
struct ggml_context;
struct ggml_backend;
struct ggml_backend_buffer;

namespace ZeroVOX
{
    class FS2Encoder;

    struct B200Hparams
    {
        uint32_t max_seq_len, emb_dim, punct_emb_dim, decoder_n_head, conv_filter_size, conv_kernel_size[2];
        uint32_t encoder_layer, encoder_head, encoder_vp_filter_size, encoder_vp_kernel_size, encoder_ve_n_bins;
        uint32_t audio_sampling_rate, audio_num_mels, audio_hop_size;
    };

    class ZeroVOXModelB200
    {
        public:
            explicit ZeroVOXModelB200(const std::string &fname, int n_threads = 0);
            ~ZeroVOXModelB200();

            // one sentence: src_seq / puncts hold MAX_N_PHONEMES entries (zerovox.h:37), style emb_dim + punct_emb_dim.
            // Returns the number of valid mel frames (FS2Encoder::eval's return value).
            uint32_t eval(const int32_t *src_seq, const int32_t *puncts, const float *style, uint32_t num_phonemes,
                          bool valid_frames_only = true);

            // B sentences; frames_out[b] = valid frames of sentence b.  Waveforms: wav(b), samples(b).
            void eval_batch(int B, const int32_t *const *src_seq, const int32_t *const *puncts, const float *const *style,
                            const uint32_t *num_phonemes, bool valid_frames_only, uint32_t *frames_out);

            bool write_wav_file(const std::string &fname, int b = 0);

            const int16_t *pcm(int b = 0) const { return pcm_[b].data(); }
            size_t         samples(int b = 0) const { return pcm_[b].size(); }
            uint32_t       frames(int b = 0) const { return frames_[b]; }
            const B200Hparams &hp() const { return hparams; }
            // seconds spent in the FastSpeech2 encoder (host, ggml) and in the GPU call during the last eval / eval_batch
            double last_encoder_s = 0.0, last_gpu_s = 0.0;

        private:
            void init_gpu();

            B200Hparams           hparams;
            FS2Encoder           *encoder = nullptr;
            zvx_ctx              *zvx = nullptr;
            ggml_backend         *backend = nullptr;
            ggml_backend_buffer  *buf_w = nullptr;
            ggml_context         *ctx_w = nullptr;
            std::vector<std::vector<float>>   hidden_;     // per sentence: [max_seq_len][emb]
            std::vector<std::vector<int16_t>> pcm_;
            std::vector<uint32_t>             frames_;
    };
}
