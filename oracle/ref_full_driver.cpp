// oracle/ref_full_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Driver around the UNMODIFIED reference class ZeroVOX::ZeroVOXModel (/root/reference/src/zerovox.cpp:21-335), i.e. the
// whole program: FastSpeech2 encoder + length regulator (fs2encoder.cpp:594-656) -> StyleTTS decoder -> HiFi-GAN.
// zerovox.cpp is compiled where it lies with -Dmain=zerovox_reference_main (its own main() loads a fixed file name) and
// against oracle/sndfile_stub/ (libsndfile is absent from this image); nothing else is changed.
//
// What it is for (SURVEY.md 8f rows f1 / f2 and BASELINE.json configs[4]):
//   * pins the length regulator: dumps the graph outputs `features` / `log_duration_prediction`, the expanded
//     hidden_state and the frame count FS2Encoder::eval returns;
//   * gives the reference's own end-to-end numbers (encoder + decoder + vocoder, max_seq_len frames) on this host.
//
// usage: zvfull <model.gguf> <inputs.bin|default> <out_prefix|-> <threads> <reps> <stages>
//   inputs.bin : int32 P, int32 src_seq[120], int32 puncts[120], float style[emb]   ("default": the sentence
//                hard-coded in ZeroVOXModel::eval, zerovox.cpp:204-314 -- run through model.eval() itself; the call
//                it makes to FS2Encoder::eval is intercepted at link time (ld --wrap, see the Makefile) and its
//                arguments are recorded, so the sentence never has to be copied out of the reference source)
//   stages     : enc | full
//   writes <out>.src.i32 .puncts.i32 .style.f32 .feat.f32 [120][emb] .logdur.f32 [120] .hidden.f32 [max_seq_len][emb]
//          and, for `full`, .mel.f32 [max_seq_len][80] .wav.f32 [max_seq_len*hop]
//   prints ONE json line with the frame count and timings to stderr.
//
// Oracle hazards: H1 vocoder compute buffer cleared before the first eval, H2 one instance per process,
// H3 stdout -> /dev/null, H4 threads through ggml_backend_cpu_set_n_threads.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

#define private public
#include "zerovox.h"
#undef private

using namespace ZeroVOX;

// ld --wrap=_ZN7ZeroVOX10FS2Encoder4evalEPKiS2_PKfjPf: every call of FS2Encoder::eval (from zerovox.cpp and from here)
// lands in __wrap_..., which records the arguments and forwards to the unmodified function (__real_...).
static std::vector<int32_t> g_src, g_puncts;
static std::vector<float> g_style;
static uint32_t g_P = 0, g_frames = 0;
extern "C" uint32_t __real__ZN7ZeroVOX10FS2Encoder4evalEPKiS2_PKfjPf(FS2Encoder *, const int32_t *, const int32_t *, const float *, uint32_t, float *);
extern "C" uint32_t __wrap__ZN7ZeroVOX10FS2Encoder4evalEPKiS2_PKfjPf(FS2Encoder *self, const int32_t *src, const int32_t *puncts,
                                                                     const float *style, uint32_t P, float *x)
{
    g_src.assign(src, src + self->max_n_phonemes);
    g_puncts.assign(puncts, puncts + self->max_n_phonemes);
    g_style.assign(style, style + self->embed_dim + self->punct_embed_dim);
    g_P = P;
    g_frames = __real__ZN7ZeroVOX10FS2Encoder4evalEPKiS2_PKfjPf(self, src, puncts, style, P, x);
    return g_frames;
}

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T>
static void dump(const std::string &path, const T *p, size_t n)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fwrite(p, sizeof(T), n, f);
    fclose(f);
}

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: %s model.gguf inputs.bin|default out_prefix|- threads reps enc|full\n", argv[0]);
        return 2;
    }
    const std::string fname = argv[1], inputs = argv[2], out = argv[3], stages = argv[6];
    const int threads = atoi(argv[4]), reps = std::max(1, atoi(argv[5]));
    const bool full = stages == "full";

    fflush(stdout);
    const int devnull = open("/dev/null", O_WRONLY);
    dup2(devnull, 1);                                                   // H3

    ZeroVOXModel model(fname);                                          // H2: the one instance of this process
    if (threads > 0) ggml_backend_cpu_set_n_threads(model.backend, threads);   // H4
    ggml_backend_buffer_clear(model.meldec->mel->buffer, 0);            // H1

    const zerovox_hparams &hp = model.hparams;
    const uint32_t emb = hp.emb_dim + hp.punct_emb_dim;
    std::vector<int32_t> src(MAX_N_PHONEMES, 0), puncts(MAX_N_PHONEMES, 0);
    std::vector<float> style(emb, 0.f);
    uint32_t P = MAX_N_PHONEMES;

    double t_model_eval = 0.0;
    if (inputs == "default") {
        // the unmodified ZeroVOXModel::eval with its hard-coded sentence (its FS2Encoder::eval call is recorded above)
        const double t0 = now_s();
        model.eval();
        t_model_eval = now_s() - t0;
        src = g_src; puncts = g_puncts; style = g_style; P = g_P;
    } else {
        FILE *f = fopen(inputs.c_str(), "rb");
        int32_t p32 = 0;
        if (!f || fread(&p32, 4, 1, f) != 1 || fread(src.data(), 4, src.size(), f) != src.size() ||
            fread(puncts.data(), 4, puncts.size(), f) != puncts.size() || fread(style.data(), 4, style.size(), f) != style.size()) {
            fprintf(stderr, "cannot read %s\n", inputs.c_str());
            return 2;
        }
        fclose(f);
        P = (uint32_t)p32;
    }

    uint32_t frames = 0;
    double t_enc = 1e30, t_dec = 1e30, t_voc = 1e30, t_tot = 1e30;
    std::string rep_list;
    for (int r = 0; r < reps; ++r) {
        const double t0 = now_s();
        frames = model.encoder->eval(src.data(), puncts.data(), style.data(), P, model.hidden_state);
        const double t1 = now_s();
        if (full) model.decoder->eval(model.hidden_state, style.data(), model.mel);
        const double t2 = now_s();
        if (full) model.meldec->eval(model.mel, model.wav);
        const double t3 = now_s();
        t_enc = std::min(t_enc, t1 - t0); t_dec = std::min(t_dec, t2 - t1); t_voc = std::min(t_voc, t3 - t2); t_tot = std::min(t_tot, t3 - t0);
        char tmp[64];
        snprintf(tmp, sizeof tmp, "%s%.6f", r ? ", " : "", t3 - t0);
        rep_list += tmp;
    }

    if (out != "-") {
        dump(out + ".src.i32", src.data(), src.size());
        dump(out + ".puncts.i32", puncts.data(), puncts.size());
        dump(out + ".style.f32", style.data(), style.size());
        dump(out + ".feat.f32", ggml_get_data_f32(model.encoder->features), (size_t)MAX_N_PHONEMES * emb);
        dump(out + ".logdur.f32", ggml_get_data_f32(model.encoder->log_duration_prediction), (size_t)MAX_N_PHONEMES);
        dump(out + ".hidden.f32", model.hidden_state, (size_t)hp.max_seq_len * emb);
        if (full) {
            dump(out + ".mel.f32", model.mel, (size_t)hp.max_seq_len * hp.audio_num_mels);
            dump(out + ".wav.f32", model.wav, (size_t)hp.max_seq_len * hp.audio_hop_size);
        }
    }
    fprintf(stderr,
            "{\"frames\": %u, \"phonemes\": %u, \"max_seq_len\": %u, \"emb\": %u, \"threads\": %d, \"reps\": %d, \"stages\": \"%s\", "
            "\"enc_s\": %.6f, \"dec_s\": %.6f, \"voc_s\": %.6f, \"total_best_s\": %.6f, \"model_eval_s\": %.6f, \"rep_s\": [%s]}\n",
            frames, P, hp.max_seq_len, emb, threads, reps, stages.c_str(), t_enc, full ? t_dec : 0.0, full ? t_voc : 0.0, t_tot,
            t_model_eval, rep_list.c_str());
    return 0;
}
