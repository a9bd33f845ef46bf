// zvx_model_main.cpp -- the reference program (zerovox.cpp:396-406: load the GGUF, eval, write a WAV) with the B200
// mel-decoder + vocoder behind ZeroVOXModelB200, the sentence(s) read from files instead of being hard-coded.
//
// usage: zvx_model model.gguf out_prefix [--reference-default] [--threads T] [--bench N] inputs.bin [inputs.bin ...]
//   inputs.bin : int32 P, int32 src_seq[120], int32 puncts[120], float style[emb]
//   writes <out_prefix>.<b>.wav (mono PCM_16) and <out_prefix>.<b>.pcm.i16 (raw samples) per sentence and prints one
//   JSON line: frames per sentence, seconds in the FastSpeech2 encoder (host, ggml) and in the GPU call; with --bench N
//   the first sentence is synthesised N more times and p50 / p99 of encoder, GPU and total latency are reported
//   (BASELINE.json configs[4]).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "zerovox_model_b200.h"

using namespace ZeroVOX;

struct Sentence {
    uint32_t P = 0;
    std::vector<int32_t> src, puncts;
    std::vector<float> style;
};

static bool read_sentence(const char *path, uint32_t emb, Sentence &s)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    int32_t p = 0;
    s.src.resize(120); s.puncts.resize(120); s.style.resize(emb);
    const bool ok = fread(&p, 4, 1, f) == 1 && fread(s.src.data(), 4, 120, f) == 120 && fread(s.puncts.data(), 4, 120, f) == 120 &&
                    fread(s.style.data(), 4, emb, f) == emb;
    fclose(f);
    s.P = (uint32_t)p;
    return ok && p > 0 && p <= 120;
}

static double pct(std::vector<double> v, double q)
{
    std::sort(v.begin(), v.end());
    return v[std::min(v.size() - 1, (size_t)(q * (double)v.size()))];
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s model.gguf out_prefix [--reference-default] [--threads T] [--bench N] inputs.bin [...]\n", argv[0]);
        return 2;
    }
    const std::string fname = argv[1], out = argv[2];
    bool valid_only = true;
    int threads = 0, bench = 0;
    std::vector<const char *> files;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--reference-default")) valid_only = false;
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--bench") && i + 1 < argc) bench = atoi(argv[++i]);
        else files.push_back(argv[i]);
    }
    try {
        ZeroVOXModelB200 model(fname, threads);
        const uint32_t emb = model.hp().emb_dim + model.hp().punct_emb_dim;
        std::vector<Sentence> s(files.size());
        for (size_t b = 0; b < files.size(); ++b)
            if (!read_sentence(files[b], emb, s[b])) { fprintf(stderr, "cannot read %s\n", files[b]); return 2; }
        const int B = (int)s.size();
        if (B == 0) { fprintf(stderr, "no input sentences\n"); return 2; }
        std::vector<const int32_t *> ps(B), pp(B);
        std::vector<const float *> pst(B);
        std::vector<uint32_t> P(B), frames(B);
        for (int b = 0; b < B; ++b) { ps[b] = s[b].src.data(); pp[b] = s[b].puncts.data(); pst[b] = s[b].style.data(); P[b] = s[b].P; }
        model.eval_batch(B, ps.data(), pp.data(), pst.data(), P.data(), valid_only, frames.data());
        std::string fr;
        for (int b = 0; b < B; ++b) {
            const std::string stem = out + "." + std::to_string(b);
            if (!model.write_wav_file(stem + ".wav", b)) return 1;
            FILE *f = fopen((stem + ".pcm.i16").c_str(), "wb");
            if (!f || fwrite(model.pcm(b), 2, model.samples(b), f) != model.samples(b)) { fprintf(stderr, "cannot write %s\n", stem.c_str()); return 1; }
            fclose(f);
            fr += (b ? ", " : "") + std::to_string(frames[b]);
        }
        printf("{\"sentences\": %d, \"mode\": \"%s\", \"frames\": [%s], \"encoder_s\": %.6f, \"gpu_s\": %.6f", B,
               valid_only ? "valid_frames" : "reference_default", fr.c_str(), model.last_encoder_s, model.last_gpu_s);
        if (bench > 0) {
            std::vector<double> te, tg, tt;
            for (int r = 0; r < bench + 3; ++r) {
                model.eval(ps[0], pp[0], pst[0], P[0], valid_only);
                if (r < 3) continue;                    // warm-up (CUDA graph capture happens on the second call of a length)
                te.push_back(model.last_encoder_s); tg.push_back(model.last_gpu_s); tt.push_back(model.last_encoder_s + model.last_gpu_s);
            }
            printf(", \"bench_runs\": %d, \"encoder_ms\": {\"p50\": %.3f, \"p99\": %.3f}, \"gpu_ms\": {\"p50\": %.3f, \"p99\": %.3f}, "
                   "\"total_ms\": {\"p50\": %.3f, \"p99\": %.3f}",
                   bench, 1e3 * pct(te, 0.5), 1e3 * pct(te, 0.99), 1e3 * pct(tg, 0.5), 1e3 * pct(tg, 0.99), 1e3 * pct(tt, 0.5), 1e3 * pct(tt, 0.99));
        }
        printf("}\n");
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
