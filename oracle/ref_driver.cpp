// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Command-line driver around the UNMODIFIED reference classes
// ZeroVOX::StyleTTSDecoder (/root/reference/src/stylettsdec.cpp:306-470) and
// ZeroVOX::HiFiGAN (/root/reference/src/hifigan.cpp:187-377).  It is compiled by
// oracle/Makefile together with the reference sources where they lie; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may execute the resulting binary.
//
// It follows the load sequence of ZeroVOXModel::ZeroVOXModel
// (/root/reference/src/zerovox.cpp:28-35,86-91,140-172): gguf_init_from_file with
// no_alloc, CPU backend, ggml_backend_alloc_ctx_tensors, then one read per tensor.
//
// usage: zvref <model.gguf> <L> <enc_seq.f32|-> <style.f32|-> <mel_in.f32|-> <out_prefix|-> <threads> <reps> <stage>
//   stage = both | dec | voc       (voc reads mel from <mel_in.f32>)
//   writes <out_prefix>.mel.f32 / <out_prefix>.wav.f32 (raw little-endian f32)
//   prints ONE json line with timings to stderr; the reference's stdout spam
//   (hifigan.cpp:368-372) is sent to /dev/null.
//
// Oracle hazards handled here: H1 the vocoder's zero-stuffed
// buffer is never initialised by the reference (hifigan.cpp:50-54) -> the
// compute buffer is cleared once before the first eval; H2 one instance per
// class per process -> one process per L; H3 stdout redirected; H4 threads set
// through ggml_backend_cpu_set_n_threads.

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <unistd.h>
#include <fcntl.h>

#define private public
#include "zerovox.h"
#undef private

using namespace ZeroVOX;

static std::vector<float> read_f32(const char *path, size_t n)
{
    std::vector<float> v(n);
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    if (fread(v.data(), sizeof(float), n, f) != n) { fprintf(stderr, "short read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

static void write_f32(const std::string &path, const float *p, size_t n)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fwrite(p, sizeof(float), n, f);
    fclose(f);
}

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
    if (argc < 10) {
        fprintf(stderr, "usage: %s model.gguf L enc_seq.f32 style.f32 mel_in.f32 out_prefix threads reps stage\n", argv[0]);
        return 2;
    }
    const char *fname   = argv[1];
    const uint32_t L    = (uint32_t)atoi(argv[2]);
    const char *enc_p   = argv[3];
    const char *sty_p   = argv[4];
    const char *mel_p   = argv[5];
    const std::string out_prefix = argv[6];
    const int threads   = atoi(argv[7]);
    const int reps      = atoi(argv[8]);
    const std::string stage = argv[9];
    const bool do_dec = stage == "both" || stage == "dec";
    const bool do_voc = stage == "both" || stage == "voc";

    // H3: keep the reference's prints off our stdout
    fflush(stdout);
    int devnull = open("/dev/null", O_WRONLY);
    dup2(devnull, 1);

    struct ggml_context *ctx_w = nullptr;
    struct gguf_init_params gparams = { /*.no_alloc =*/ true, /*.ctx =*/ &ctx_w };
    struct gguf_context *ctx_gguf = gguf_init_from_file(fname, gparams);
    if (!ctx_gguf) { fprintf(stderr, "gguf_init_from_file failed\n"); return 2; }

    const int kid_mels = gguf_find_key(ctx_gguf, HPARAM_AUDIO_NUM_MELS);
    const int kid_hop  = gguf_find_key(ctx_gguf, HPARAM_AUDIO_HOP_SIZE);
    const int kid_emb  = gguf_find_key(ctx_gguf, HPARAM_EMB_DIM);
    const int kid_pun  = gguf_find_key(ctx_gguf, HPARAM_PUNCT_EMB_DIM);
    if (kid_mels < 0 || kid_hop < 0 || kid_emb < 0 || kid_pun < 0) { fprintf(stderr, "missing hparams\n"); return 2; }
    const uint32_t num_mels = gguf_get_val_u32(ctx_gguf, kid_mels);
    const uint32_t hop      = gguf_get_val_u32(ctx_gguf, kid_hop);
    const uint32_t emb_size = gguf_get_val_u32(ctx_gguf, kid_emb) + gguf_get_val_u32(ctx_gguf, kid_pun);

    ggml_backend_t backend = ggml_backend_cpu_init();
    if (threads > 0) ggml_backend_cpu_set_n_threads(backend, threads);   // H4
    ggml_backend_buffer_t buf_w = ggml_backend_alloc_ctx_tensors(ctx_w, backend);
    if (!buf_w) { fprintf(stderr, "alloc weights failed\n"); return 2; }

    // H1/H2: construct right after start, one instance per class
    StyleTTSDecoder *decoder = nullptr;
    HiFiGAN *meldec = nullptr;
    if (do_dec)
        decoder = new StyleTTSDecoder(*ctx_w, backend, L, emb_size, emb_size, 64, num_mels);
    const int upsample_scales[4] = {5, 5, 4, 3};
    const int64_t dilations[9]   = {1, 3, 5, 1, 3, 5, 1, 3, 5};
    if (do_voc) {
        meldec = new HiFiGAN(*ctx_w, backend, L, num_mels, hop, 7, 4, upsample_scales, 3, 3, dilations);
        // H1: the graph's compute buffer holds the never-written zero-stuffing gaps
        ggml_backend_buffer_clear(meldec->mel->buffer, 0);
    }

    FILE *f = fopen(fname, "rb");
    const int n_tensors = gguf_get_n_tensors(ctx_gguf);
    for (int i = 0; i < n_tensors; i++) {
        const char *name = gguf_get_tensor_name(ctx_gguf, i);
        struct ggml_tensor *t = ggml_get_tensor(ctx_w, name);
        size_t offs = gguf_get_data_offset(ctx_gguf) + gguf_get_tensor_offset(ctx_gguf, i);
        std::vector<uint8_t> b(ggml_nbytes(t));
        if (fseek(f, (long)offs, SEEK_SET) != 0 || fread(b.data(), 1, b.size(), f) != b.size()) {
            fprintf(stderr, "read tensor %s failed\n", name); return 2;
        }
        ggml_backend_tensor_set(t, b.data(), 0, b.size());
    }
    fclose(f);
    gguf_free(ctx_gguf);

    std::vector<float> enc, sty, mel((size_t)L * num_mels), wav((size_t)L * hop);
    if (do_dec) {
        enc = read_f32(enc_p, (size_t)L * emb_size);
        sty = read_f32(sty_p, emb_size);
    } else {
        mel = read_f32(mel_p, (size_t)L * num_mels);
    }

    double t_dec_best = 1e30, t_voc_best = 1e30, t_tot_best = 1e30, t_tot_sum = 0;
    std::string rep_list;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        if (do_dec) decoder->eval(enc.data(), sty.data(), mel.data());
        double t1 = now_s();
        if (do_voc) meldec->eval(mel.data(), wav.data());
        double t2 = now_s();
        if (t1 - t0 < t_dec_best) t_dec_best = t1 - t0;
        if (t2 - t1 < t_voc_best) t_voc_best = t2 - t1;
        if (t2 - t0 < t_tot_best) t_tot_best = t2 - t0;
        if (r > 0 || reps == 1) t_tot_sum += t2 - t0;
        char tmp[64];
        snprintf(tmp, sizeof tmp, "%s%.6f", r ? ", " : "", t2 - t0);
        rep_list += tmp;
    }
    const int timed = reps > 1 ? reps - 1 : 1;

    if (out_prefix != "-") {
        if (do_dec) write_f32(out_prefix + ".mel.f32", mel.data(), mel.size());
        if (do_voc) write_f32(out_prefix + ".wav.f32", wav.data(), wav.size());
    }

    fprintf(stderr,
            "{\"L\": %u, \"threads\": %d, \"reps\": %d, \"stage\": \"%s\", \"dec_s\": %.6f, \"voc_s\": %.6f, "
            "\"total_best_s\": %.6f, \"total_mean_s\": %.6f, \"audio_s\": %.6f, \"rep_s\": [%s]}\n",
            L, threads, reps, stage.c_str(), do_dec ? t_dec_best : 0.0, do_voc ? t_voc_best : 0.0,
            t_tot_best, t_tot_sum / timed, (double)L * hop / 24000.0, rep_list.c_str());
    return 0;
}
