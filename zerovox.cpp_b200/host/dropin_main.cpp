// dropin_main.cpp -- the reference's own load-and-run sequence with the B200 classes dropped in.
//
// Mirrors ZeroVOXModel::ZeroVOXModel / ::eval (/root/reference/src/zerovox.cpp:28-35,86-91,
// 104-138,140-172,330-334) for the mel-decoder + vocoder part: gguf_init_from_file(no_alloc),
// CPU backend, ggml_backend_alloc_ctx_tensors, construct the stages, THEN read the tensor data,
// decoder->eval, meldec->eval.  The only difference from the reference program is which header
// declares the two classes.  Used by tests/test_dropin_gpu.py; built only where the ggml sources
// of the reference tree are available (this container), the binary travels to the GPU box.
//
// usage: zvx_dropin model.gguf L enc_seq.f32 style.f32 out_prefix
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"
#include "ggml-cpu.h"

#include "zerovox_b200.h"

using namespace ZeroVOX;

static std::vector<float> read_f32(const char *path, size_t n)
{
    std::vector<float> v(n);
    FILE *f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(float), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}
static void write_f32(const std::string &path, const std::vector<float> &v)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f || fwrite(v.data(), sizeof(float), v.size(), f) != v.size()) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fclose(f);
}
static uint32_t kv_u32(gguf_context *g, const char *key)
{
    const int id = gguf_find_key(g, key);
    if (id < 0) { fprintf(stderr, "error: key not found in model: %s\n", key); exit(1); }   // die_fmt, zerovox.h:435-455
    return gguf_get_val_u32(g, id);
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s model.gguf L enc_seq.f32 style.f32 out_prefix\n", argv[0]); return 2; }
    const char *fname = argv[1];
    const uint32_t L = (uint32_t)atoi(argv[2]);

    ggml_context *ctx_w = nullptr;
    gguf_init_params gp = {/*no_alloc*/ true, /*ctx*/ &ctx_w};
    gguf_context *g = gguf_init_from_file(fname, gp);
    if (!g) { fprintf(stderr, "gguf_init_from_file failed\n"); return 2; }
    const uint32_t num_mels = kv_u32(g, "zerovox-resnet-fs2-styletts.audio.num_mels");
    const uint32_t hop      = kv_u32(g, "zerovox-resnet-fs2-styletts.audio.hop_size");
    const uint32_t emb_size = kv_u32(g, "zerovox-resnet-fs2-styletts.emb_dim") + kv_u32(g, "zerovox-resnet-fs2-styletts.punct_emb_dim");

    ggml_backend_t backend = ggml_backend_cpu_init();
    ggml_backend_buffer_t buf_w = ggml_backend_alloc_ctx_tensors(ctx_w, backend);
    if (!buf_w) { fprintf(stderr, "alloc weights failed\n"); return 2; }

    // stages first ... (zerovox.cpp:119-138)
    StyleTTSDecoder decoder(*ctx_w, backend, L, emb_size, emb_size, 64, num_mels);
    const int upsample_scales[4] = {5, 5, 4, 3};
    const int64_t dilations[9] = {1, 3, 5, 1, 3, 5, 1, 3, 5};
    HiFiGAN meldec(*ctx_w, backend, L, num_mels, hop, 7, 4, upsample_scales, 3, 3, dilations);

    // ... then the tensor data (zerovox.cpp:140-172)
    FILE *f = fopen(fname, "rb");
    for (int i = 0; i < (int)gguf_get_n_tensors(g); i++) {
        const char *name = gguf_get_tensor_name(g, i);
        ggml_tensor *t = ggml_get_tensor(ctx_w, name);
        const size_t offs = gguf_get_data_offset(g) + gguf_get_tensor_offset(g, i);
        std::vector<uint8_t> b(ggml_nbytes(t));
        if (fseek(f, (long)offs, SEEK_SET) != 0 || fread(b.data(), 1, b.size(), f) != b.size()) { fprintf(stderr, "read %s failed\n", name); return 2; }
        ggml_backend_tensor_set(t, b.data(), 0, b.size());
    }
    fclose(f);

    const std::vector<float> enc = read_f32(argv[3], (size_t)L * emb_size), sty = read_f32(argv[4], emb_size);
    std::vector<float> mel((size_t)L * num_mels), wav((size_t)L * hop);
    try {
        decoder.eval(enc.data(), sty.data(), mel.data());     // zerovox.cpp:330
        meldec.eval(mel.data(), wav.data());                  // zerovox.cpp:334
        // zerovox.cpp:337-391 (write_wav_file): PCM_16 conversion on the GPU, RIFF file without libsndfile
        std::vector<int16_t> pcm((size_t)L * hop);
        meldec.eval_pcm16(mel.data(), pcm.data(), (uint32_t)L);
        if (!ZeroVOX::write_wav_file_pcm16(std::string(argv[5]) + ".wav", pcm.data(), pcm.size(), 24000)) return 1;
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    write_f32(std::string(argv[5]) + ".mel.f32", mel);
    write_f32(std::string(argv[5]) + ".wav.f32", wav);
    gguf_free(g);
    return 0;
}
