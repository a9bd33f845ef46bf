// oracle/sndfile_stub/sndfile_stub.cpp -- TEST INFRASTRUCTURE, not product code.
//
// The part of libsndfile the reference's write_wav_file (zerovox.cpp:354-384) exercises, restated from
// libsndfile's published behaviour (it is not under /root/reference): sf_open(SFM_WRITE, WAV | PCM_16) writes the
// canonical 44-byte RIFF header, sf_write_float converts with normalisation on and clipping off --
// pcm.c f2s_array: lrintf(x * 0x7FFF), the result narrowed to short -- and sf_close patches the two sizes.
// Pinned by tests/test_output_stage_cpu.py against zv_oracle.pcm16 / Python's `wave` reader.
#include "sndfile.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

struct SNDFILE_tag {
    FILE *f;
    int rate;
    int64_t samples;
};

static const char *g_err = "no error";

static void put_u32(unsigned char *p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; }
static void put_u16(unsigned char *p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; }

static bool write_header(SNDFILE *s)
{
    unsigned char h[44];
    const uint32_t data = (uint32_t)(2 * s->samples);
    memcpy(h, "RIFF", 4); put_u32(h + 4, 36u + data); memcpy(h + 8, "WAVE", 4);
    memcpy(h + 12, "fmt ", 4); put_u32(h + 16, 16u); put_u16(h + 20, 1u); put_u16(h + 22, 1u);
    put_u32(h + 24, (uint32_t)s->rate); put_u32(h + 28, 2u * (uint32_t)s->rate); put_u16(h + 32, 2u); put_u16(h + 34, 16u);
    memcpy(h + 36, "data", 4); put_u32(h + 40, data);
    if (fseek(s->f, 0, SEEK_SET) != 0) return false;
    return fwrite(h, 1, sizeof h, s->f) == sizeof h;
}

extern "C" SNDFILE *sf_open(const char *path, int mode, SF_INFO *info)
{
    if (mode != SFM_WRITE || !info || info->channels != 1 || info->format != (SF_FORMAT_WAV | SF_FORMAT_PCM_16)) {
        g_err = "sndfile stub: only mono WAV/PCM_16 writing is implemented";
        return nullptr;
    }
    FILE *f = fopen(path, "wb");
    if (!f) { g_err = "sndfile stub: cannot open file"; return nullptr; }
    SNDFILE *s = new SNDFILE_tag{f, info->samplerate, 0};
    if (!write_header(s)) { fclose(f); delete s; g_err = "sndfile stub: header write failed"; return nullptr; }
    return s;
}

extern "C" sf_count_t sf_write_float(SNDFILE *s, const float *ptr, sf_count_t items)
{
    if (!s || !ptr) return 0;
    std::vector<unsigned char> buf((size_t)items * 2);
    for (sf_count_t i = 0; i < items; ++i) {
        const short v = (short)lrintf(ptr[i] * (1.0f * 0x7FFF));
        put_u16(buf.data() + 2 * i, (uint16_t)v);
    }
    if (fseek(s->f, 44 + 2 * s->samples, SEEK_SET) != 0) return 0;
    if (fwrite(buf.data(), 1, buf.size(), s->f) != buf.size()) return 0;
    s->samples += items;
    return items;
}

extern "C" int sf_close(SNDFILE *s)
{
    if (!s) return 1;
    const bool ok = write_header(s);
    const int rc = fclose(s->f);
    delete s;
    return (ok && rc == 0) ? 0 : 1;
}

extern "C" const char *sf_strerror(SNDFILE *) { return g_err; }
