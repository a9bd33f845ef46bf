/* oracle/sndfile_stub/sndfile.h -- TEST INFRASTRUCTURE, not product code.
 *
 * libsndfile is an unvendored system dependency of the reference (/root/reference/CMakeLists.txt:7-8,
 * src/zerovox.cpp:12) and is absent from this image.  This header declares exactly the six names
 * zerovox.cpp:354-384 uses, with libsndfile's public values (sndfile.h of libsndfile 1.x: SF_FORMAT_WAV
 * 0x010000, SF_FORMAT_PCM_16 0x0002, SFM_WRITE 0x20), so that the UNMODIFIED zerovox.cpp compiles;
 * sndfile_stub.cpp implements them for the one case the reference needs (mono WAV / PCM_16 writing). */
#ifndef ZVX_ORACLE_SNDFILE_STUB_H
#define ZVX_ORACLE_SNDFILE_STUB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t sf_count_t;
typedef struct SNDFILE_tag SNDFILE;

typedef struct SF_INFO {
    sf_count_t frames;
    int samplerate;
    int channels;
    int format;
    int sections;
    int seekable;
} SF_INFO;

enum { SF_FORMAT_WAV = 0x010000, SF_FORMAT_PCM_16 = 0x0002 };
enum { SFM_READ = 0x10, SFM_WRITE = 0x20, SFM_RDWR = 0x30 };

SNDFILE *sf_open(const char *path, int mode, SF_INFO *sfinfo);
sf_count_t sf_write_float(SNDFILE *sndfile, const float *ptr, sf_count_t items);
int sf_close(SNDFILE *sndfile);
const char *sf_strerror(SNDFILE *sndfile);

#ifdef __cplusplus
}
#endif
#endif
