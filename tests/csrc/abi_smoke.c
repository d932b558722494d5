/* Plain-C consumer of include/ppgs_b200.h: what a maintainer binding the library from another
 * language would write.  Only host-side entry points are called, so it runs without a GPU:
 * WAVE probe / decode, the resampler's filter table, the torch.load-compatible writer, and
 * the error conventions (status code + ppgs_last_error). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ppgs_b200.h"

static int fail(const char* what) {
    fprintf(stderr, "abi_smoke: %s failed: %s\n", what, ppgs_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: abi_smoke in.wav out.pt\n");
        return 2;
    }
    if (ppgs_abi_version() != PPGS_ABI_VERSION) return fail("abi version");

    ppgs_model_config cfg;
    ppgs_default_config(&cfg);
    if (cfg.hidden_channels != 256 || cfg.output_channels != 40) return fail("default config");

    int64_t frames = 0;
    int rate = 0, channels = 0, bits = 0, is_float = 0;
    if (ppgs_wav_info(argv[1], &frames, &rate, &channels, &bits, &is_float) != PPGS_OK) return fail("wav_info");
    float* audio = (float*)malloc((size_t)frames * sizeof(float));
    int64_t got = 0;
    if (ppgs_wav_read_f32(argv[1], audio, frames, &got, &rate) != PPGS_OK || got != frames) return fail("wav_read");

    int ntaps = 0, phases = 0, width = 0;
    if (ppgs_resample_taps(44100, 16000, NULL, 0, &ntaps, &phases, &width) != PPGS_OK) return fail("resample_taps");
    if (ppgs_resample_length(44100, 44100, 16000) != 16000) return fail("resample_length");

    /* the "posteriorgram": 40 rows over the decoded samples, cropped to frames / 160 columns */
    const int64_t cols = frames / 160;
    if (ppgs_pt_write_f32(argv[2], audio, 40, cols, cols) != PPGS_OK) return fail("pt_write");

    /* error convention: status + message, no abort */
    if (ppgs_wav_info("/nonexistent/file.wav", &frames, NULL, NULL, NULL, NULL) != PPGS_E_INVALID) return fail("error code");
    if (!strstr(ppgs_last_error(), "cannot open")) return fail("error text");
    if (ppgs_engine_finalize(NULL) != PPGS_E_INVALID) return fail("NULL engine");

    printf("%lld %d %d %d %d %d %d %d %lld\n", (long long)got, rate, channels, bits, is_float, ntaps, phases, width,
           (long long)cols);
    free(audio);
    return 0;
}
