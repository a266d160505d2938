/* Plain-C caller of the C ABI (include/polar_b200.h): builds the float SC decoder of the reference
 * (PD/src/SCDecoder.cpp) and its CA-SCL decoder for an N=64 code, encodes random messages with pd_sim_encode, sends
 * them over a noiseless BPSK channel and checks that pd_decode returns them.  No Python, no torch:
 *     gcc -std=c99 -Iinclude examples/sc_roundtrip.c -Lquantized_decoder_polar_codes_b200 -lpolar_b200 \
 *         -Wl,-rpath,$PWD/quantized_decoder_polar_codes_b200 -o sc_roundtrip && ./sc_roundtrip
 * `./sc_roundtrip --version` only prints pd_version() (works without a GPU). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "polar_b200.h"

#define CODE_N 64
#define FRAMES 1000

static int fail(const char *what) {
    fprintf(stderr, "%s: %s\n", what, pd_last_error());
    return 1;
}

int main(int argc, char **argv) {
    if (argc > 1 && strcmp(argv[1], "--version") == 0) {
        printf("%s\n", pd_version());
        return 0;
    }
    /* frozen set: the 24 positions of lowest Hamming weight / index (a Reed-Muller-like rule is enough for a demo) */
    int32_t frozen[CODE_N];
    int K = 0;
    for (int i = 0; i < CODE_N; ++i) {
        int w = 0;
        for (int b = 0; b < 6; ++b) w += (i >> b) & 1;
        frozen[i] = (w < 3) ? 1 : 0;
        K += !frozen[i];
    }
    const int A = K - 8;                       /* CA-SCL: 8 CRC bits, x^8+x^2+x+1 */
    const int32_t crc_loc[4] = {8, 2, 1, 0};

    pd_sim_config sc;
    memset(&sc, 0, sizeof sc);
    sc.N = CODE_N; sc.K = K; sc.A = A; sc.frozen_bits = frozen; sc.crc_n = 8; sc.crc_loc = crc_loc; sc.crc_loc_len = 4;
    pd_sim *enc = NULL;
    if (pd_sim_create(&sc, &enc) != PD_OK) return fail("pd_sim_create");

    uint8_t *msg = malloc((size_t)FRAMES * A), *word = malloc((size_t)FRAMES * K), *code = malloc((size_t)FRAMES * CODE_N), *out = malloc((size_t)FRAMES * K);
    double *llr = malloc(sizeof(double) * FRAMES * CODE_N);
    srand(1);
    for (int i = 0; i < FRAMES * A; ++i) msg[i] = (uint8_t)(rand() & 1);
    if (pd_sim_encode(enc, PD_ENC_CRC, msg, FRAMES, word) != PD_OK) return fail("pd_sim_encode(CRC)");
    if (pd_sim_encode(enc, PD_ENC_CRC_POLAR, msg, FRAMES, code) != PD_OK) return fail("pd_sim_encode(CRC+polar)");
    for (int i = 0; i < FRAMES * CODE_N; ++i) llr[i] = code[i] ? -4.0 : 4.0;

    pd_config c;
    memset(&c, 0, sizeof c);
    c.kind = PD_SC; c.N = CODE_N; c.K = K; c.frozen_bits = frozen;
    pd_decoder *sc_dec = NULL;
    if (pd_create(&c, &sc_dec) != PD_OK) return fail("pd_create(SC)");
    if (pd_decode(sc_dec, llr, PD_F64, FRAMES, out) != PD_OK) return fail("pd_decode(SC)");
    if (memcmp(out, word, (size_t)FRAMES * K) != 0) { fprintf(stderr, "SC: decoded words differ\n"); return 1; }

    c.kind = PD_CASCL; c.A = A; c.L = 4; c.crc_n = 8; c.crc_loc = crc_loc; c.crc_loc_len = 4;
    pd_decoder *ca_dec = NULL;
    if (pd_create(&c, &ca_dec) != PD_OK) return fail("pd_create(CASCL)");
    if (pd_out_len(ca_dec) != A) { fprintf(stderr, "CASCL: out_len %d != A\n", pd_out_len(ca_dec)); return 1; }
    if (pd_decode(ca_dec, llr, PD_F64, FRAMES, out) != PD_OK) return fail("pd_decode(CASCL)");
    if (memcmp(out, msg, (size_t)FRAMES * A) != 0) { fprintf(stderr, "CASCL: decoded messages differ\n"); return 1; }

    printf("ok: %d frames, N=%d K=%d A=%d, kernels %s / %s, %lld launches\n", FRAMES, CODE_N, K, A, pd_kernel_name(sc_dec), pd_kernel_name(ca_dec),
           (long long)pd_launch_count());
    pd_destroy(sc_dec); pd_destroy(ca_dec); pd_sim_destroy(enc);
    free(msg); free(word); free(code); free(out); free(llr);
    return 0;
}
