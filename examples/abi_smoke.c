/* Plain C99 caller of the boundary (include/vkzg.h): proves that the header is valid C, that the library links from C and —
 * with a B200 present — commits one width-4 vector.  Built and run by tests/test_abi_and_host.py (the GPU part only when a
 * device is there: without one vkzg_ctx_create must fail with VKZG_ERR_CUDA, never fall back to a CPU path).
 *   gcc -std=c99 -Wall -Iinclude examples/abi_smoke.c -Lverkle_kzg_b200 -lvkzg -Wl,-rpath,$PWD/verkle_kzg_b200 -o /tmp/abi_smoke */
#include <stdio.h>
#include <string.h>
#include "vkzg.h"

int main(void) {
    vkzg_ctx* ctx = NULL;
    int32_t st;
    printf("abi %u\n", vkzg_abi_version());
    if (strlen(vkzg_strerror(VKZG_ERR_RANGE)) == 0) return 2;
    st = vkzg_ctx_create(&ctx, 0);
    if (st != VKZG_OK) {
        printf("no device: %s\n", vkzg_strerror(st));
        return st == VKZG_ERR_CUDA ? 0 : 3;
    }
    {
        /* the first points of the default IPA CRS as bases, scalars 1, 0, 0, 0 (Montgomery form of 1 = R mod r) */
        vkzg_g1_affine bases[4], out;
        vkzg_fr s[4];
        uint32_t key = 0;
        static const uint32_t one_mont[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        const char* seed = "eth_verkle_oct_2021";
        memset(s, 0, sizeof s);
        memcpy(s[0].l, one_mont, sizeof one_mont);
        if (vkzg_ipa_crs_generate(ctx, (const uint8_t*)seed, strlen(seed), 4, bases, NULL) != VKZG_OK) return 4;
        if (vkzg_key_load(ctx, bases, 4, NULL, VKZG_KEY_WINDOW, 8, &key) != VKZG_OK) return 5;
        if (vkzg_commit_batch(ctx, key, s, 4, 1, &out) != VKZG_OK) return 6;
        if (memcmp(&out, &bases[0], sizeof out) != 0) return 7; /* 1 * G_0 */
        printf("commit ok\n");
        vkzg_key_free(ctx, key);
    }
    vkzg_ctx_destroy(ctx);
    return 0;
}
