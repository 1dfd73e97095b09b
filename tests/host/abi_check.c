/* abi_check.c -- TEST INFRASTRUCTURE: include/eetq_b200.h must be a plain C header (C99, no C++ in the signatures) and a C host must
 * be able to link the library and call it.  Built and run by tests/test_cabi.py (no GPU needed: only validation paths are called). */
#include <stdio.h>
#include <string.h>

#include "../../include/eetq_b200.h"

int main(void)
{
    static char buf[1 << 16];
    int fails = 0;
    if (eetq_b200_version() != EETQ_B200_VERSION) { printf("version mismatch\n"); ++fails; }
    /* empty batch: nothing to enqueue, must succeed without a device */
    if (eetq_b200_w8a16_gemm(buf, (const int8_t*)buf, buf, NULL, buf, 0, 64, 64, EETQ_B200_F16, NULL, 0, NULL) != EETQ_B200_OK) { printf("empty batch\n"); ++fails; }
    /* K not a multiple of 64 -> EINVAL with a message */
    if (eetq_b200_w8a16_gemm(buf, (const int8_t*)buf, buf, NULL, buf, 1, 64, 100, EETQ_B200_F16, NULL, 0, NULL) != EETQ_B200_EINVAL) { printf("K check\n"); ++fails; }
    if (strstr(eetq_b200_last_error(), "multiples of 64") == NULL) { printf("message: %s\n", eetq_b200_last_error()); ++fails; }
    if (eetq_b200_quantize4(NULL, EETQ_B200_F16, 64, 64, NULL, NULL, NULL, NULL, NULL) != EETQ_B200_EINVAL) { printf("quantize4 null\n"); ++fails; }
    if (eetq_b200_w4a16_workspace_bytes(1, 4096, 4096) != 0) { printf("w4 workspace\n"); ++fails; }
    if (eetq_b200_workspace_bytes(0, 4096, 4096) != 0) { printf("workspace\n"); ++fails; }
    printf(fails ? "FAIL\n" : "OK\n");
    return fails;
}
