/* The C ABI consumed from plain C (no CUDA headers, no C++): compiles include/dnaf_b200.h with gcc -std=c99 -pedantic,
 * links libdnaf_b200.so and calls the entry points that need no GPU.  Run by tests/test_host_mirror.py. */
#include <stdio.h>
#include <string.h>

#include "dnaf_b200.h"

int main(void) {
    uint8_t eof[28];
    uint32_t csize[4], usize[4];
    uint64_t n_blocks = 0;
    static const uint8_t want[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43,
                                     0x02, 0x00, 0x1b, 0x00, 0x03, 0x00, 0, 0, 0, 0, 0, 0, 0, 0};
    if (dnaf_abi_version() != DNAF_ABI_VERSION) return 1;
    if (dnaf_bgzf_eof(eof) != DNAF_OK || memcmp(eof, want, 28) != 0) return 2;
    if (dnaf_bgzf_scan(eof, 28, csize, usize, 4, &n_blocks) != DNAF_OK || n_blocks != 1 || csize[0] != 28 || usize[0] != 0) return 3;
    if (dnaf_bgzf_bound(65280) < 65280 + 26) return 4;
    if (dnaf_device_count() < 0) return 5;
    {   /* without a GPU dnaf_create must fail loudly, with a message, and leave *out NULL */
        dnaf_ctx* ctx = (dnaf_ctx*)1;
        int rc = dnaf_create(0, &ctx);
        if (dnaf_device_count() == 0 && (rc == DNAF_OK || ctx != NULL || strlen(dnaf_last_error(NULL)) == 0)) return 6;
        if (rc == DNAF_OK) dnaf_destroy(ctx);
    }
    printf("abi %d ok, %d device(s)\n", dnaf_abi_version(), dnaf_device_count());
    return 0;
}
