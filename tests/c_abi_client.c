/* A plain C11 client of include/loghisto_b200.h: proves the header is valid C (not just C++), that the library links
 * from C, and that without a GPU lh_create fails with LH_ERR_NO_DEVICE instead of falling back.  Built and run by
 * tests/test_abi.py::test_plain_c_client. */
#include <stdio.h>
#include <string.h>
#include "loghisto_b200.h"

int main(void) {
    if (lh_abi_version() != LH_ABI_VERSION) { printf("abi mismatch\n"); return 2; }
    lh_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (uint32_t)sizeof cfg;
    cfg.max_histograms = 4;
    cfg.max_counters = 4;
    lh_ctx *ctx = NULL;
    lh_status st = lh_create(&cfg, &ctx);
    if (st == LH_OK) {
        /* a GPU is present: run one tiny end-to-end pass through the C ABI */
        double vals[3] = {33.0, 59.0, 330000.0};   /* metrics_test.go:297-299 */
        uint64_t counts[4]; double sums[4], avgs[4], pvals[4]; int32_t pkeys[4];
        double p50 = 0.5;
        if (lh_ingest_f64_host(ctx, 1, vals, 3) != LH_OK) return 3;
        if (lh_snapshot_begin(ctx) != LH_OK) return 4;
        if (lh_snapshot_reduce(ctx, &p50, 1, counts, sums, avgs, pkeys, pvals) != LH_OK) return 5;
        if (lh_snapshot_end(ctx) != LH_OK) return 6;
        printf("gpu count=%llu sum=%d p50key=%d\n", (unsigned long long)counts[1], (int)sums[1], pkeys[1]);
        lh_destroy(ctx);
        return (counts[1] == 3 && (int)sums[1] == 331132 && pkeys[1] == 409) ? 0 : 7;
    }
    printf("no gpu: %s\n", lh_strerror(st));
    return st == LH_ERR_NO_DEVICE ? 0 : 8;
}
