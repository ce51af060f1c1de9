/* A plain C11 host that drives N GPUs without Python, torch or NCCL: one pthread and one lh_ctx per GPU (both ranks on
 * device 0 when the box has a single GPU), the sample stream sharded as SURVEY.md section 8e says, the bucket arrays
 * summed by lh_snapshot_allreduce (the library's peer-memory kernel).  Every rank's view of the snapshot must equal,
 * bucket for bucket, what ONE context produces from the whole stream.  Built and run by tests/test_gpu_comm.py.
 *
 *   c_comm_client <ranks> <n_total> <n_gpus>        prints "C_COMM_OK ..." and exits 0 on success */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "loghisto_b200.h"

#define CHECK(call) do { lh_status _s = (call); if (_s != LH_OK) { fprintf(stderr, "%s -> %s\n", #call, lh_strerror(_s)); exit(10); } } while (0)
#define SEED 0x10C415C0ull
#define NP 9
static const double PS[NP] = {0.0, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999, 0.9999, 1.0};

typedef struct {
    int rank, world, device;
    size_t start, n;
    lh_ctx *ctx;
    pthread_barrier_t *bar;
    lh_peer_handle *handles;          /* [world], shared */
    uint64_t *dense;                  /* [65536] this rank's view after the all-reduce */
    uint64_t count; int32_t pkeys[NP];
    float ar_ms;
} rank_t;

static lh_ctx *make_ctx(int device) {
    lh_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (uint32_t)sizeof cfg; cfg.device = device; cfg.max_histograms = 2; cfg.max_counters = 2;
    lh_ctx *ctx = NULL;
    lh_status st = lh_create(&cfg, &ctx);
    if (st != LH_OK) { fprintf(stderr, "lh_create(device %d): %s\n", device, lh_strerror(st)); exit(st == LH_ERR_NO_DEVICE ? 77 : 11); }
    return ctx;
}

/* cudaMalloc / cudaFree synchronise the whole DEVICE: when two ranks share one GPU (single-GPU box) they must not be
 * called while the other rank's all-reduce kernel may be waiting for this rank, so the buffer is allocated once up
 * front and freed after the last barrier. */
static void ingest_range(lh_ctx *ctx, double *d, size_t start, size_t n) {
    CHECK(lh_gen_stream_f64(ctx, 2 /* S: signed / edge mix */, SEED, start, n, d, NULL));
    CHECK(lh_ingest_f64(ctx, 1, d, n, NULL));
}

static void *rank_main(void *arg) {
    rank_t *r = (rank_t *)arg;
    r->ctx = make_ctx(r->device);
    CHECK(lh_comm_export(r->ctx, &r->handles[r->rank]));
    pthread_barrier_wait(r->bar);                         /* the "exchange": the handle array is shared memory here */
    CHECK(lh_comm_import(r->ctx, (uint32_t)r->rank, (uint32_t)r->world, r->handles));
    void *d = NULL;
    CHECK(lh_device_alloc(r->ctx, r->n * 8, &d));
    pthread_barrier_wait(r->bar);                         /* every rank has mapped every peer and holds its buffer */
    for (int interval = 0; interval < 2; interval++) {
        ingest_range(r->ctx, (double *)d, r->start, r->n);
        uint64_t seq = 0, counts[2]; double sums[2], avgs[2], pvals[2 * NP]; int32_t pkeys[2 * NP];
        CHECK(lh_snapshot_begin(r->ctx));
        CHECK(lh_snapshot_allreduce(r->ctx, 0, &seq));
        CHECK(lh_snapshot_reduce(r->ctx, PS, NP, counts, sums, avgs, pkeys, pvals));
        CHECK(lh_snapshot_copy_histogram(r->ctx, 1, r->dense));
        CHECK(lh_snapshot_end(r->ctx));
        CHECK(lh_comm_allreduce_ms(r->ctx, seq, &r->ar_ms));
        r->count = counts[1];
        memcpy(r->pkeys, pkeys + NP, sizeof r->pkeys);
    }
    lh_comm_stats cs;
    CHECK(lh_comm_info(r->ctx, &cs));
    if (cs.status != 0 || cs.world != (uint32_t)r->world || cs.allreduces != 2) { fprintf(stderr, "rank %d: comm status %u\n", r->rank, cs.status); exit(12); }
    pthread_barrier_wait(r->bar);
    CHECK(lh_device_free(r->ctx, d));
    lh_destroy(r->ctx);
    return NULL;
}

int main(int argc, char **argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 2;
    const size_t n_total = argc > 2 ? (size_t)atoll(argv[2]) : 4000001;
    const int n_gpus = argc > 3 ? atoi(argv[3]) : 1;
    if (world < 2 || world > LH_MAX_RANKS) return 2;
    /* single-context reference over the whole stream */
    uint64_t *want = calloc(65536, 8);
    uint64_t wcount[2]; int32_t wkeys[2 * NP];
    {
        lh_ctx *ctx = make_ctx(0);
        void *d = NULL;
        CHECK(lh_device_alloc(ctx, n_total * 8, &d));
        ingest_range(ctx, (double *)d, 0, n_total);
        CHECK(lh_snapshot_begin(ctx));
        CHECK(lh_snapshot_reduce(ctx, PS, NP, wcount, NULL, NULL, wkeys, NULL));
        CHECK(lh_snapshot_copy_histogram(ctx, 1, want));
        CHECK(lh_snapshot_end(ctx));
        CHECK(lh_device_free(ctx, d));
        lh_destroy(ctx);
    }
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)world);
    lh_peer_handle *handles = calloc((size_t)world, sizeof *handles);
    rank_t *ranks = calloc((size_t)world, sizeof *ranks);
    pthread_t *th = calloc((size_t)world, sizeof *th);
    const size_t base = n_total / (size_t)world, extra = n_total % (size_t)world;
    size_t start = 0;
    for (int r = 0; r < world; r++) {
        ranks[r].rank = r; ranks[r].world = world; ranks[r].device = r % n_gpus; ranks[r].bar = &bar; ranks[r].handles = handles;
        ranks[r].start = start; ranks[r].n = base + ((size_t)r < extra ? 1 : 0); start += ranks[r].n;
        ranks[r].dense = calloc(65536, 8);
        pthread_create(&th[r], NULL, rank_main, &ranks[r]);
    }
    int bad = 0;
    for (int r = 0; r < world; r++) {
        pthread_join(th[r], NULL);
        if (ranks[r].count != n_total || ranks[r].count != wcount[1]) { fprintf(stderr, "rank %d: count %llu\n", r, (unsigned long long)ranks[r].count); bad = 1; }
        if (memcmp(ranks[r].dense, want, 65536 * 8)) { fprintf(stderr, "rank %d: bucket arrays differ from the single-context run\n", r); bad = 1; }
        if (memcmp(ranks[r].pkeys, wkeys + NP, sizeof ranks[r].pkeys)) { fprintf(stderr, "rank %d: percentile keys differ\n", r); bad = 1; }
    }
    if (bad) return 1;
    printf("C_COMM_OK ranks=%d gpus=%d samples=%zu allreduce_ms=%.3f\n", world, n_gpus, n_total, ranks[0].ar_ms);
    return 0;
}
