/*
 * lh_stub.c -- TEST-ONLY stand-in for libloghisto_b200.so, backed by the CPU oracle.
 *
 * It exists so that the C++ host mirror (loghisto_b200/host/metric_system.cc: name interning, staging batches,
 * RawMetricSet / ProcessedMetricSet reconstruction, reaper, channels) can be exercised by the CPU test suite.  It is
 * built only by tests/test_host_logic_cpu.py into tests/_build/, implements only the entry points the host mirror
 * calls, and is never loaded by the product package (which has no CPU fallback).  Bucket arithmetic comes from
 * oracle/loghisto_oracle.c, compiled into the same shared object.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "loghisto_b200.h"

/* from oracle/loghisto_oracle.c */
int16_t lho_compress(double value);
uint64_t lho_process_histogram(const uint64_t *counts65536, const double *ps, int np, double *out_stats,
                               double *out_pvals, int32_t *out_pkeys);

#define STUB_SLOTS 64

/* one big lock: the real library is thread-safe, and the host mirror relies on that */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
#define LOCKED(expr) do { pthread_mutex_lock(&g_mu); lh_status _st = (expr); pthread_mutex_unlock(&g_mu); return _st; } while (0)

struct lh_ctx {
    lh_config cfg;
    uint64_t *buckets[2];   /* [H][65536] */
    uint64_t *counters[2];  /* [C] */
    int active, frozen;
    void *slot_mem[STUB_SLOTS];
    int slot_busy[STUB_SLOTS];
    uint32_t nslots;
    uint64_t slot_bytes;
    uint64_t dropped, samples, counter_ops, snapshots;
    /* export scratch */
    uint32_t *offsets; int16_t *keys; uint64_t *counts; uint64_t *deltas;
    char err[128];
};

LH_API uint32_t lh_abi_version(void) { return LH_ABI_VERSION; }
LH_API const char *lh_strerror(lh_status st) { return st == LH_OK ? "ok" : "stub error"; }
LH_API const char *lh_last_error(const lh_ctx *ctx) { return ctx ? ctx->err : ""; }

LH_API lh_status lh_create(const lh_config *cfg, lh_ctx **out) {
    if (!cfg || !out || cfg->struct_size != sizeof(lh_config) || !cfg->max_histograms || !cfg->max_counters) return LH_ERR_INVALID;
    lh_ctx *c = (lh_ctx *)calloc(1, sizeof(lh_ctx));
    c->cfg = *cfg;
    for (int b = 0; b < 2; b++) {
        c->buckets[b] = (uint64_t *)calloc((size_t)cfg->max_histograms * 65536u, 8);
        c->counters[b] = (uint64_t *)calloc(cfg->max_counters, 8);
    }
    c->nslots = cfg->staging_slots ? cfg->staging_slots : 3;
    if (c->nslots > STUB_SLOTS) c->nslots = STUB_SLOTS;
    c->slot_bytes = cfg->staging_bytes ? cfg->staging_bytes : (32u << 20);
    for (uint32_t i = 0; i < c->nslots; i++) c->slot_mem[i] = malloc(c->slot_bytes);
    c->offsets = (uint32_t *)calloc((size_t)cfg->max_histograms + 1, 4);
    c->keys = (int16_t *)malloc((size_t)cfg->max_histograms * 65536u * 2);
    c->counts = (uint64_t *)malloc((size_t)cfg->max_histograms * 65536u * 8);
    c->deltas = (uint64_t *)calloc(cfg->max_counters, 8);
    *out = c;
    return LH_OK;
}

LH_API lh_status lh_destroy(lh_ctx *c) {
    if (!c) return LH_OK;
    for (int b = 0; b < 2; b++) { free(c->buckets[b]); free(c->counters[b]); }
    for (uint32_t i = 0; i < c->nslots; i++) free(c->slot_mem[i]);
    free(c->offsets); free(c->keys); free(c->counts); free(c->deltas);
    free(c);
    return LH_OK;
}

static lh_status lh_staging_acquire_impl(lh_ctx *c, lh_staging *out) {
    for (uint32_t i = 0; i < c->nslots; i++)
        if (!c->slot_busy[i]) {
            c->slot_busy[i] = 1;
            out->host = c->slot_mem[i]; out->bytes = c->slot_bytes; out->slot = i; out->reserved = 0;
            return LH_OK;
        }
    snprintf(c->err, sizeof c->err, "every staging slot is acquired");
    return LH_ERR_STATE;
}

static lh_status lh_staging_commit_keyed_f64_u16_impl(lh_ctx *c, const lh_staging *s, size_t n, uint64_t ids_offset) {
    const double *v = (const double *)c->slot_mem[s->slot];
    const uint16_t *ids = (const uint16_t *)((const char *)c->slot_mem[s->slot] + ids_offset);
    uint64_t *b = c->buckets[c->active];
#ifndef LH_STUB_DISCARD                 /* -DLH_STUB_DISCARD: host-path profiling only (tools/host_path_profile.sh) */
    for (size_t i = 0; i < n; i++) {
        if (ids[i] >= c->cfg.max_histograms) { c->dropped++; continue; }
        b[(size_t)ids[i] * 65536u + (uint16_t)lho_compress(v[i])]++;
    }
#else
    (void)v; (void)ids; (void)b;
#endif
    c->samples += n;
    c->slot_busy[s->slot] = 0;
    return LH_OK;
}

static lh_status lh_staging_commit_counter_u16_impl(lh_ctx *c, const lh_staging *s, size_t n, uint64_t ids_offset) {
    const uint64_t *a = (const uint64_t *)c->slot_mem[s->slot];
    const uint16_t *ids = (const uint16_t *)((const char *)c->slot_mem[s->slot] + ids_offset);
    for (size_t i = 0; i < n; i++) {
        if (ids[i] >= c->cfg.max_counters) { c->dropped++; continue; }
        c->counters[c->active][ids[i]] += a[i];
    }
    c->counter_ops += n;
    c->slot_busy[s->slot] = 0;
    return LH_OK;
}

static lh_status lh_snapshot_begin_impl(lh_ctx *c) {
    if (c->frozen) { snprintf(c->err, sizeof c->err, "previous snapshot not ended"); return LH_ERR_STATE; }
    c->active ^= 1;
    c->frozen = 1;
    c->snapshots++;
    return LH_OK;
}

static lh_status lh_snapshot_reduce_impl(lh_ctx *c, const double *ps, uint32_t np, uint64_t *counts, double *sums, double *avgs,
                                         int32_t *pkeys, double *pvals) {
    if (!c->frozen) return LH_ERR_STATE;
    const uint64_t *fb = c->buckets[c->active ^ 1];
    for (uint32_t h = 0; h < c->cfg.max_histograms; h++) {
        double stats[3], pv[LH_MAX_PERCENTILES];
        int32_t pk[LH_MAX_PERCENTILES];
        uint64_t total = lho_process_histogram(fb + (size_t)h * 65536u, ps, (int)np, stats, pv, pk);
        if (counts) counts[h] = total;
        if (sums) sums[h] = stats[1];
        if (avgs) avgs[h] = stats[2];
        for (uint32_t j = 0; j < np; j++) {
            if (pkeys) pkeys[(size_t)h * np + j] = pk[j];
            if (pvals) pvals[(size_t)h * np + j] = pv[j];
        }
    }
    return LH_OK;
}

static lh_status lh_snapshot_export_impl(lh_ctx *c, lh_sparse *out) {
    if (!c->frozen) return LH_ERR_STATE;
    const uint64_t *fb = c->buckets[c->active ^ 1];
    uint32_t pos = 0;
    for (uint32_t h = 0; h < c->cfg.max_histograms; h++) {
        c->offsets[h] = pos;
        for (int key = -32768; key <= 32767; key++) {
            uint64_t v = fb[(size_t)h * 65536u + (uint16_t)(int16_t)key];
            if (v) { c->keys[pos] = (int16_t)key; c->counts[pos] = v; pos++; }
        }
    }
    c->offsets[c->cfg.max_histograms] = pos;
    memcpy(c->deltas, c->counters[c->active ^ 1], (size_t)c->cfg.max_counters * 8);
    out->offsets = c->offsets; out->keys = c->keys; out->counts = c->counts; out->counter_deltas = c->deltas;
    out->total_entries = pos;
    return LH_OK;
}

static lh_status lh_snapshot_end_impl(lh_ctx *c) {
    if (!c->frozen) return LH_ERR_STATE;
    memset(c->buckets[c->active ^ 1], 0, (size_t)c->cfg.max_histograms * 65536u * 8);
    memset(c->counters[c->active ^ 1], 0, (size_t)c->cfg.max_counters * 8);
    c->frozen = 0;
    return LH_OK;
}

static lh_status lh_get_stats_impl(lh_ctx *c, lh_stats *out) {
    memset(out, 0, sizeof *out);
    out->samples = c->samples; out->counter_ops = c->counter_ops; out->dropped = c->dropped; out->snapshots = c->snapshots;
    return LH_OK;
}

static lh_status lh_staging_abandon_impl(lh_ctx *c, const lh_staging *s) {
    if (!s || s->slot >= c->nslots) return LH_ERR_INVALID;
    c->slot_busy[s->slot] = 0;
    return LH_OK;
}
LH_API lh_status lh_staging_abandon(lh_ctx *c, const lh_staging *s) { LOCKED(lh_staging_abandon_impl(c, s)); }
LH_API lh_status lh_staging_acquire(lh_ctx *c, lh_staging *out) { LOCKED(lh_staging_acquire_impl(c, out)); }

LH_API lh_status lh_staging_commit_keyed_f64_u16(lh_ctx *c, const lh_staging *s, size_t n, uint64_t ids_offset) { LOCKED(lh_staging_commit_keyed_f64_u16_impl(c, s, n, ids_offset)); }

LH_API lh_status lh_staging_commit_counter_u16(lh_ctx *c, const lh_staging *s, size_t n, uint64_t ids_offset) { LOCKED(lh_staging_commit_counter_u16_impl(c, s, n, ids_offset)); }

LH_API lh_status lh_snapshot_begin(lh_ctx *c) { LOCKED(lh_snapshot_begin_impl(c)); }

LH_API lh_status lh_snapshot_end(lh_ctx *c) { LOCKED(lh_snapshot_end_impl(c)); }

LH_API lh_status lh_snapshot_export(lh_ctx *c, lh_sparse *out) { LOCKED(lh_snapshot_export_impl(c, out)); }

LH_API lh_status lh_get_stats(lh_ctx *c, lh_stats *out) { LOCKED(lh_get_stats_impl(c, out)); }

LH_API lh_status lh_snapshot_reduce(lh_ctx *c, const double *ps, uint32_t np, uint64_t *counts, double *sums, double *avgs,
                                    int32_t *pkeys, double *pvals) {
    LOCKED(lh_snapshot_reduce_impl(c, ps, np, counts, sums, avgs, pkeys, pvals));
}
