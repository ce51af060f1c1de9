// lh_kernels.cuh -- sm_100a kernels of the loghisto hot path.
//
//   K1   k_ingest_single_*  one histogram, float64 stream -> bucket counts
//                           (compress + Histogram increment, metrics.go:273-295, 316-322)
//          _bulk  : cp.async.bulk (TMA 1-D, UBLKCP) into a shared-memory ring guarded by mbarriers, one producer
//                   warp + N consumer warps; with V2 = packed-FP32 bucket arithmetic this is the shipped default
//          _ldg   : 256-bit ld.global.nc loads, software-pipelined in registers (first version)
//          _v2/_v3: register-only variants of the lean arithmetic (kept for the comparison in profiles/)
//        all privatise the histogram in shared memory (uint32 sub-histograms, ATOMS.POPC.INC) and flush once per
//        CTA with one global 64-bit atomic per non-empty bucket.
//   K1k  k_ingest_keyed_small   (id,value) pairs, <= 11 ids per pass privatised in shared memory (HBM-bound)
//        k_ingest_keyed_vec     any number of ids: one L2 RED per sample into a compact, replicated uint32 window
//        k_ingest_keyed_part    owner-partitioned cooperative kernel (opt-in experiment)
//        k_ingest_keyed         scalar fallback for ragged / misaligned pieces;  k_fold_hot drains the window
//   K2   k_counter_add{_smem}   (id,amount) pairs -> counters[id]        (metrics.go:251-269)
//   K3   k_reduce               per histogram: count, sum, avg, percentiles (processHistograms + percentile,
//                               metrics.go:336-418)
//   K4   k_scan_nnz, k_export   sparse (key,count) lists (RawMetricSet.Histograms);  k_merge_sparse is the inverse
//   misc k_fill_decompress, k_compress_probe, k_fastpath_margin, k_stream_probe, k_gen_stream, k_gen_ids_u16
#pragma once
#include "lh_device.cuh"

namespace lh {

// ---------------------------------------------------------------- helpers
// Streaming loads: read-once data, keep it out of L1 and first in line for L2 eviction.
// sm_100 has 256-bit global loads (ld.global.v4.b64 -> LDG.E.256); the L2
// eviction-priority qualifier is only accepted on those.
struct f64x4 { double a, b, c, d; };
__device__ __forceinline__ f64x4 ldg_stream_f64x4(const void *p) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    f64x4 r;
    r.a = __longlong_as_double((long long)a); r.b = __longlong_as_double((long long)b);
    r.c = __longlong_as_double((long long)c); r.d = __longlong_as_double((long long)d);
    return r;
}
__device__ __forceinline__ f64x4 ldg_stream_f64x2x2(const void *p) {   // two 128-bit loads (comparison variant)
    f64x4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.a), "=d"(r.b) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.c), "=d"(r.d) : "l"((const char *)p + 16));
    return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LH_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LH_DONE_%=;\n\t"
        "bra LH_WAIT_%=;\n\t"
        "LH_DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// One sample -> shared sub-histogram slot.  Samples outside the window are
// counted straight into the global array and redirected to a trash slot so
// the shared atomic below stays unconditional.
constexpr uint32_t LH_TRASH = LH_SUBHIST;          // slot never flushed
constexpr uint32_t LH_SUBHIST_ALLOC = LH_SUBHIST + 8;

template <int NS>
__device__ __forceinline__ void bucket_samples(const double (&v)[NS], uint32_t *hist,
                                               unsigned long long *__restrict__ counts) {
    uint32_t idx[NS];
    bool slow[NS];
    bool any = false;
#pragma unroll
    for (int i = 0; i < NS; i++) { fast_candidate(v[i], idx[i], slow[i]); any |= slow[i]; }
    if (__any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
        for (int i = 0; i < NS; i++) {
            if (slow[i]) {
                uint32_t key = exact_key16(v[i]);
                uint32_t slot = key16_to_slot(key);
                if (slot == 0xFFFFFFFFu) { atomicAdd(&counts[key], 1ull); slot = LH_TRASH; }
                idx[i] = slot;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NS; i++) atomicAdd(&hist[idx[i]], 1u);
}

__device__ __forceinline__ void flush_subhist(const uint32_t *hist, int copies, int tid, int nthreads,
                                              unsigned long long *__restrict__ counts) {
    for (int slot = tid; slot < LH_SUBHIST; slot += nthreads) {
        uint32_t c = 0;
        for (int k = 0; k < copies; k++) c += hist[k * LH_SUBHIST_ALLOC + slot];
        if (c) atomicAdd(&counts[slot_to_key16((uint32_t)slot)], (unsigned long long)c);
    }
}

// ------------------------------------------------------------------- K1/ldg
// vals32: 32-byte aligned, nvec 32-byte vectors (4 samples each).  Up to three
// scalar stragglers on either side (misaligned head, ragged tail) come separately.
__device__ __forceinline__ void bucket_stragglers(const double *p, int n, unsigned long long *__restrict__ counts) {
    for (int i = 0; i < n; i++) atomicAdd(&counts[key16_of(p[i])], 1ull);
}

template <int THREADS, int UNROLL, int COPIES, int MINB, bool WIDE>
__global__ void __launch_bounds__(THREADS, MINB)
k_ingest_single_ldg(const double *__restrict__ vals32, size_t nvec, const double *head, int nhead,
                    const double *tail, int ntail, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) uint32_t s_hist[];
    for (int i = threadIdx.x; i < COPIES * (int)LH_SUBHIST_ALLOC; i += THREADS) s_hist[i] = 0;
    __syncthreads();
    uint32_t *my = s_hist + ((threadIdx.x >> 5) % COPIES) * LH_SUBHIST_ALLOC;
    const char *base = reinterpret_cast<const char *>(vals32);

    constexpr size_t TILE = (size_t)THREADS * UNROLL;   // 32-byte vectors per tile
    const size_t ntiles = nvec / TILE;
    f64x4 cur[UNROLL], nxt[UNROLL];
    size_t tile = blockIdx.x;
    if (tile < ntiles) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const char *p = base + (tile * TILE + (size_t)u * THREADS + threadIdx.x) * 32;
            cur[u] = WIDE ? ldg_stream_f64x4(p) : ldg_stream_f64x2x2(p);
        }
    }
    while (tile < ntiles) {
        size_t nt = tile + gridDim.x;
        if (nt < ntiles) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const char *p = base + (nt * TILE + (size_t)u * THREADS + threadIdx.x) * 32;
                nxt[u] = WIDE ? ldg_stream_f64x4(p) : ldg_stream_f64x2x2(p);
            }
        }
        double v[4 * UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { v[4 * u] = cur[u].a; v[4 * u + 1] = cur[u].b; v[4 * u + 2] = cur[u].c; v[4 * u + 3] = cur[u].d; }
        bucket_samples<4 * UNROLL>(v, my, counts);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) cur[u] = nxt[u];
        tile = nt;
    }
    // partial last tile + stragglers: one CTA, bounds-checked (rare, < TILE vectors)
    if (blockIdx.x == ntiles % gridDim.x) {
        for (size_t j = ntiles * TILE * 4 + threadIdx.x; j < nvec * 4; j += THREADS) {
            uint32_t k0 = key16_of(vals32[j]);
            uint32_t s0 = key16_to_slot(k0);
            if (s0 == 0xFFFFFFFFu) atomicAdd(&counts[k0], 1ull); else atomicAdd(&my[s0], 1u);
        }
        if (threadIdx.x == 0) { bucket_stragglers(head, nhead, counts); bucket_stragglers(tail, ntail, counts); }
    }
    __syncthreads();
    flush_subhist(s_hist, COPIES, threadIdx.x, THREADS, counts);
}

// ---------------------------------------------------------------- K1/ldg v2
// Same algorithm as k_ingest_single_ldg with the per-sample instruction count cut from ~30 to ~20 so that the
// kernel stays HBM-bound under sustained load (profiles/r01/sustained_probe.txt):
//   * samples are processed in pairs with Blackwell's packed FP32 ops (fma.rn.f32x2 / add.f32x2);
//   * (float)e comes from one I2FP and the -1023 bias rides in the FMA addend;
//   * ONE flag per sample -- estimate too close to a bucket boundary, OR v's high word >= 0x43E00000 unsigned
//     (|v| >= 2^63, Inf, NaN, and every negative value) -- sends it to a fix-up that re-derives the slot
//     (negative values cost a second fast evaluation there, not the FP64 path);
//   * the shared-memory byte offset is built as e*276 + (rounded bits << 2) + const: one IMAD and one LEA;
//   * two register buffers alternate, so no copies.
// Extra estimate error vs fast_candidate(): the float constant -1023*c2 (|err| <= 1.6e-5 bucket units), still
// well inside LH_FAST_EPS.
__device__ __forceinline__ uint32_t fixup_slot(double v, unsigned long long *__restrict__ counts) {
    uint32_t key = key16_of(v);
    uint32_t slot = key16_to_slot(key);
    if (slot == 0xFFFFFFFFu) { atomicAdd(&counts[key], 1ull); slot = LH_TRASH; }
    return slot;
}

template <int NS>
__device__ __forceinline__ void bucket_samples_v2(const double (&v)[NS], uint32_t *hist, uint32_t one_bits,
                                                  unsigned long long *__restrict__ counts) {
    static_assert(NS % 2 == 0, "pairs");
    constexpr float C1 = 69.31471805599453f, C2 = 0.31471805599453f;
    constexpr float KB = (float)(-1023.0 * (double)C2);
    constexpr float MAGIC = 12582912.0f;
    // byte offset = 4*(69*(eb-1023) + (bits(r) - 0x4B400000))  (mod 2^32)
    constexpr uint32_t COFF = 0u - (1023u * 69u * 4u) - (0x4B400000u << 2);
    uint32_t off[NS];
    bool flag[NS];
    bool any = false;
#pragma unroll
    for (int i = 0; i < NS; i += 2) {
        const double x0 = __dadd_rn(1.0, fabs(v[i])), x1 = __dadd_rn(1.0, fabs(v[i + 1]));
        const uint32_t h0 = (uint32_t)__double2hiint(x0), h1 = (uint32_t)__double2hiint(x1);
        const uint32_t t0 = __funnelshift_l((uint32_t)__double2loint(x0), h0, 3);
        const uint32_t t1 = __funnelshift_l((uint32_t)__double2loint(x1), h1, 3);
        uint32_t m0, m1;
        asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(m0) : "r"(t0), "r"(one_bits));   // (t & mask) | 1.0f
        asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(m1) : "r"(t1), "r"(one_bits));
        float2 lg;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg.x) : "f"(__uint_as_float(m0)));
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg.y) : "f"(__uint_as_float(m1)));
        const uint32_t e0 = h0 >> 20, e1 = h1 >> 20;                    // 1023 + e
        const float2 ef = make_float2(__uint2float_rn(e0), __uint2float_rn(e1));
        const float2 a = __ffma2_rn(ef, make_float2(C2, C2), make_float2(KB, KB));
        const float2 w = __ffma2_rn(lg, make_float2(C1, C1), a);
        const float2 r = __fadd2_rn(w, make_float2(MAGIC, MAGIC));
        const float2 s = __fadd2_rn(r, make_float2(-MAGIC, -MAGIC));
        const float2 d = __ffma2_rn(s, make_float2(-1.0f, -1.0f), w);   // w - s, one rounding
        flag[i] = (fabsf(d.x) > 0.5f - LH_FAST_EPS) | ((uint32_t)__double2hiint(v[i]) >= 0x43E00000u);
        flag[i + 1] = (fabsf(d.y) > 0.5f - LH_FAST_EPS) | ((uint32_t)__double2hiint(v[i + 1]) >= 0x43E00000u);
        off[i] = e0 * 276u + (__float_as_uint(r.x) << 2) + COFF;
        off[i + 1] = e1 * 276u + (__float_as_uint(r.y) << 2) + COFF;
        any |= flag[i] | flag[i + 1];
    }
    if (__any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
        for (int i = 0; i < NS; i++)
            if (flag[i]) off[i] = fixup_slot(v[i], counts) * 4u;
    }
#pragma unroll
    for (int i = 0; i < NS; i++) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + off[i]), 1u);
}

template <int THREADS, int UNROLL, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_ingest_single_v2(const double *__restrict__ vals32, size_t nvec, const double *head, int nhead,
                   const double *tail, int ntail, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) uint32_t s_hist[];
    for (int i = threadIdx.x; i < (int)LH_SUBHIST_ALLOC; i += THREADS) s_hist[i] = 0;
    __syncthreads();
    uint32_t one_bits;
    asm volatile("mov.b32 %0, 0x3F800000;" : "=r"(one_bits));   // opaque to constant folding: keeps LOP3 at one instruction
    const char *base = reinterpret_cast<const char *>(vals32);
    constexpr size_t TILE = (size_t)THREADS * UNROLL;           // 32-byte vectors per tile
    const size_t ntiles = nvec / TILE;
    const size_t step = gridDim.x;

    f64x4 A[UNROLL], B[UNROLL];
    auto load = [&](f64x4(&buf)[UNROLL], size_t tile) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) buf[u] = ldg_stream_f64x4(base + (tile * TILE + (size_t)u * THREADS + threadIdx.x) * 32);
    };
    auto process = [&](const f64x4(&buf)[UNROLL]) {
        double v[4 * UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { v[4 * u] = buf[u].a; v[4 * u + 1] = buf[u].b; v[4 * u + 2] = buf[u].c; v[4 * u + 3] = buf[u].d; }
        bucket_samples_v2<4 * UNROLL>(v, s_hist, one_bits, counts);
    };
    size_t tile = blockIdx.x;
    if (tile < ntiles) load(A, tile);
    while (tile < ntiles) {
        const size_t t1 = tile + step;
        if (t1 < ntiles) load(B, t1);
        process(A);
        if (t1 >= ntiles) break;
        const size_t t2 = t1 + step;
        if (t2 < ntiles) load(A, t2);
        process(B);
        tile = t2;
    }
    if (blockIdx.x == ntiles % gridDim.x) {   // partial last tile + stragglers
        for (size_t j = ntiles * TILE * 4 + threadIdx.x; j < nvec * 4; j += THREADS) {
            uint32_t slot = fixup_slot(vals32[j], counts);
            atomicAdd(&s_hist[slot], 1u);
        }
        if (threadIdx.x == 0) { bucket_stragglers(head, nhead, counts); bucket_stragglers(tail, ntail, counts); }
    }
    __syncthreads();
    flush_subhist(s_hist, 1, threadIdx.x, THREADS, counts);
}

// ---------------------------------------------------------------- K1/ldg v3
// v2's arithmetic with a three-deep register rotation: every thread always has DEPTH-1 256-bit loads in flight
// (64 KB per SM at 1024 threads) while it buckets the 4 samples of the oldest one, inside 64 registers.
template <int THREADS, int MINB, int DEPTH>
__global__ void __launch_bounds__(THREADS, MINB)
k_ingest_single_v3(const double *__restrict__ vals32, size_t nvec, const double *head, int nhead,
                   const double *tail, int ntail, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) uint32_t s_hist[];
    for (int i = threadIdx.x; i < (int)LH_SUBHIST_ALLOC; i += THREADS) s_hist[i] = 0;
    __syncthreads();
    uint32_t one_bits;
    asm volatile("mov.b32 %0, 0x3F800000;" : "=r"(one_bits));
    const char *base = reinterpret_cast<const char *>(vals32) + (size_t)threadIdx.x * 32;
    constexpr size_t TILE_BYTES = (size_t)THREADS * 32;
    const size_t ntiles = nvec / THREADS;
    const size_t step = gridDim.x;

    f64x4 buf[DEPTH];
    size_t tile = blockIdx.x;          // tile whose data sits in buf[0]
#pragma unroll
    for (int d = 0; d < DEPTH - 1; d++)
        if (tile + d * step < ntiles) buf[d] = ldg_stream_f64x4(base + (tile + d * step) * TILE_BYTES);
    while (tile < ntiles) {
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {          // rotation unrolled: buffer indices are compile-time
            const size_t cur = tile + (size_t)d * step;
            if (cur >= ntiles) break;
            const size_t pre = cur + (size_t)(DEPTH - 1) * step;
            if (pre < ntiles) buf[(d + DEPTH - 1) % DEPTH] = ldg_stream_f64x4(base + pre * TILE_BYTES);
            const double v[4] = {buf[d].a, buf[d].b, buf[d].c, buf[d].d};
            bucket_samples_v2<4>(v, s_hist, one_bits, counts);
        }
        tile += (size_t)DEPTH * step;
    }
    if (blockIdx.x == ntiles % gridDim.x) {   // partial last tile + stragglers
        for (size_t j = ntiles * THREADS * 4 + threadIdx.x; j < nvec * 4; j += THREADS) {
            uint32_t slot = fixup_slot(vals32[j], counts);
            atomicAdd(&s_hist[slot], 1u);
        }
        if (threadIdx.x == 0) { bucket_stragglers(head, nhead, counts); bucket_stragglers(tail, ntail, counts); }
    }
    __syncthreads();
    flush_subhist(s_hist, 1, threadIdx.x, THREADS, counts);
}

// --------------------------------------------------------------- read probe
// Diagnostic only (lh_tune "k1" = last variant): the same 256-bit streaming loads as K1/ldg with the
// bucket arithmetic replaced by an XOR fold, to separate memory-side from SM-side limits.  Counts are NOT
// produced; one word per CTA is written so the loads cannot be elided.
template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS, 2)
k_stream_probe(const double *__restrict__ vals32, size_t nvec, const double *, int, const double *, int,
               unsigned long long *__restrict__ counts) {
    const char *base = reinterpret_cast<const char *>(vals32);
    constexpr size_t TILE = (size_t)THREADS * UNROLL;
    const size_t ntiles = nvec / TILE;
    unsigned long long acc = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        f64x4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) v[u] = ldg_stream_f64x4(base + (tile * TILE + (size_t)u * THREADS + threadIdx.x) * 32);
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            acc ^= f64_as_u64(v[u].a) ^ f64_as_u64(v[u].b) ^ f64_as_u64(v[u].c) ^ f64_as_u64(v[u].d);
    }
    if (acc == 0x123456789ABCDEFull) counts[65535] = acc;   // never true in practice; keeps the loads alive
}

// ------------------------------------------------------------------ K1/bulk
// Producer warp streams STAGE_BYTES tiles into a STAGES-deep shared ring with
// cp.async.bulk; CW consumer warps bucket them.  Tiles are dealt round-robin.
template <int NS>
__device__ __forceinline__ void bucket_samples_v2(const double (&v)[NS], uint32_t *hist, uint32_t one_bits,
                                                  unsigned long long *__restrict__ counts);

template <int CW, int STAGES, int STAGE_BYTES, int COPIES, int MINB, bool V2 = false>
__global__ void __launch_bounds__((CW + 1) * 32, MINB)
k_ingest_single_bulk(const double *__restrict__ vals32, size_t nvec32, const double *head, int nhead,
                     const double *tail, int ntail, unsigned long long *__restrict__ counts) {
    const double2 *vals16 = reinterpret_cast<const double2 *>(vals32);
    const size_t nvec = nvec32 * 2;   // 16-byte vectors
    constexpr int CT = CW * 32;                       // consumer threads
    constexpr int STAGE_VEC = STAGE_BYTES / 16;
    constexpr int PER_THREAD = STAGE_VEC / CT;        // double2 per consumer thread per stage
    static_assert(STAGE_VEC % CT == 0, "stage must divide evenly over consumer threads");
    extern __shared__ __align__(128) unsigned char s_raw[];
    double2 *s_data = reinterpret_cast<double2 *>(s_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(s_raw + (size_t)STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(empty + STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    for (int i = tid; i < COPIES * (int)LH_SUBHIST_ALLOC; i += (CW + 1) * 32) s_hist[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const size_t ntiles = (nvec + STAGE_VEC - 1) / STAGE_VEC;
    if (warp == CW) {
        // ===== producer =====
        if ((tid & 31) == 0) {
            const uint64_t pol = policy_evict_first();
            uint32_t it = 0;
            for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                size_t first = tile * (size_t)STAGE_VEC;
                size_t rem = nvec - first;
                uint32_t bytes = (uint32_t)((rem < (size_t)STAGE_VEC ? rem : (size_t)STAGE_VEC) * 16);
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(s_data + (size_t)s * STAGE_VEC, vals16 + first, bytes, &full[s], pol);
            }
        }
    } else {
        // ===== consumers =====
        uint32_t *my = s_hist + (warp % COPIES) * LH_SUBHIST_ALLOC;
        uint32_t it = 0;
        for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            const double2 *st = s_data + (size_t)s * STAGE_VEC;
            size_t first = tile * (size_t)STAGE_VEC;
            size_t rem = nvec - first;
            if (rem >= (size_t)STAGE_VEC) {
                double v[2 * PER_THREAD];
#pragma unroll
                for (int u = 0; u < PER_THREAD; u++) {
                    double2 d = st[u * CT + tid];
                    v[2 * u] = d.x; v[2 * u + 1] = d.y;
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty[s]);   // registers hold the data: release early
                if constexpr (V2) {
                    uint32_t one_bits;
                    asm volatile("mov.b32 %0, 0x3F800000;" : "=r"(one_bits));
#pragma unroll
                    for (int q = 0; q < 2 * PER_THREAD; q += 4) {
                        const double v4[4] = {v[q], v[q + 1], v[q + 2], v[q + 3]};
                        bucket_samples_v2<4>(v4, my, one_bits, counts);
                    }
                } else {
                    bucket_samples<2 * PER_THREAD>(v, my, counts);
                }
            } else {
                for (int j = tid; j < (int)rem; j += CT) {
                    double2 d = st[j];
                    uint32_t k0 = key16_of(d.x), k1 = key16_of(d.y);
                    uint32_t s0 = key16_to_slot(k0), s1 = key16_to_slot(k1);
                    if (s0 == 0xFFFFFFFFu) atomicAdd(&counts[k0], 1ull); else atomicAdd(&my[s0], 1u);
                    if (s1 == 0xFFFFFFFFu) atomicAdd(&counts[k1], 1ull); else atomicAdd(&my[s1], 1u);
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty[s]);
            }
        }
        if (blockIdx.x == 0 && tid == 0) { bucket_stragglers(head, nhead, counts); bucket_stragglers(tail, ntail, counts); }
    }
    __syncthreads();
    flush_subhist(s_hist, COPIES, tid, (CW + 1) * 32, counts);
}

// --------------------------------------------------------------------- K1k
// (id,value) pairs.  1024 histograms x ~4.4K live buckets cannot be privatised
// in shared memory, so the cells live in L2: a compact uint32 "hot window"
// [H][LH_SUBHIST] (36 MB at H = 1024, vs 512 MB for the dense uint64 arrays)
// updated with no-return atomics that carry an L2 evict_last policy, while the
// sample stream is read once with 256-bit evict_first loads so it does not push
// the cells out of the 126 MB L2.  Keys outside the window go straight to the
// uint64 array.  k_fold_hot drains the window into the uint64 buckets at every
// snapshot (and before any cell could reach 2^32).
template <typename T> __device__ __forceinline__ double sample_to_f64(T v);
template <> __device__ __forceinline__ double sample_to_f64<double>(double v) { return v; }
// float64(duration.Nanoseconds()): CVTSQ2SD, round-to-nearest-even
template <> __device__ __forceinline__ double sample_to_f64<long long>(long long v) { return __ll2double_rn(v); }

__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void red_add_u32_keep(unsigned int *addr, unsigned int v, uint64_t policy) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(addr), "r"(v), "l"(policy) : "memory");
}

template <typename ValT>
__device__ __forceinline__ void keyed_one(uint32_t id, ValT raw, uint32_t H, unsigned int *__restrict__ hot,
                                          unsigned long long *__restrict__ buckets,
                                          unsigned long long *__restrict__ dropped, uint64_t pol) {
    if (id >= H) { atomicAdd(dropped, 1ull); return; }
    double v = sample_to_f64<ValT>(raw);
    uint32_t idx; bool slow;
    fast_candidate(v, idx, slow);
    if (slow) {
        uint32_t key = exact_key16(v);
        idx = key16_to_slot(key);
        if (idx == 0xFFFFFFFFu) { atomicAdd(&buckets[(size_t)id * 65536u + key], 1ull); return; }
    }
    red_add_u32_keep(&hot[(size_t)id * LH_SUBHIST + idx], 1u, pol);
}

// Vector body: every thread takes 4 consecutive pairs (one 256-bit value load, one 64/128-bit id load).
// vals must be 32-byte aligned and ids 4*sizeof(IdT)-aligned; n4 = number of 4-sample groups.
// `hot` holds `replicas` copies of the window ([replicas][H][LH_SUBHIST]); CTA b updates copy b % replicas, which
// divides the same-address pressure on hot cells (clustered, latency-like data) by the replica count while every
// copy stays L2-resident.  k_fold_hot sums the copies.
template <typename IdT, typename ValT, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ingest_keyed_vec(const IdT *__restrict__ ids, const ValT *__restrict__ vals, size_t n4, uint32_t H,
                   unsigned int *__restrict__ hot, uint32_t replicas, unsigned long long *__restrict__ buckets,
                   unsigned long long *__restrict__ dropped) {
    hot += (size_t)(blockIdx.x % replicas) * H * LH_SUBHIST;
    const uint64_t pol = policy_evict_last();
    const size_t stride = (size_t)gridDim.x * THREADS;
    for (size_t g = (size_t)blockIdx.x * THREADS + threadIdx.x; g < n4; g += stride) {
        unsigned long long a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                     : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(reinterpret_cast<const char *>(vals) + g * 32));
        uint32_t i0, i1, i2, i3;
        if (sizeof(IdT) == 2) {
            unsigned int lo, hi;
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi)
                         : "l"(reinterpret_cast<const char *>(ids) + g * 8));
            i0 = lo & 0xFFFFu; i1 = lo >> 16; i2 = hi & 0xFFFFu; i3 = hi >> 16;
        } else {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(i0), "=r"(i1), "=r"(i2), "=r"(i3)
                         : "l"(reinterpret_cast<const char *>(ids) + g * 16));
        }
        ValT r0, r1, r2, r3;
        memcpy(&r0, &a, 8); memcpy(&r1, &b, 8); memcpy(&r2, &c, 8); memcpy(&r3, &d, 8);
        keyed_one<ValT>(i0, r0, H, hot, buckets, dropped, pol);
        keyed_one<ValT>(i1, r1, H, hot, buckets, dropped, pol);
        keyed_one<ValT>(i2, r2, H, hot, buckets, dropped, pol);
        keyed_one<ValT>(i3, r3, H, hot, buckets, dropped, pol);
    }
}

// Scalar version for ragged heads/tails and misaligned inputs.
template <typename IdT, typename ValT, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ingest_keyed(const IdT *__restrict__ ids, const ValT *__restrict__ vals, size_t n, uint32_t H,
               unsigned int *__restrict__ hot, unsigned long long *__restrict__ buckets,
               unsigned long long *__restrict__ dropped) {
    const uint64_t pol = policy_evict_last();
    const size_t stride = (size_t)gridDim.x * THREADS;
    for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride)
        keyed_one<ValT>((uint32_t)ids[i], vals[i], H, hot, buckets, dropped, pol);
}

// ------------------------------------------------------------------ K1k/small
// Keyed ingest when only a few histograms are configured (H <= KS_MAX_H): all their positive windows
// (uint32[H][4368]) are privatised per CTA in shared memory, exactly like K1, so the kernel is HBM-bound
// (10 B/sample) instead of L2-atomic-bound.  Same packed-FP32 bucket arithmetic as bucket_samples_v2; the ONE
// flag per sample also covers id >= H, and flagged samples (boundary-close estimates, negatives, |v| >= 2^63,
// NaN/Inf, bad ids) take the L2 route of keyed_one().  Windows are added into the uint32 hot window at the end.
constexpr int KS_MAX_H = 11;                 // 11 * 4368 * 4 B = 192 KB of shared memory
constexpr int KS_THREADS = 1024;

template <typename IdT, typename ValT>
__global__ void __launch_bounds__(KS_THREADS, 1)
k_ingest_keyed_small(const IdT *__restrict__ ids, const ValT *__restrict__ vals, size_t n4, uint32_t H,
                     uint32_t id_lo, uint32_t id_cnt, unsigned int *__restrict__ hot,
                     unsigned long long *__restrict__ buckets, unsigned long long *__restrict__ dropped) {
    // This launch owns ids [id_lo, id_lo + id_cnt) (id_cnt <= KS_MAX_H); with more histograms than fit, the host
    // runs one pass per id sub-range over the same batch.  Samples of other valid ids are skipped; ids >= H are
    // dropped (and counted) by the pass that starts at id 0.
    extern __shared__ __align__(16) uint32_t ks_hist[];          // [id_cnt][LH_WIN] + trash word
    const uint32_t words = id_cnt * (uint32_t)LH_WIN;
    for (uint32_t i = threadIdx.x; i <= words; i += KS_THREADS) ks_hist[i] = 0;
    __syncthreads();
    const uint64_t pol = policy_evict_last();
    uint32_t one_bits;
    asm volatile("mov.b32 %0, 0x3F800000;" : "=r"(one_bits));
    constexpr float C1 = 69.31471805599453f, C2 = 0.31471805599453f;
    constexpr float KB = (float)(-1023.0 * (double)C2);
    constexpr float MAGIC = 12582912.0f;
    constexpr uint32_t COFF = 0u - (1023u * 69u * 4u) - (0x4B400000u << 2);
    const uint32_t trash_off = words * 4u;

    // The loop bound is warp-uniform (base index of the CTA's row of groups); lanes past the end are predicated
    // off, because the fix-up below votes with the full warp mask.
    const size_t stride = (size_t)gridDim.x * KS_THREADS;
    size_t base = (size_t)blockIdx.x * KS_THREADS;
    unsigned long long cur[4] = {0, 0, 0, 0}, nxt[4] = {0, 0, 0, 0};
    uint32_t cur_id[4] = {0, 0, 0, 0}, nxt_id[4] = {0, 0, 0, 0};
    auto load = [&](unsigned long long(&raw)[4], uint32_t(&id4)[4], size_t gi) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                     : "=l"(raw[0]), "=l"(raw[1]), "=l"(raw[2]), "=l"(raw[3]) : "l"(reinterpret_cast<const char *>(vals) + gi * 32));
        if (sizeof(IdT) == 2) {
            unsigned int lo, hi;
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi)
                         : "l"(reinterpret_cast<const char *>(ids) + gi * 8));
            id4[0] = lo & 0xFFFFu; id4[1] = lo >> 16; id4[2] = hi & 0xFFFFu; id4[3] = hi >> 16;
        } else {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(id4[0]), "=r"(id4[1]), "=r"(id4[2]), "=r"(id4[3])
                         : "l"(reinterpret_cast<const char *>(ids) + gi * 16));
        }
    };
    if (base + threadIdx.x < n4) load(cur, cur_id, base + threadIdx.x);
    for (; base < n4; base += stride) {
        const bool valid = base + threadIdx.x < n4;
        const size_t gn = base + stride + threadIdx.x;
        if (gn < n4) load(nxt, nxt_id, gn);
        uint32_t off[4];
        bool flag[4];
        bool any = false;
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
            ValT r0, r1;
            memcpy(&r0, &cur[i], 8); memcpy(&r1, &cur[i + 1], 8);
            const double v0 = sample_to_f64<ValT>(r0), v1 = sample_to_f64<ValT>(r1);
            const double x0 = __dadd_rn(1.0, fabs(v0)), x1 = __dadd_rn(1.0, fabs(v1));
            const uint32_t h0 = (uint32_t)__double2hiint(x0), h1 = (uint32_t)__double2hiint(x1);
            const uint32_t t0 = __funnelshift_l((uint32_t)__double2loint(x0), h0, 3);
            const uint32_t t1 = __funnelshift_l((uint32_t)__double2loint(x1), h1, 3);
            uint32_t m0, m1;
            asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(m0) : "r"(t0), "r"(one_bits));
            asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(m1) : "r"(t1), "r"(one_bits));
            float2 lg;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg.x) : "f"(__uint_as_float(m0)));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg.y) : "f"(__uint_as_float(m1)));
            const uint32_t e0 = h0 >> 20, e1 = h1 >> 20;
            const float2 ef = make_float2(__uint2float_rn(e0), __uint2float_rn(e1));
            const float2 a = __ffma2_rn(ef, make_float2(C2, C2), make_float2(KB, KB));
            const float2 w = __ffma2_rn(lg, make_float2(C1, C1), a);
            const float2 r = __fadd2_rn(w, make_float2(MAGIC, MAGIC));
            const float2 sv = __fadd2_rn(r, make_float2(-MAGIC, -MAGIC));
            const float2 d = __ffma2_rn(sv, make_float2(-1.0f, -1.0f), w);
            // v's high word >= 0x43E00000 unsigned: |v| >= 2^63, Inf, NaN and every negative value
            const uint32_t l0 = cur_id[i] - id_lo, l1 = cur_id[i + 1] - id_lo;          // local ids (wrap when below id_lo)
            const bool mine0 = valid & (l0 < id_cnt), mine1 = valid & (l1 < id_cnt);
            const bool bad0 = valid & (cur_id[i] >= H) & (id_lo == 0), bad1 = valid & (cur_id[i + 1] >= H) & (id_lo == 0);
            flag[i] = (mine0 & ((fabsf(d.x) > 0.5f - LH_FAST_EPS) | ((uint32_t)__double2hiint(v0) >= 0x43E00000u))) | bad0;
            flag[i + 1] = (mine1 & ((fabsf(d.y) > 0.5f - LH_FAST_EPS) | ((uint32_t)__double2hiint(v1) >= 0x43E00000u))) | bad1;
            off[i] = mine0 ? e0 * 276u + (__float_as_uint(r.x) << 2) + COFF + l0 * (uint32_t)(LH_WIN * 4) : trash_off;
            off[i + 1] = mine1 ? e1 * 276u + (__float_as_uint(r.y) << 2) + COFF + l1 * (uint32_t)(LH_WIN * 4) : trash_off;
            any |= flag[i] | flag[i + 1];
        }
        if (__any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (!flag[i]) continue;
                ValT rv;
                memcpy(&rv, &cur[i], 8);
                // uncertain-but-positive samples of a valid id could stay in shared memory; the L2 route is exact too
                // and keeps this path trivial (it handles ~0.05 % of the samples)
                keyed_one<ValT>(cur_id[i], rv, H, hot, buckets, dropped, pol);
                off[i] = trash_off;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(ks_hist) + off[i]), 1u);
#pragma unroll
        for (int i = 0; i < 4; i++) { cur[i] = nxt[i]; cur_id[i] = nxt_id[i]; }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < words; i += KS_THREADS) {
        const uint32_t c = ks_hist[i];
        if (!c) continue;
        const uint32_t lid = i / (uint32_t)LH_WIN, slot = i - lid * (uint32_t)LH_WIN;
        atomicAdd(&hot[(size_t)(id_lo + lid) * LH_SUBHIST + slot], c);
    }
}

// ------------------------------------------------------------- K1k/partitioned
// Gets the keyed path past the L2 atomic rate (one RED sector per sample).  One persistent cooperative CTA per
// SM; CTA p OWNS the histogram ids {p, p+P, p+2P, ...} and keeps their positive windows (uint32[ids_per][4368])
// in shared memory for the whole launch.  The stream is processed in chunks; per chunk
//   phase A  every CTA ("writer" w) bins its slice: bucket index via the fast path, a 16-bit record
//            (lid*4368 + slot) per sample, counting-sorted by owner in shared memory and appended as contiguous
//            runs to the (owner, writer) sub-queue in global memory -- every pair has its own region, so the
//            append offsets live in shared memory and no global atomic is needed; at the end of the slice the
//            writer publishes its P record counts;
//   barrier  grid-wide (one per chunk; sub-queues are double-buffered by chunk parity);
//   phase B  every owner drains its P sub-queues (L2 hits) into its shared-memory windows with shared atomics.
// Samples the window does not cover (negative, |v| >= 2^63, NaN/Inf), ids >= H and records that do not fit their
// sub-queue take the L2-atomic route of k_ingest_keyed.  At the end each CTA adds its windows into the hot window.
constexpr int KP_MAX_PARTS = 320;             // owners = CTAs: up to 2 per SM
constexpr int KP_SCAN_PER_LANE = KP_MAX_PARTS / 32;

struct KpParams {
    const void *ids;                 // IdT[n], 4*sizeof(IdT)-aligned
    const void *vals;                // ValT[n], 32-byte aligned
    size_t n;                        // multiple of 4
    uint32_t H;
    uint32_t ids_per;                // ceil(H / P)
    uint32_t cap;                    // records per (owner, writer) sub-queue per parity, multiple of 8
    uint32_t slice_tiles;            // tiles per CTA per chunk
    uint32_t inv_p;                  // floor(2^32 / P) + 1: id / P == __umulhi(id, inv_p) for id < 65536
    unsigned short *queues;          // [2][P owners][P writers][cap]
    unsigned int *q_cnt;             // [2][P owners][P writers]
    unsigned int *barrier;           // grid barrier counter, zeroed by the host before the launch
    unsigned int *hot;               // [H][LH_SUBHIST]
    unsigned long long *buckets;     // [H][65536]
    unsigned long long *dropped;
};

__device__ __forceinline__ void kp_grid_barrier(unsigned int *bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

template <typename IdT, typename ValT, int KP_THREADS, int KP_MINB>
__global__ void __launch_bounds__(KP_THREADS, KP_MINB)
k_ingest_keyed_part(KpParams prm) {
    constexpr int KP_TILE = KP_THREADS * 8;       // samples per binning tile (8 per thread)
    extern __shared__ __align__(16) unsigned char kp_smem[];
    const int P = gridDim.x, p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(kp_smem);                       // [ids_per][LH_WIN]
    const size_t hist_words = (size_t)prm.ids_per * LH_WIN;
    unsigned int *s_cnt = s_hist + hist_words;                                              // per-owner count, this tile
    unsigned int *s_start = s_cnt + KP_MAX_PARTS;                                           // exclusive scan of s_cnt
    unsigned int *s_off = s_start + KP_MAX_PARTS;                                           // records appended this chunk
    unsigned int *s_dst = s_off + KP_MAX_PARTS;                                             // [KP_TILE] global record index
    unsigned short *s_rec = reinterpret_cast<unsigned short *>(s_dst + KP_TILE);            // [KP_TILE]

    for (size_t i = tid; i < hist_words; i += KP_THREADS) s_hist[i] = 0;
    for (unsigned int i = tid; i < KP_TILE; i += KP_THREADS) s_dst[i] = 0xFFFFFFFFu;
    const uint64_t pol = policy_evict_last();
    __syncthreads();

    const size_t tiles_total = (prm.n + KP_TILE - 1) / KP_TILE;
    const size_t chunk_tiles = (size_t)prm.slice_tiles * P;
    const size_t nchunks = (tiles_total + chunk_tiles - 1) / chunk_tiles;
    const IdT *ids = reinterpret_cast<const IdT *>(prm.ids);
    const char *vals = reinterpret_cast<const char *>(prm.vals);
    const unsigned int cap = prm.cap;

    for (size_t c = 0; c < nchunks; c++) {
        const size_t par = c & 1;
        unsigned short *qset = prm.queues + par * (size_t)P * P * cap;      // [owner][writer][cap]
        unsigned int *cset = prm.q_cnt + par * (size_t)P * P;              // [owner][writer]
        if (tid < KP_MAX_PARTS) s_off[tid] = 0;
        // ---------------- phase A: bin my slice of chunk c (I am writer p)
        for (uint32_t t = 0; t < prm.slice_tiles; t++) {
            const size_t tile = c * chunk_tiles + (size_t)p * prm.slice_tiles + t;
            if (tile >= tiles_total) break;                       // uniform per CTA
            const size_t s0 = tile * KP_TILE;
            if (tid < KP_MAX_PARTS) s_cnt[tid] = 0;
            __syncthreads();
            // Common path is branch-free: owner/record/position for every sample; samples that need anything
            // else (id >= H, estimate too close to a boundary, |v| >= 2^63 / NaN / Inf, negative) raise one flag
            // and are handled -- entirely, on the L2 route -- in a warp-voted fix-up.
            uint32_t part[8], pos[8], rec[8];
            bool any_rare = false;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const size_t g = s0 + (size_t)h * (KP_TILE / 2) + (size_t)tid * 4;   // 4 consecutive samples
                uint32_t id4[4];
                unsigned long long raw[4] = {0, 0, 0, 0};
                const bool in = g < prm.n;                         // n is a multiple of 4
                id4[0] = id4[1] = id4[2] = id4[3] = 0;
                if (in) {
                    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                                 : "=l"(raw[0]), "=l"(raw[1]), "=l"(raw[2]), "=l"(raw[3]) : "l"(vals + g * 8));
                    if (sizeof(IdT) == 2) {
                        unsigned int lo, hi;
                        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi)
                                     : "l"(reinterpret_cast<const char *>(ids) + g * 2));
                        id4[0] = lo & 0xFFFFu; id4[1] = lo >> 16; id4[2] = hi & 0xFFFFu; id4[3] = hi >> 16;
                    } else {
                        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(id4[0]), "=r"(id4[1]), "=r"(id4[2]), "=r"(id4[3])
                                     : "l"(reinterpret_cast<const char *>(ids) + g * 4));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int q = h * 4 + j;
                    const uint32_t id = id4[j];
                    ValT rv;
                    memcpy(&rv, &raw[j], 8);
                    const double v = sample_to_f64<ValT>(rv);
                    uint32_t idx; bool slow;
                    fast_candidate(v, idx, slow);
                    const bool rare = slow | (idx >= (uint32_t)LH_WIN) | (id >= prm.H);
                    const uint32_t lid = __umulhi(id, prm.inv_p), owner = id - lid * (uint32_t)P;   // id / P, id % P
                    rec[q] = lid * (uint32_t)LH_WIN + idx;
                    // 0xFFFFFFFF = no record; 0xFFFFFFFE = rare (pending fix-up)
                    part[q] = !in ? 0xFFFFFFFFu : rare ? 0xFFFFFFFEu : owner;
                    any_rare |= in & rare;
                    pos[q] = 0;
                    if (in & !rare) pos[q] = atomicAdd(&s_cnt[owner], 1u);
                }
                if (__any_sync(0xFFFFFFFFu, any_rare)) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int q = h * 4 + j;
                        if (part[q] != 0xFFFFFFFEu) continue;
                        part[q] = 0xFFFFFFFFu;
                        ValT rv;
                        memcpy(&rv, &raw[j], 8);
                        keyed_one<ValT>(id4[j], rv, prm.H, prm.hot, prm.buckets, prm.dropped, pol);
                    }
                    any_rare = false;
                }
            }
            __syncthreads();
            // exclusive scan of the per-owner counts (one warp, KP_SCAN_PER_LANE owners per lane)
            if (warp == 0) {
                unsigned int loc[KP_SCAN_PER_LANE], sum = 0;
#pragma unroll
                for (int k = 0; k < KP_SCAN_PER_LANE; k++) { int o = lane * KP_SCAN_PER_LANE + k; loc[k] = (o < P) ? s_cnt[o] : 0; sum += loc[k]; }
                unsigned int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned int y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
                unsigned int run = incl - sum;
#pragma unroll
                for (int k = 0; k < KP_SCAN_PER_LANE; k++) { int o = lane * KP_SCAN_PER_LANE + k; if (o < P) s_start[o] = run; run += loc[k]; }
            }
            __syncthreads();
            {
                const unsigned int wbase = (unsigned int)p * cap;
                bool any_full = false;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint32_t o = part[q];
                    if (o < (uint32_t)KP_MAX_PARTS) {                   // has a record
                        const unsigned int at = s_off[o] + pos[q];      // position inside my sub-queue for owner o
                        const bool full = at >= cap;
                        any_full |= full;
                        if (!full) {
                            const unsigned int si = s_start[o] + pos[q];
                            s_rec[si] = (unsigned short)rec[q];
                            s_dst[si] = o * ((unsigned int)P * cap) + wbase + at;
                        }
                    }
                }
                if (__any_sync(0xFFFFFFFFu, any_full)) {                // sub-queue full: those records take the L2 route
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t o = part[q];
                        if (o < (uint32_t)KP_MAX_PARTS && s_off[o] + pos[q] >= cap) {
                            const uint32_t lid = rec[q] / (uint32_t)LH_WIN, slot = rec[q] - lid * (uint32_t)LH_WIN;
                            red_add_u32_keep(&prm.hot[(size_t)(lid * (uint32_t)P + o) * LH_SUBHIST + slot], 1u, pol);
                        }
                    }
                }
            }
            __syncthreads();
            {   // copy out: consecutive threads write consecutive records of one owner's run
                const unsigned int total = s_start[P - 1] + s_cnt[P - 1];
                for (unsigned int i = tid; i < total; i += KP_THREADS) {
                    const unsigned int d = s_dst[i];
                    if (d != 0xFFFFFFFFu) { qset[d] = s_rec[i]; s_dst[i] = 0xFFFFFFFFu; }
                }
            }
            if (tid < P) s_off[tid] += s_cnt[tid];
            __syncthreads();
        }
        // publish my P counts (zero for owners I sent nothing to), then the grid-wide barrier
        if (tid < P) cset[(size_t)tid * P + p] = min(s_off[tid], cap);
        kp_grid_barrier(prm.barrier, (unsigned int)((c + 1) * (size_t)P));
        // ---------------- phase B: drain the P sub-queues I own; warp w takes writers w, w+32, ...
        for (int w = warp; w < P; w += KP_THREADS / 32) {
            unsigned int nrec = 0;
            if (lane == 0) nrec = __ldcg(&cset[(size_t)p * P + w]);
            nrec = __shfl_sync(0xFFFFFFFFu, nrec, 0);
            const unsigned short *q = qset + ((size_t)p * P + w) * cap;
            const unsigned int nvec = nrec / 8;
            for (unsigned int i = lane; i < nvec; i += 32) {
                const uint4 v4 = __ldcg(reinterpret_cast<const uint4 *>(q) + i);
                const unsigned int ww[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    atomicAdd(&s_hist[ww[k] & 0xFFFFu], 1u);
                    atomicAdd(&s_hist[ww[k] >> 16], 1u);
                }
            }
            for (unsigned int i = nvec * 8 + lane; i < nrec; i += 32) atomicAdd(&s_hist[__ldcg(q + i)], 1u);
        }
        // no barrier here: the next chunk writes the other parity; this parity is rewritten only after the
        // next grid barrier, which every CTA reaches after finishing this drain
    }
    __syncthreads();
    // ---------------- flush my windows into the uint32 hot window
    for (size_t i = tid; i < hist_words; i += KP_THREADS) {
        const unsigned int cnt = s_hist[i];
        if (!cnt) continue;
        const uint32_t lid = (uint32_t)(i / LH_WIN), slot = (uint32_t)(i - (size_t)lid * LH_WIN);
        const uint32_t id = lid * (uint32_t)P + (uint32_t)p;
        if (id < prm.H) atomicAdd(&prm.hot[(size_t)id * LH_SUBHIST + slot], cnt);
    }
}

// Drain the hot window into the uint64 buckets.  atomicExch/atomicAdd so that ingest on other
// streams may keep running against the same buffer.
__global__ void k_fold_hot(unsigned int *__restrict__ hot, unsigned long long *__restrict__ buckets, size_t cells,
                           uint32_t replicas) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        unsigned long long sum = 0;
        for (uint32_t r = 0; r < replicas; r++) {
            unsigned int *cell = hot + (size_t)r * cells + i;
            if (*cell) sum += atomicExch(cell, 0u);
        }
        if (sum) {
            size_t h = i / LH_SUBHIST;
            uint32_t slot = (uint32_t)(i - h * LH_SUBHIST);
            atomicAdd(&buckets[h * 65536u + slot_to_key16(slot)], sum);
        }
    }
}

// ---------------------------------------------------------------------- K2
// Counter(name, amount): counters[id] += amount (wrapping uint64).  Up to
// K2_SMEM_COUNTERS ids are privatised per CTA as lo/hi uint32 halves in shared
// memory: one returning shared atomic on the low half, a second one on the
// high half only when the amount has high bits or the low half carried.
constexpr int K2_SMEM_COUNTERS = 8192;

template <typename IdT, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_counter_add_smem(const IdT *__restrict__ ids, const unsigned long long *__restrict__ amounts, size_t n,
                   unsigned long long *__restrict__ counters, uint32_t C,
                   unsigned long long *__restrict__ dropped) {
    extern __shared__ unsigned int s_cnt[];          // [C] low halves, [C] high halves
    unsigned int *lo = s_cnt, *hi = s_cnt + C;
    for (uint32_t i = threadIdx.x; i < 2 * C; i += THREADS) s_cnt[i] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * THREADS;
    for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
        uint32_t id = (uint32_t)ids[i];
        unsigned long long amt = amounts[i];
        if (id >= C) { atomicAdd(dropped, 1ull); continue; }
        unsigned int a_lo = (unsigned int)amt, a_hi = (unsigned int)(amt >> 32);
        unsigned int old = atomicAdd(&lo[id], a_lo);
        a_hi += (old + a_lo < old) ? 1u : 0u;        // carry out of the low half
        if (a_hi) atomicAdd(&hi[id], a_hi);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < C; i += THREADS) {
        unsigned long long v = ((unsigned long long)hi[i] << 32) | lo[i];
        if (v) atomicAdd(&counters[i], v);
    }
}

template <typename IdT, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_counter_add(const IdT *__restrict__ ids, const unsigned long long *__restrict__ amounts, size_t n,
              unsigned long long *__restrict__ counters, uint32_t C,
              unsigned long long *__restrict__ dropped) {
    const size_t stride = (size_t)gridDim.x * THREADS;
    for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
        uint32_t id = (uint32_t)ids[i];
        if (id < C) atomicAdd(&counters[id], amounts[i]);
        else atomicAdd(dropped, 1ull);
    }
}

// ------------------------------------------------------------------- merge
// Adds sparse (histogram id, int16 key, uint64 count) triples -- the wire format lh_snapshot_export produces --
// into the bucket arrays: merging snapshots from other hosts / GPUs is the same commutative uint64 sum.
__global__ void k_merge_sparse(const uint32_t *__restrict__ ids, const short *__restrict__ keys,
                               const unsigned long long *__restrict__ counts, size_t n, uint32_t H,
                               unsigned long long *__restrict__ buckets, unsigned long long *__restrict__ dropped) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t id = ids[i];
        if (id >= H) { atomicAdd(dropped, 1ull); continue; }
        atomicAdd(&buckets[(size_t)id * 65536u + ((uint32_t)(int)keys[i] & 0xFFFFu)], counts[i]);
    }
}

// ---------------------------------------------------------------------- K3
// One CTA per histogram, 32 warps.  Warp w owns the 2048 consecutive keys
// [-32768 + 2048 w, +2047] (ascending key == ascending value, the order
// percentile() sorts into, metrics.go:409) and reads them as 64 coalesced
// 256-byte rows.  Pass 1: per-warp count / sum / non-empty totals, then a scan
// over the 32 warp totals.  Pass 2: the warp whose range contains a
// percentile's crossing walks its rows again (L2 hits) with a warp prefix sum
// and applies the reference's rule float64(sofar)/float64(total) >= p
// (metrics.go:413) to non-empty buckets only.
constexpr int K3_THREADS = 1024;
constexpr int K3_WARP_KEYS = 2048;

__global__ void __launch_bounds__(K3_THREADS)
k_reduce(const unsigned long long *__restrict__ buckets, const double *__restrict__ decomp,
         const double *__restrict__ ps, int np, unsigned long long *__restrict__ out_count,
         double *__restrict__ out_sum, double *__restrict__ out_avg, int *__restrict__ out_pkeys,
         double *__restrict__ out_pvals, uint32_t *__restrict__ out_nnz) {
    __shared__ unsigned long long s_cnt[32];    // per-warp totals, then exclusive prefix
    __shared__ unsigned long long s_tot[32];    // per-warp totals (kept)
    __shared__ double s_sum[32];
    __shared__ unsigned int s_nnz[32];
    __shared__ int s_owner[LH_MAX_PCT];
    __shared__ unsigned long long s_total;
    const int h = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned long long *hb = buckets + (size_t)h * 65536u;
    const int key0 = -32768 + warp * K3_WARP_KEYS;

    unsigned long long mine = 0;
    double msum = 0.0;
    unsigned int nnz = 0;
#pragma unroll 16
    for (int r = 0; r < K3_WARP_KEYS / 32; r++) {      // 16 independent 256-byte rows in flight per warp
        unsigned int slot = (unsigned int)(key0 + r * 32 + lane) & 0xFFFFu;
        unsigned long long c = hb[slot];
        if (c) { mine += c; msum += decomp[slot] * (double)c; nnz++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
        msum += __shfl_xor_sync(0xFFFFFFFFu, msum, o);
        nnz += __shfl_xor_sync(0xFFFFFFFFu, nnz, o);
    }
    if (lane == 0) { s_cnt[warp] = mine; s_tot[warp] = mine; s_sum[warp] = msum; s_nnz[warp] = nnz; }
    if (t < LH_MAX_PCT) s_owner[t] = 0x7FFFFFFF;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = s_cnt[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += y;
        }
        s_cnt[lane] = wi - w;
        if (lane == 31) s_total = wi;
        double ts = s_sum[lane];
        unsigned int tn = s_nnz[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ts += __shfl_xor_sync(0xFFFFFFFFu, ts, o);
            tn += __shfl_xor_sync(0xFFFFFFFFu, tn, o);
        }
        if (lane == 0) { s_sum[0] = ts; s_nnz[0] = tn; }
    }
    __syncthreads();
    const unsigned long long total = s_total;
    const double ftotal = (double)total;
    // owner of percentile j = first non-empty warp whose inclusive prefix satisfies the rule
    if (lane == 0 && s_tot[warp]) {
        const unsigned long long end_incl = s_cnt[warp] + s_tot[warp];
        for (int j = 0; j < np; j++)
            if (__ddiv_rn((double)end_incl, ftotal) >= ps[j]) atomicMin(&s_owner[j], warp);
    }
    __syncthreads();
    unsigned int pending = 0;
    for (int j = 0; j < np; j++) if (s_owner[j] == warp) pending |= 1u << j;
    if (pending) {
        unsigned long long sofar = s_cnt[warp];
        for (int r = 0; r < K3_WARP_KEYS / 32 && pending; r++) {
            int key = key0 + r * 32 + lane;
            unsigned long long c = hb[(unsigned int)key & 0xFFFFu];
            unsigned long long incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += y;
            }
            const double frac = __ddiv_rn((double)(sofar + incl), ftotal);
            for (int j = 0; j < np; j++) {
                if (!(pending >> j & 1u)) continue;
                unsigned int hit = __ballot_sync(0xFFFFFFFFu, c != 0 && frac >= ps[j]);
                if (hit) {
                    if (lane == __ffs(hit) - 1) {
                        out_pkeys[(size_t)h * np + j] = key;
                        out_pvals[(size_t)h * np + j] = decomp[(unsigned int)key & 0xFFFFu];
                    }
                    pending &= ~(1u << j);
                }
            }
            sofar += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    if (t == 0) {
        for (int j = 0; j < np; j++)
            if (s_owner[j] == 0x7FFFFFFF) {   // percentile() error (p > 1, NaN, empty): key omitted by the caller
                out_pkeys[(size_t)h * np + j] = (int)0x80000000;
                out_pvals[(size_t)h * np + j] = __longlong_as_double(0x7FF8000000000000ll);
            }
        out_count[h] = total;
        out_sum[h] = s_sum[0];
        out_avg[h] = __ddiv_rn(s_sum[0], ftotal);    // metrics.go:356 (NaN when empty)
        out_nnz[h] = s_nnz[0];
    }
}

// ---------------------------------------------------------------------- K4
// offsets[h] = exclusive prefix of nnz (computed by k_scan_nnz); entries are
// written in ascending key order.
__global__ void k_scan_nnz(const uint32_t *__restrict__ nnz, uint32_t H, uint32_t *__restrict__ offsets) {
    // single CTA, H is small (<= a few thousand): serial per-chunk scan is fine
    __shared__ uint32_t s_part[1024];
    const int t = threadIdx.x;
    const uint32_t per = (H + 1023u) / 1024u;
    uint32_t a = 0;
    for (uint32_t i = 0; i < per; i++) { uint32_t idx = t * per + i; if (idx < H) a += nnz[idx]; }
    s_part[t] = a;
    __syncthreads();
    if (t == 0) { uint32_t run = 0; for (int i = 0; i < 1024; i++) { uint32_t x = s_part[i]; s_part[i] = run; run += x; } offsets[H] = run; }
    __syncthreads();
    uint32_t run = s_part[t];
    for (uint32_t i = 0; i < per; i++) { uint32_t idx = t * per + i; if (idx < H) { offsets[idx] = run; run += nnz[idx]; } }
}

__global__ void __launch_bounds__(K3_THREADS)
k_export(const unsigned long long *__restrict__ buckets, const uint32_t *__restrict__ offsets,
         short *__restrict__ out_keys, unsigned long long *__restrict__ out_counts) {
    __shared__ unsigned int s_warp[32];
    const int h = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned long long *hb = buckets + (size_t)h * 65536u;
    const int key0 = -32768 + warp * K3_WARP_KEYS;
    unsigned int nnz = 0;
#pragma unroll 4
    for (int r = 0; r < K3_WARP_KEYS / 32; r++)
        nnz += __popc(__ballot_sync(0xFFFFFFFFu, hb[(unsigned int)(key0 + r * 32 + lane) & 0xFFFFu] != 0));
    if (lane == 0) s_warp[warp] = nnz;     // every lane holds the warp total
    __syncthreads();
    if (warp == 0) {
        unsigned int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int y = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += y; }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    if (!nnz) return;
    unsigned int pos = offsets[h] + s_warp[warp];
    for (int r = 0; r < K3_WARP_KEYS / 32; r++) {
        int key = key0 + r * 32 + lane;
        unsigned long long c = hb[(unsigned int)key & 0xFFFFu];
        unsigned int m = __ballot_sync(0xFFFFFFFFu, c != 0);
        if (c) {
            unsigned int p = pos + __popc(m & ((1u << lane) - 1u));
            out_keys[p] = (short)key;
            out_counts[p] = c;
        }
        pos += __popc(m);
    }
}

// ----------------------------------------------------------- probes / tables
__global__ void k_fill_decompress(double *__restrict__ table) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536) table[i] = go_decompress((int)(short)i);
}

__global__ void k_compress_probe(const double *__restrict__ v, size_t n, short *__restrict__ out, int mode) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (short)(mode == 1 ? exact_key16(v[i]) : key16_of(v[i]));
}

// max | estimate - 100 ln(x) | over samples inside the fast window, for both estimators that ship:
//   [0] fast_candidate()      (k_ingest_keyed*, probes, ragged tails)
//   [2] bucket_samples_v2()   (packed-FP32 form with the -1023*c2 constant folded into the FMA; K1, keyed_small)
// plus [1] the tally of samples fast_candidate() sends to the exact path.
__global__ void k_fastpath_margin(const double *__restrict__ v, size_t n, unsigned long long *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = __dadd_rn(1.0, fabs(v[i]));
    uint32_t hi = (uint32_t)__double2hiint(x), lo = (uint32_t)__double2loint(x);
    uint32_t idx; bool slow;
    fast_candidate(v[i], idx, slow);
    if (slow) atomicAdd(&out[1], 1ull);
    if (hi >= 0x43E00000u) return;
    uint32_t t = __funnelshift_l(lo, hi, 3);
    float m = __uint_as_float((t & 0x007FFFFFu) | 0x3F800000u);
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(m));
    uint32_t eb = hi >> 20;
    const double truth = 100.0 * log(x);
    const double base = (double)(int)(eb - 1023u) * 69.0;
    {   // estimator 1
        float ef = __fadd_rn(__uint_as_float(0x4B000000u | eb), -(8388608.0f + 1023.0f));
        float w = __fmaf_rn(lg, 69.31471805599453f, __fmul_rn(ef, 0.31471805599453f));
        atomicMax(&out[0], (unsigned long long)__double_as_longlong(fabs(base + (double)w - truth)));
    }
    {   // estimator 2 (same constants as bucket_samples_v2 / k_ingest_keyed_small)
        constexpr float C1 = 69.31471805599453f, C2 = 0.31471805599453f;
        constexpr float KB = (float)(-1023.0 * (double)C2);
        float a = __fmaf_rn(__uint2float_rn(eb), C2, KB);
        float w = __fmaf_rn(lg, C1, a);
        atomicMax(&out[2], (unsigned long long)__double_as_longlong(fabs(base + (double)w - truth)));
    }
}

// ---------------------------------------------------------- synthetic streams
__device__ __constant__ unsigned char c_streamL_exp[16] = {17, 18, 18, 19, 19, 19, 20, 20, 20, 20, 21, 21, 21, 22, 22, 23};

__device__ __forceinline__ uint64_t stream_bits(int kind, uint64_t seed, uint64_t i) {
    uint64_t u = splitmix64(seed + i);
    uint64_t mant = u & 0x000FFFFFFFFFFFFFull;
    switch (kind) {
    case 0: return ((uint64_t)(1023 + (u >> 52) % 63) << 52) | mant;
    case 1: return ((uint64_t)(1023 + c_streamL_exp[(u >> 52) & 15]) << 52) | mant;
    case 2: {
        uint32_t sel = (uint32_t)(u >> 52) & 0xFFFu;
        if (sel < 41) return 0x8000000000000000ull | ((uint64_t)(1023 + (u >> 40) % 63) << 52) | mant;
        if (sel < 60) return splitmix64(u);
        if (sel < 80) return ((u >> 11) & 0x8000000000000000ull) | ((uint64_t)(1023 - 10 + (u >> 40) % 12) << 52) | mant;
        if (sel < 90) return ((u >> 13) & 0x8000000000000000ull) | ((uint64_t)(1023 + 63 + (u >> 40) % 961) << 52) | mant;
        return ((uint64_t)(1023 + (u >> 40) % 63) << 52) | mant;
    }
    case 3: return 0x40F86A0000000000ull;
    case 4:
        if (u >> 63) return 0x40F86A0000000000ull;
        return ((uint64_t)(1023 + c_streamL_exp[(u >> 52) & 15]) << 52) | mant;
    case 6: {   // timer durations as int64 nanoseconds: the L stream truncated toward zero
        double d = u64_as_f64(((uint64_t)(1023 + c_streamL_exp[(u >> 52) & 15]) << 52) | mant);
        return (uint64_t)__double2ll_rz(d);
    }
    case 7: return 1 + (u >> 60);   // counter amounts 1..16
    default: return u;
    }
}

__global__ void k_gen_stream(int kind, uint64_t seed, uint64_t start, size_t n, double *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = u64_as_f64(stream_bits(kind, seed, start + i));
}

__global__ void k_gen_ids_u16(int kind, uint64_t seed, uint64_t start, size_t n, uint32_t H,
                              unsigned short *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t u = splitmix64((seed ^ 0xA5A5A5A5DEADBEEFull) + start + i);
        uint32_t a = (uint32_t)((u & 0xFFFFFFFFu) % H), b = (uint32_t)((u >> 32) % H);
        out[i] = (unsigned short)(kind == 0 ? a : (a < b ? a : b));
    }
}

}  // namespace lh
