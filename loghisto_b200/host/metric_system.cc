// metric_system.cc -- see metric_system.h.  Host glue only: interning, batching into the pinned staging
// ring, rebuilding RawMetricSet / ProcessedMetricSet from what the device returns.
#include "metric_system.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sys/syscall.h>
#include <unistd.h>
#include <cstring>
#include <limits>
#include <stdexcept>

#include "../../include/loghisto_b200.h"

namespace loghisto {

namespace {

std::string format_label(const std::string &label, const std::string &name) {   // fmt.Sprintf(label, name), one %s
    size_t p = label.find("%s");
    if (p == std::string::npos) return label;
    return label.substr(0, p) + name + label.substr(p + 2);
}

// uint64(float64) as the Go compiler lowers it on amd64 (metrics.go:374): CVTTSD2SQ below 2^63
// (negative values wrap two's-complement), subtract-2^63 path above.
uint64_t go_f64_to_u64(double x) {
    if (x < 9223372036854775808.0) {
        if (!(x > -9223372036854777856.0)) return 0x8000000000000000ull;
        return (uint64_t)(int64_t)x;
    }
    if (!(x < 18446744073709551616.0)) return 0;
    return (uint64_t)(int64_t)(x - 9223372036854775808.0) ^ 0x8000000000000000ull;
}

void check(lh_ctx *ctx, lh_status st, const char *what) {
    if (st != LH_OK) {
        std::string msg = std::string(what) + ": " + lh_strerror(st);
        if (ctx) msg += std::string(" (") + lh_last_error(ctx) + ")";
        throw std::runtime_error(msg);
    }
}

// ---- asymmetric owner / reaper exclusion ------------------------------------------------------------------
// A staging shard that belongs to ONE thread is entered by that thread with plain stores (no locked instruction on
// the per-call path: on x86 a locked RMW waits for every store in flight, i.e. for the cache misses of the pinned
// buffer it just wrote).  The reaper, which needs the shard once per interval, raises `reaper_wants`, forces a full
// barrier on every thread of the process with membarrier(PRIVATE_EXPEDITED), and then waits for `owner_busy` to
// drop: the Dekker handshake with the expensive half moved to the side that runs once per interval.  Where the
// syscall is unavailable (or LOGHISTO_B200_SHARD_LOCK=1) every shard falls back to its spinlock.
constexpr int kMembarrierQuery = 0, kMembarrierPrivateExpedited = 1 << 3, kMembarrierRegisterPrivateExpedited = 1 << 4;
bool asym_available() {
    if (const char *e = getenv("LOGHISTO_B200_SHARD_LOCK")) if (e[0] == '1') return false;   // read per system: A/B runs in one process
    static const bool ok = [] {
#ifdef __NR_membarrier
        const long q = syscall(__NR_membarrier, kMembarrierQuery, 0);
        if (q < 0 || !(q & kMembarrierPrivateExpedited)) return false;
        return syscall(__NR_membarrier, kMembarrierRegisterPrivateExpedited, 0) == 0;
#else
        return false;
#endif
    }();
    return ok;
}
void asym_barrier() {
#ifdef __NR_membarrier
    syscall(__NR_membarrier, kMembarrierPrivateExpedited, 0);
#endif
}

// live systems by id: a thread that ends (or moves to another system) hands its exclusive shard back through this
std::mutex g_reg_mu;
std::unordered_map<uint64_t, MetricSystem *> g_systems;

size_t next_thread_slot() {
    static std::atomic<size_t> next{0};
    return next.fetch_add(1);
}

// ---- per-thread name -> id cache --------------------------------------------------------------------------
// The reference pays an RWMutex RLock/RUnlock (two atomic RMWs on ONE shared cache line) plus two string-hashed
// map lookups per sample (metrics.go:275-279); that line ping-pongs between cores and caps multi-core ingest.
// Here the steady-state lookup touches only thread-local memory: a direct-mapped cache keyed by the hash of the
// name's bytes, verified byte for byte (a collision can never send a sample to the wrong histogram).  A miss falls
// back to the shared intern table (the reference's RLock / double-checked Lock idiom, metrics.go:275-294).
// Fixed-size loads only: a memcpy of a run-time length is a library call, and it was most of the cost of a lookup.
inline uint64_t load8(const char *p) { uint64_t w; memcpy(&w, p, 8); return w; }
inline uint64_t load_1to7(const char *p, size_t n) {        // n in 1..7, reads only p[0..n)
    if (n >= 4) {
        uint32_t a, b;
        memcpy(&a, p, 4);
        memcpy(&b, p + n - 4, 4);
        return (uint64_t)a | ((uint64_t)b << 32);
    }
    return (uint64_t)(uint8_t)p[0] | ((uint64_t)(uint8_t)p[n >> 1] << 8) | ((uint64_t)(uint8_t)p[n - 1] << 16);
}
// the (at most) 16 bytes of a short name as two words that determine them given the length
inline void short_words(const char *p, size_t n, uint64_t *w0, uint64_t *w1) {
    if (n >= 8) { *w0 = load8(p); *w1 = load8(p + n - 8); }
    else { *w0 = n ? load_1to7(p, n) : 0; *w1 = 0; }
}
inline uint64_t hash_bytes(const char *p, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xFF51AFD7ED558CCDull);
    const char *q = p;
    size_t m = n;
    while (m >= 8) {
        h = (h ^ load8(q)) * 0xC2B2AE3D27D4EB4Full;
        h ^= h >> 29;
        q += 8; m -= 8;
    }
    if (m) {                                                 // the tail: an overlapping 8-byte load when the name allows it
        const uint64_t w = n >= 8 ? load8(p + n - 8) : load_1to7(q, m);
        h = (h ^ w) * 0xC2B2AE3D27D4EB4Full;
        h ^= h >> 29;
    }
    // full avalanche: the table index is the LOW bits, and the bytes that tell "histogram7" from "histogram1023" sit in
    // the high half of the overlapping tail word
    h ^= h >> 33;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}

struct NameCache {
    // open addressing, linear probing, load factor <= 1/2: a name that was interned once is found without ever going
    // back to the shared table (a direct-mapped cache thrashes as soon as two hot names collide).  Entries are 24
    // bytes; the names' bytes live in an append-only arena owned by the cache.
    struct Entry {                               // id_plus1 == 0: empty
        uint64_t hash; const char *name; uint32_t len; uint32_t id_plus1;
        uint64_t w0, w1;                         // short_words() of a name of at most 16 bytes: it never touches the arena
        bool matches(uint64_t h, const char *p, size_t n) const {
            if (hash != h || len != n) return false;
            if (n <= 16) {
                uint64_t a0, a1;
                short_words(p, n, &a0, &a1);
                return a0 == w0 && a1 == w1;
            }
            return memcmp(name, p, n) == 0;
        }
    };
    std::vector<Entry> e;
    std::vector<std::unique_ptr<char[]>> arena;
    size_t arena_left = 0;
    char *arena_next = nullptr;
    size_t count = 0;
    const Entry *last = nullptr;                 // most recently found entry: a thread that repeats one name skips the hash
    bool find_last(const char *p, size_t n, uint32_t *id) const {
        const Entry *x = last;
        if (x && x->len == n) {
            bool same;
            if (n <= 16) { uint64_t a0, a1; short_words(p, n, &a0, &a1); same = a0 == x->w0 && a1 == x->w1; }
            else same = memcmp(x->name, p, n) == 0;
            if (same) { *id = x->id_plus1 - 1; return true; }
        }
        return false;
    }
    bool find(uint64_t h, const char *p, size_t n, uint32_t *id) {
        if (e.empty()) return false;
        const size_t mask = e.size() - 1;
        for (size_t i = h & mask;; i = (i + 1) & mask) {
            const Entry &x = e[i];
            if (!x.id_plus1) return false;
            if (x.matches(h, p, n)) { *id = x.id_plus1 - 1; last = &x; return true; }
        }
    }
    void insert_raw(const Entry &en) {
        const size_t mask = e.size() - 1;
        size_t i = en.hash & mask;
        while (e[i].id_plus1) i = (i + 1) & mask;
        e[i] = en;
    }
    void put(uint64_t h, const char *p, size_t n, uint32_t id) {
        if (e.empty()) e.assign(256, Entry{});
        last = nullptr;
        if ((count + 1) * 2 > e.size()) {        // grow and rehash
            std::vector<Entry> old(e.size() * 2, Entry{});
            old.swap(e);
            for (const Entry &x : old) if (x.id_plus1) insert_raw(x);
        }
        if (n > arena_left) {
            const size_t block = std::max<size_t>(n, 16384);
            arena.emplace_back(new char[block]);
            arena_next = arena.back().get();
            arena_left = block;
        }
        memcpy(arena_next, p, n);
        Entry en{};
        en.hash = h; en.name = arena_next; en.len = (uint32_t)n; en.id_plus1 = id + 1;
        if (n <= 16) short_words(p, n, &en.w0, &en.w1);
        insert_raw(en);
        arena_next += n; arena_left -= n;
        count++;
    }
    void reset() { last = nullptr; e.clear(); arena.clear(); arena_left = 0; arena_next = nullptr; count = 0; }
};

// Everything a thread needs on the per-call path, reached through ONE thread-local pointer (a plain pointer has no
// initialisation guard; the object behind it is created on the thread's first call and freed when the thread ends).
struct ThreadState {
    uint64_t system_id = 0;                      // which MetricSystem `shard` and the caches belong to
    void *shard = nullptr;                       // MetricSystem::Shard * of this thread
    bool exclusive = false;                      // the shard is this thread's alone (handed back when the thread ends)
    size_t slot = 0;                             // process-wide thread number
    NameCache h, c;
};
void release_thread_shard(ThreadState *ts) {
    if (!ts->shard || !ts->exclusive) return;
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_systems.find(ts->system_id);
    if (it != g_systems.end()) it->second->release_shard(ts->shard);
    ts->shard = nullptr;
    ts->exclusive = false;
}
thread_local ThreadState *tl_state = nullptr;
struct ThreadStateOwner {                        // destroyed at thread exit
    ThreadState *p = nullptr;
    ~ThreadStateOwner() {
        if (p) release_thread_shard(p);
        delete p;
        p = nullptr;
        tl_state = nullptr;                      // a later thread-local destructor that still records a sample starts over
    }
};
thread_local ThreadStateOwner tl_owner;
ThreadState *make_thread_state() {
    tl_owner.p = new ThreadState();
    tl_owner.p->slot = next_thread_slot();
    tl_state = tl_owner.p;
    return tl_state;
}

std::atomic<uint64_t> g_system_ids{1};

}  // namespace

// One staging shard: a pinned staging slot for (id, value) samples and one for (id, amount) counter ops, filled
// with plain stores.  Threads map onto shards round-robin (one shard per hardware thread by default), so in the
// steady state a shard has ONE writer and its spinlock is uncontended (an exchange and a store, ~10 ns); the reaper
// takes it once per interval to commit whatever is open.
struct alignas(128) MetricSystem::Shard {     // its own cache-line pair: nothing else on the heap shares the lines its owner writes on every call
    // exclusive shards (one owner thread): asymmetric handshake, see asym_available()
    std::atomic<uint32_t> owner_busy{0};     // written by the owner with plain stores
    std::atomic<uint32_t> reaper_wants{0};   // raised by a collecting thread
    bool exclusive = false;
    // shared shards (more threads than shards, or no membarrier): a spinlock
    std::atomic_flag busy = ATOMIC_FLAG_INIT;
    void lock() { while (busy.test_and_set(std::memory_order_acquire)) { __builtin_ia32_pause(); } }
    void unlock() { busy.clear(std::memory_order_release); }
    // histogram / timer samples
    lh_staging hs{};
    bool h_open = false;
    size_t h_n = 0, h_cap = 0;
    double *h_vals = nullptr;
    uint16_t *h_ids = nullptr;
    // counter ops
    lh_staging cs{};
    bool c_open = false;
    size_t c_n = 0, c_cap = 0;
    uint64_t *c_amounts = nullptr;
    uint16_t *c_ids = nullptr;
    std::vector<uint8_t> c_touched;      // [max_counters] Counter(name, x) was called this interval, even with x == 0
    bool c_any_touched = false;
    std::atomic<uint64_t> dropped{0};    // samples lost to a failed staging call (never silent: dropped_samples())
    char pad[128];                       // keep neighbouring shards (and the adjacent-line prefetcher) off this one's cache lines
};

namespace {
struct ShardGuard {                      // the calling thread's OWN shard (or a shared one)
    MetricSystem::Shard &s;
    explicit ShardGuard(MetricSystem::Shard &sh) : s(sh) {
        if (!s.exclusive) { s.lock(); return; }
        for (;;) {
            s.owner_busy.store(1, std::memory_order_relaxed);
            std::atomic_signal_fence(std::memory_order_seq_cst);             // compiler only; the reaper's membarrier supplies the fence
            if (__builtin_expect(s.reaper_wants.load(std::memory_order_acquire) == 0, 1)) return;
            s.owner_busy.store(0, std::memory_order_release);                // a collect is flushing this shard (microseconds)
            while (s.reaper_wants.load(std::memory_order_acquire)) __builtin_ia32_pause();
        }
    }
    ~ShardGuard() {
        if (s.exclusive) s.owner_busy.store(0, std::memory_order_release);
        else s.unlock();
    }
};
struct ForeignGuard {                    // another thread's shard, after reaper_wants was raised and asym_barrier() ran
    MetricSystem::Shard &s;
    explicit ForeignGuard(MetricSystem::Shard &sh) : s(sh) {
        if (!s.exclusive) { s.lock(); return; }
        while (s.owner_busy.load(std::memory_order_acquire)) __builtin_ia32_pause();
    }
    ~ForeignGuard() {
        if (s.exclusive) s.reaper_wants.store(0, std::memory_order_release);
        else s.unlock();
    }
};
void log_once(std::atomic<bool> &flag, lh_ctx *ctx, lh_status st, const char *what) {
    if (!flag.exchange(true))
        fprintf(stderr, "loghisto: %s failed: %s (%s); samples are being dropped and counted\n", what, lh_strerror(st),
                ctx ? lh_last_error(ctx) : "");
}
std::atomic<bool> g_logged_staging{false};
}  // namespace

std::chrono::nanoseconds TimerToken::Stop() {
    auto d = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - Start);
    if (id_valid) System->histogram_id(id, (double)d.count());   // float64(duration.Nanoseconds()); name interned by StartTimer
    else System->Histogram(Name, (double)d.count());
    return d;
}

MetricSystem::MetricSystem(std::chrono::nanoseconds interval, bool /*sysStats*/, const Options &opt)
    : interval_(interval.count() > 0 ? interval : std::chrono::nanoseconds(1)), opt_(opt) {
    percentiles_ = {{"%s_min", 0.0}, {"%s_50", .5}, {"%s_75", .75}, {"%s_90", .9}, {"%s_95", .95},
                    {"%s_99", .99}, {"%s_99.9", .999}, {"%s_99.99", .9999}, {"%s_max", 1.0}};   // metrics.go:145-155
    // one shard per hardware thread: goroutines of print_benchmark.go:59-67 become OS threads here, each with its own
    // operational overrides (no recompile): LOGHISTO_B200_SHARDS, LOGHISTO_B200_STAGING_BYTES
    if (const char *e = getenv("LOGHISTO_B200_SHARDS")) { long v = atol(e); if (v > 0 && v <= 4096) opt_.shards = (uint32_t)v; }
    if (const char *e = getenv("LOGHISTO_B200_STAGING_BYTES")) { long long v = atoll(e); if (v >= 4096) opt_.staging_bytes = (uint64_t)v; }
    uint32_t nshards = opt_.shards ? opt_.shards : std::min<uint32_t>(std::max(1u, std::thread::hardware_concurrency()), 256u);
    system_id_ = g_system_ids.fetch_add(1);
    lh_config cfg{};
    cfg.struct_size = sizeof(cfg);
    cfg.device = opt.device;
    cfg.max_histograms = opt.max_histograms;
    cfg.max_counters = opt.max_counters;
    cfg.staging_bytes = opt_.staging_bytes;
    // exclusive shards for the first `nshards` live threads; a few shared, spin-locked ones for the threads beyond that
    asym_ = asym_available();
    const uint32_t nshared = asym_ ? 8u : 0u;
    cfg.staging_slots = 2 * (nshards + nshared) + 2;   // every shard may hold one histogram and one counter slot (memory is allocated on first use)
    cfg.precision = opt.precision;
    lh_status st = lh_create(&cfg, &ctx_);
    if (st != LH_OK) throw std::runtime_error(std::string("lh_create: ") + lh_strerror(st));
    for (uint32_t i = 0; i < nshards + nshared; i++) {
        shards_.emplace_back(new Shard());
        shards_.back()->c_touched.assign(opt.max_counters, 0);
        if (asym_ && i < nshards) {
            shards_.back()->exclusive = true;
            free_exclusive_.push_back(shards_.back().get());
        } else {
            shared_.push_back(shards_.back().get());
        }
    }
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_systems[system_id_] = this;
}

// the calling thread's shard: an exclusive one while any is free, otherwise one of the shared ones (round-robin)
void *MetricSystem::assign_shard(size_t thread_slot, bool *exclusive) {
    {
        std::lock_guard<std::mutex> lk(assign_mu_);
        if (!free_exclusive_.empty()) {
            Shard *s = free_exclusive_.back();
            free_exclusive_.pop_back();
            *exclusive = true;
            return s;
        }
    }
    *exclusive = false;
    return shared_[thread_slot % shared_.size()];
}
void MetricSystem::release_shard(void *shard) {
    std::lock_guard<std::mutex> lk(assign_mu_);
    free_exclusive_.push_back(static_cast<Shard *>(shard));   // whatever it still holds is committed by the next collect
}

MetricSystem::~MetricSystem() {
    {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        g_systems.erase(system_id_);
    }
    try { Stop(); } catch (...) {}
    if (reaper_thread_.joinable()) reaper_thread_.join();
    lh_destroy(ctx_);
}

void MetricSystem::SpecifyPercentiles(const std::map<std::string, double> &percentiles) {
    std::lock_guard<std::mutex> lk(percentiles_mu_);
    percentiles_.assign(percentiles.begin(), percentiles.end());
    if (percentiles_.size() > LH_MAX_PERCENTILES) percentiles_.resize(LH_MAX_PERCENTILES);
}

uint16_t MetricSystem::intern(std::shared_mutex &mu, std::unordered_map<std::string, uint32_t> &ids,
                              std::vector<std::string> &names, const std::string &name, uint32_t limit, bool *ok) {
    {   // read-lock fast path, then write-lock and re-check: the idiom of metrics.go:275-294
        std::shared_lock<std::shared_mutex> rl(mu);
        auto it = ids.find(name);
        if (it != ids.end()) { *ok = true; return (uint16_t)it->second; }
    }
    std::unique_lock<std::shared_mutex> wl(mu);
    auto it = ids.find(name);
    if (it != ids.end()) { *ok = true; return (uint16_t)it->second; }
    if (names.size() >= limit || names.size() >= 65536) { *ok = false; return 0; }
    uint32_t id = (uint32_t)names.size();
    ids.emplace(name, id);
    names.push_back(name);
    *ok = true;
    return (uint16_t)id;
}

// name -> dense id through the calling thread's cache; false when the name table is full (sample dropped, counted)
// The calling thread's state for THIS system (a thread that moves to another MetricSystem starts over).
static inline ThreadState *thread_state(MetricSystem &ms, uint64_t system_id) {
    ThreadState *ts = tl_state;
    if (__builtin_expect(ts == nullptr, 0)) ts = make_thread_state();
    if (__builtin_expect(ts->system_id != system_id, 0)) {
        release_thread_shard(ts);                // a shard of the system this thread used before
        ts->system_id = system_id;
        ts->shard = ms.assign_shard(ts->slot, &ts->exclusive);
        ts->h.reset();
        ts->c.reset();
    }
    return ts;
}

bool MetricSystem::lookup_histogram(const char *p, size_t n, uint32_t *id) {
    ThreadState *ts = thread_state(*this, system_id_);
    const uint64_t h = hash_bytes(p, n);
    if (ts->h.find(h, p, n, id)) return true;
    bool ok;
    const uint16_t v = intern(histo_mu_, histo_ids_, histo_names_, std::string(p, n), opt_.max_histograms, &ok);
    if (!ok) { dropped_over_limit_.fetch_add(1, std::memory_order_relaxed); return false; }
    ts->h.put(h, p, n, v);
    *id = v;
    return true;
}
bool MetricSystem::lookup_counter(const char *p, size_t n, uint32_t *id) {
    ThreadState *ts = thread_state(*this, system_id_);
    const uint64_t h = hash_bytes(p, n);
    if (ts->c.find(h, p, n, id)) return true;
    bool ok;
    const uint16_t v = intern(counter_mu_, counter_ids_, counter_names_, std::string(p, n), opt_.max_counters, &ok);
    if (!ok) { dropped_over_limit_.fetch_add(1, std::memory_order_relaxed); return false; }
    ts->c.put(h, p, n, v);
    *id = v;
    return true;
}

// Ingest never fails the caller and never throws (metrics.go:570-573, 632-636: problems are logged and data is
// dropped): a staging call that fails resets the shard, counts the samples it held as dropped and logs once.
void MetricSystem::append_histogram(Shard &s, uint32_t id, double value) noexcept {
    ShardGuard g(s);
    if (!s.h_open) {
        lh_status st = lh_staging_acquire(ctx_, &s.hs);
        if (st != LH_OK) { s.dropped.fetch_add(1, std::memory_order_relaxed); log_once(g_logged_staging, ctx_, st, "lh_staging_acquire"); return; }
        s.h_cap = ((size_t)s.hs.bytes / 10) & ~(size_t)15;
        s.h_vals = reinterpret_cast<double *>(s.hs.host);
        s.h_ids = reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(s.hs.host) + s.h_cap * 8);
        s.h_n = 0;
        s.h_open = true;
    }
    s.h_vals[s.h_n] = value;
    s.h_ids[s.h_n] = (uint16_t)id;
    if (++s.h_n == s.h_cap) commit_histograms(s);
}
void MetricSystem::histogram_id(uint32_t id, double value) noexcept {
    append_histogram(*static_cast<Shard *>(thread_state(*this, system_id_)->shard), id, value);
}

void MetricSystem::commit_histograms(Shard &s) noexcept {
    if (!s.h_open) return;
    lh_status st = lh_staging_commit_keyed_f64_u16(ctx_, &s.hs, s.h_n, s.h_cap * 8);
    if (st != LH_OK) {
        s.dropped.fetch_add(s.h_n, std::memory_order_relaxed);
        log_once(g_logged_staging, ctx_, st, "lh_staging_commit_keyed_f64_u16");
        lh_staging_abandon(ctx_, &s.hs);      // harmless if the commit already recycled the slot
    }
    s.h_open = false;
    s.h_n = 0;
}

void MetricSystem::commit_counters(Shard &s) noexcept {
    if (!s.c_open) return;
    lh_status st = lh_staging_commit_counter_u16(ctx_, &s.cs, s.c_n, s.c_cap * 8);
    if (st != LH_OK) {
        s.dropped.fetch_add(s.c_n, std::memory_order_relaxed);
        log_once(g_logged_staging, ctx_, st, "lh_staging_commit_counter_u16");
        lh_staging_abandon(ctx_, &s.cs);
    }
    s.c_open = false;
    s.c_n = 0;
}

void MetricSystem::Histogram(const std::string &name, double value) noexcept { Histogram(name.data(), name.size(), value); }
void MetricSystem::Histogram(const char *name, size_t len, double value) noexcept {
    // steady state: one thread-local pointer, one hash of the name's bytes, one probe of this thread's name cache, one
    // uncontended spinlock, two stores into pinned memory
    ThreadState *ts = thread_state(*this, system_id_);
    uint32_t id;
    if (!ts->h.find_last(name, len, &id) && __builtin_expect(!ts->h.find(hash_bytes(name, len), name, len, &id), 0)) {
        if (!lookup_histogram(name, len, &id)) return;              // name table full: dropped and counted
    }
    append_histogram(*static_cast<Shard *>(ts->shard), id, value);
}

void MetricSystem::Counter(const std::string &name, uint64_t amount) noexcept {
    ThreadState *ts = thread_state(*this, system_id_);
    uint32_t id;
    if (!ts->c.find_last(name.data(), name.size(), &id) &&
        __builtin_expect(!ts->c.find(hash_bytes(name.data(), name.size()), name.data(), name.size(), &id), 0)) {
        if (!lookup_counter(name.data(), name.size(), &id)) return;  // name table full: dropped and counted
    }
    Shard &s = *static_cast<Shard *>(ts->shard);
    ShardGuard g(s);
    s.c_touched[id] = 1;                    // Counter(name, 0) still makes the name appear in Rates (metrics.go:430-433)
    s.c_any_touched = true;
    if (amount == 0) return;                // nothing to add on the device
    if (!s.c_open) {
        lh_status st = lh_staging_acquire(ctx_, &s.cs);
        if (st != LH_OK) { s.dropped.fetch_add(1, std::memory_order_relaxed); log_once(g_logged_staging, ctx_, st, "lh_staging_acquire"); return; }
        s.c_cap = ((size_t)s.cs.bytes / 10) & ~(size_t)15;
        s.c_amounts = reinterpret_cast<uint64_t *>(s.cs.host);
        s.c_ids = reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(s.cs.host) + s.c_cap * 8);
        s.c_n = 0;
        s.c_open = true;
    }
    s.c_amounts[s.c_n] = amount;
    s.c_ids[s.c_n] = (uint16_t)id;
    if (++s.c_n == s.c_cap) commit_counters(s);
}

TimerToken MetricSystem::StartTimer(const std::string &name) {
    TimerToken t;
    t.Name = name;
    t.System = this;
    t.id_valid = lookup_histogram(name.data(), name.size(), &t.id);   // interned once here; Stop() skips the lookup
    t.Start = std::chrono::steady_clock::now();
    return t;
}

void MetricSystem::RegisterGaugeFunc(const std::string &name, std::function<double()> f) {
    std::lock_guard<std::mutex> lk(gauge_mu_);
    gauge_funcs_[name] = std::move(f);
}
void MetricSystem::DeregisterGaugeFunc(const std::string &name) {
    std::lock_guard<std::mutex> lk(gauge_mu_);
    gauge_funcs_.erase(name);
}

// commit whatever the shard holds and hand over (and clear) its touched-counter marks
void MetricSystem::flush_shard(Shard &s, std::vector<uint8_t> *touched) {
    ForeignGuard g(s);
    commit_histograms(s);
    commit_counters(s);
    if (s.c_any_touched) {
        for (size_t i = 0; i < s.c_touched.size(); i++)
            if (s.c_touched[i]) { (*touched)[i] = 1; s.c_touched[i] = 0; }
        s.c_any_touched = false;
    }
}

uint64_t MetricSystem::dropped_samples() {
    lh_stats st{};
    lh_get_stats(ctx_, &st);
    uint64_t d = st.dropped + dropped_over_limit_.load();
    for (auto &s : shards_) d += s->dropped.load(std::memory_order_relaxed);
    return d;
}

// collectRawMetrics, metrics.go:420-479.
std::shared_ptr<RawMetricSet> MetricSystem::collectRawMetrics() {
    std::lock_guard<std::mutex> snap(snapshot_mu_);
    auto raw = std::make_shared<RawMetricSet>();
    const int64_t now = std::chrono::duration_cast<std::chrono::nanoseconds>(
                            std::chrono::system_clock::now().time_since_epoch()).count();
    const int64_t iv = interval_.count();
    raw->Time = TimePoint(std::chrono::duration_cast<TimePoint::duration>(std::chrono::nanoseconds(now / iv * iv)));   // :421-423
    raw->origin = this;
    {
        std::lock_guard<std::mutex> lk(percentiles_mu_);
        raw->percentile_labels = percentiles_;
    }

    std::vector<uint8_t> touched(opt_.max_counters, 0);
    // exclusive shards: announce, one process-wide barrier, then every shard is entered as soon as its owner is out
    if (asym_) {
        for (auto &s : shards_) if (s->exclusive) s->reaper_wants.store(1, std::memory_order_seq_cst);
        asym_barrier();
    }
    for (auto &s : shards_) flush_shard(*s, &touched);
    check(ctx_, lh_snapshot_begin(ctx_), "lh_snapshot_begin");   // the cache swaps of :425-428 and :460-463

    std::vector<std::string> hnames, cnames;
    {
        std::shared_lock<std::shared_mutex> rl(histo_mu_);
        hnames = histo_names_;
    }
    {
        std::shared_lock<std::shared_mutex> rl(counter_mu_);
        cnames = counter_names_;
    }
    const uint32_t H = opt_.max_histograms, np = (uint32_t)raw->percentile_labels.size();
    std::vector<double> ps(np);
    for (uint32_t j = 0; j < np; j++) ps[j] = raw->percentile_labels[j].second;
    std::vector<uint64_t> counts(H);
    std::vector<double> sums(H), avgs(H), pvals((size_t)H * np);
    std::vector<int32_t> pkeys((size_t)H * np);
    lh_sparse sp{};
    try {
        check(ctx_, lh_snapshot_reduce(ctx_, ps.data(), np, counts.data(), sums.data(), avgs.data(), pkeys.data(), pvals.data()),
              "lh_snapshot_reduce");
        check(ctx_, lh_snapshot_export(ctx_, &sp), "lh_snapshot_export");
    } catch (...) {
        lh_snapshot_end(ctx_);
        throw;
    }
    // histograms: present only when touched this interval (the swapped-out cache only holds touched names)
    for (size_t h = 0; h < hnames.size(); h++) {
        if (sp.offsets[h] == sp.offsets[h + 1]) continue;
        auto &m = raw->Histograms[hnames[h]];
        for (uint32_t i = sp.offsets[h]; i < sp.offsets[h + 1]; i++) m[sp.keys[i]] = sp.counts[i];
        ReducedHistogram r;
        r.count = counts[h]; r.sum = sums[h]; r.avg = avgs[h];
        r.pkeys.assign(pkeys.begin() + h * np, pkeys.begin() + (h + 1) * np);
        r.pvals.assign(pvals.begin() + h * np, pvals.begin() + (h + 1) * np);
        raw->reduced[hnames[h]] = std::move(r);
    }
    // counters: Rates = interval deltas of the names touched (:430-433); Counters = cumulative store (:435-458).
    // "Touched" is tracked on the host (a name appears in Rates even when only Counter(name, 0) was called, or when
    // its amounts wrapped to a zero delta); the deltas themselves come from the device.
    {
        std::lock_guard<std::mutex> lk(counter_store_mu_);
        for (size_t c = 0; c < cnames.size(); c++) {
            uint64_t d = sp.counter_deltas[c];
            if (d || touched[c]) {
                raw->Rates[cnames[c]] = d;
                counter_store_[cnames[c]] += d;
            }
        }
        raw->Counters = counter_store_;
    }
    check(ctx_, lh_snapshot_end(ctx_), "lh_snapshot_end");
    {
        std::lock_guard<std::mutex> lk(gauge_mu_);
        for (auto &g : gauge_funcs_) raw->Gauges[g.first] = g.second();   // :465-470
    }
    return raw;
}

// processMetrics + processHistograms, metrics.go:483-506 and :336-387.
std::shared_ptr<ProcessedMetricSet> MetricSystem::processMetrics(const RawMetricSet &raw) {
    if (raw.origin != this)
        throw std::invalid_argument("processMetrics: RawMetricSet was not produced by this MetricSystem's collectRawMetrics");
    auto out = std::make_shared<ProcessedMetricSet>();
    out->Time = raw.Time;
    auto &m = out->Metrics;
    for (auto &c : raw.Counters) m[c.first] = (double)c.second;
    for (auto &r : raw.Rates) m[r.first + "_rate"] = (double)r.second;
    for (auto &h : raw.Histograms) {
        const std::string &name = h.first;
        auto it = raw.reduced.find(name);
        if (it == raw.reduced.end()) continue;
        const ReducedHistogram &r = it->second;
        const std::string sumName = name + "_sum", countName = name + "_count", avgName = name + "_avg";
        m[countName] = (double)r.count;
        m[sumName] = r.sum;
        m[avgName] = r.avg;
        {   // aggregate store, :359-376
            std::lock_guard<std::mutex> lk(histogram_count_mu_);
            histogram_count_store_[sumName] += go_f64_to_u64(r.sum);
            histogram_count_store_[countName] += r.count;
        }
        for (size_t j = 0; j < raw.percentile_labels.size(); j++) {
            if (r.pkeys[j] == std::numeric_limits<int32_t>::min()) {   // percentile() error: logged, key omitted (:380-382)
                fprintf(stderr, "loghisto: unable to calculate percentile: Invalid percentile.  Should be between 0 and 1.\n");
                continue;
            }
            m[format_label(raw.percentile_labels[j].first, name)] = r.pvals[j];
        }
    }
    for (auto &g : raw.Gauges) m[g.first] = g.second;
    return out;
}

// the reaper's "add aggregate mean" step, metrics.go:590-608 (integer division)
void MetricSystem::add_aggregates(const RawMetricSet &raw, ProcessedMetricSet &out) {
    for (auto &h : raw.Histograms) {
        uint64_t aggCount = 0, aggSum = 0;
        bool countPresent, sumPresent;
        {
            std::lock_guard<std::mutex> lk(histogram_count_mu_);
            auto c = histogram_count_store_.find(h.first + "_count");
            auto s = histogram_count_store_.find(h.first + "_sum");
            countPresent = c != histogram_count_store_.end();
            sumPresent = s != histogram_count_store_.end();
            if (countPresent) aggCount = c->second;
            if (sumPresent) aggSum = s->second;
        }
        if (countPresent && sumPresent && aggCount > 0) {
            out.Metrics[h.first + "_agg_avg"] = (double)(aggSum / aggCount);
            out.Metrics[h.first + "_agg_count"] = (double)aggCount;
            out.Metrics[h.first + "_agg_sum"] = (double)aggSum;
        }
    }
}

void MetricSystem::SubscribeToRawMetrics(std::shared_ptr<Channel<std::shared_ptr<RawMetricSet>>> ch) {
    std::lock_guard<std::mutex> lk(subscribers_mu_);
    raw_subscribers_.push_back(std::move(ch));
}
void MetricSystem::UnsubscribeFromRawMetrics(std::shared_ptr<Channel<std::shared_ptr<RawMetricSet>>> ch) {
    std::lock_guard<std::mutex> lk(subscribers_mu_);
    raw_subscribers_.erase(std::remove(raw_subscribers_.begin(), raw_subscribers_.end(), ch), raw_subscribers_.end());
    raw_bad_.erase(ch.get());
}
void MetricSystem::SubscribeToProcessedMetrics(std::shared_ptr<Channel<std::shared_ptr<ProcessedMetricSet>>> ch) {
    std::lock_guard<std::mutex> lk(subscribers_mu_);
    processed_subscribers_.push_back(std::move(ch));
}
void MetricSystem::UnsubscribeFromProcessedMetrics(std::shared_ptr<Channel<std::shared_ptr<ProcessedMetricSet>>> ch) {
    std::lock_guard<std::mutex> lk(subscribers_mu_);
    processed_subscribers_.erase(std::remove(processed_subscribers_.begin(), processed_subscribers_.end(), ch),
                                 processed_subscribers_.end());
    processed_bad_.erase(ch.get());
}

// reaper, metrics.go:530-639: wake at wall-clock multiples of the interval, collect, broadcast raw,
// process, add aggregates, broadcast processed; never block on a subscriber, close one that misses twice.
void MetricSystem::reaper() {
    for (;;) {
        const int64_t iv = interval_.count();
        const int64_t now = std::chrono::duration_cast<std::chrono::nanoseconds>(
                                std::chrono::system_clock::now().time_since_epoch()).count();
        const int64_t tts = iv - (now % iv);
        {
            std::unique_lock<std::mutex> lk(run_mu_);
            if (run_cv_.wait_for(lk, std::chrono::nanoseconds(tts), [&] { return shutdown_; })) {
                reaping_ = false;
                return;
            }
        }
        std::shared_ptr<RawMetricSet> raw;
        try {
            raw = collectRawMetrics();
        } catch (const std::exception &e) {
            fprintf(stderr, "loghisto: collectRawMetrics failed: %s\n", e.what());
            continue;
        }
        {
            std::lock_guard<std::mutex> lk(subscribers_mu_);
            for (size_t i = 0; i < raw_subscribers_.size();) {
                auto &ch = raw_subscribers_[i];
                if (ch->TrySend(raw)) { raw_bad_.erase(ch.get()); i++; continue; }
                fprintf(stderr, "loghisto: a raw subscriber has allowed their channel to fill up. dropping their metrics on the floor rather than blocking.\n");
                if (++raw_bad_[ch.get()] >= 2) {
                    ch->Close();
                    raw_bad_.erase(ch.get());
                    raw_subscribers_.erase(raw_subscribers_.begin() + i);
                } else {
                    i++;
                }
            }
        }
        std::shared_ptr<ProcessedMetricSet> processed;
        try {
            processed = processMetrics(*raw);
        } catch (const std::exception &e) {
            fprintf(stderr, "loghisto: processMetrics failed: %s\n", e.what());
            continue;
        }
        add_aggregates(*raw, *processed);
        {
            std::lock_guard<std::mutex> lk(subscribers_mu_);
            for (size_t i = 0; i < processed_subscribers_.size();) {
                auto &ch = processed_subscribers_[i];
                if (ch->TrySend(processed)) { processed_bad_.erase(ch.get()); i++; continue; }
                fprintf(stderr, "loghisto: a subscriber has allowed their channel to fill up. dropping their metrics on the floor rather than blocking.\n");
                if (++processed_bad_[ch.get()] >= 2) {
                    ch->Close();
                    processed_bad_.erase(ch.get());
                    processed_subscribers_.erase(processed_subscribers_.begin() + i);
                } else {
                    i++;
                }
            }
        }
    }
}

void MetricSystem::Start() {
    std::lock_guard<std::mutex> lk(run_mu_);
    if (reaping_ || shutdown_) return;
    reaping_ = true;
    reaper_thread_ = std::thread([this] { reaper(); });
}

void MetricSystem::Stop() {   // idempotent, unlike the reference's double close (metrics.go:652)
    {
        std::lock_guard<std::mutex> lk(run_mu_);
        shutdown_ = true;
    }
    run_cv_.notify_all();
    if (reaper_thread_.joinable() && std::this_thread::get_id() != reaper_thread_.get_id()) reaper_thread_.join();
}

}  // namespace loghisto

// ---------------------------------------------------------------------------------------------
// C shim so that ctypes tests can replay the reference's metrics_test.go against the C++ mirror.
using namespace loghisto;

extern "C" {
#define LHMS_API __attribute__((visibility("default")))
typedef void (*lhms_emit_fn)(void *ctx, int kind, const char *name, int key, uint64_t u, double f);

LHMS_API void *lhms_new(int64_t interval_ns, int device, uint32_t max_histograms, uint32_t max_counters, char *err, int errlen) {
    try {
        Options o;
        o.device = device;
        o.max_histograms = max_histograms;
        o.max_counters = max_counters;
        return new MetricSystem(std::chrono::nanoseconds(interval_ns), false, o);
    } catch (const std::exception &e) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.what());
        return nullptr;
    }
}
LHMS_API void lhms_free(void *ms) { delete static_cast<MetricSystem *>(ms); }
// Nothing may unwind through these C entry points (ctypes / cgo callers): the ingest methods are noexcept, the
// remaining allocations are guarded.
LHMS_API void lhms_histogram(void *ms, const char *name, double v) {
    if (!ms || !name) return;
    static_cast<MetricSystem *>(ms)->Histogram(name, strlen(name), v);
}
LHMS_API void lhms_counter(void *ms, const char *name, uint64_t a) {
    if (!ms || !name) return;
    try { static_cast<MetricSystem *>(ms)->Counter(std::string(name), a); } catch (...) {}
}
LHMS_API void lhms_histogram_many(void *ms, const char *name, const double *v, size_t n) {
    if (!ms || !name) return;
    auto *m = static_cast<MetricSystem *>(ms);
    const size_t len = strlen(name);
    for (size_t i = 0; i < n; i++) m->Histogram(name, len, v[i]);
}
LHMS_API void *lhms_start_timer(void *ms, const char *name) {
    if (!ms || !name) return nullptr;
    try { return new TimerToken(static_cast<MetricSystem *>(ms)->StartTimer(name)); } catch (...) { return nullptr; }
}
// consumes the token; a NULL token (failed start, or a second stop through a cleared handle) returns 0
LHMS_API int64_t lhms_timer_stop(void *token) {
    if (!token) return 0;
    auto *t = static_cast<TimerToken *>(token);
    int64_t ns = t->Stop().count();
    delete t;
    return ns;
}
LHMS_API void lhms_timer_free(void *token) { delete static_cast<TimerToken *>(token); }
LHMS_API void lhms_specify_percentiles(void *ms, int n, const char *const *labels, const double *ps) {
    std::map<std::string, double> m;
    for (int i = 0; i < n; i++) m[labels[i]] = ps[i];
    static_cast<MetricSystem *>(ms)->SpecifyPercentiles(m);
}
LHMS_API void lhms_register_constant_gauge(void *ms, const char *name, double v) {
    static_cast<MetricSystem *>(ms)->RegisterGaugeFunc(name, [v] { return v; });
}

static void emit_raw(const RawMetricSet &raw, lhms_emit_fn emit, void *ctx) {
    for (auto &c : raw.Counters) emit(ctx, 0, c.first.c_str(), 0, c.second, 0);
    for (auto &r : raw.Rates) emit(ctx, 1, r.first.c_str(), 0, r.second, 0);
    for (auto &h : raw.Histograms)
        for (auto &b : h.second) emit(ctx, 2, h.first.c_str(), b.first, b.second, 0);
    for (auto &g : raw.Gauges) emit(ctx, 4, g.first.c_str(), 0, 0, g.second);
}
static void emit_processed(const ProcessedMetricSet &p, lhms_emit_fn emit, void *ctx) {
    for (auto &m : p.Metrics) emit(ctx, 3, m.first.c_str(), 0, 0, m.second);
}

// processMetrics(collectRawMetrics()), as metrics_test.go:195 does.  Returns 0, or -1 with err filled.
LHMS_API int lhms_collect_and_process(void *ms, lhms_emit_fn emit, void *ctx, char *err, int errlen) {
    try {
        auto *m = static_cast<MetricSystem *>(ms);
        auto raw = m->collectRawMetrics();
        auto p = m->processMetrics(*raw);
        emit_raw(*raw, emit, ctx);
        emit_processed(*p, emit, ctx);
        return 0;
    } catch (const std::exception &e) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.what());
        return -1;
    }
}
LHMS_API void lhms_start(void *ms) { static_cast<MetricSystem *>(ms)->Start(); }
LHMS_API void lhms_stop(void *ms) { static_cast<MetricSystem *>(ms)->Stop(); }
LHMS_API uint64_t lhms_dropped(void *ms) { return static_cast<MetricSystem *>(ms)->dropped_samples(); }

using RawCh = std::shared_ptr<Channel<std::shared_ptr<RawMetricSet>>>;
using ProcCh = std::shared_ptr<Channel<std::shared_ptr<ProcessedMetricSet>>>;

LHMS_API void *lhms_subscribe_processed(void *ms, int capacity) {
    auto *ch = new ProcCh(std::make_shared<Channel<std::shared_ptr<ProcessedMetricSet>>>((size_t)capacity));
    static_cast<MetricSystem *>(ms)->SubscribeToProcessedMetrics(*ch);
    return ch;
}
LHMS_API void lhms_unsubscribe_processed(void *ms, void *ch) {
    static_cast<MetricSystem *>(ms)->UnsubscribeFromProcessedMetrics(*static_cast<ProcCh *>(ch));
}
// 1 = received, 0 = timeout, -1 = channel closed by the reaper
LHMS_API int lhms_recv_processed(void *ch, int64_t timeout_ns, lhms_emit_fn emit, void *ctx) {
    auto &c = *static_cast<ProcCh *>(ch);
    std::shared_ptr<ProcessedMetricSet> p;
    if (c->Receive(&p, std::chrono::nanoseconds(timeout_ns))) { emit_processed(*p, emit, ctx); return 1; }
    return c->Closed() ? -1 : 0;
}
LHMS_API void lhms_free_processed_channel(void *ch) { delete static_cast<ProcCh *>(ch); }

LHMS_API void *lhms_subscribe_raw(void *ms, int capacity) {
    auto *ch = new RawCh(std::make_shared<Channel<std::shared_ptr<RawMetricSet>>>((size_t)capacity));
    static_cast<MetricSystem *>(ms)->SubscribeToRawMetrics(*ch);
    return ch;
}
LHMS_API void lhms_unsubscribe_raw(void *ms, void *ch) {
    static_cast<MetricSystem *>(ms)->UnsubscribeFromRawMetrics(*static_cast<RawCh *>(ch));
}
LHMS_API int lhms_recv_raw(void *ch, int64_t timeout_ns, lhms_emit_fn emit, void *ctx) {
    auto &c = *static_cast<RawCh *>(ch);
    std::shared_ptr<RawMetricSet> r;
    if (c->Receive(&r, std::chrono::nanoseconds(timeout_ns))) { emit_raw(*r, emit, ctx); return 1; }
    return c->Closed() ? -1 : 0;
}
LHMS_API void lhms_free_raw_channel(void *ch) { delete static_cast<RawCh *>(ch); }
}  // extern "C"
