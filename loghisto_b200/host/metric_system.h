// metric_system.h -- C++ mirror of loghisto's MetricSystem for the ingest + reduction path.
//
// Same names, argument meaning and error behaviour as the Go type (reference metrics.go:80-653), written in
// C++ because the Go toolchain is absent from the build image (see INTEGRATION.md for the cgo form).
// The bodies of Histogram / Counter / TimerToken::Stop / collectRawMetrics / processMetrics go through the
// C ABI in include/loghisto_b200.h; nothing here computes a bucket, a count or a percentile on the CPU.
//
// Differences from the Go type, all outside the hot path:
//   * sysStats gauges (sys.Alloc, sys.NumGC, ... metrics.go:172-193) are Go-runtime facts and are not provided;
//     RegisterGaugeFunc / DeregisterGaugeFunc work as in the reference.
//   * channels are loghisto::Channel<T>: bounded, non-blocking send, closable (Go's `select { case ch <- x: default: }`).
//   * processMetrics() accepts only RawMetricSets produced by this system's collectRawMetrics(): the
//     per-histogram statistics were reduced on the GPU for exactly that snapshot and travel with it.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

struct lh_ctx;

namespace loghisto {

template <typename T>
class Channel {   // make(chan T, capacity)
 public:
    explicit Channel(size_t capacity) : cap_(capacity) {}
    // non-blocking send: false when the buffer is full or the channel is closed
    bool TrySend(T v) {
        std::lock_guard<std::mutex> lk(mu_);
        if (closed_ || q_.size() >= (cap_ ? cap_ : 1)) return false;
        q_.push_back(std::move(v));
        cv_.notify_one();
        return true;
    }
    // blocking receive with timeout; false on timeout or when closed and drained
    bool Receive(T *out, std::chrono::nanoseconds timeout) {
        std::unique_lock<std::mutex> lk(mu_);
        if (!cv_.wait_for(lk, timeout, [&] { return !q_.empty() || closed_; })) return false;
        if (q_.empty()) return false;
        *out = std::move(q_.front());
        q_.pop_front();
        return true;
    }
    void Close() { std::lock_guard<std::mutex> lk(mu_); closed_ = true; cv_.notify_all(); }
    bool Closed() { std::lock_guard<std::mutex> lk(mu_); return closed_; }
    size_t Len() { std::lock_guard<std::mutex> lk(mu_); return q_.size(); }

 private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

using TimePoint = std::chrono::system_clock::time_point;

// metrics.go:47-50
struct ProcessedMetricSet {
    TimePoint Time;
    std::map<std::string, double> Metrics;
};

struct ReducedHistogram {   // GPU results of processHistograms for one histogram of one snapshot
    uint64_t count = 0;
    double sum = 0, avg = 0;
    std::vector<int32_t> pkeys;   // INT32_MIN: percentile() error, key omitted
    std::vector<double> pvals;
};

// metrics.go:54-60 (map[int16]*uint64 becomes map<int16_t,uint64_t>: the snapshot owns plain values)
struct RawMetricSet {
    TimePoint Time;
    std::map<std::string, uint64_t> Counters;
    std::map<std::string, uint64_t> Rates;
    std::map<std::string, std::map<int16_t, uint64_t>> Histograms;
    std::map<std::string, double> Gauges;
    // travels with the snapshot: what the device reduced for it, and for which percentile labels
    std::map<std::string, ReducedHistogram> reduced;
    std::vector<std::pair<std::string, double>> percentile_labels;
    const void *origin = nullptr;
};

class MetricSystem;

// metrics.go:63-67
struct TimerToken {
    std::string Name;
    std::chrono::steady_clock::time_point Start;
    MetricSystem *System = nullptr;
    uint32_t id = 0;                   // dense histogram id interned by StartTimer (not in the Go type): Stop() skips the lookup
    bool id_valid = false;
    std::chrono::nanoseconds Stop();   // metrics.go:242-246
};

struct Options {
    int device = 0;
    uint32_t max_histograms = 1024;
    uint32_t max_counters = 1024;
    uint32_t shards = 0;             // staging shards (0 = one per hardware thread, at most 256)
    uint64_t staging_bytes = 4u << 20;
    uint32_t precision = 0;          // compress/decompress precision (0 = the reference's 100, metrics.go:40-43)
};

class MetricSystem {
 public:
    // NewMetricSystem(interval, sysStats), metrics.go:143.  Throws std::runtime_error if no B200 is usable.
    MetricSystem(std::chrono::nanoseconds interval, bool sysStats, const Options &opt = Options());
    ~MetricSystem();
    MetricSystem(const MetricSystem &) = delete;
    MetricSystem &operator=(const MetricSystem &) = delete;

    void SpecifyPercentiles(const std::map<std::string, double> &percentiles);              // :199
    void SubscribeToRawMetrics(std::shared_ptr<Channel<std::shared_ptr<RawMetricSet>>> ch);  // :205
    void UnsubscribeFromRawMetrics(std::shared_ptr<Channel<std::shared_ptr<RawMetricSet>>> ch);
    void SubscribeToProcessedMetrics(std::shared_ptr<Channel<std::shared_ptr<ProcessedMetricSet>>> ch);   // :218
    void UnsubscribeFromProcessedMetrics(std::shared_ptr<Channel<std::shared_ptr<ProcessedMetricSet>>> ch);
    TimerToken StartTimer(const std::string &name);                       // :232
    // Ingest never throws and never fails the caller (problems are logged, samples dropped and counted).
    void Counter(const std::string &name, uint64_t amount) noexcept;      // :251
    void Histogram(const std::string &name, double value) noexcept;       // :273
    void Histogram(const char *name, size_t len, double value) noexcept;  // same, without building a std::string
    void *assign_shard(size_t thread_slot, bool *exclusive);             // internal: a thread's staging shard (exclusive while any is free)
    void release_shard(void *shard);                                      // internal: a finished thread hands its exclusive shard back
    void histogram_id(uint32_t id, double value) noexcept;                // body of Histogram once the name is interned
    void RegisterGaugeFunc(const std::string &name, std::function<double()> f);   // :299
    void DeregisterGaugeFunc(const std::string &name);                    // :306
    void Start();                                                         // :644
    void Stop();                                                          // :651

    // unexported in Go, called directly by metrics_test.go; public here for the same purpose
    std::shared_ptr<RawMetricSet> collectRawMetrics();                                  // :420
    std::shared_ptr<ProcessedMetricSet> processMetrics(const RawMetricSet &raw);        // :483

    lh_ctx *context() const { return ctx_; }
    uint64_t dropped_samples();   // ids beyond max_histograms / max_counters (never silent)

 public:
    struct Shard;

 private:
    uint16_t intern(std::shared_mutex &mu, std::unordered_map<std::string, uint32_t> &ids,
                    std::vector<std::string> &names, const std::string &name, uint32_t limit, bool *ok);
    bool lookup_histogram(const char *p, size_t n, uint32_t *id);
    bool lookup_counter(const char *p, size_t n, uint32_t *id);
    void append_histogram(Shard &s, uint32_t id, double value) noexcept;
    void commit_histograms(Shard &s) noexcept;
    void commit_counters(Shard &s) noexcept;
    void flush_shard(Shard &s, std::vector<uint8_t> *touched);
    void reaper();
    void add_aggregates(const RawMetricSet &raw, ProcessedMetricSet &out);

    lh_ctx *ctx_ = nullptr;
    std::chrono::nanoseconds interval_;
    Options opt_;

    std::mutex percentiles_mu_;
    std::vector<std::pair<std::string, double>> percentiles_;   // label (with %s) -> p

    std::shared_mutex histo_mu_, counter_mu_;
    std::unordered_map<std::string, uint32_t> histo_ids_, counter_ids_;
    std::vector<std::string> histo_names_, counter_names_;

    std::vector<std::unique_ptr<Shard>> shards_;
    std::mutex assign_mu_;
    std::vector<Shard *> free_exclusive_, shared_;
    bool asym_ = false;             // exclusive shards use the membarrier handshake instead of a lock

    std::mutex counter_store_mu_;
    std::map<std::string, uint64_t> counter_store_;            // metrics.go:111-113
    std::mutex histogram_count_mu_;
    std::map<std::string, uint64_t> histogram_count_store_;    // metrics.go:122-126

    std::mutex gauge_mu_;
    std::map<std::string, std::function<double()>> gauge_funcs_;

    std::mutex subscribers_mu_;
    std::vector<std::shared_ptr<Channel<std::shared_ptr<RawMetricSet>>>> raw_subscribers_;
    std::vector<std::shared_ptr<Channel<std::shared_ptr<ProcessedMetricSet>>>> processed_subscribers_;
    std::map<void *, int> raw_bad_, processed_bad_;

    std::mutex snapshot_mu_;   // one collectRawMetrics at a time
    std::thread reaper_thread_;
    std::mutex run_mu_;
    std::condition_variable run_cv_;
    bool reaping_ = false, shutdown_ = false;
    std::atomic<uint64_t> dropped_over_limit_{0};
    uint64_t system_id_ = 0;
};

// print_benchmark.go:49: run `op` from `concurrency` threads between StartTimer/Stop and print every interval's
// metrics; returns the last interval's <name>_count after `seconds` (the reference runs forever).
double PrintBenchmark(const std::string &name, unsigned concurrency, std::function<void()> op, double seconds,
                      std::chrono::nanoseconds interval = std::chrono::seconds(1), const Options &opt = Options(),
                      bool print = true);

}  // namespace loghisto
