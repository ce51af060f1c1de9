// print_benchmark.cc -- C++ form of the reference's load generator (print_benchmark.go:49-106) over the host
// mirror: `concurrency` threads loop { StartTimer(name); op(); Stop() } against one MetricSystem with a 1 s (here:
// configurable) interval, a receiver prints the interesting keys of every ProcessedMetricSet.  Unlike the reference
// it stops after `seconds` and returns the last interval's <name>_count, so it can be used as a measurement.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "metric_system.h"

namespace loghisto {

double PrintBenchmark(const std::string &name, unsigned concurrency, std::function<void()> op, double seconds,
                      std::chrono::nanoseconds interval, const Options &opt, bool print) {
    MetricSystem ms(interval, true, opt);
    auto mc = std::make_shared<Channel<std::shared_ptr<ProcessedMetricSet>>>(1);
    ms.SubscribeToProcessedMetrics(mc);
    ms.Start();
    std::atomic<bool> stop{false};
    std::vector<std::thread> workers;
    for (unsigned i = 0; i < concurrency; i++)
        workers.emplace_back([&] {
            try {
                while (!stop.load(std::memory_order_relaxed)) {
                    TimerToken timer = ms.StartTimer(name);   // print_benchmark.go:62-64
                    op();
                    timer.Stop();
                }
            } catch (const std::exception &e) {               // an exception from op() must not reach std::terminate
                fprintf(stderr, "PrintBenchmark worker: %s\n", e.what());
            }
        });
    static const char *suffixes[] = {"_count", "_max", "_99.99", "_99.9", "_99", "_95", "_90", "_75", "_50", "_min",
                                     "_sum", "_avg", "_agg_avg", "_agg_count", "_agg_sum"};
    double last_count = 0;
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
    while (std::chrono::steady_clock::now() < t_end) {
        std::shared_ptr<ProcessedMetricSet> m;
        if (!mc->Receive(&m, std::chrono::milliseconds(50))) {
            if (mc->Closed()) break;
            continue;
        }
        auto it = m->Metrics.find(name + "_count");
        if (it != m->Metrics.end()) last_count = it->second;
        if (print) {
            for (const char *sfx : suffixes) {
                auto e = m->Metrics.find(name + sfx);
                printf("%s%s:\t%.17g\n", name.c_str(), sfx, e == m->Metrics.end() ? 0.0 : e->second);
            }
            printf("\n");
        }
    }
    stop.store(true);
    for (auto &w : workers) w.join();
    ms.Stop();
    return last_count;
}

}  // namespace loghisto

// ---------------------------------------------------------------------------------------------------------------
// Per-call API load generator.  `threads` OS threads call MetricSystem::Histogram(name, value) -- or the
// StartTimer/Stop pair of print_benchmark.go:62-64 -- in a tight loop, one call per sample, the way an instrumented
// service would.  Values and name indices come from the repo's synthetic streams (splitmix64 of
// the sample index, integer-only), so a checker can regenerate exactly what was fed.
namespace {
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
const unsigned char kStreamLExp[16] = {17, 18, 18, 19, 19, 19, 20, 20, 20, 20, 21, 21, 21, 22, 22, 23};
inline double stream_value(int kind, uint64_t seed, uint64_t i) {     // kinds 0 (U) and 1 (L)
    const uint64_t u = splitmix64(seed + i), mant = u & 0x000FFFFFFFFFFFFFull;
    const uint64_t bits = kind == 0 ? (((uint64_t)(1023 + (u >> 52) % 63) << 52) | mant)
                                    : (((uint64_t)(1023 + kStreamLExp[(u >> 52) & 15]) << 52) | mant);
    double d;
    memcpy(&d, &bits, 8);
    return d;
}
inline uint32_t stream_id(uint64_t seed, uint64_t i, uint32_t H) {    // ids kind 0: uniform
    const uint64_t u = splitmix64((seed ^ 0xA5A5A5A5DEADBEEFull) + i);
    return (uint32_t)((u & 0xFFFFFFFFu) % H);
}
}  // namespace

// Feeds samples [start, start + n) of stream `kind` through ms.Histogram(names[id_i], value_i) from `threads` threads
// (contiguous slices).  Every thread works in blocks of 1024 samples: it regenerates the block's (name index, value)
// pairs, then makes the 1024 Histogram() calls.  Returns the wall seconds of the whole run (generator included,
// thread start-up excluded); *call_seconds (optional) receives the largest per-thread time spent inside the call
// loops alone -- the cost of the API path without the synthetic generator.  dry != 0 skips the calls (generator only).
static double histogram_stream_impl(loghisto::MetricSystem *ms, const char *const *names, uint32_t n_names, int kind, uint64_t seed,
                                    uint64_t start, uint64_t n, unsigned threads, int dry, double *call_seconds) {
    using clk = std::chrono::steady_clock;
    std::vector<std::string> nm(names, names + n_names);
    std::atomic<unsigned> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> workers;
    std::vector<double> in_calls(threads, 0.0);
    std::atomic<uint64_t> sink{0};
    const uint64_t per = n / threads;
    for (unsigned t = 0; t < threads; t++)
        workers.emplace_back([&, t] {
            constexpr int B = 1024;
            double vals[B];
            uint32_t ids[B];
            const uint64_t a = start + per * t, b = (t + 1 == threads) ? start + n : a + per;
            double spent = 0;
            uint64_t acc = 0;
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            for (uint64_t i0 = a; i0 < b; i0 += B) {
                const int m = (int)std::min<uint64_t>(B, b - i0);
                for (int k = 0; k < m; k++) {
                    vals[k] = stream_value(kind, seed, i0 + k);
                    ids[k] = n_names == 1 ? 0u : stream_id(seed, i0 + k, n_names);
                }
                if (dry) { for (int k = 0; k < m; k++) acc += ids[k] + (uint64_t)vals[k]; continue; }
                const auto t0 = clk::now();
                for (int k = 0; k < m; k++) ms->Histogram(nm[ids[k]], vals[k]);
                spent += std::chrono::duration<double>(clk::now() - t0).count();
            }
            in_calls[t] = spent;
            sink.fetch_add(acc);
        });
    while (ready.load() < threads) std::this_thread::yield();
    const auto t0 = clk::now();
    go.store(true, std::memory_order_release);
    for (auto &w : workers) w.join();
    const double wall = std::chrono::duration<double>(clk::now() - t0).count();
    if (call_seconds) { double mx = 0; for (double x : in_calls) mx = std::max(mx, x); *call_seconds = mx; }
    return wall;
}

extern "C" __attribute__((visibility("default")))
double lhms_histogram_stream(void *msp, const char *const *names, uint32_t n_names, int kind, uint64_t seed, uint64_t start,
                             uint64_t n, unsigned threads) {
    auto *ms = static_cast<loghisto::MetricSystem *>(msp);
    if (!ms || !names || !n_names || !threads) return -1.0;
    return histogram_stream_impl(ms, names, n_names, kind, seed, start, n, threads, 0, nullptr);
}

extern "C" __attribute__((visibility("default")))
double lhms_histogram_stream2(void *msp, const char *const *names, uint32_t n_names, int kind, uint64_t seed, uint64_t start,
                              uint64_t n, unsigned threads, int dry, double *call_seconds) {
    auto *ms = static_cast<loghisto::MetricSystem *>(msp);
    if (!ms || !names || !n_names || !threads) return -1.0;
    return histogram_stream_impl(ms, names, n_names, kind, seed, start, n, threads, dry, call_seconds);
}

extern "C" __attribute__((visibility("default")))
// print_benchmark.go:59-67 as a measurement: `threads` threads loop { t := StartTimer(name); Stop() } for `seconds`
// against a running MetricSystem (reaper at `interval_ns`); returns calls per second summed over the threads and
// writes the total number of calls and the sum of every interval's <name>_count as the reaper reported them.
double lhms_timer_loop(const char *name, unsigned threads, double seconds, int64_t interval_ns, int device,
                       uint64_t *total_calls, double *reported_count) {
    loghisto::Options o;
    o.device = device;
    o.max_histograms = 16;
    o.max_counters = 16;
    try {
        loghisto::MetricSystem ms(std::chrono::nanoseconds(interval_ns), true, o);
        auto mc = std::make_shared<loghisto::Channel<std::shared_ptr<loghisto::ProcessedMetricSet>>>(64);
        ms.SubscribeToProcessedMetrics(mc);
        ms.Start();
        const std::string nm(name);
        std::atomic<bool> stop{false};
        std::atomic<uint64_t> calls{0};
        std::vector<std::thread> workers;
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned i = 0; i < threads; i++)
            workers.emplace_back([&] {
                while (!stop.load(std::memory_order_relaxed)) {
                    for (int k = 0; k < 256; k++) {
                        loghisto::TimerToken timer = ms.StartTimer(nm);
                        timer.Stop();
                    }
                    calls.fetch_add(256, std::memory_order_relaxed);
                }
            });
        // the rate is taken over the last 60 % of the run: the first part pins the staging slots (a one-time cost)
        std::this_thread::sleep_for(std::chrono::duration<double>(seconds * 0.4));
        const uint64_t c0 = calls.load();
        const auto t1 = std::chrono::steady_clock::now();
        std::this_thread::sleep_for(std::chrono::duration<double>(seconds * 0.6));
        const uint64_t c1 = calls.load();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
        stop.store(true);
        for (auto &w : workers) w.join();
        (void)t0;
        ms.Stop();
        // what the reaper delivered, plus one last collection for the tail of the run
        double reported = 0;
        std::shared_ptr<loghisto::ProcessedMetricSet> m;
        while (mc->Receive(&m, std::chrono::milliseconds(1))) {
            auto it = m->Metrics.find(nm + "_count");
            if (it != m->Metrics.end()) reported += it->second;
        }
        auto raw = ms.collectRawMetrics();
        auto last = ms.processMetrics(*raw);
        auto it = last->Metrics.find(nm + "_count");
        if (it != last->Metrics.end()) reported += it->second;
        if (total_calls) *total_calls = calls.load();
        if (reported_count) *reported_count = reported;
        return (double)(c1 - c0) / dt;
    } catch (const std::exception &e) {
        fprintf(stderr, "lhms_timer_loop: %s\n", e.what());
        return -1.0;
    }
}

extern "C" __attribute__((visibility("default")))
double lhms_print_benchmark(const char *name, unsigned concurrency, double seconds, int64_t interval_ns, int device, int print) {
    loghisto::Options o;
    o.device = device;
    o.max_histograms = 16;
    o.max_counters = 16;
    try {
        return loghisto::PrintBenchmark(name, concurrency, [] {}, seconds, std::chrono::nanoseconds(interval_ns), o, print != 0);
    } catch (const std::exception &e) {
        fprintf(stderr, "lhms_print_benchmark: %s\n", e.what());
        return -1.0;
    }
}
