// print_benchmark.cc -- C++ form of the reference's load generator (print_benchmark.go:49-106) over the host
// mirror: `concurrency` threads loop { StartTimer(name); op(); Stop() } against one MetricSystem with a 1 s (here:
// configurable) interval, a receiver prints the interesting keys of every ProcessedMetricSet.  Unlike the reference
// it stops after `seconds` and returns the last interval's <name>_count, so it can be used as a measurement.
#include <atomic>
#include <chrono>
#include <cstdio>
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
            while (!stop.load(std::memory_order_relaxed)) {
                TimerToken timer = ms.StartTimer(name);   // print_benchmark.go:62-64
                op();
                timer.Stop();
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
