// UNVERIFIED SOURCE: this image has no Go toolchain, so this file has never been
// compiled.  It shows the binding a loghisto maintainer would add: it replaces the
// bodies of Histogram / Counter / collectRawMetrics / processHistograms in
// metrics.go (same exported API, same RawMetricSet / ProcessedMetricSet types) with
// calls into libloghisto_b200.so.  Everything else in metrics.go (reaper,
// subscriptions, gauges, Submitter, serializers, PrintBenchmark) is unchanged.
//
// Build: CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<repo>/loghisto_b200 -lloghisto_b200" go build -tags b200

//go:build b200

package loghisto

/*
#include <stdint.h>
#include "loghisto_b200.h"
*/
import "C"

import (
	"fmt"
	"math"
	"runtime"
	"sync"
	"sync/atomic"
	"unsafe"
)

// b200Engine owns the device context, the name->id tables and the staging shards.
type b200Engine struct {
	ctx *C.lh_ctx

	histoMu  sync.RWMutex // same RLock fast path / Lock-and-recheck idiom as metrics.go:275-294
	histoIDs map[string]uint16
	histoNames []string

	counterMu  sync.RWMutex
	counterIDs map[string]uint16
	counterNames []string

	shards []*stagingShard // one per P; picked by a cheap per-goroutine hash
}

// stagingShard batches (id,value) pairs into a C-owned pinned slot (cgo forbids C code from keeping
// Go pointers past the call, so Go writes into C memory through unsafe.Slice).
type stagingShard struct {
	mu    sync.Mutex
	slot  C.lh_staging
	vals  []float64 // view of slot.host[0 : cap*8]
	ids   []uint16  // view of slot.host[idsOff : idsOff+cap*2]
	n     int
	cap   int
	idsOff uint64
}

const maxHistograms, maxCounters = 1024, 1024

func newB200Engine(device int) (*b200Engine, error) {
	cfg := C.lh_config{struct_size: C.uint32_t(unsafe.Sizeof(C.lh_config{})), device: C.int32_t(device),
		max_histograms: maxHistograms, max_counters: maxCounters}
	e := &b200Engine{histoIDs: map[string]uint16{}, counterIDs: map[string]uint16{}}
	if st := C.lh_create(&cfg, &e.ctx); st != C.LH_OK {
		return nil, fmt.Errorf("lh_create: %s", C.GoString(C.lh_strerror(st)))
	}
	e.shards = make([]*stagingShard, runtime.GOMAXPROCS(0))
	for i := range e.shards {
		e.shards[i] = &stagingShard{}
	}
	return e, nil
}

func (e *b200Engine) acquire(s *stagingShard) {
	C.lh_staging_acquire(e.ctx, &s.slot) // blocks only if every slot is still in flight
	s.cap = int(uint64(s.slot.bytes)/10) &^ 15
	s.idsOff = uint64(s.cap) * 8
	base := unsafe.Pointer(s.slot.host)
	s.vals = unsafe.Slice((*float64)(base), s.cap)
	s.ids = unsafe.Slice((*uint16)(unsafe.Add(base, s.idsOff)), s.cap)
	s.n = 0
}

func (e *b200Engine) flush(s *stagingShard) { // caller holds s.mu
	if s.n > 0 {
		C.lh_staging_commit_keyed_f64_u16(e.ctx, &s.slot, C.size_t(s.n), C.uint64_t(s.idsOff)) // ONE cgo call per batch
		s.vals, s.ids, s.n = nil, nil, 0
	}
}

func (e *b200Engine) histoID(name string) uint16 {
	e.histoMu.RLock()
	id, ok := e.histoIDs[name]
	e.histoMu.RUnlock()
	if ok {
		return id
	}
	e.histoMu.Lock()
	defer e.histoMu.Unlock()
	if id, ok = e.histoIDs[name]; !ok {
		id = uint16(len(e.histoNames))
		e.histoIDs[name] = id
		e.histoNames = append(e.histoNames, name)
	}
	return id
}

var shardPick uint32

// Histogram keeps metrics.go:273's signature and "never fails, never blocks on consumers" contract.
func (ms *MetricSystem) Histogram(name string, value float64) {
	e := ms.b200
	id := e.histoID(name)
	s := e.shards[atomic.AddUint32(&shardPick, 1)%uint32(len(e.shards))]
	s.mu.Lock()
	if s.vals == nil {
		e.acquire(s)
	}
	s.vals[s.n], s.ids[s.n] = value, id
	s.n++
	if s.n == s.cap {
		e.flush(s)
	}
	s.mu.Unlock()
}

// Counter keeps metrics.go:251's signature; amounts ride in a second staging slot per shard and are committed with
// lh_staging_commit_counter_u16 (uint64 amounts at offset 0, uint16 ids at idsOff), exactly like Histogram.
func (ms *MetricSystem) Counter(name string, amount uint64) {
	e := ms.b200
	id := e.counterID(name) // same interning idiom as histoID, over counterIDs / counterNames
	s := e.counterShards[atomic.AddUint32(&shardPick, 1)%uint32(len(e.counterShards))]
	s.mu.Lock()
	if s.amounts == nil {
		e.acquireCounter(s)
	}
	s.amounts[s.n], s.ids[s.n] = amount, id
	s.n++
	if s.n == s.cap {
		C.lh_staging_commit_counter_u16(e.ctx, &s.slot, C.size_t(s.n), C.uint64_t(s.idsOff))
		s.amounts, s.ids, s.n = nil, nil, 0
	}
	s.mu.Unlock()
}

// StartTimer / TimerToken.Stop (metrics.go:232-246) are unchanged Go: Stop() calls Histogram(name, float64(ns)).

// collectRawMetrics keeps metrics.go:420's contract: interval-delta histograms (absent when untouched),
// Rates = interval deltas, Counters = cumulative store.
func (ms *MetricSystem) collectRawMetrics() *RawMetricSet {
	e := ms.b200
	for _, s := range e.shards {
		s.mu.Lock()
		e.flush(s)
		s.mu.Unlock()
	}
	C.lh_snapshot_begin(e.ctx) // the cache swap of metrics.go:425-428 / 460-463
	var sp C.lh_sparse
	C.lh_snapshot_export(e.ctx, &sp)
	H := len(e.histoNames)
	offs := unsafe.Slice((*uint32)(unsafe.Pointer(sp.offsets)), maxHistograms+1)
	keys := unsafe.Slice((*int16)(unsafe.Pointer(sp.keys)), int(sp.total_entries))
	cnts := unsafe.Slice((*uint64)(unsafe.Pointer(sp.counts)), int(sp.total_entries))
	histograms := make(map[string]map[int16]*uint64)
	for h := 0; h < H; h++ {
		if offs[h] == offs[h+1] {
			continue // untouched this interval: absent, like a swapped-out empty cache
		}
		m := make(map[int16]*uint64, offs[h+1]-offs[h])
		for i := offs[h]; i < offs[h+1]; i++ {
			c := cnts[i]
			m[keys[i]] = &c
		}
		histograms[e.histoNames[h]] = m
	}
	// ... counter deltas from sp.counter_deltas -> rates; fold into ms.counterStore exactly as metrics.go:435-458
	// ... gauges, normalized timestamp: unchanged Go code
	// processMetrics may call lh_snapshot_reduce on the same frozen buffers; lh_snapshot_end releases them.
	return &RawMetricSet{Histograms: histograms /* Time, Counters, Rates, Gauges as before */}
}

// processHistograms keeps metrics.go:336's signature and output keys; the numbers come from lh_snapshot_reduce,
// which collectRawMetrics called once for the whole snapshot (results cached per histogram id in ms.b200.reduced).
func (ms *MetricSystem) processHistograms(name string, valuesToCounts map[int16]*uint64) map[string]float64 {
	e := ms.b200
	r := e.reduced[name] // {count uint64; sum, avg float64; pkeys []int32; pvals []float64}
	output := map[string]float64{
		fmt.Sprintf("%s_count", name): float64(r.count),
		fmt.Sprintf("%s_sum", name):   r.sum,
		fmt.Sprintf("%s_avg", name):   r.avg,
	}
	// aggregate store: unchanged Go (metrics.go:359-376), including uint64(totalSum)
	ms.addToHistogramCountStore(name, uint64(r.sum), r.count)
	i := 0
	for label := range ms.percentiles { // e.percentileOrder fixes the label -> column mapping used at reduce time
		if r.pkeys[i] != math.MinInt32 { // percentile() error (p > 1, NaN): logged and omitted, metrics.go:380-382
			output[fmt.Sprintf(label, name)] = r.pvals[i]
		}
		i++
	}
	return output
}

// In words: processHistograms (metrics.go:336) becomes a lookup into the arrays lh_snapshot_reduce filled:
// <name>_count, _sum, _avg and one entry per percentile label whose pkeys[] is not INT32_MIN
// (percentile()'s error case: key omitted, metrics.go:380-382).  The cumulative store update
// (uint64(totalSum), metrics.go:374) and the reaper's integer _agg_avg (metrics.go:601-606) stay in Go.
