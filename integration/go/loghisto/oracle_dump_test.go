// UNVERIFIED SOURCE (no Go toolchain in the build image).  Run inside the reference package on any machine
// with Go (amd64, GOAMD64=v1) to close the oracle loop: it prints the bucket counts of the same seeded
// synthetic streams oracle/loghisto_oracle.c generates, in the format tests/golden/README.md describes.
package loghisto

import (
	"fmt"
	"math"
	"sort"
	"testing"
)

func splitmix64(x uint64) uint64 {
	x += 0x9E3779B97F4A7C15
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9
	x = (x ^ (x >> 27)) * 0x94D049BB133111EB
	return x ^ (x >> 31)
}

func TestDumpOracleStreamU(t *testing.T) {
	const seed, n = 0x10C415C0, 1000000
	counts := map[int16]uint64{}
	for i := uint64(0); i < n; i++ {
		u := splitmix64(seed + i)
		bits := (uint64(1023+(u>>52)%63) << 52) | (u & 0x000FFFFFFFFFFFFF)
		counts[compress(math.Float64frombits(bits))]++
	}
	keys := make([]int, 0, len(counts))
	for k := range counts {
		keys = append(keys, int(k))
	}
	sort.Ints(keys)
	for _, k := range keys {
		fmt.Printf("%d %d\n", k, counts[int16(k)])
	}
}
