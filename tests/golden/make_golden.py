"""Regenerates the committed golden fixtures from the CPU oracle.

The Go reference cannot run in the build image (no Go toolchain), so these vectors come from the oracle's
restatement; integration/go/loghisto/oracle_dump_test.go prints the same file from the real compress() on any
machine with Go, which is how the loop gets closed.  Format of *.counts: "<int16 key> <count>" per line,
ascending key; header lines start with '#'.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as o

N = 1_000_000
PS = [0.0, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999, 0.9999, 1.0]


def main():
    o.build()
    summary = {}
    for name, kind in (("U", o.STREAM_U), ("L", o.STREAM_L), ("S", o.STREAM_S)):
        vals = o.gen_stream(kind, N, o.DEFAULT_SEED)
        counts = o.ingest(vals)
        with open(os.path.join(HERE, "stream_%s_1e6.counts" % name), "w") as f:
            f.write("# stream %s, seed 0x%X, n %d; key count\n" % (name, o.DEFAULT_SEED, N))
            for key in range(-32768, 32768):
                c = int(counts[key & 0xFFFF])
                if c:
                    f.write("%d %d\n" % (key, c))
        ref = o.process_histogram(counts, PS)
        summary[name] = {"count": ref["total"], "sum_hex": float(ref["sum"]).hex(), "avg_hex": float(ref["avg"]).hex(),
                         "pkeys": [int(k) for k in ref["pkeys"]], "pvals_hex": [float(v).hex() for v in ref["pvals"]],
                         "first_values_hex": [float(v).hex() for v in vals[:4]]}
    with open(os.path.join(HERE, "stream_summaries.json"), "w") as f:
        json.dump({"percentiles": PS, "streams": summary}, f, indent=1)
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    main()
