#!/bin/bash
# 1 GPU: ncu full-set capture of the default write-combining kernel at the bench's batch size (traffic.json c3)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_wc -s 1 -c 1 -o gpurun_out/prof_kwc_r02n \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-parity --no-e2e --no-cpu-baseline --no-api > gpurun_out/ncu_kwc_r02n.log 2>&1
tail -2 gpurun_out/ncu_kwc_r02n.log | cut -c1-200
