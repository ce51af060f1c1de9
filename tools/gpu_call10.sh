#!/bin/bash
# 1 GPU: tile-shape sweep of the write-combining kernel, batch-size probe, GPU tests of the keyed paths
mkdir -p gpurun_out
timeout 600 python tools/keyed_sweep.py 1000000000 1024 quick > gpurun_out/keyed_sweep_r02l.txt 2>&1
cut -c1-165 gpurun_out/keyed_sweep_r02l.txt
timeout 300 python tools/keyed_batch_probe.py > gpurun_out/keyed_batch_probe_r02l.txt 2>&1
cat gpurun_out/keyed_batch_probe_r02l.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "keyed or wc or mixed" > gpurun_out/pytest_keyed_r02l.txt 2>&1
tail -4 gpurun_out/pytest_keyed_r02l.txt
