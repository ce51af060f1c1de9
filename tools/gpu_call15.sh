#!/bin/bash
mkdir -p gpurun_out
{
timeout 100 python tools/keyed_pf_probe.py wc_spt=6,3,5 wc_pf=1,2
timeout 100 python tools/keyed_pf_probe.py wc_spt=5 wc_pf=0
timeout 100 python tools/keyed_pf_probe.py wc_flush=12288,16384,24576
timeout 100 python tools/keyed_pf_probe.py kp_chunk=16777216,33554432,67108864
} > gpurun_out/keyed_pf_probe_r02q.txt 2>&1
cat gpurun_out/keyed_pf_probe_r02q.txt
