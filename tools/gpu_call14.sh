#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/keyed_pf_probe.py wc_spt=6,4 wc_pf=0,1,2,4 > gpurun_out/keyed_pf_probe_r02p.txt 2>&1
cat gpurun_out/keyed_pf_probe_r02p.txt
