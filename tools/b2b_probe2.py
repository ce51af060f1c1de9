"""Which ingredient of bench.py slows K1 from 1.145 to 1.23 ms?"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

n = 1_000_000_000
eng = lh.Engine(device=0, max_histograms=1, max_counters=1)
d = eng.gen_stream(0, n, lh.DEFAULT_SEED)
eng.sync()
PS = [0.0, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999, 0.9999, 1.0]

def pipelined(tag, k=12, stream=None):
    seqs = []
    eng.ingest_f64(0, d, n, stream=stream); seqs.append(eng.ingest_seq())
    t0 = time.perf_counter()
    for i in range(k):
        eng.snapshot_begin(); h = eng.snapshot_reduce_async(PS); eng.snapshot_end()
        if i + 1 < k:
            eng.ingest_f64(0, d, n, stream=stream); seqs.append(eng.ingest_seq())
        eng.snapshot_result(h)
    wall = (time.perf_counter() - t0) / k * 1e3
    eng.sync()
    print("%-34s wall/step %.3f  kernels" % (tag, wall), " ".join("%.3f" % eng.kernel_ms(s) for s in seqs[-8:]), flush=True)

pipelined("plain")
pipelined("plain again")

# NVML polling thread
stop = False
def poll():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetPowerUsage(h)
        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(0.002)
t = threading.Thread(target=poll, daemon=True); t.start(); time.sleep(0.3)
pipelined("with NVML polling @2ms")
stop = True; t.join()
pipelined("NVML stopped")

import torch
torch.cuda.set_device(0)
x = torch.zeros(1, device="cuda")
pipelined("after torch CUDA init")
ext = torch.cuda.ExternalStream(eng.ingest_stream, device=0)
pipelined("ExternalStream(ctx stream)", stream=ext)
pipelined("torch default stream", stream=torch.cuda.current_stream())
