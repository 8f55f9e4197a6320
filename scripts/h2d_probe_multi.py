#!/usr/bin/env python
"""Concurrent pinned host -> device copy bandwidth over k GPUs of the box (k = 1, 2, 4, 8): what bounds the e2e leg of the sharded
job (every rank pushes its compressed bytes over PCIe at the same time).  One process per GPU (like the job), all starting their
timed copies at an agreed wall-clock instant.  Prints per-GPU and aggregate GB/s per k, the PCIe / NUMA tree and the CPU affinity
hints of every GPU.   usage: h2d_probe_multi.py            (parent)   |   h2d_probe_multi.py --child GPU START_EPOCH"""
import os, subprocess, sys, time


def child(gpu, start):
    import torch
    torch.cuda.set_device(gpu)
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < start:
        pass
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(12):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    print("gpu %d: %.1f GB/s" % (gpu, 12 * n / (a.elapsed_time(b) * 1e-3) / 1e9), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "--child":
    child(int(sys.argv[2]), float(sys.argv[3]))
    sys.exit(0)
import torch
ng = torch.cuda.device_count()
print("GPUs:", ng, " host cores:", os.cpu_count())
for cmd in (["nvidia-smi", "topo", "-m"], ["bash", "-c", "lspci -tv 2>/dev/null | head -60"], ["bash", "-c", "numactl -H 2>/dev/null | head -20; for d in /sys/bus/pci/devices/*; do if [ -e $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo $(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist); fi; done | head -20"]):
    try:
        print(subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60).stdout)
    except Exception as e:
        print(cmd, "failed:", e)
for k in (1, 2, 4, 8):
    if k > ng:
        break
    start = time.time() + 14.0   # process start-up + pinning 1 GiB takes seconds
    ps = [subprocess.Popen([sys.executable, __file__, "--child", str(g), str(start)], stdout=subprocess.PIPE, text=True) for g in range(k)]
    outs = [p.communicate()[0].strip() for p in ps]
    rates = [float(o.split(":")[1].split()[0]) for o in outs if "GB/s" in o]
    print("k=%d concurrent: %s  -> aggregate %.1f GB/s, per GPU %.1f" % (k, "  ".join(outs), sum(rates), sum(rates) / max(len(rates), 1)), flush=True)
