#!/bin/bash
# Box probe: topology, PCIe link, host cores/RAM/NUMA, pin limits,
# practical pinned H2D/D2H ceiling, and whether vLLM's own allocator loads.
O=gpurun_out/probe; mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max,memory.total --format=csv > $O/pcie.csv 2>&1
lscpu > $O/lscpu.txt 2>&1
nproc > $O/nproc.txt; free -g > $O/free.txt; ulimit -a > $O/ulimit.txt
cat /proc/meminfo > $O/meminfo.txt
ls /sys/devices/system/node/ > $O/numa_nodes.txt 2>&1
for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor 2>/dev/null); if [ "$v" = "0x10de" ]; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done > $O/gpu_numa.txt 2>&1
cat /sys/kernel/mm/transparent_hugepage/enabled > $O/thp.txt 2>&1
cat /proc/self/status | grep -i -E 'cpus_allowed|mems_allowed' > $O/affinity.txt
cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/cpu.max > $O/cgroup.txt 2>&1
python - > $O/torch_bw.txt 2>&1 <<'PY'
import torch, time
print(torch.cuda.get_device_name(0), torch.cuda.device_count())
free, tot = torch.cuda.mem_get_info(); print("mem", free, tot)
t0=time.time(); h = torch.empty(4<<30, dtype=torch.uint8, pin_memory=True); print("pin 4GiB s", time.time()-t0)
d = torch.empty(4<<30, dtype=torch.uint8, device="cuda")
for sz in [1<<20, 2<<20, 8<<20, 32<<20, 128<<20, 1<<30, 4<<30]:
    for name,(dst,src) in {"h2d":(d,h),"d2h":(h,d)}.items():
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        n = max(1,(2<<30)//sz)
        dst[:sz].copy_(src[:sz], non_blocking=True); torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            o=(i*sz)%((4<<30)-sz+1)
            dst[o:o+sz].copy_(src[o:o+sz], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1); print(f"{name} chunk={sz>>20}MiB n={n} GB/s={n*sz/ms/1e6:.2f}")
# two streams, both directions simultaneously
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
h2=torch.empty(2<<30, dtype=torch.uint8, pin_memory=True); d2=torch.empty(2<<30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0=time.time()
with torch.cuda.stream(s1): d[:2<<30].copy_(h[:2<<30], non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt=time.time()-t0; print("bidir 2GiB each: s", dt, "GB/s per dir", (2<<30)/dt/1e9)
PY
python - > $O/vllm_cumem.txt 2>&1 <<'PY'
import time; t0=time.time()
import torch
from vllm.device_allocator.cumem import CuMemAllocator, cumem_available
print("cumem_available", cumem_available, "import s", time.time()-t0)
a = CuMemAllocator.get_instance()
with a.use_memory_pool("weights"):
    xs=[torch.full((64<<20,), i+1, dtype=torch.uint8, device="cuda") for i in range(8)]
torch.cuda.synchronize()
print("usage", a.get_current_usage(), "nseg", len(a.pointer_to_data))
ptrs=[x.data_ptr() for x in xs]
t0=time.time(); a.sleep(offload_tags=("weights",)); torch.cuda.synchronize(); print("sleep s", time.time()-t0)
t0=time.time(); a.wake_up(); torch.cuda.synchronize(); print("wake s", time.time()-t0)
print("ok", all(int(x[0])==i+1 and int(x[-1])==i+1 for i,x in enumerate(xs)), [x.data_ptr() for x in xs]==ptrs)
PY
echo done > $O/done.txt
