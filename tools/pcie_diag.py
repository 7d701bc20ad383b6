import os, time, torch, subprocess
print(subprocess.run("nvidia-smi topo -m | head -8; lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'; nproc", shell=True, capture_output=True, text=True).stdout)
print("affinity", sorted(os.sched_getaffinity(0)))
prop=torch.cuda.get_device_properties(0)
bus=f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0" if hasattr(prop,'pci_bus_id') else None
print("bus", bus)
try:
    print("numa_node", open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip(), "local_cpulist", open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip())
except Exception as e: print("sysfs", e)
d=torch.empty(100*1024*1024//8, dtype=torch.float64, device='cuda')
for trial in range(3):
    h=torch.empty(100*1024*1024//8, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()
    for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(10): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
    t=time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt2=(time.perf_counter()-t)/10
    print("D2H GB/s", 0.1048576/dt, "H2D GB/s", 0.1048576/dt2)
