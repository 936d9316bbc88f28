import os, time, torch, subprocess
print(subprocess.run("nvidia-smi topo -m | head -20; lscpu | grep -i -E 'numa|socket|model name|^CPU.s.'; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo $d $(cat $d/numa_node) $(cat $d/current_link_speed) $(cat $d/current_link_width); fi; done; nproc; cat /proc/self/status | grep -i allowed", shell=True, capture_output=True, text=True).stdout)
def bw(tag):
    x = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
    h = torch.empty(256<<20, dtype=torch.uint8).pin_memory()
    for _ in range(2): h.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(10): h.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    dt=time.perf_counter()-t
    print(tag, "D2H GB/s", 10*(256<<20)/dt/1e9)
    t=time.perf_counter()
    for _ in range(10): x.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt=time.perf_counter()-t
    print(tag, "H2D GB/s", 10*(256<<20)/dt/1e9)
bw("default affinity")
for node in sorted(os.listdir("/sys/devices/system/node")):
    if not node.startswith("node"): continue
    cl=open(f"/sys/devices/system/node/{node}/cpulist").read().strip()
    cpus=set()
    for part in cl.split(","):
        a,_,b=part.partition("-"); cpus.update(range(int(a), int(b or a)+1))
    allowed=os.sched_getaffinity(0)
    use=cpus & allowed
    print(node, cl, "allowed", len(use))
    if use:
        old=os.sched_getaffinity(0)
        os.sched_setaffinity(0, use)
        bw(node)
        os.sched_setaffinity(0, old)
