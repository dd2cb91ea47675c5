"""Where do pinned host buffers land (NUMA node) and what does the link give for them?  Repeated allocations of
the two 1.9 GB blocks of the e2e step; duplex copy rate and the node distribution from /proc/self/numa_maps."""
import os, re, subprocess, sys, time
import torch
print(subprocess.run("lscpu | grep -i 'numa\\|socket\\|model name'; cat /sys/devices/system/node/node*/cpulist 2>/dev/null", shell=True,
                     capture_output=True, text=True).stdout)
print("affinity", sorted(os.sched_getaffinity(0)))
dev = torch.device("cuda:0")
n = 231376 * 512
d1 = torch.empty(n, dtype=torch.complex128, device=dev); d2 = torch.empty(n, dtype=torch.complex128, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def nodes(t):
    addr = t.data_ptr()
    for line in open("/proc/self/numa_maps"):
        a = int(line.split()[0], 16)
        if a <= addr < a + t.numel() * 16 + (1 << 21) and ("N0=" in line or "N1=" in line):
            if abs(a - addr) < (1 << 30):
                return " ".join(re.findall(r"N\d+=\d+", line))
    return "?"
def rate(fn, reps=3):
    fn(); torch.cuda.synchronize(); t = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time() - t) / reps
for trial in range(5):
    a = torch.empty(n, dtype=torch.complex128).pin_memory(); b = torch.zeros(n, dtype=torch.complex128).pin_memory()
    gb = n * 16 / 1e9
    def both():
        with torch.cuda.stream(s1): d1.copy_(a, non_blocking=True)
        with torch.cuda.stream(s2): b.copy_(d2, non_blocking=True)
    h2d = gb / rate(lambda: d1.copy_(a, non_blocking=True)); d2h = gb / rate(lambda: b.copy_(d2, non_blocking=True))
    dup = 2 * gb / rate(both)
    print(f"trial {trial}: H2D {h2d:.1f} D2H {d2h:.1f} duplex {dup:.1f} GB/s | a on {nodes(a)} | b on {nodes(b)}", flush=True)
    del a, b
