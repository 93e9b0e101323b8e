"""Host-to-device bandwidth of the box (pinned and pageable), for reading bench.py's e2e numbers."""
import torch, time, numpy as np
a = torch.empty(100_000_000, dtype=torch.float64).pin_memory()
d = torch.empty_like(a, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("pinned H2D GB/s", 0.8 / dt)
b = np.empty(100_000_000)
t = time.perf_counter(); d.copy_(torch.from_numpy(b)); torch.cuda.synchronize(); print("pageable H2D GB/s", 0.8 / (time.perf_counter() - t))
import subprocess; print(subprocess.run("nvidia-smi topo -m | head -4; lscpu | grep -i -e numa -e 'model name' | head -5", shell=True, capture_output=True, text=True).stdout)
