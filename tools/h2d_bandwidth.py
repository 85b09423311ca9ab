import torch, time
x = torch.empty(20 * 1024 * 1024, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (1, 4):
    chunks = x.chunk(n); dch = d.chunk(n)
    for _ in range(3):
        for a, b in zip(chunks, dch): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        for a, b in zip(chunks, dch): b.copy_(a, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"pinned H2D 20 MiB in {n} copies: {ms:.3f} ms = {20 * 1.048576 / ms:.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max", "--format=csv"], capture_output=True, text=True).stdout)
