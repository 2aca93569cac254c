import torch, time
x = torch.empty(39321600, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (1, 3):
    streams = [torch.cuda.Stream() for _ in range(n)]
    hs = [torch.empty(39321600, dtype=torch.uint8).pin_memory() for _ in range(n)]
    ds = [torch.empty(39321600, dtype=torch.uint8, device="cuda") for _ in range(n)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(30):
        with torch.cuda.stream(streams[i % n]):
            ds[i % n].copy_(hs[i % n], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D pinned, {n} stream(s): {30 * 39.3216 / dt / 1e3:.1f} GB/s ({dt / 30 * 1e3:.3f} ms per 39.3 MB batch)")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
