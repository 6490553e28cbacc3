import torch, time, subprocess, json, sys
torch.cuda.init()
x = torch.empty(1 << 29, dtype=torch.float64, device="cuda")   # 4 GB
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("copy GB/s", 2 * x.numel() * 8 / ms / 1e6)
e0.record()
for _ in range(10): y.zero_()
e1.record(); torch.cuda.synchronize()
print("memset GB/s", x.numel() * 8 / (e0.elapsed_time(e1) / 10) / 1e6)
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.mem,clocks.max.mem,power.draw,power.limit,temperature.gpu,ecc.mode.current", "--format=csv"], capture_output=True, text=True).stdout)
