import torch, time
x = torch.empty(8*1024*1024*16, dtype=torch.float16, device="cuda")
y = torch.empty_like(x)
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
nb = x.numel() * 2
us = t(lambda: x.fill_(1.0)); print(f"fill 268MB: {us:.1f} us  {nb/us/1e6:.2f} TB/s write")
us = t(lambda: y.copy_(x)); print(f"copy 268MB: {us:.1f} us  {2*nb/us/1e6:.2f} TB/s r+w")
us = t(lambda: x.sum()); print(f"read 268MB: {us:.1f} us  {nb/us/1e6:.2f} TB/s read")
big = torch.empty(1<<30, dtype=torch.float16, device="cuda")
us = t(lambda: big.fill_(1.0), 5); print(f"fill 2GB: {us:.1f} us  {big.numel()*2/us/1e6:.2f} TB/s write")
