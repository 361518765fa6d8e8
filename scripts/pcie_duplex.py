import torch, time
n = 280_000_000 // 4
h1 = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d1 = torch.empty(n, dtype=torch.float32, device="cuda"); d2 = torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
a = torch.randn(8192, 8192, device="cuda")
def d2h_and_kernel():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    with torch.cuda.stream(s1):
        for _ in range(20): b = a @ a
def kernel():
    with torch.cuda.stream(s1):
        for _ in range(20): b = a @ a
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both), ("kernel", kernel), ("d2h+kernel", d2h_and_kernel)):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, round(1e3 * (time.perf_counter() - t0) / 5, 2), "ms")
