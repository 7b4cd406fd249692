import torch, time
n = 1 << 31
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
for name, fn in [("h2d", h2d), ("d2h", d2h), ("both", both)]:
    dt = t(fn)
    print(name, "%.1f ms  %.1f GB/s per direction" % (dt * 1e3, n / dt / 1e9))
