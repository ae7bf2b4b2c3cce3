import torch, time
n = 26214400 // 4
h1 = torch.empty(n).pin_memory(); h2 = torch.empty(n).pin_memory()
d1 = torch.empty(n, device="cuda"); d2 = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("H2D 26MB ms", t(lambda: d1.copy_(h1, non_blocking=True)))
print("D2H 26MB ms", t(lambda: h2.copy_(d2, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("H2D+D2H concurrent ms", t(both))
