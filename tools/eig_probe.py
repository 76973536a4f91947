import time, torch
n = 10000
torch.manual_seed(0)
a = torch.randn(n, 2000, dtype=torch.float64, device="cuda")
c = a @ a.T
torch.cuda.synchronize()
for name, fn in (("eigh", lambda: torch.linalg.eigh(c)), ("eigvalsh", lambda: torch.linalg.eigvalsh(c))):
    t0 = time.time(); r = fn(); torch.cuda.synchronize(); print(name, time.time() - t0, flush=True)
# subspace iteration: top-32 with 64-wide block
k, b = 32, 96
t0 = time.time()
q = torch.linalg.qr(torch.randn(n, b, dtype=torch.float64, device="cuda"))[0]
for it in range(60):
    q = torch.linalg.qr(c @ q)[0]
t = q.T @ (c @ q)
w, v = torch.linalg.eigh(t)
torch.cuda.synchronize(); print("subspace 60 it", time.time() - t0)
wr = torch.linalg.eigvalsh(c)[-k:]
print("eigval err", (w[-k:] - wr).abs().max().item() / wr.max().item())
t0 = time.time(); y = c @ q; torch.cuda.synchronize(); print("one block matvec", time.time() - t0)
