"""Where the time of the streamed ingest goes (diagnostic, not a test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import snprelate_b200 as S

N, M = 10000, 1000000
c = S.Context(0)
c.geno_begin(N, M); c.geno_synth(M)
rb = (N + 255) // 256 * 256 // 4
host = torch.empty((M, rb), dtype=torch.uint8, pin_memory=True)
c.geno_copy_2b(host.numpy())
out = torch.empty((N, N), dtype=torch.float64, pin_memory=True)
hg, ho = host.numpy(), out.numpy()

def t(fn, reps=3):
    r = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); r.append((time.perf_counter() - t0) * 1e3)
    return [round(x, 1) for x in r]

def sync_path():
    c.geno_begin(N, M); c.geno_push_2b(hg); c.pca(genmat_only=True, genmat_out=ho)
def copy_only_async():
    c.geno_begin(N, M); c.geno_push_2b_async(hg); c.geno_wait()
def copy_only_sync():
    c.geno_begin(N, M); c.geno_push_2b(hg)
def streamed():
    c.geno_begin(N, M); c.geno_push_2b_async(hg); c.pca(genmat_only=True, genmat_out=ho)
def async_wait_then_pca():
    c.geno_begin(N, M); c.geno_push_2b_async(hg); c.geno_wait(); c.pca(genmat_only=True, genmat_out=ho)

print("copy only, blocking push ", t(copy_only_sync))
print("copy only, async + wait  ", t(copy_only_async))
print("blocking path            ", t(sync_path))
print("async, wait, then pca    ", t(async_wait_then_pca))
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c.geno_begin(N, M); c.geno_push_2b_async(hg); t1 = time.perf_counter()
    c.pca(genmat_only=True, genmat_out=ho); t2 = time.perf_counter()
    print(f"streamed: push returns after {1e3*(t1-t0):.2f} ms, pca {1e3*(t2-t1):.1f} ms; device step {c.last_step_ms():.1f} ms, "
          f"first tensor pass -> last {c.last_hot_kernel()[0]:.1f} ms in {c.last_hot_kernel()[1]} launches, copies {c.stream_last_copy_ms():.1f} ms, stats {c.stream_stats()}")
# pca only on resident data, for reference
c.geno_begin(N, M); c.geno_push_2b(hg)
print("pca on resident data     ", t(lambda: c.pca(genmat_only=True, genmat_out=ho)))
