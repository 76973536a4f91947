"""Where the end-to-end step (host 2-bit rows -> covariance on the host) spends its time:
wall clock of every C-ABI call of bench.py's e2e_step, bench size by default."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snprelate_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = S.Context(0)
ctx.geno_begin(n, m)
ctx.geno_synth(m, miss_rate=0.005)
rb = (n + 255) // 256 * 256 // 4
host_geno = torch.empty((m, rb), dtype=torch.uint8, pin_memory=True)
ctx.geno_copy_2b(host_geno.numpy())
host_out = torch.empty((n, n), dtype=torch.float64, pin_memory=True)


def timed(name, fn, acc):
    t0 = time.perf_counter()
    r = fn()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return r


for rep in range(reps + 1):
    acc = {}
    t0 = time.perf_counter()
    timed("geno_begin", lambda: ctx.geno_begin(n, m), acc)
    timed("push_2b", lambda: ctx.geno_push_2b(host_geno.numpy()), acc)
    timed("plan_local", lambda: ctx.plan_local(0), acc)
    ctx.invalidate()
    timed("pca(genmat)", lambda: ctx.pca(genmat_only=True, genmat_out=host_out.numpy()), acc)
    total = (time.perf_counter() - t0) * 1e3
    if rep:
        print(f"rep {rep}: total {total:.1f} ms (incl. one extra plan_local) | " +
              " | ".join(f"{k} {v:.1f}" for k, v in acc.items()) +
              f" | device step {ctx.last_step_ms():.1f} hot {ctx.last_hot_kernel()[0]:.1f}", flush=True)
