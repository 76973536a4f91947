"""BASELINE configs 3 and 4 at full size (CLI around snprelate_b200.configs.run_pair_config):

    python tools/config_run.py --est ibs  --samples 50000 --snps 500000            # config 3, 1 GPU
    torchrun --nproc-per-node 4 tools/config_run.py --est ibs --samples 50000 --snps 500000
    torchrun --nproc-per-node 8 tools/config_run.py --est king --samples 100000 --snps 1000000   # config 4

Prints one JSON line on rank 0 with pair-SNPs/s, the phases, the algorithmic HBM bytes and an oracle check at
scattered samples.  (Option names avoid prefixes of torchrun's own options: argparse abbreviation matching
would otherwise swallow them.)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snprelate_b200.configs import run_pair_config

ap = argparse.ArgumentParser()
ap.add_argument("--est", default="ibs", choices=["ibs", "king"])
ap.add_argument("--samples", dest="n", type=int, default=50000)
ap.add_argument("--snps", dest="m", type=int, default=500000)
ap.add_argument("--rows", type=int, default=-1, help="row-window height (multiple of 256), 0 = whole matrix, -1 = auto")
ap.add_argument("--miss", type=float, default=0.005)
ap.add_argument("--check", type=int, default=24, help="oracle check at K scattered samples")
ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"], help="the library's peer-memory reduction or NCCL")
ap.add_argument("--reduce-everywhere", action="store_true", help="all-reduce instead of reducing to the finishing rank")
ap.add_argument("--engine", default="bits", choices=["bits", "tensor"], help="pair-counter engine (snprel_set_count_engine)")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as td
    td.init_process_group("nccl", device_id=torch.device("cuda", local))

from oracle import snprel_oracle as O        # checker only
idx = O.scattered_samples(args.n, args.check, seed=7) if args.check > 0 else None
res, kept = run_pair_config(local, args.est, args.n, args.m, rank, world, args.engine, args.reduce, args.rows, args.miss,
                            idx, not args.reduce_everywhere)
if idx is not None:
    err, cnt = O.check_pair_rows(args.est, idx, args.m, kept, miss_rate=args.miss)
    if world > 1:
        t = torch.tensor([err], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        c = torch.tensor([float(cnt)], dtype=torch.float64, device="cuda")
        td.all_reduce(c)
        err, cnt = float(t[0]), int(c[0])
    res["parity"] = {"scattered_samples": int(len(idx)), "entries_checked": cnt, "max_abs_err": err, "ok": bool(err < 1e-12)}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    td.destroy_process_group()
