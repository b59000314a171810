"""N-rank check on real GPUs (torchrun, NCCL): contiguous year shards + the collectives of sharding.py reproduce the
single-GPU run of the same experiment -- accumulators, per-hour failure counts, convergence history, VaR / CVaR."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from powersystemsreliabilityassessment_b200 import Engine, rts79, sharding, indices_from_raw

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
cap, mttf, mttr = rts79.units(); load = rts79.load_curve_int()
total = 2_000_000
a, b = sharding.shard_range(total, rank, world, 10)
with Engine(device=local) as e:
    e.set_system(cap, mttf, mttr); e.set_load(load)
    r = e.seq_mc(b - a, seed=99, year0=a, per_year=True, fail_count=True, group=10)
    red = sharding.allreduce_raw(r.raw, device=dev)
    fail = sharding.allreduce_counts(r.fail_count, device=dev)
    hist = sharding.merged_history(r.group_lol, 10, device=dev)
    tail = sharding.tail_all_ranks(e, r.raw["ens_fp_vector"], device=dev)
    if rank == 0:
        full = e.seq_mc(total, seed=99, per_year=True, fail_count=True, history=10, keep_on_device=True)
        ref_tail = e.tail(None, alphas=(0.95, 0.99))
        for k in ("sum_lol_hours", "sum_ens_fp", "sum_entries", "sum_lol_sq", "sum_ens_sq", "years_with_loss", "events", "years"):
            assert red[k] == full.raw[k], k
        assert np.array_equal(fail, full.fail_count.astype(np.int64))
        assert np.allclose(hist, full.history, rtol=1e-13, atol=0)
        for t, u in zip(tail, ref_tail):
            assert t["var"] == u["var"] and t["cvar"] == u["cvar"] and t["n_tail"] == u["n_tail"]
        idx = indices_from_raw(red)
        print(f"MULTI_RANK_OK world={world} years={idx.years} LOLE={idx.lole:.4f} EENS={idx.eens:.2f} VaR95={tail[0]['var']:.0f} CVaR95={tail[0]['cvar']:.1f}")
dist.barrier()
dist.destroy_process_group()
