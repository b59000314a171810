"""world_size-2 gloo run of the multi-GPU host logic on CPU: contiguous year sharding + one all-reduce
of the integer accumulators reproduces the single-process sums."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from powersystemsreliabilityassessment_b200 import sharding, indices_from_raw
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    total = 1000
    rng = np.random.default_rng(0)                       # same synthetic per-year vectors on every rank
    lol = rng.integers(0, 30, total) * (rng.random(total) < 0.5)
    ens = lol * rng.integers(1, 10**12, total)           # large enough that ENS^2 sums exceed 64 bits
    ent = np.minimum(lol, rng.integers(0, 4, total))
    a, b = sharding.shard_range(total, rank, world, 10)
    def raw_of(sl):
        return dict(years=len(lol[sl]), sum_lol_hours=int(lol[sl].sum()), sum_ens_fp=int(ens[sl].sum()),
                    sum_entries=int(ent[sl].sum()), years_with_loss=int((lol[sl] > 0).sum()),
                    sum_lol_sq=int((lol[sl].astype(object) ** 2).sum()),
                    sum_ens_sq=int((ens[sl].astype(object) ** 2).sum()), events=int(lol[sl].sum()) * 3)
    red = sharding.allreduce_raw(raw_of(slice(a, b)))
    full = raw_of(slice(0, total))
    assert red == full, (red, full)
    assert full["sum_ens_sq"] > 2**64
    r = indices_from_raw(red)
    assert r.years == total and abs(r.lole - lol.mean()) < 1e-12
    # per-hour failure counts / histogram bins: element-wise sums; per-year vectors: concatenation in year order
    fc = np.bincount(rng.integers(0, 48, 500)[rank::world], minlength=48)
    tot = sharding.allreduce_counts(fc)
    rng2 = np.random.default_rng(0)                      # replay the generator up to the draw of the 500 hours
    rng2.integers(0, 30, total); rng2.random(total); rng2.integers(1, 10**12, total); rng2.integers(0, 4, total)
    assert tot.sum() == 500 and np.array_equal(tot, np.bincount(rng2.integers(0, 48, 500), minlength=48))
    a2, b2 = sharding.shard_range(total, rank, world, 10)
    assert np.array_equal(sharding.gather_years(ens[a2:b2]), ens)
    assert np.array_equal(sharding.gather_years(lol[a2:b2].astype(np.uint32)), lol.astype(np.uint32))
    # PSA.jl:263-265 running mean every 10 years from the ranks' per-group sums
    glocal = lol[a2:b2].reshape(-1, 10).sum(axis=1)
    hist = sharding.merged_history(glocal, 10)
    assert np.allclose(hist, np.cumsum(lol)[9::10] / np.arange(10, total + 1, 10), rtol=1e-14)
    if rank == 0:
        print("GLOO_OK", json.dumps({"world": world, "span": [a, b]}))
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_gloo_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout
