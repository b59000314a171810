"""CPU simulation of the seq_fast.cu generation schedule on RTS-79 (no GPU, no oracle): how many waves / jobs per
year a choice of static blocks per unit and of the wave-fill policy costs.  Durations are numpy exponentials (the
schedule statistics do not depend on the sampler's exact bits).  Used to pick static_blocks = 3 and to rule out
per-unit static tables and wider wave fills.  Cost model from the SASS of the r01j build:
255 warp-instructions per static block, 381 per wave."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powersystemsreliabilityassessment_b200 import rts79
cap, mttf, mttr = rts79.units()
U = len(cap); H = 8736
rng = np.random.default_rng(1)
FOR = mttr / (mttf + mttr)
ispan = 0.5 / (mttf + mttr)

def gen_year():
    """returns per unit the cumulative time after each block (list of arrays)"""
    out = []
    for u in range(U):
        s0up = rng.random() >= FOR[u]
        # draws: 0 = init (no duration), then alternate: first duration is of state s0
        t = 0.0; ends = []
        nblk = 0
        while nblk < 40:
            # block of 4 draws; draw j overall index = 4*nblk + q ; draw 0 is init
            for q in range(4):
                j = 4 * nblk + q
                if j == 0: continue
                # state entered: durations alternate starting with s0 state
                up = s0up if (j % 2 == 1) else (not s0up)
                m = mttf[u] if up else mttr[u]
                t += rng.exponential(m)
            ends.append(t); nblk += 1
        out.append(np.array(ends))
    return out

def simulate(S, years=1500, nbmax=4, verbose=False):
    """S[u] = static blocks of unit u. returns avg waves, avg wave jobs, avg static wasted"""
    waves = 0; jobs = 0; needed = 0
    for y in range(years):
        ends = gen_year()
        nb = np.array(S).copy()
        need = np.array([int(np.searchsorted(ends[u], H, side='right')) + 1 for u in range(U)])  # blocks until t > H
        needed += need.sum()
        while True:
            tl = np.array([ends[u][nb[u] - 1] for u in range(U)])
            short = tl <= H
            if not short.any(): break
            rem = np.floor(np.maximum(H - tl, 0))
            want = np.minimum(nbmax, 1 + (rem * ispan).astype(int))
            n_m = np.where(short, want, 0)
            J1 = n_m.sum()
            if J1 < 32:
                extra = short & (want < nbmax)
                if J1 + extra.sum() <= 32:
                    n_m = n_m + extra
            off = np.concatenate([[0], np.cumsum(n_m)[:-1]])
            n_u = np.maximum(0, np.minimum(n_m, 32 - off))
            nb += n_u
            jobs += n_u.sum(); waves += 1
    return waves / years, jobs / years, needed / years


if __name__ == "__main__":
    years = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    allends = [gen_year() for _ in range(years)]
    def run(S, rounds=1, nbmax=4, nbtot=5, alpha=1.0, partial=False):
        waves = 0; jobs = 0
        hist = {}
        for y in range(years):
            ends = allends[y]
            nb = np.array(S).copy(); w = 0
            while True:
                tl = np.array([ends[u][nb[u] - 1] for u in range(U)])
                short = tl <= H
                if not short.any(): break
                rem = np.floor(np.maximum(H - tl, 0))
                want = np.minimum(nbmax, 1 + (rem * ispan * alpha).astype(int))
                n_m = np.where(short, want, 0)
                for r in range(rounds):
                    J1 = n_m.sum()
                    if J1 >= 32: break
                    extra = short & (n_m < nbtot)
                    if J1 + extra.sum() <= 32:
                        n_m = n_m + extra
                    elif partial:
                        # give +1 to the first (32 - J1) candidates in lane order
                        idx = np.flatnonzero(extra)[: 32 - J1]
                        n_m[idx] += 1
                        break
                    else:
                        break
                off = np.concatenate([[0], np.cumsum(n_m)[:-1]])
                n_u = np.maximum(0, np.minimum(n_m, 32 - off))
                nb += n_u
                jobs += n_u.sum(); waves += 1; w += 1
            hist[w] = hist.get(w, 0) + 1
        return round(waves / years, 3), round(jobs / years, 1), dict(sorted(hist.items()))
    S3 = [3] * U
    print("current          ", run(S3))
    print("2 rounds         ", run(S3, rounds=2, nbtot=6))
    print("3 rounds         ", run(S3, rounds=3, nbtot=7))
    print("3 rounds partial ", run(S3, rounds=3, nbtot=7, partial=True))
    print("6 rounds partial ", run(S3, rounds=6, nbtot=7, partial=True))
    print("alpha1.5 r1      ", run(S3, alpha=1.5))
    print("alpha1.5 r3p nb7 ", run(S3, alpha=1.5, rounds=3, nbmax=6, nbtot=7, partial=True))
    print("alpha2 r3p nb7   ", run(S3, alpha=2.0, rounds=3, nbmax=6, nbtot=7, partial=True))
    print("nbmax7 r3p       ", run(S3, rounds=3, nbmax=7, nbtot=7, partial=True))
    S2 = [2] * U
    print("static2 nbmax7 r3p", run(S2, rounds=3, nbmax=7, nbtot=7, partial=True))
    S4 = [4] * U
    print("static4 current   ", run(S4))
    print("static4 r3p nb7   ", run(S4, rounds=3, nbmax=7, nbtot=7, partial=True))
