"""TEST INFRASTRUCTURE -- a second, deliberately naive restatement of the reference's Monte Carlo loops in pure Python,
written independently of oracle/psra_oracle.c (no shared code, no event form, no integer ticks): plain Float64
arithmetic in the reference's own order of operations.  Used only by tests (tests/test_reference_pin.py) to cross-check
the C oracle; small cases only (it is a Python hour loop).

  sequential_mc      GeneratingAdequacy/PowerSystemAdequacy.jl:214-269 with the three draws (:224,243,246) read from
                     per-unit duration lists D[u][k] (k = 0: initial time to failure, then repair, failure, ...)
  non_sequential_mc  PowerSystemAdequacy.jl:169-208 with rand() replaced by the matrix r[i][u]
  calnlc             Montecarlo_seq/calnlc.m:22-34 (number of load-curtailment events of a 0/1 hour series)
  matlab_unit_series Montecarlo_seq/seq_mcsampling.m:40-74 for one unit with injected uniforms
"""
import math


def sequential_mc(cap, load, years, D, record_hours=False):
    n = len(cap)
    up = [True] * n
    used = [1] * n
    ttf = [D[u][0] for u in range(n)]                      # :224
    per_year_lole, per_year_eue, per_year_nlc, history = [], [], [], []
    cum = 0.0
    for y in range(1, years + 1):
        yl, ye = 0.0, 0.0
        flags = []
        for h in range(len(load)):
            avail = 0.0
            for u in range(n):
                ttf[u] -= 1.0                              # :237
                while ttf[u] <= 0:                         # :238
                    up[u] = not up[u]                      # :240 / :244
                    ttf[u] += D[u][used[u]]                # :243 / :246 (repair after a failure, failure after a repair)
                    used[u] += 1
                if up[u]:
                    avail += cap[u]                        # :249
            lost = avail < load[h]                         # :253 strict
            if lost:
                yl += 1.0
                ye += load[h] - avail
            flags.append(1 if lost else 0)
        per_year_lole.append(yl)
        per_year_eue.append(ye)
        per_year_nlc.append(calnlc(flags))
        cum += yl
        if y % 10 == 0:
            history.append(cum / y)                        # :264-266
    return per_year_lole, per_year_eue, per_year_nlc, history


def non_sequential_mc(cap, for_rate, load, r):
    lole, eue, history = [], [], []
    cum = 0.0
    for i, row in enumerate(r, start=1):
        avail = 0.0
        for u in range(len(cap)):
            if row[u] >= for_rate[u]:                      # :183
                avail += cap[u]
        il, ie = 0.0, 0.0
        for L in load:
            if avail < L:                                  # :192
                il += 1.0
                ie += L - avail
        lole.append(il)
        eue.append(ie)
        cum += il
        if i % 100 == 0:
            history.append(cum / i)                        # :203-205
    return lole, eue, history


def calnlc(series):
    starts = sum(1 for a, b in zip(series[:-1], series[1:]) if b - a == 1)      # diff(...) == 1
    if series and series[0] == 1:
        starts += 1
    return starts


def matlab_unit_series(mttf, mttr, total_hours, uniforms):
    """0/1 failure series of one unit (1 = DOWN), seq_mcsampling.m:40-74; MATLAB round() is half away from zero."""
    s = [0] * total_hours
    t = 0
    is_up = True
    k = 0
    while t < total_hours:
        u = uniforms[k]
        k += 1
        if is_up:
            d = -mttf * math.log(u)
            t += int(math.floor(d + 0.5))                  # round(duration), duration >= 0
        else:
            d = -mttr * math.log(u)
            di = int(math.ceil(d))
            start = t + 1                                  # 1-based start_idx
            end = min(start + di - 1, total_hours)
            if start <= total_hours:
                for j in range(start, end + 1):
                    s[j - 1] = 1
            t += di
        is_up = not is_up
    return s, k
