"""Exhaustive check of the device logarithm / tick duration against the CPU specification: ALL 2^32 draws, in chunks
(oracle on a thread pool).  Writes gpurun_out/sampler_exhaustive.json.  usage: python scripts/sampler_exhaustive.py [log2 chunk]"""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from powersystemsreliabilityassessment_b200 import Engine

LC = int(sys.argv[1]) if len(sys.argv) > 1 else 26
chunk = 1 << LC
nthreads = min(16, os.cpu_count() or 1)
means = (450.0, 2940.0)
t0 = time.time()
bad_e = 0; bad_t = 0; n = 0
xor_e = np.uint32(0)
O.lib()
with Engine() as e, ThreadPoolExecutor(nthreads) as pool:
    for c in range(1 << (32 - LC)):
        x = np.arange(c * chunk, (c + 1) * chunk, dtype=np.uint64).astype(np.uint32)
        mean = means[c & 1]
        tg, eg = e.sampler_durations(mean, x)
        parts = np.array_split(np.arange(chunk), nthreads)
        res = list(pool.map(lambda p: O.sampler_durations(mean, x[p[0]:p[-1] + 1]), parts))
        tc = np.concatenate([r[0] for r in res]); ec = np.concatenate([r[1] for r in res])
        bad_e += int((eg != ec).sum()); bad_t += int((tg != tc).sum()); n += chunk
        xor_e ^= np.bitwise_xor.reduce(eg)
        if c % 8 == 0:
            print(f"chunk {c}: {n} draws, mismatches e {bad_e} ticks {bad_t}, {time.time() - t0:.0f} s", flush=True)
out = dict(draws=n, e_bits_mismatches=bad_e, tick_mismatches=bad_t, means_alternating=list(means), xor_of_all_e_bits=int(xor_e),
           seconds=time.time() - t0, oracle_threads=nthreads)
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sampler_exhaustive.json", "w"), indent=1)
