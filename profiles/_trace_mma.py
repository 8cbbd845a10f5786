"""Anatomy of the MMA phase of K1 from the trace saved by _trace_chain.py k1 (gpurun_out/trace_k1.npy):
who waits for whom between the epilogue's operand-ready signal and the accumulator being seen again."""
import sys
import numpy as np
t = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_k1.npy").astype(np.int64)
n_steps = 8
slices = [2, 8, 8, 8, 8, 8, 8, 8]
per_tile = sum(slices)
n_tiles = int((t[3, :, 1] > 0).sum()) // n_steps
print("step | sig(thread 64) -> a_ready seen by MMA warp | a_ready -> first slice seen | sum of full-waits of the other slices | "
      "first slice seen -> last MMA issued | last issue -> commit stamp | commit -> acc seen by epilogue | total")
for st in range(n_steps):
    rows = []
    for tile in range(2, n_tiles - 1):
        g = tile * n_steps + st
        it0 = tile * per_tile + sum(slices[:st])
        sig = t[3, g, 2]                       # thread 64's signal that PRECEDES this step's accumulator wait
        a_seen = t[2, g, 1]
        first_full = t[1, it0, 1]
        waits = sum(t[1, it0 + k, 1] - t[1, it0 + k, 0] for k in range(1, slices[st]))
        last_issue = t[1, it0 + slices[st] - 1, 2]
        commit = t[2, g, 2]
        acc_seen = t[3, g, 1]
        rows.append([a_seen - sig, first_full - a_seen, waits, last_issue - first_full, commit - last_issue, acc_seen - commit, acc_seen - sig])
    m = np.array(rows).mean(0)
    print(f"{st:4d} | " + " | ".join(f"{v:7.0f}" for v in m))
