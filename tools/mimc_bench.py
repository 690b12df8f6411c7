"""MiMC-7 / Merkle throughput on the GPU (device-resident, CUDA events) with the C++ restatement on the host cores beside it."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import zelana_b200  # noqa: E402
from oracle import cpu as orc  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = zelana_b200.Context(0, stream=st.cuda_stream)
peak = max(ctx.int32_peak(0)[0], ctx.int32_peak(1)[0])


def rnd(n):
    g = torch.Generator(device=dev)
    g.manual_seed(n)
    x = torch.randint(0, 1 << 32, (n, 8), dtype=torch.int64, device=dev, generator=g)
    x[:, 7] %= 0x30644E72
    return x.to(torch.int32)


def timed(fn, steps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


rows = []
n = 1 << 22
a, out = rnd(2 * n), torch.empty((n, 8), dtype=torch.int32, device=dev)
ms = timed(lambda: ctx.mimc_hash_dev(2, a, n, out))
mul32 = n * (2 * 364 + 3) * 136.0       # two permutations per hash (the arity absorption is computed once per context) + conversions
threads = orc.max_threads()
m = 1 << 16
host = a[:2 * m].cpu().numpy().tobytes()
t0 = time.perf_counter()
ref = orc.mimc_hash(2, host, threads=threads)
t_cpu = time.perf_counter() - t0
assert bytes(out[:m].cpu().numpy().tobytes()) == ref
rows.append({"what": "hash_2 (one tree level)", "n": n, "gpu_ms": ms, "gpu_hashes_per_s": n / (ms * 1e-3),
             "int32_frac": mul32 / (ms * 1e-3) / peak, "cpu_hashes_per_s": m / t_cpu, "cpu_cores": threads})
n = 1 << 18
leaves, sibs = rnd(n), rnd(n * 32)
bits = torch.randint(0, 2, (n * 32,), dtype=torch.uint8, device=dev)
out = torch.empty((n, 8), dtype=torch.int32, device=dev)
ms = timed(lambda: ctx.mimc_merkle_roots_dev(leaves, sibs, bits, n, 32, out), steps=3)
m = 1 << 11
t0 = time.perf_counter()
ref = orc.mimc_merkle_roots(leaves[:m].cpu().numpy().tobytes(), sibs[:m * 32].cpu().numpy().tobytes(), bits[:m * 32].cpu().numpy().tobytes(), 32, threads=threads)
t_cpu = time.perf_counter() - t0
assert bytes(out[:m].cpu().numpy().tobytes()) == ref
rows.append({"what": "depth-32 path roots (AccountMerklePath::compute_root)", "n": n, "gpu_ms": ms, "gpu_roots_per_s": n / (ms * 1e-3),
             "int32_frac": n * (64 * 364 + 34) * 136.0 / (ms * 1e-3) / peak, "cpu_roots_per_s": m / t_cpu, "cpu_cores": threads})
print(json.dumps({"workload": "mimc7_merkle", "rows": rows, "int32_peak_tmul32": peak / 1e12,
                  "cpu": "oracle/cpu_oracle.cpp (C++ restatement; the reference uses BigUint, account_tree.rs:56-90)"}))
ctx.close()
