"""SURVEY.md 8f.3 tail, measured: the independent Poseidon LEAF hashes of the L2 batch circuit (prover/src/l2_circuit.rs:315-330,
477-490) on the GPU versus the host's native sponge.  Scaled shape = 64 transfers over 128 accounts: 128 pre-state + 128 post-state
account leaves (2 elements) + 64 transfer leaves (3 elements) = 320 independent hashes per proof; the ~390 fold hashes are
sequential chains and stay on the host.  Reports, for K proofs' worth of leaves: host threads (the walkers' own code) vs
GPU through host buffers (H2D + kernel + D2H) vs GPU device-resident."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import zelana_b200  # noqa: E402

lib = zelana_b200.load_library()
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = zelana_b200.Context(0, stream=st.cuda_stream)
threads = len(os.sched_getaffinity(0))
rows = []
for K in (1, 16, 256, 4096):
    n2, n3 = 256 * K, 64 * K
    rs = np.random.RandomState(K)

    def rnd(n):
        a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        a[:, 7] %= 0x30644E72
        return a.tobytes()

    in2, in3 = rnd(2 * n2), rnd(3 * n3)
    out2, out3 = C.create_string_buffer(32 * n2), C.create_string_buffer(32 * n3)
    t0 = time.perf_counter()
    assert lib.zkb_l2_poseidon_hash_batch_host(2, in2, n2, threads, out2) == 0
    assert lib.zkb_l2_poseidon_hash_batch_host(3, in3, n3, threads, out3) == 0
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    assert lib.zkb_l2_poseidon_hash_batch_host(2, in2, min(n2, 256), 1, out2) == 0
    t_host1 = (time.perf_counter() - t0) * (n2 / min(n2, 256)) * (1 + 2 * n3 / n2)    # one thread, extrapolated (3-element hashes = 2 permutations)
    for _ in range(2):
        g2, g3 = ctx.l2_poseidon_hash_batch(2, in2), ctx.l2_poseidon_hash_batch(3, in3)
    assert lib.zkb_l2_poseidon_hash_batch_host(2, in2, n2, threads, out2) == 0
    assert g2 == out2.raw and g3 == out3.raw
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.l2_poseidon_hash_batch(2, in2)
        ctx.l2_poseidon_hash_batch(3, in3)
    t_gpu_host = (time.perf_counter() - t0) / reps
    d2 = torch.from_numpy(np.frombuffer(in2, dtype=np.int32).copy()).to(dev)
    d3 = torch.from_numpy(np.frombuffer(in3, dtype=np.int32).copy()).to(dev)
    o2 = torch.empty(n2 * 8, dtype=torch.int32, device=dev)
    o3 = torch.empty(n3 * 8, dtype=torch.int32, device=dev)
    for _ in range(2):
        ctx.l2_poseidon_hash_batch_dev(2, d2, n2, o2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        ctx.l2_poseidon_hash_batch_dev(2, d2, n2, o2)
        ctx.l2_poseidon_hash_batch_dev(3, d3, n3, o3)
    e1.record(st)
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / reps * 1e-3
    rows.append({"proofs": K, "leaf_hashes": n2 + n3, "host_%d_threads_ms" % threads: t_host * 1e3, "host_1_thread_ms_extrapolated": t_host1 * 1e3,
                 "gpu_through_host_buffers_ms": t_gpu_host * 1e3, "gpu_device_resident_ms": t_dev * 1e3,
                 "gpu_hashes_per_s_device": (n2 + n3) / t_dev, "host_hashes_per_s": (n2 + n3) / t_host})
print(json.dumps({"workload": "l2_poseidon_leaf_hashes", "per_proof": "256 two-element + 64 three-element hashes (64 transfers, 128 accounts)",
                  "host_threads": threads, "rows": rows}))
ctx.close()
