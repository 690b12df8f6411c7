"""Round-2 profiling driver for `ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`:
   python tools/prof_r2.py msm24   -> one 2^24 G1 MSM inside the profiled range (after two warm-ups)
   python tools/prof_r2.py batch   -> one zkb_prove_batch of 128 L2-circuit proofs inside the profiled range
Without ncu it prints CUDA-event timings of the same work."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import zelana_b200  # noqa: E402
from tools.diag_batch import rand_fr  # noqa: E402


def msm(log_n):
    n = 1 << log_n
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    ctx = zelana_b200.Context(0, stream=st.cuda_stream)
    if os.environ.get("MSM_C"):
        ctx.set_msm_window(int(os.environ["MSM_C"]))
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    k = torch.randint(0, 1 << 32, (n, 8), dtype=torch.int64, device=dev, generator=g)
    k[:, 7] %= 0x30644E72
    k = k.to(torch.int32)
    bases = ctx.g1_bases_generate(k, n)
    s = torch.randint(0, 1 << 32, (n, 8), dtype=torch.int64, device=dev, generator=g)
    s[:, 7] %= 0x30644E72
    s = s.to(torch.int32)
    out = torch.zeros(64, dtype=torch.uint8, device=dev)
    for _ in range(2):
        ctx.msm_g1_dev(bases, s, n, out_affine_dev=out)
    torch.cuda.synchronize()
    ctx.profile(True)
    ctx.profile_reset()
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    ctx.msm_g1_dev(bases, s, n, out_affine_dev=out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.profiler.stop()
    print("msm 2^%d: %.3f ms wall" % (log_n, dt * 1e3), {k2: round(v[0], 3) for k2, v in ctx.profile_read().items()})


def batch(K):
    from zelana_b200 import l2_circuit as P2
    ctx = zelana_b200.Context(0)
    circ, pk_bytes, vk_bytes, raw = P2.keygen(ctx)
    dpk = ctx.proving_key_compressed(pk_bytes, validate=False)
    a, b, c = circ.matrices()
    m = ctx.r1cs(circ.num_instance, circ.num_witness, a, b, c)
    zs, rs = [], []
    for bid in range(1, K + 1):
        ck = P2.L2BlockCircuit(transactions=[P2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 3 * bid + 1)],
                               initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): bid}, batch_id=bid)
        zs.append(circ.assign(ck.with_inputs(P2.satisfying_inputs(ck))))
        r, s = P2.prover_randomness(bid)
        rs.append(r + s)
    z, rsb = b"".join(zs), b"".join(rs)
    for _ in range(2):
        ctx.prove_batch(dpk, m, z, rsb)
    ctx.profile(True)
    ctx.profile_reset()
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    ctx.prove_batch(dpk, m, z, rsb)
    dt = time.perf_counter() - t0
    torch.cuda.profiler.stop()
    print("prove_batch K=%d: %.3f ms wall = %.1f proofs/s" % (K, dt * 1e3, K / dt), {k2: round(v[0], 3) for k2, v in ctx.profile_read().items()})
    ctx.profile(False)
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.prove_batch(dpk, m, z, rsb)
    dt = (time.perf_counter() - t0) / 5
    print("prove_batch K=%d unprofiled: %.3f ms = %.1f proofs/s" % (K, dt * 1e3, K / dt))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "msm24":
        msm(int(sys.argv[2]) if len(sys.argv) > 2 else 24)
    else:
        batch(int(sys.argv[2]) if len(sys.argv) > 2 else 128)
