"""L2-circuit proofs/s through zkb_l2_batch_prove for one (slots, sub-batch, proofs per call, host threads) setting.
   python tools/l2_batch_sweep.py SLOTS SUBBATCH PER_CALL [LANES] [STEPS]      (env is read once per process: one setting per run)"""
import os
import sys
import time

slots, sub, per_call = sys.argv[1], sys.argv[2], int(sys.argv[3])
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 16
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 6
os.environ["ZKB_L2_SLOTS"] = slots
os.environ["ZKB_L2_SUBBATCH"] = sub
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zelana_b200  # noqa: E402
from zelana_b200 import l2_circuit as l2  # noqa: E402

ctx = zelana_b200.Context(0)
circ, pk_bytes, vk_bytes, _raw = l2.keygen(ctx)
if os.environ.get("MSM_C"):
    ctx.set_msm_window(int(os.environ["MSM_C"]))     # window width of the key's tables (default: chosen from the key size)
pk = ctx.proving_key_compressed(pk_bytes, validate=False)
prover = l2.L2BatchProver(ctx, circ, pk, lanes=lanes)


def batch(bid):
    ckt = l2.L2BlockCircuit(transactions=[l2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 1 + bid % 1000)],
                            initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): bid}, batch_id=bid)
    return ckt.with_inputs(l2.satisfying_inputs(ckt))


pack = prover.marshal([batch(i + 1) for i in range(per_call)])
for _ in range(2):
    prover.prove_marshalled(pack)
t0 = time.perf_counter()
for _ in range(steps):
    prover.prove_marshalled(pack)
dt = (time.perf_counter() - t0) / steps
print("slots=%s subbatch=%s per_call=%d lanes=%d: %.1f ms per call = %.0f proofs/s" % (slots, sub, per_call, lanes, dt * 1e3, per_call / dt), flush=True)
