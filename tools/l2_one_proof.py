"""Three L2BlockCircuit proofs on one context with direct launches (graphs off), for `ncu --metrics gpu__time_duration.sum`:
the launch list of the LAST proof is the per-kernel timeline of a small proof.  Prints the per-phase CUDA-event times too."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zelana_b200  # noqa: E402
from zelana_b200 import l2_circuit as l2  # noqa: E402

ctx = zelana_b200.Context(0)
ctx.set_graphs(False)
circ, pk_bytes, vk_bytes, _raw = l2.keygen(ctx)
prover = l2.L2Prover(ctx, circ, ctx.proving_key_compressed(pk_bytes, validate=False), vk_bytes)
ckt = l2.L2BlockCircuit.dummy()
ckt = ckt.with_inputs(l2.satisfying_inputs(ckt))
prover.prove_circuit(ckt)
prover.prove_circuit(ckt)
ctx.synchronize()
print("MARK last proof starts after %d library launches" % ctx.launch_count(), flush=True)
ctx.profile(True)
ctx.profile_reset()
t0 = time.perf_counter()
prover.prove_circuit(ckt)
wall = (time.perf_counter() - t0) * 1e3
print(json.dumps({"wall_ms": wall, "phases_ms": {k: v[0] for k, v in ctx.profile_read().items()},
                  "launches": ctx.launch_count()}))
