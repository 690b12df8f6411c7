"""The L2 batch circuit at its scaled shape (64 transfers = MAX_TXS over 128 accounts; ~292 k
constraints, domain 2^19): keygen on the GPU, BatchProver::prove end to end, pairing check of one proof by the oracle."""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zelana_b200  # noqa: E402
from zelana_b200 import l2_circuit as l2  # noqa: E402


def main():
    rnd = random.Random(1)
    keys = [bytes(rnd.randrange(256) for _ in range(32)) for _ in range(128)]

    def batch(bid):
        r = random.Random(bid)
        txs = [l2.TransactionWitness(keys[rnd2], keys[(rnd2 * 7 + 3) % 128], r.randrange(1000)) for rnd2 in range(64)]
        c = l2.L2BlockCircuit(transactions=txs, initial_accounts={k: 10 ** 9 + bid for k in keys}, batch_id=bid)
        return c.with_inputs(l2.satisfying_inputs(c))

    ctx = zelana_b200.Context(0)
    t0 = time.perf_counter()
    circ, pk_bytes, vk_bytes, _raw = l2.keygen(ctx, batch(0))
    t_keygen = time.perf_counter() - t0
    t0 = time.perf_counter()
    pk = ctx.proving_key_compressed(pk_bytes, validate=True)
    t_load = time.perf_counter() - t0
    prover = l2.L2Prover(ctx, circ, pk, vk_bytes)
    circuits = [batch(i + 1) for i in range(6)]
    for c in circuits[:3]:
        prover.prove_circuit(c)
    t0 = time.perf_counter()
    proofs = [prover.prove_circuit(c).proof_bytes for c in circuits]
    t_prove = (time.perf_counter() - t0) / len(circuits)
    t0 = time.perf_counter()
    for c in circuits:
        circ.assign(c)
    t_assign = (time.perf_counter() - t0) / len(circuits)
    from oracle import bn254 as bn, groth16 as g16
    vk = g16.VerifyingKey.deserialize_compressed(vk_bytes)
    z = circ.assign(circuits[0])
    pub = [int.from_bytes(z[32 * i:32 * i + 32], "little") for i in range(1, 8)]
    pb = proofs[0]
    proof = g16.Proof(bn.G1.neg(bn.g1_from_raw(pb[:64])), bn.g2_from_raw(pb[64:192]), bn.g1_from_raw(pb[192:]))
    ok = g16.verify(vk, pub, proof)
    print(json.dumps({"shape": "64 transfers over 128 accounts", "constraints": circ.num_constraints, "witness": circ.num_witness,
                      "pk_bytes": len(pk_bytes), "keygen_s": t_keygen, "pk_load_validate_s": t_load,
                      "prove_ms_end_to_end": t_prove * 1e3, "assign_ms": t_assign * 1e3,
                      "assign_threads": os.environ.get("ZKB_L2_ASSIGN_THREADS", "default min(8, cores)"),
                      "host_cores": os.cpu_count(), "proof_verifies_by_pairing": bool(ok), "satisfied": circ.is_satisfied(z)[0]}))


if __name__ == "__main__":
    main()
