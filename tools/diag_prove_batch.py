"""GPU diagnostic: zkb_prove_batch against zkb_prove, component by component."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zelana_b200
from zelana_b200 import l2_circuit as P2


def main():
    ctx = zelana_b200.Context(0)
    circ, pk_bytes, vk_bytes, raw = P2.keygen(ctx)
    dpk = ctx.proving_key_compressed(pk_bytes, validate=False)
    a, b, c = circ.matrices()
    m = ctx.r1cs(circ.num_instance, circ.num_witness, a, b, c)

    def assignment(bid):
        ck = P2.L2BlockCircuit(transactions=[P2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 3 * bid + 1)],
                               initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): bid}, batch_id=bid)
        return circ.assign(ck.with_inputs(P2.satisfying_inputs(ck)))

    for K in (1, 2, 5):
        zs = [assignment(i + 1) for i in range(K)]
        rs = [P2.prover_randomness(i + 1) for i in range(K)]
        expect = [ctx.prove(dpk, m, z, r, s) for z, (r, s) in zip(zs, rs)]
        got = ctx.prove_batch(dpk, m, b"".join(zs), b"".join(r + s for r, s in rs))
        print("K=%d" % K, [(g[0] == e[0], g[1] == e[1], g[2] == e[2]) for g, e in zip(got, expect)])
        # r = s = 0: C = L + H, A = alpha + sum, B = beta + sum
        zero = bytes(32)
        e0 = [ctx.prove(dpk, m, z, zero, zero) for z in zs]
        g0 = ctx.prove_batch(dpk, m, b"".join(zs), b"".join(zero + zero for _ in zs))
        print("   r=s=0:", [(g[0] == e[0], g[1] == e[1], g[2] == e[2]) for g, e in zip(g0, e0)])
        one = (1).to_bytes(32, "little")
        e1 = [ctx.prove(dpk, m, z, one, zero) for z in zs]
        g1 = ctx.prove_batch(dpk, m, b"".join(zs), b"".join(one + zero for _ in zs))
        print("   r=1,s=0:", [(g[0] == e[0], g[1] == e[1], g[2] == e[2]) for g, e in zip(g1, e1)])
        e2 = [ctx.prove(dpk, m, z, zero, one) for z in zs]
        g2 = ctx.prove_batch(dpk, m, b"".join(zs), b"".join(zero + one for _ in zs))
        print("   r=0,s=1:", [(g[0] == e[0], g[1] == e[1], g[2] == e[2]) for g, e in zip(g2, e2)])
    ctx.close()


if __name__ == "__main__":
    main()
