"""Writes tests/golden/l2_dummy_circuit.json: counts, roots and SHA-256 digests of the L2BlockCircuit::dummy() constraint
matrices and assignment as the oracle (oracle/l2_circuit.py) produces them.

NOT a reference fixture (the reference cannot be run here and commits none for this circuit): it freezes the oracle so that
drift is noticed, and it is the comparison point for a maintainer with cargo -- the digest is defined so that ten lines of
Rust over `cs.to_matrices()` reproduce it:

    sha256( for M in (A, B, C): for row in M: u32_le(len(row)) || for (coeff, col) in row sorted by col: u32_le(col) || coeff_le32 )

with col = instance index (0 = ONE) or num_instance + witness index, as ark-relations numbers them.

    python tests/golden/make_l2_golden.py
"""
import hashlib
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import l2_circuit as O  # noqa: E402


def matrices_digest(r1cs):
    h = hashlib.sha256()
    for m in (r1cs.a, r1cs.b, r1cs.c):
        for row in m:
            h.update(struct.pack("<I", len(row)))
            for co, col in sorted(row, key=lambda e: e[1]):
                h.update(struct.pack("<I", col) + int(co).to_bytes(32, "little"))
    return h.hexdigest()


def assignment_digest(z):
    return hashlib.sha256(b"".join(int(v).to_bytes(32, "little") for v in z)).hexdigest()


def build():
    c = O.with_satisfying_roots(O.L2BlockCircuit.dummy())
    r1cs, z = O.synthesize(c)
    cfg = O.get_poseidon_config()
    return {
        "circuit": "L2BlockCircuit::dummy() (prover/src/l2_circuit.rs:147-170), batch_id 0, roots replaced by the satisfying Poseidon values",
        "num_constraints": r1cs.num_constraints, "num_instance": r1cs.num_instance, "num_witness": r1cs.num_witness,
        "nnz": [sum(len(r) for r in m) for m in (r1cs.a, r1cs.b, r1cs.c)],
        "matrices_sha256": matrices_digest(r1cs), "assignment_sha256": assignment_digest(z),
        "public_inputs_le_hex": [int(v).to_bytes(32, "little").hex() for v in z[1:8]],
        "poseidon_ark_0_0": hex(cfg.ark[0][0]), "poseidon_mds_0_0": hex(cfg.mds[0][0]),
        "poseidon_hash_1_2": hex(O.poseidon_hash([1, 2])),
    }


if __name__ == "__main__":
    with open(os.path.join(HERE, "l2_dummy_circuit.json"), "w") as f:
        json.dump(build(), f, indent=1)
        f.write("\n")
    print(open(os.path.join(HERE, "l2_dummy_circuit.json")).read())
