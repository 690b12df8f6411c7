"""gnark / sunspot byte formats of the forge stack (zelana_b200/gnark_format.py) against the sizes and the parser the reference
holds: docs/PROVER_LAYER.md:139-140,226-227; forge/crates/prover-worker/src/prover.rs:575-596;
forge/crates/prover-coordinator/src/solana_client.rs:159-187; prover-coordinator/src/ownership_api.rs:355-382."""
import hashlib
import random

import pytest

from oracle import bn254 as bn
from zelana_b200 import gnark_format as gf

R = bn.R


def test_public_witness_236_bytes_and_reference_parser():
    rnd = random.Random(3)
    inputs = [rnd.randrange(R) for _ in range(7)]
    pw = gf.write_public_witness(inputs)
    assert len(pw) == 236                                   # "4-byte header + 8-byte padding + 7 x 32-byte inputs"
    assert pw[:4] == bytes([0, 0, 0, 7])
    assert gf.parse_public_witness(pw) == ["0x%064x" % v for v in inputs]
    assert gf.parse_public_witness(pw[:11]) == []
    assert gf.parse_public_witness(pw[:12 + 32 + 5]) == ["0x%064x" % inputs[0]]     # a truncated element is dropped (:588)
    # the mock the reference's own API builds (ownership_api.rs:367-376): count 3, 8 bytes of padding, 3 x 32 B -> 108 B
    h = hashlib.sha256(b"x").digest()
    mock = bytes([0, 0, 0, 3]) + bytes(8) + h * 3
    assert len(mock) == 108 and gf.parse_public_witness(mock) == ["0x" + h.hex()] * 3
    assert len(gf.write_public_witness([1, 2, 3])) == 108


def test_proof_388_bytes_round_trip_and_layout():
    g1 = lambda k: bn.g1_to_raw(bn.G1.mul(bn.G1_GEN, k))
    g2 = lambda k: bn.g2_to_raw(bn.G2.mul(bn.G2_GEN, k))
    a, b, c, cm, pok = g1(5), g2(7), g1(11), g1(13), g1(17)
    proof = gf.write_proof(a, b, c, [cm], pok)
    assert len(proof) == gf.PROOF_BYTES_ONE_COMMITMENT == 388          # 256 + 4 + 64 + 64
    assert proof[256:260] == bytes([0, 0, 0, 1])
    assert gf.parse_proof(proof) == (a, b, c, [cm], pok)
    # big-endian coordinates, G2 imaginary part first: A.x of 5 G as an integer
    ax = bn.G1.mul(bn.G1_GEN, 5)[0]
    assert int.from_bytes(proof[:32], "big") == ax
    bx = bn.G2.mul(bn.G2_GEN, 7)[0]
    assert int.from_bytes(proof[64:96], "big") == bx[1] and int.from_bytes(proof[96:128], "big") == bx[0]
    assert len(gf.write_proof(a, b, c)) == 324
    with pytest.raises(ValueError):
        gf.parse_proof(proof[:-1])
    data = gf.verifier_instruction_data(proof, gf.write_public_witness(range(7)))
    assert len(data) == 388 + 236 and data[:388] == proof
    with pytest.raises(ValueError):
        gf.verifier_instruction_data(proof[:-1], gf.write_public_witness(range(7)))
    with pytest.raises(ValueError):
        gf.verifier_instruction_data(proof, gf.write_public_witness(range(3)))
