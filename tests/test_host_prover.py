"""Host-side mirror of the reference's prover interface (zelana_b200/prover.py) against the oracle: the RNG stream behind
`StdRng::seed_from_u64(batch_id)`, Fr::rand, the 256-byte Solana proof layout, the VK hash."""
import base64
import json
import os

import pytest

from conftest import REF_FIXTURES
from oracle import bn254 as bn
from oracle import groth16 as g16
from oracle import rng as orng
from zelana_b200 import prover as zp


@pytest.mark.parametrize("seed", [0, 1, 42, 7, (1 << 64) - 1, 0x1234567890ABCDEF])
def test_stdrng_stream_matches_oracle(seed):
    a, b = zp.StdRng.seed_from_u64(seed), orng.StdRng.seed_from_u64(seed)
    # mixed u32 / u64 draws cross the 64-word buffer boundary in every alignment
    for i in range(400):
        if i % 7 == 3:
            assert a.next_u32() == b.next_u32()
        else:
            assert a.next_u64() == b.next_u64()


@pytest.mark.parametrize("seed", [0, 5, 42, 70])
def test_fr_rand_matches_oracle(seed):
    a, b = zp.StdRng.seed_from_u64(seed), orng.StdRng.seed_from_u64(seed)
    for _ in range(20):
        assert zp.fr_rand(a) == orng.rand_fr(b)


def test_proof_to_solana_bytes_negates_a_and_matches_oracle():
    pc = json.load(open(os.path.join(REF_FIXTURES, "proof_for_onchain.json")))["proof_components"]
    a = bn.g1_deserialize(bytes(pc["pi_a"]), compressed=False)
    b = bn.g2_deserialize(bytes(pc["pi_b"]), compressed=False)
    c = bn.g1_deserialize(bytes(pc["pi_c"]), compressed=False)
    got = zp.proof_to_solana_bytes(bn.g1_to_raw(a), bn.g2_to_raw(b), bn.g1_to_raw(c))
    assert len(got) == 256
    assert got == g16.Proof(a, b, c).to_solana_bytes()
    assert got[:64] == bn.g1_to_raw(bn.G1.neg(a))
    assert zp.proof_to_solana_bytes(bytes(64), bn.g2_to_raw(b), bn.g1_to_raw(c))[:64] == bytes(64)
    with pytest.raises(ValueError):
        zp.proof_to_solana_bytes(b"", b"", b"")


def test_vk_hash_is_blake3_of_the_compressed_vk():
    import blake3
    raw = base64.b64decode(json.load(open(os.path.join(REF_FIXTURES, "l2_vk.json")))["verifying_key"])
    p = zp.Groth16Prover(ctx=None, pk=None, vk_bytes=raw)
    assert p.verification_key_hash() == blake3.blake3(raw).digest() and len(p.verification_key_hash()) == 32
    # verify(): the reference only checks the length (prover.rs:427-442)
    inputs = zp.BatchPublicInputs(batch_id=3)
    assert p.verify(zp.BatchProof(inputs, bytes(256), 0)) is True
    assert p.verify(zp.BatchProof(inputs, bytes(255), 0)) is False
    assert p.verify(zp.BatchProof(inputs, bytes(388 + 236), 0)) is False
    with pytest.raises(NotImplementedError):
        p.prove(inputs, witness=None)


@pytest.mark.gpu
def test_groth16_prover_from_bytes_prove_matches_reference_flow():
    """from_bytes(compressed pk, compressed vk) -> prove(batch_id) == the oracle running the reference's flow:
    StdRng::seed_from_u64(batch_id) -> Groth16::prove -> proof_to_solana_bytes."""
    from helpers import fr_bytes, mimc7_chain
    r1cs, z = mimc7_chain(num_perm=2, seed=42, rounds=20)
    pk = g16.circuit_specific_setup(r1cs, orng.StdRng.seed_from_u64(0))       # keygen.rs:87: seed 0

    def synth(inputs, witness):
        return r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c, fr_bytes(witness)

    p = zp.Groth16Prover.from_bytes(pk.serialize_compressed(), pk.vk.serialize_compressed(), device=0, synthesizer=synth)
    try:
        for batch_id in (0, 1, 70):
            inputs = zp.BatchPublicInputs(batch_id=batch_id)
            got = p.prove(inputs, z)
            ref = g16.prove(pk, r1cs, z, orng.StdRng.seed_from_u64(batch_id))
            assert got.proof_bytes == ref.to_solana_bytes() and len(got.proof_bytes) == 256
            assert p.verify(got)
            assert g16.verify(pk.vk, [z[1]], ref)
        import blake3
        assert p.verification_key_hash() == blake3.blake3(pk.vk.serialize_compressed()).digest()
    finally:
        p.close()
