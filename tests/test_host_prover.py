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
    with pytest.raises(TypeError):            # neither an L2BlockCircuit nor a custom synthesizer
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


# ----------------------------------------------------------------------------- keygen host pieces (zelana_b200/keygen.py)
def test_keygen_sampling_and_compression_match_oracle():
    from zelana_b200 import keygen as kg
    assert kg.B_G2 == bn.B_G2 and kg.G2_COFACTOR == bn.G2_COFACTOR
    for seed in (0, 42, 9):
        a, b = zp.StdRng.seed_from_u64(seed), orng.StdRng.seed_from_u64(seed)
        for _ in range(4):
            assert zp.fr_rand(a) == orng.rand_fr(b)
        assert kg.g1_rand(a) == orng.rand_g1(b)
        g2u = kg.g2_rand_uncleared(a)
        assert bn.G2.mul(g2u, bn.G2_COFACTOR) == orng.rand_g2(b)      # the product clears the cofactor on the GPU
        assert zp.fr_rand(a) == orng.rand_fr(b)                       # streams still aligned afterwards
    g = bn.G1.mul(bn.G1_GEN, 12345)
    for pt in (g, bn.G1.neg(g), None):
        assert kg.g1_compress(bn.g1_to_raw(pt)) == bn.g1_serialize(pt)
    h = bn.G2.mul(bn.G2_GEN, 777)
    for pt in (h, bn.G2.neg(h), None):
        assert kg.g2_compress(bn.g2_to_raw(pt)) == bn.g2_serialize(pt)


@pytest.mark.gpu
def test_gpu_keygen_reproduces_reference_vk_and_proof():
    """keygen on the GPU with StdRng(42), as prover/src/snarkjs.rs:141-176 does for SquareCircuit: the verifying key equals
    the reference's committed vk_snarkjs.json, the key bytes equal the oracle's, and proving with it on the same RNG stream
    reproduces the committed proof."""
    import zelana_b200
    from zelana_b200 import keygen as kg
    from helpers import fr_bytes
    r1cs, z = g16.square_circuit(7)
    ctx = zelana_b200.Context(0)
    try:
        rng = zp.StdRng.seed_from_u64(42)
        pk_bytes, vk_bytes, _ = kg.circuit_specific_setup(ctx, r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c, rng)
        orc_rng = orng.StdRng.seed_from_u64(42)
        opk = g16.circuit_specific_setup(r1cs, orc_rng)
        assert vk_bytes == opk.vk.serialize_compressed()
        assert pk_bytes == opk.serialize_compressed()
        v = json.load(open(os.path.join(REF_FIXTURES, "vk_snarkjs.json")))
        vk = g16.VerifyingKey.deserialize_compressed(vk_bytes)
        assert vk.alpha_g1 == (int(v["vk_alpha_1"][0]), int(v["vk_alpha_1"][1]))
        assert [p for p in vk.gamma_abc_g1] == [(int(p[0]), int(p[1])) for p in v["IC"]]
        # prove on the SAME stream (snarkjs.rs:156-159) from the GPU-made key
        r, s = zp.fr_rand(rng), zp.fr_rand(rng)
        m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
        a, b, c = ctx.prove(ctx.proving_key_compressed(pk_bytes), m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
        proof = g16.Proof(bn.g1_from_raw(a), bn.g2_from_raw(b), bn.g1_from_raw(c))
        pc = json.load(open(os.path.join(REF_FIXTURES, "proof_for_onchain.json")))["proof_components"]
        assert proof.serialize_uncompressed() == bytes(pc["pi_a"]) + bytes(pc["pi_b"]) + bytes(pc["pi_c"])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_keygen_mimc_matches_oracle_and_proofs_verify():
    import zelana_b200
    from zelana_b200 import keygen as kg
    from helpers import fr_bytes, mimc7_chain
    r1cs, z = mimc7_chain(num_perm=2, seed=42, rounds=20)
    ctx = zelana_b200.Context(0)
    try:
        pk_bytes, vk_bytes, _ = kg.circuit_specific_setup(ctx, r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c,
                                                          zp.StdRng.seed_from_u64(0))          # keygen.rs:87: seed 0
        opk = g16.circuit_specific_setup(r1cs, orng.StdRng.seed_from_u64(0))
        assert pk_bytes == opk.serialize_compressed() and vk_bytes == opk.vk.serialize_compressed()
        p = zp.Groth16Prover(ctx, ctx.proving_key_compressed(pk_bytes), vk_bytes)
        m = p.circuit(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
        got = p.prove_assignment(zp.BatchPublicInputs(batch_id=5), m, fr_bytes(z))
        ref = g16.prove(opk, r1cs, z, orng.StdRng.seed_from_u64(5))
        assert got.proof_bytes == ref.to_solana_bytes()
        assert g16.verify(g16.VerifyingKey.deserialize_compressed(vk_bytes), [z[1]], ref)
        # a trapdoor inside the domain is refused
        with pytest.raises(zelana_b200.ZkbError):
            ctx.setup(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c, alpha=1, beta=2, gamma=3, delta=4, tau=1,
                      g1_generator=bn.g1_to_raw(bn.G1_GEN), g2_generator=bn.g2_to_raw(bn.G2_GEN))
    finally:
        ctx.close()


@pytest.mark.gpu
def test_groth16_prover_proves_the_l2_circuit_natively():
    """`Groth16Prover::from_bytes(pk, vk)` + `prove(inputs, witness)` with the witness as the reference's L2BlockCircuit struct:
    constraints come from the library's native synthesiser, the proof verifies under the key's VK and equals L2Prover's."""
    import zelana_b200
    from zelana_b200 import l2_circuit as l2
    ctx = zelana_b200.Context(0)
    try:
        circ, pk_bytes, vk_bytes, _raw = l2.keygen(ctx)
        circ.free()
    finally:
        ctx.close()
    p = zp.Groth16Prover.from_bytes(pk_bytes, vk_bytes, device=0)
    try:
        w = l2.L2BlockCircuit.dummy()
        w.batch_id = 11
        inputs = l2.satisfying_inputs(w)
        proof = p.prove(inputs, w)
        assert p.verify(proof) and len(proof.proof_bytes) == 256
        again = p.prove(inputs, w)                         # cached circuit, CUDA graph from here on
        assert again.proof_bytes == proof.proof_bytes      # StdRng(batch_id): deterministic
        vk = g16.VerifyingKey.deserialize_compressed(vk_bytes)
        pb = proof.proof_bytes
        pr = g16.Proof(bn.G1.neg(bn.g1_from_raw(pb[:64])), bn.g2_from_raw(pb[64:192]), bn.g1_from_raw(pb[192:]))
        z = l2.L2Circuit(w).assign(w.with_inputs(inputs))
        assert g16.verify(vk, [int.from_bytes(z[32 * i:32 * i + 32], "little") for i in range(1, 8)], pr)
    finally:
        p.close()
