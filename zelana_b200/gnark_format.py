"""Byte formats of the forge ("stack B") prover outputs -- what `sunspot prove` (gnark Groth16 over BN254) writes and what the
reference's coordinator parses and ships to the Solana verifier:

  * the 388-byte proof `target/zelana_batch.proof` (docs/PROVER_LAYER.md:139,226; size checked at
    forge/crates/prover-coordinator/src/solana_client.rs:170-175): gnark's `Proof.WriteRawTo` for BN254 --
        Ar (G1, 64 B) | Bs (G2, 128 B) | Krs (G1, 64 B) | u32 BE number of commitments | commitments (G1, 64 B each) |
        CommitmentPok (G1, 64 B)
    = 256 + 4 + 64 + 64 with the one BSB22 commitment the Noir circuit's range checks need.  Coordinates are 32-byte
    BIG-endian; a G2 coordinate is written imaginary part first (X.A1 | X.A0 | Y.A1 | Y.A0).
  * the 236-byte public witness `target/zelana_batch.pw` (docs/PROVER_LAYER.md:140,227): gnark's witness binary for a
    public-only witness -- u32 BE public count | u32 BE secret count (0) | u32 BE vector length | count x 32 B BE elements;
    the reference reads it as "4-byte count + 8 bytes of padding/metadata" (forge/crates/prover-worker/src/prover.rs:575-596).
  * the verifier instruction = proof bytes followed by the public-witness bytes (solana_client.rs:183-187).

Host-side formatting only (like proof_to_solana_bytes for stack A): the C ABI emits canonical little-endian affine points.
What is NOT here: gnark's proving-key / constraint-system readers and its quotient convention -- the reference keeps none of
those files (`forge/.gitignore:2,37-38`), so there is nothing to pin them on.
"""
from typing import List, Sequence, Tuple

PROOF_BYTES_ONE_COMMITMENT = 388
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _be32(le: bytes) -> bytes:
    return bytes(reversed(le))


def g1_le_to_gnark(p: bytes) -> bytes:
    """64 B canonical little-endian affine (x | y, the C ABI's form) -> gnark raw: x BE | y BE; infinity (zeros) stays zeros."""
    assert len(p) == 64
    return _be32(p[:32]) + _be32(p[32:])


def g2_le_to_gnark(p: bytes) -> bytes:
    """128 B little-endian (x.c0 | x.c1 | y.c0 | y.c1) -> gnark raw: X.A1 | X.A0 | Y.A1 | Y.A0, each 32 B BE."""
    assert len(p) == 128
    return _be32(p[32:64]) + _be32(p[:32]) + _be32(p[96:128]) + _be32(p[64:96])


def g1_gnark_to_le(p: bytes) -> bytes:
    return _be32(p[:32]) + _be32(p[32:64])


def g2_gnark_to_le(p: bytes) -> bytes:
    return _be32(p[32:64]) + _be32(p[:32]) + _be32(p[96:128]) + _be32(p[64:96])


def write_proof(a: bytes, b: bytes, c: bytes, commitments: Sequence[bytes] = (), commitment_pok: bytes = bytes(64)) -> bytes:
    """(A, B, C) as zkb_prove returns them (+ the BSB22 commitments and their proof of knowledge, little-endian G1) -> gnark
    `Proof.WriteRawTo` bytes: 388 B with one commitment, 324 B with none."""
    out = g1_le_to_gnark(a) + g2_le_to_gnark(b) + g1_le_to_gnark(c) + len(commitments).to_bytes(4, "big")
    for cm in commitments:
        out += g1_le_to_gnark(cm)
    return out + g1_le_to_gnark(commitment_pok)


def parse_proof(data: bytes) -> Tuple[bytes, bytes, bytes, List[bytes], bytes]:
    """Inverse of write_proof -> (A, B, C, commitments, pok) in the C ABI's little-endian form."""
    if len(data) < 256 + 4 + 64:
        raise ValueError("gnark proof: %d bytes is shorter than the fixed part" % len(data))
    n = int.from_bytes(data[256:260], "big")
    if len(data) != 256 + 4 + 64 * n + 64:
        raise ValueError("gnark proof: %d bytes does not match %d commitments" % (len(data), n))
    cms = [g1_gnark_to_le(data[260 + 64 * i:324 + 64 * i]) for i in range(n)]
    return (g1_gnark_to_le(data[:64]), g2_gnark_to_le(data[64:192]), g1_gnark_to_le(data[192:256]), cms,
            g1_gnark_to_le(data[260 + 64 * n:]))


def write_public_witness(inputs: Sequence[int]) -> bytes:
    """Public inputs (field elements) -> gnark public-witness bytes: 12-byte header + 32 B big-endian each (236 B for the batch
    circuit's seven inputs, 108 B for the ownership circuit's three: prover-coordinator/src/ownership_api.rs:367-376)."""
    n = len(inputs)
    out = n.to_bytes(4, "big") + (0).to_bytes(4, "big") + n.to_bytes(4, "big")
    for v in inputs:
        out += (int(v) % R).to_bytes(32, "big")
    return out


def parse_public_witness(data: bytes) -> List[str]:
    """forge/crates/prover-worker/src/prover.rs:575-596, statement for statement: "0x"-prefixed hex of every 32-byte input that
    is completely present; fewer than 12 bytes -> nothing."""
    if len(data) < 12:
        return []
    count = int.from_bytes(data[:4], "big")
    out = []
    for i in range(count):
        off = 12 + 32 * i
        if off + 32 <= len(data):
            out.append("0x" + data[off:off + 32].hex())
    return out


def verifier_instruction_data(proof_bytes: bytes, public_witness_bytes: bytes) -> bytes:
    """solana_client.rs:159-187: the sunspot verifier's instruction data, with the size checks the client makes."""
    if len(proof_bytes) != 388:
        raise ValueError("Expected 388 bytes proof, got %d" % len(proof_bytes))
    if len(public_witness_bytes) != 236:
        raise ValueError("Expected 236 bytes public witness, got %d" % len(public_witness_bytes))
    return bytes(proof_bytes) + bytes(public_witness_bytes)
