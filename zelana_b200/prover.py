"""Host-side mirror of the reference's prover interface for the accelerated path.

Mirrors core/src/sequencer/settlement/prover.rs (same names, argument meaning and error behaviour):
  * BatchPublicInputs / BatchProof                      prover.rs:48-74
  * trait BatchProver { prove, verify, verification_key_hash }   prover.rs:160-169
  * Groth16Prover::{from_bytes, from_files, proof_to_solana_bytes}  prover.rs:252-334
  * Groth16Prover::prove: `StdRng::seed_from_u64(inputs.batch_id)` then `Groth16::<Bn254>::prove`   prover.rs:350-425
All group / field / NTT arithmetic runs in libzkb200.so on the GPU; this file only does what the reference's own Rust does
on the host around the arkworks call: seed the RNG, draw (r, s), format bytes, hash the verifying key.

Constraint synthesis (`prover::L2BlockCircuit::generate_constraints`, prover/src/l2_circuit.rs:179-505) is host work.  `prove`
takes the witness as a `zelana_b200.l2_circuit.L2BlockCircuit` (the struct prover.rs:395-405 builds from BatchWitness) and
synthesises it with the library's native L2 synthesiser (zkb_l2_*).  Other circuits: pass `synthesizer`, a callable
`(inputs, witness) -> (num_instance, num_witness, A, B, C, full_assignment_bytes)`.
"""
import struct
import time
from dataclasses import dataclass, field
from typing import Callable, Optional

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_M32 = 0xFFFFFFFF
_M64 = 0xFFFFFFFFFFFFFFFF


# ----------------------------------------------------------------------------- rand 0.8.5 StdRng (ChaCha12), ark-ff UniformRand
class StdRng:
    """`rand::rngs::StdRng` of rand 0.8.5 = ChaCha12 with a 4-block (64-word) output buffer; `seed_from_u64` is
    rand_core 0.6.4's PCG32 expansion.  The reference seeds it with the batch id (prover.rs:354)."""

    def __init__(self, seed: bytes):
        if len(seed) != 32:
            raise ValueError("StdRng seed must be 32 bytes")
        self._key = struct.unpack("<8I", seed)
        self._block = 0
        self._words = []
        self._pos = 64

    @classmethod
    def seed_from_u64(cls, state: int) -> "StdRng":
        seed = bytearray()
        for _ in range(8):
            state = (state * 6364136223846793005 + 11634580027462260723) & _M64
            x = (((state >> 18) ^ state) >> 27) & _M32
            rot = state >> 59
            seed += struct.pack("<I", ((x >> rot) | (x << (32 - rot))) & _M32 if rot else x)
        return cls(bytes(seed))

    def _refill(self):
        out = []
        for b in range(4):
            ctr = self._block + b
            init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *self._key, ctr & _M32, (ctr >> 32) & _M32, 0, 0]
            w = list(init)
            for _ in range(6):  # 12 rounds = 6 double rounds
                for a, b_, c, d in ((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
                                    (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14)):
                    w[a] = (w[a] + w[b_]) & _M32; t = w[d] ^ w[a]; w[d] = ((t << 16) | (t >> 16)) & _M32
                    w[c] = (w[c] + w[d]) & _M32; t = w[b_] ^ w[c]; w[b_] = ((t << 12) | (t >> 20)) & _M32
                    w[a] = (w[a] + w[b_]) & _M32; t = w[d] ^ w[a]; w[d] = ((t << 8) | (t >> 24)) & _M32
                    w[c] = (w[c] + w[d]) & _M32; t = w[b_] ^ w[c]; w[b_] = ((t << 7) | (t >> 25)) & _M32
            out += [(x + y) & _M32 for x, y in zip(w, init)]
        self._block += 4
        self._words = out

    def next_u32(self) -> int:
        if self._pos >= 64:
            self._refill()
            self._pos = 0
        v = self._words[self._pos]
        self._pos += 1
        return v

    def next_u64(self) -> int:
        # rand_core BlockRng::next_u64: two consecutive words, low first; a read straddling the buffer end refills between
        if self._pos < 63:
            lo, hi = self._words[self._pos], self._words[self._pos + 1]
            self._pos += 2
        elif self._pos >= 64:
            self._refill()
            lo, hi = self._words[0], self._words[1]
            self._pos = 2
        else:
            lo = self._words[63]
            self._refill()
            hi = self._words[0]
            self._pos = 1
        return (hi << 32) | lo


_R_INV_FR = pow(1 << 256, -1, FR_MODULUS)


def fr_rand(rng: StdRng) -> int:
    """`Fr::rand` (ark-ff 0.5.0): four u64 limbs, top two bits cleared, read as the MONTGOMERY representation, rejected if
    >= r.  Returns the canonical value."""
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        raw = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | ((limbs[3] & (_M64 >> 2)) << 192)
        if raw < FR_MODULUS:
            return raw * _R_INV_FR % FR_MODULUS


# ----------------------------------------------------------------------------- reference data types
@dataclass
class BatchPublicInputs:
    """prover.rs:48-63"""
    pre_state_root: bytes = bytes(32)
    post_state_root: bytes = bytes(32)
    pre_shielded_root: bytes = bytes(32)
    post_shielded_root: bytes = bytes(32)
    withdrawal_root: bytes = bytes(32)
    batch_hash: bytes = bytes(32)
    batch_id: int = 0


@dataclass
class BatchProof:
    """prover.rs:66-74"""
    public_inputs: BatchPublicInputs
    proof_bytes: bytes
    proving_time_ms: int
    phase_ms: dict = field(default_factory=dict)   # extra: per-phase device time (zkb_prof_*)


def proof_to_solana_bytes(a_raw: bytes, b_raw: bytes, c_raw: bytes) -> bytes:
    """prover.rs:304-334: (-A).x || (-A).y || B.x.c0 || B.x.c1 || B.y.c0 || B.y.c1 || C.x || C.y, 32-byte LE each = 256 B.
    Inputs are the raw affine outputs of zkb_prove (A not negated)."""
    if len(a_raw) != 64 or len(b_raw) != 128 or len(c_raw) != 64:
        raise ValueError("proof components must be 64 / 128 / 64 bytes")
    if a_raw == bytes(64):
        neg_a = a_raw                                  # -infinity = infinity
    else:
        y = int.from_bytes(a_raw[32:], "little")
        neg_a = a_raw[:32] + ((FQ_MODULUS - y) % FQ_MODULUS).to_bytes(32, "little")
    return neg_a + b_raw + c_raw


# ----------------------------------------------------------------------------- the prover
class Groth16Prover:
    """`Groth16Prover: BatchProver` (prover.rs:252-447) on one B200.  Not thread-safe (one prove in flight, as
    pipeline.rs:369-375 guarantees); make one per GPU."""

    def __init__(self, ctx, pk, vk_bytes: bytes, synthesizer: Optional[Callable] = None):
        self.ctx, self.pk, self.vk_bytes, self.synthesizer = ctx, pk, bytes(vk_bytes), synthesizer
        self._vk_hash = None
        self._shapes = {}

    @classmethod
    def from_bytes(cls, pk_bytes: bytes, vk_bytes: bytes, device: int = 0, synthesizer=None) -> "Groth16Prover":
        """prover.rs:263-277: both keys are ark-serialize COMPRESSED; the proving key is decompressed + validated on the GPU.
        Raises ZkbError (the anyhow::Error of the reference) on a malformed key."""
        from .api import Context
        ctx = Context(device)
        try:
            pk = ctx.proving_key_compressed(pk_bytes, validate=True)
        except Exception:
            ctx.close()
            raise
        return cls(ctx, pk, vk_bytes, synthesizer)

    @classmethod
    def from_files(cls, pk_path: str, vk_path: str, device: int = 0, synthesizer=None) -> "Groth16Prover":
        """prover.rs:280-286"""
        with open(pk_path, "rb") as f:
            pk_bytes = f.read()
        with open(vk_path, "rb") as f:
            vk_bytes = f.read()
        return cls.from_bytes(pk_bytes, vk_bytes, device, synthesizer)

    def verification_key_hash(self) -> bytes:
        """prover.rs:289-294: blake3 of the compressed verifying key."""
        if self._vk_hash is None:
            import blake3
            self._vk_hash = blake3.blake3(self.vk_bytes).digest()
        return self._vk_hash

    def circuit(self, num_instance, num_witness, a, b, c, key=None):
        """Upload (and cache per shape) the constraint matrices `cs.to_matrices()` yields."""
        key = key if key is not None else (num_instance, num_witness, len(a[0]) if isinstance(a, tuple) else len(a))
        m = self._shapes.get(key)
        if m is None:
            m = self._shapes[key] = self.ctx.r1cs(num_instance, num_witness, a, b, c)
        return m

    def prove_assignment(self, inputs: BatchPublicInputs, matrices, full_assignment: bytes) -> BatchProof:
        """prover.rs:350-425 after constraint synthesis: seed StdRng with the batch id, draw r then s (what Groth16::prove
        does first), run the GPU prover, format the 256-byte Solana proof."""
        start = time.perf_counter()
        rng = StdRng.seed_from_u64(inputs.batch_id & _M64)
        r = fr_rand(rng)
        s = fr_rand(rng)
        a, b, c = self.ctx.prove(self.pk, matrices, full_assignment, r.to_bytes(32, "little"), s.to_bytes(32, "little"))
        proof_bytes = proof_to_solana_bytes(a, b, c)
        return BatchProof(inputs, proof_bytes, int((time.perf_counter() - start) * 1000))

    def prove(self, inputs: BatchPublicInputs, witness) -> BatchProof:
        """prover.rs:350-425.  `witness`: an L2BlockCircuit carrying the private fields (its own public fields are replaced
        by `inputs`), or whatever the custom `synthesizer` understands."""
        if self.synthesizer is not None:
            ni, nw, a, b, c, z = self.synthesizer(inputs, witness)
            return self.prove_assignment(inputs, self.circuit(ni, nw, a, b, c), z)
        from .l2_circuit import L2BlockCircuit, L2Prover
        if not isinstance(witness, L2BlockCircuit):
            raise TypeError("Groth16Prover.prove: witness must be an L2BlockCircuit (or pass synthesizer=...)")
        l2 = self._shapes.get("l2")
        if l2 is None:
            from .l2_circuit import L2Circuit
            # the shape the key was generated for: L2BlockCircuit::dummy() in the reference (keygen.rs:84); a key made for
            # another shape is served by passing a witness of that shape first
            l2 = self._shapes["l2"] = L2Prover(self.ctx, L2Circuit(witness), self.pk, self.vk_bytes)
        return l2.prove(inputs, witness)

    def verify(self, proof: BatchProof) -> bool:
        """prover.rs:427-442: the reference only checks that the proof is well-formed (exactly 256 bytes)."""
        return len(proof.proof_bytes) == 256

    def close(self):
        if self.ctx is not None:
            self.pk.free()
            for m in self._shapes.values():
                m.close() if hasattr(m, "close") else m.free()
            self.ctx.close()
            self.ctx = None
