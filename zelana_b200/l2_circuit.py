"""Host-side mirror of the reference's `prover` crate types for the L2 batch circuit, over the native synthesiser in
libzkb200.so (zelana_b200/csrc/l2_circuit.cpp).

Mirrors, with the same names and argument meaning:
  prover/src/l2_circuit.rs:38-62      TransactionWitness, ShieldedCommitmentWitness, WithdrawalWitness
  prover/src/l2_circuit.rs:92-170     L2BlockCircuit { 7 public inputs, transactions, initial_accounts, shielded_commitments,
                                      withdrawals }, L2BlockCircuit::dummy()
  prover/src/l2_circuit.rs:179-505    generate_constraints            -> L2Circuit.assign / .matrices (native)
  prover/src/bin/keygen.rs:81-131     keygen: dummy circuit, StdRng::seed_from_u64(0), circuit_specific_setup -> keygen()
  core/src/sequencer/settlement/prover.rs:350-425   Groth16Prover::prove -> L2Prover.prove (zkb_l2_prove)
Nothing here computes: every field operation is in the C++/CUDA library.
"""
import ctypes as C
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from ._lib import L2PublicInputs, L2Witness, R1csDesc, ZkbError, load_library
from .prover import BatchProof, BatchPublicInputs, StdRng


@dataclass
class TransactionWitness:
    """l2_circuit.rs:42-47"""
    sender_pk: bytes
    recipient_pk: bytes
    amount: int


@dataclass
class ShieldedCommitmentWitness:
    """l2_circuit.rs:50-53"""
    commitment: bytes


@dataclass
class WithdrawalWitness:
    """l2_circuit.rs:56-60"""
    recipient: bytes
    amount: int


@dataclass
class L2BlockCircuit:
    """l2_circuit.rs:92-120.  The Poseidon configuration (get_poseidon_config, :68-83) is fixed inside the library."""
    pre_state_root: bytes = bytes(32)
    post_state_root: bytes = bytes(32)
    pre_shielded_root: bytes = bytes(32)
    post_shielded_root: bytes = bytes(32)
    withdrawal_root: bytes = bytes(32)
    batch_hash: bytes = bytes(32)
    batch_id: int = 0
    transactions: List[TransactionWitness] = field(default_factory=list)
    initial_accounts: Dict[bytes, int] = field(default_factory=dict)
    shielded_commitments: List[ShieldedCommitmentWitness] = field(default_factory=list)
    withdrawals: List[WithdrawalWitness] = field(default_factory=list)

    @classmethod
    def dummy(cls) -> "L2BlockCircuit":
        """l2_circuit.rs:147-170"""
        return cls(transactions=[TransactionWitness(bytes([1] * 32), bytes([2] * 32), 100)],
                   initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): 0})

    def public_inputs(self) -> BatchPublicInputs:
        return BatchPublicInputs(self.pre_state_root, self.post_state_root, self.pre_shielded_root, self.post_shielded_root,
                                 self.withdrawal_root, self.batch_hash, self.batch_id)

    def with_inputs(self, p: BatchPublicInputs) -> "L2BlockCircuit":
        return L2BlockCircuit(p.pre_state_root, p.post_state_root, p.pre_shielded_root, p.post_shielded_root, p.withdrawal_root,
                              p.batch_hash, p.batch_id, self.transactions, self.initial_accounts, self.shielded_commitments,
                              self.withdrawals)


# ----------------------------------------------------------------------------- marshalling
def _k32(b) -> bytes:
    b = bytes(b)
    if len(b) != 32:
        raise ValueError("keys, roots and commitments are 32 bytes")
    return b


def _u64s(vals):
    arr = (C.c_uint64 * max(1, len(vals)))(*[int(v) for v in vals])
    return arr


def _c_witness(c: L2BlockCircuit):
    """-> (L2Witness, keep-alive list)"""
    w = L2Witness()
    pks = b"".join(_k32(k) for k in c.initial_accounts)
    bal = _u64s(list(c.initial_accounts.values()))
    snd = b"".join(_k32(t.sender_pk) for t in c.transactions)
    rcp = b"".join(_k32(t.recipient_pk) for t in c.transactions)
    amt = _u64s([t.amount for t in c.transactions])
    com = b"".join(_k32(s.commitment) for s in c.shielded_commitments)
    wdr = b"".join(_k32(x.recipient) for x in c.withdrawals)
    wda = _u64s([x.amount for x in c.withdrawals])
    w.account_pks, w.account_balances, w.n_accounts = pks, bal, len(c.initial_accounts)
    w.tx_senders, w.tx_recipients, w.tx_amounts, w.n_txs = snd, rcp, amt, len(c.transactions)
    w.commitments, w.n_commitments = com, len(c.shielded_commitments)
    w.wd_recipients, w.wd_amounts, w.n_withdrawals = wdr, wda, len(c.withdrawals)
    return w, [pks, bal, snd, rcp, amt, com, wdr, wda]


def _c_inputs(p: BatchPublicInputs) -> L2PublicInputs:
    x = L2PublicInputs()
    for name in ("pre_state_root", "post_state_root", "pre_shielded_root", "post_shielded_root", "withdrawal_root", "batch_hash"):
        getattr(x, name)[:] = _k32(getattr(p, name))
    x.batch_id = int(p.batch_id) & 0xFFFFFFFFFFFFFFFF
    return x


def _py_inputs(x: L2PublicInputs) -> BatchPublicInputs:
    return BatchPublicInputs(bytes(x.pre_state_root), bytes(x.post_state_root), bytes(x.pre_shielded_root),
                             bytes(x.post_shielded_root), bytes(x.withdrawal_root), bytes(x.batch_hash), int(x.batch_id))


def _check(rc):
    if rc != 0:
        raise ZkbError(rc, (load_library().zkb_l2_last_error() or b"").decode())


def poseidon_hash(elems) -> bytes:
    """Poseidon(get_poseidon_config()) of up to three 32-byte LE field elements."""
    out = (C.c_uint8 * 32)()
    buf = b"".join(_k32(e) for e in elems)
    _check(load_library().zkb_l2_poseidon_hash(buf, len(elems), out))
    return bytes(out)


def satisfying_inputs(c: L2BlockCircuit) -> BatchPublicInputs:
    """The public inputs a valid batch carries: the roots the circuit recomputes (main.rs.bak:93-154)."""
    w, _keep = _c_witness(c)
    out = L2PublicInputs()
    _check(load_library().zkb_l2_roots(C.byref(w), int(c.batch_id), _k32(c.pre_shielded_root), C.byref(out)))
    return _py_inputs(out)


def prover_randomness(batch_id: int):
    """(r, s) of `StdRng::seed_from_u64(batch_id)` + two `Fr::rand` (prover.rs:354), 32 B canonical each."""
    r, s = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
    _check(load_library().zkb_l2_prover_randomness(int(batch_id) & 0xFFFFFFFFFFFFFFFF, r, s))
    return bytes(r), bytes(s)


class L2Circuit:
    """The constraint matrices of one circuit shape (zkb_l2_circuit): what keygen fixes and every later batch must match."""

    def __init__(self, shape: L2BlockCircuit):
        self.lib = load_library()
        w, _keep = _c_witness(shape)
        h = C.c_void_p()
        _check(self.lib.zkb_l2_circuit_create(C.byref(w), C.byref(h)))
        self.h = h
        d = R1csDesc()
        _check(self.lib.zkb_l2_circuit_desc(h, C.byref(d)))
        self.desc = d
        self.num_constraints, self.num_instance, self.num_witness = int(d.num_constraints), int(d.num_instance), int(d.num_witness)

    def matrices(self):
        """(A, B, C) as CSR triples (row_ptr u64, col u32, coeff u8[nnz*32]) copied out of the library."""
        out = []
        for m in (self.desc.a, self.desc.b, self.desc.c):
            rp = np.ctypeslib.as_array(C.cast(m.row_ptr, C.POINTER(C.c_uint64)), (self.num_constraints + 1,)).copy()
            nnz = int(rp[-1])
            col = np.ctypeslib.as_array(C.cast(m.col, C.POINTER(C.c_uint32)), (max(nnz, 1),))[:nnz].copy()
            co = np.ctypeslib.as_array(C.cast(m.coeff, C.POINTER(C.c_uint8)), (max(nnz * 32, 1),))[:nnz * 32].copy()
            out.append((rp, col, co))
        return tuple(out)

    def assign(self, c: L2BlockCircuit) -> bytes:
        """Full assignment z = [1, public inputs, witness] of `c`, (num_instance + num_witness) x 32 B canonical."""
        w, _keep = _c_witness(c)
        x = _c_inputs(c.public_inputs())
        z = C.create_string_buffer((self.num_instance + self.num_witness) * 32)
        _check(self.lib.zkb_l2_circuit_assign(self.h, C.byref(x), C.byref(w), z))
        return z.raw

    def is_satisfied(self, z: bytes):
        want = (self.num_instance + self.num_witness) * 32
        if len(z) != want:
            raise ZkbError(-6, "assignment has %d bytes, this circuit needs %d" % (len(z), want))
        ok, row = C.c_int(0), C.c_uint64(0)
        _check(self.lib.zkb_l2_circuit_is_satisfied(self.h, z, C.byref(ok), C.byref(row)))
        return bool(ok.value), (None if ok.value else int(row.value))

    def free(self):
        if self.h:
            self.lib.zkb_l2_circuit_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def keygen(ctx, shape: Optional[L2BlockCircuit] = None, seed: int = 0):
    """prover/src/bin/keygen.rs:81-131: `Groth16::circuit_specific_setup(L2BlockCircuit::dummy(), StdRng::seed_from_u64(0))`
    with the QAP evaluation and fixed-base multiplications on the GPU (zkb_setup).  -> (L2Circuit, pk bytes, vk bytes, raw),
    the byte strings being the ark-serialize compressed forms keygen.rs writes, raw the uncompressed affine outputs."""
    from .keygen import circuit_specific_setup
    circ = L2Circuit(shape or L2BlockCircuit.dummy())
    a, b, c = circ.matrices()
    pk_bytes, vk_bytes, raw = circuit_specific_setup(ctx, circ.num_instance, circ.num_witness, a, b, c,
                                                     StdRng.seed_from_u64(seed))
    return circ, pk_bytes, vk_bytes, raw


class L2Prover:
    """`Groth16Prover: BatchProver` for the L2 circuit (prover.rs:252-447) on one B200: the key and the matrices are resident in
    HBM; `prove` is one C call (zkb_l2_prove) that assigns the witness on the host and proves on the GPU."""

    def __init__(self, ctx, circ: L2Circuit, pk, vk_bytes: bytes = b""):
        self.ctx, self.circ, self.pk, self.vk_bytes = ctx, circ, pk, bytes(vk_bytes)
        a, b, c = circ.matrices()
        self.m = ctx.r1cs(circ.num_instance, circ.num_witness, a, b, c)

    @classmethod
    def from_bytes(cls, pk_bytes: bytes, vk_bytes: bytes, device: int = 0, shape: Optional[L2BlockCircuit] = None) -> "L2Prover":
        """prover.rs:263-277; the circuit shape defaults to L2BlockCircuit::dummy(), the one keygen.rs uses."""
        from .api import Context
        ctx = Context(device)
        try:
            pk = ctx.proving_key_compressed(pk_bytes, validate=True)
            return cls(ctx, L2Circuit(shape or L2BlockCircuit.dummy()), pk, vk_bytes)
        except Exception:
            ctx.close()
            raise

    def prove_circuit(self, c: L2BlockCircuit) -> BatchProof:
        start = time.perf_counter()
        w, _keep = _c_witness(c)
        x = _c_inputs(c.public_inputs())
        out = (C.c_uint8 * 256)()
        rc = self.circ.lib.zkb_l2_prove(self.ctx.h, self.pk.h, self.m.h, self.circ.h, C.byref(x), C.byref(w), out)
        if rc != 0:
            raise ZkbError(rc, (self.circ.lib.zkb_l2_last_error() or b"").decode())
        return BatchProof(c.public_inputs(), bytes(out), int((time.perf_counter() - start) * 1000))

    def prove(self, inputs: BatchPublicInputs, witness: L2BlockCircuit) -> BatchProof:
        """prover.rs:350-425: `witness` carries the private fields (transactions, initial_accounts, withdrawals,
        shielded_commitments) that prove() extracts from BatchWitness; `inputs` the seven public values."""
        return self.prove_circuit(witness.with_inputs(inputs))

    def verify(self, proof: BatchProof) -> bool:
        """prover.rs:427-442"""
        return len(proof.proof_bytes) == 256

    def verification_key_hash(self) -> bytes:
        import blake3
        return blake3.blake3(self.vk_bytes).digest()

    def close(self):
        self.m.free()
        self.circ.free()


class L2BatchProver:
    """zkb_l2_batch: `lanes` contexts + host threads inside the library proving independent batches side by side on one GPU
    (BASELINE.json config 5).  One call proves a list of circuits; the key, matrices and circuit shape are shared."""

    def __init__(self, ctx, circ: L2Circuit, pk, lanes: int = 16, device: Optional[int] = None):
        self.ctx, self.circ, self.pk = ctx, circ, pk
        a, b, c = circ.matrices()
        self.m = ctx.r1cs(circ.num_instance, circ.num_witness, a, b, c)
        ctx.synchronize()
        h = C.c_void_p()
        _check(circ.lib.zkb_l2_batch_create(ctx.device if device is None else device, lanes, C.byref(h)))
        self.h = h
        self.lanes = int(circ.lib.zkb_l2_batch_lanes(h))

    def marshal(self, circuits: List[L2BlockCircuit]):
        """-> opaque argument pack for prove_marshalled (lets a caller keep the Python-side packing out of a timed region)."""
        n = len(circuits)
        xs, ws, keep = (L2PublicInputs * max(n, 1))(), (L2Witness * max(n, 1))(), []
        for i, c in enumerate(circuits):
            xs[i] = _c_inputs(c.public_inputs())
            w, k = _c_witness(c)
            ws[i] = w
            keep.append(k)
        return n, xs, ws, keep, C.create_string_buffer(256 * max(n, 1)), (C.c_int * max(n, 1))()

    def prove_marshalled(self, pack):
        n, xs, ws, _keep, out, status = pack
        rc = self.circ.lib.zkb_l2_batch_prove(self.h, self.pk.h, self.m.h, self.circ.h, xs, ws, n, out, status)
        if rc != 0:
            raise ZkbError(rc, (self.circ.lib.zkb_l2_last_error() or b"").decode())
        return [out.raw[256 * i:256 * i + 256] for i in range(n)]

    def prove(self, circuits: List[L2BlockCircuit]) -> List[bytes]:
        """256-byte Solana proofs, one per circuit, in order."""
        return self.prove_marshalled(self.marshal(circuits))

    def close(self):
        if self.h:
            self.circ.lib.zkb_l2_batch_destroy(self.h)
            self.h = None
        self.m.free()
