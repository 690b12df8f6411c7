"""L2BlockCircuit (prover/src/l2_circuit.rs): the oracle restatement's own invariants, and the native synthesiser in
libzkb200.so against it, bit for bit.  Host logic only -- no GPU needed.  (GPU: tests/test_gpu_parity.py::test_l2_*.)

The reference pins nothing for this circuit except the instance count (l2_circuit.rs:512-541); the oracle is therefore
"parity unpinned" against arkworks and these tests pin (a) its internal consistency and (b) product == oracle."""
import ctypes as C
import random

import pytest

from oracle import l2_circuit as O
from oracle import rng as orng
from oracle.bn254 import R

from helpers import unpack32


def _k(i):
    return bytes([i]) * 32


def _shapes():
    """name -> (transactions, initial_accounts, commitments, withdrawals, batch_id)"""
    rnd = random.Random(5)
    rk = lambda: bytes(rnd.randrange(256) for _ in range(32))
    a, b, c, d = rk(), rk(), rk(), rk()
    return {
        "dummy": ([(_k(1), _k(2), 100)], {_k(1): 1000, _k(2): 0}, [], [], 0),
        "no_transfers": ([], {_k(9): 5}, [], [], 3),
        "creates_recipient_and_spends_from_it": ([(a, b, 7), (b, c, 7), (a, a, 1)], {a: 50, d: 2 ** 63}, [], [], 11),
        "commitments_and_withdrawals": ([(_k(1), _k(2), 1000)], {_k(2): 0, _k(1): 1000}, [rk(), bytes([255]) * 32],
                                        [(rk(), 12345), (rk(), 0)], 2 ** 64 - 1),
    }


def _oracle_circuit(shape, satisfy=True):
    txs, acc, com, wd, bid = shape
    c = O.L2BlockCircuit(transactions=txs, initial_accounts=acc, shielded_commitments=com, withdrawals=wd, batch_id=bid,
                         pre_shielded_root=bytes(range(32)))
    return O.with_satisfying_roots(c) if satisfy else c


def _product_circuit(shape, oc):
    from zelana_b200 import l2_circuit as P
    txs, acc, com, wd, bid = shape
    return P.L2BlockCircuit(oc.pre_state_root, oc.post_state_root, oc.pre_shielded_root, oc.post_shielded_root,
                            oc.withdrawal_root, oc.batch_hash, bid,
                            [P.TransactionWitness(*t) for t in txs], dict(acc),
                            [P.ShieldedCommitmentWitness(x) for x in com], [P.WithdrawalWitness(*x) for x in wd])


def _rows(csr):
    rp, col, co = csr
    data = co.tobytes()
    return [[(int.from_bytes(data[32 * k:32 * k + 32], "little"), int(col[k])) for k in range(int(rp[i]), int(rp[i + 1]))]
            for i in range(len(rp) - 1)]


# ----------------------------------------------------------------------------- oracle invariants
def test_poseidon_parameters_have_the_generator_s_shape():
    cfg = O.get_poseidon_config()
    assert len(cfg.ark) == 64 and all(len(r) == 3 and all(0 <= x < R for x in r) for r in cfg.ark)
    assert len({x for r in cfg.ark for x in r}) == 192
    # Cauchy matrix 1 / (x_i + y_j): every 2x2 minor of the entrywise inverse has the rank-2 additive structure
    inv = [[pow(v, -1, R) for v in row] for row in cfg.mds]
    for i in range(2):
        for j in range(2):
            assert (inv[i][j] - inv[i + 1][j] - inv[i][j + 1] + inv[i + 1][j + 1]) % R == 0
    # and it is invertible (MDS)
    m = cfg.mds
    det = (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
           + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0])) % R
    assert det != 0


def test_grain_lfsr_is_the_80_bit_register_of_the_poseidon_paper():
    g = O.GrainLFSR(False, 254, 3, 8, 56)
    # after the 160 warm-up clocks the register must not be stuck and must follow its recurrence
    before = list(g.state[g.head:] + g.state[:g.head])
    nb = g.update()
    assert nb == before[62] ^ before[51] ^ before[38] ^ before[23] ^ before[13] ^ before[0]
    bits = g.get_bits(512)
    assert 150 < sum(bits) < 362


def test_dummy_circuit_counts_match_the_reference_s_own_test():
    r1cs, z = O.synthesize(O.L2BlockCircuit.dummy())
    assert r1cs.num_instance == 8                      # l2_circuit.rs:527-531, :540
    assert len(z) == r1cs.num_instance + r1cs.num_witness
    assert (r1cs.num_constraints, r1cs.num_witness) == (6415, 5958)
    assert 4096 < r1cs.num_constraints + r1cs.num_instance <= 8192     # QAP domain 2^13
    # "This will fail because the dummy values don't actually compute to valid roots" (l2_circuit.rs:518-519)
    assert not r1cs.is_satisfied(z)


@pytest.mark.parametrize("name", list(_shapes()))
def test_oracle_is_satisfied_exactly_by_the_offchain_roots(name):
    shape = _shapes()[name]
    oc = _oracle_circuit(shape)
    r1cs, z = O.synthesize(oc)
    assert r1cs.is_satisfied(z)
    assert z[1:8] == [O.fr_from_le_bytes_mod_order(b) for b in (oc.pre_state_root, oc.post_state_root, oc.pre_shielded_root,
                                                                oc.post_shielded_root, oc.withdrawal_root, oc.batch_hash)] \
        + [oc.batch_id % R]
    for attr in ("pre_state_root", "post_state_root", "withdrawal_root", "batch_hash"):
        bad = _oracle_circuit(shape)
        setattr(bad, attr, bytes(31) + b"\x01")
        r2, z2 = O.synthesize(bad)
        assert (r2.a, r2.b, r2.c) == (r1cs.a, r1cs.b, r1cs.c)          # structure does not depend on values
        assert not r2.is_satisfied(z2)


def test_overspending_transfer_is_unsatisfiable():
    c = O.L2BlockCircuit(transactions=[(_k(1), _k(2), 1001)], initial_accounts={_k(1): 1000, _k(2): 0})
    # roots of the (wrapped-around) balances so that only the comparison can fail
    final = {_k(1): (1000 - 1001) % R, _k(2): 1001}
    c.pre_state_root = O.accounts_root(0, c.initial_accounts)
    c.post_state_root = O.accounts_root(0, final)
    c.withdrawal_root, c.batch_hash = O.withdrawal_root([]), O.batch_hash(0, c.transactions)
    r1cs, z = O.synthesize(c)
    bad = [i for i, (ra, rb, rc) in enumerate(zip(r1cs.a, r1cs.b, r1cs.c))
           if sum(co * z[v] for co, v in ra) * sum(co * z[v] for co, v in rb) % R != sum(co * z[v] for co, v in rc) % R]
    assert bad and max(bad) < 2000                     # inside enforce_cmp, before the first Poseidon fold
    ok = O.L2BlockCircuit(transactions=[(_k(1), _k(2), 1000)], initial_accounts={_k(1): 1000, _k(2): 0})
    r1cs, z = O.synthesize(O.with_satisfying_roots(ok))
    assert r1cs.is_satisfied(z)                        # equality allowed (enforce_cmp(.., Greater, true))


def test_gadget_sponge_equals_native_sponge():
    cs = O.ConstraintSystem()
    xs = [O.Fp.new_witness(cs, v) for v in (5, R - 1, 12345678901234567890)]
    for n in (1, 2, 3):
        s = O.PoseidonSpongeVar(cs, O.get_poseidon_config())
        s.absorb(xs[:n])
        assert s.squeeze_field_elements(1)[0].value == O.poseidon_hash([x.value for x in xs[:n]])
    r1cs = cs.to_r1cs()
    assert r1cs.is_satisfied(cs.assignment())
    # 240 constraints per permutation once every state element is a variable; the first round's constant lanes are free
    assert r1cs.num_constraints == 4 * 240 - 3 * (2 + 1 + 1)


# ----------------------------------------------------------------------------- native synthesiser == oracle
@pytest.mark.parametrize("name", list(_shapes()))
def test_native_matrices_and_assignment_equal_the_oracle(name):
    from zelana_b200 import l2_circuit as P
    shape = _shapes()[name]
    oc = _oracle_circuit(shape)
    r1cs, z = O.synthesize(oc)
    pc = _product_circuit(shape, oc)
    circ = P.L2Circuit(pc)
    assert (circ.num_constraints, circ.num_instance, circ.num_witness) == (r1cs.num_constraints, 8, r1cs.num_witness)
    a, b, c = circ.matrices()
    assert _rows(a) == r1cs.a and _rows(b) == r1cs.b and _rows(c) == r1cs.c
    zb = circ.assign(pc)
    assert unpack32(zb) == z
    assert circ.is_satisfied(zb) == (True, None)
    # the off-circuit roots (main.rs.bak:114-154) from the library equal the oracle's
    got = P.satisfying_inputs(pc)
    assert (got.pre_state_root, got.post_state_root, got.post_shielded_root, got.withdrawal_root, got.batch_hash) == \
        (oc.pre_state_root, oc.post_state_root, oc.post_shielded_root, oc.withdrawal_root, oc.batch_hash)
    # unsatisfying inputs are assigned all the same (release-mode arkworks does not check) and reported by is_satisfied
    bad = pc.with_inputs(P.BatchPublicInputs(batch_id=pc.batch_id))
    zbad = circ.assign(bad)
    r0, z0 = O.synthesize(O.L2BlockCircuit(transactions=shape[0], initial_accounts=shape[1], shielded_commitments=shape[2],
                                           withdrawals=shape[3], batch_id=shape[4]))
    assert unpack32(zbad) == z0
    ok, row = circ.is_satisfied(zbad)
    assert not ok and row == next(i for i in range(r0.num_constraints) if not O.R1CS(
        r0.num_instance, r0.num_witness, [r0.a[i]], [r0.b[i]], [r0.c[i]]).is_satisfied(z0))


def test_account_order_does_not_matter_but_shape_does():
    from zelana_b200 import ZkbError, l2_circuit as P
    d = P.L2BlockCircuit.dummy()
    circ = P.L2Circuit(d)
    rev = P.L2BlockCircuit(transactions=d.transactions, initial_accounts=dict(reversed(list(d.initial_accounts.items()))))
    assert circ.assign(rev) == circ.assign(d)          # BTreeMap order, whatever the insertion order
    other_amounts = P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(1), _k(2), 5)], initial_accounts={_k(1): 9, _k(2): 1})
    assert len(circ.assign(other_amounts)) == len(circ.assign(d))
    # different keys with the same debit/credit pattern are the same circuit
    renamed = P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(7), _k(8), 5)], initial_accounts={_k(7): 9, _k(8): 1})
    circ.assign(renamed)
    for bad in (P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(2), _k(1), 0)], initial_accounts=d.initial_accounts),
                P.L2BlockCircuit(transactions=d.transactions, initial_accounts={_k(1): 1000}),
                P.L2BlockCircuit(transactions=d.transactions, initial_accounts=d.initial_accounts,
                                 withdrawals=[P.WithdrawalWitness(_k(3), 1)])):
        with pytest.raises(ZkbError) as e:
            circ.assign(bad)
        assert e.value.code == -6                      # ZKB_ERR_SHAPE
    with pytest.raises(ZkbError) as e:                 # SynthesisError::AssignmentMissing (l2_circuit.rs:263-266)
        P.L2Circuit(P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(5), _k(1), 1)], initial_accounts={_k(1): 1}))
    assert e.value.code == -6
    with pytest.raises(ValueError):
        P.L2Circuit(P.L2BlockCircuit(initial_accounts={b"short": 1}))


def test_prover_randomness_is_stdrng_seeded_with_the_batch_id():
    from zelana_b200 import l2_circuit as P
    for batch_id in (0, 1, 42, 2 ** 64 - 1):
        rng = orng.StdRng.seed_from_u64(batch_id)
        r, s = orng.rand_fr(rng), orng.rand_fr(rng)
        pr, ps = P.prover_randomness(batch_id)
        assert (int.from_bytes(pr, "little"), int.from_bytes(ps, "little")) == (r, s)


def test_native_poseidon_hash_equals_oracle():
    from zelana_b200 import l2_circuit as P
    rnd = random.Random(8)
    for n in (0, 1, 2, 3):
        for _ in range(4):
            vals = [rnd.randrange(2 ** 256) for _ in range(n)]
            got = P.poseidon_hash([v.to_bytes(32, "little") for v in vals])
            assert int.from_bytes(got, "little") == O.poseidon_hash([v % R for v in vals])


def _mid_shape():
    """16 transfers over 24 accounts (some created by transfers), one commitment, two withdrawals: > 40 000 witness variables,
    i.e. above the threshold where the library assigns with parallel walkers."""
    rnd = random.Random(31)
    rk = lambda: bytes(rnd.randrange(256) for _ in range(32))
    keys = [rk() for _ in range(24)]
    acc = {k: 10 ** 6 + i for i, k in enumerate(keys[:20])}
    txs = [(keys[rnd.randrange(20)], keys[rnd.randrange(24)], rnd.randrange(1, 500)) for _ in range(16)]
    return txs, acc, [rk()], [(rk(), 77), (rk(), 1)], 123456789


def test_parallel_assignment_equals_the_oracle_and_the_sequential_pass():
    """Circuits above PAR_MIN_WITNESS are assigned by two passes of parallel walkers (leaf hashes + comparisons, then the five
    fold chains): the bytes must equal the oracle's assignment, and those of a single-threaded run of the same library."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from zelana_b200 import l2_circuit as P
    shape = _mid_shape()
    oc = _oracle_circuit(shape)
    r1cs, z = O.synthesize(oc)
    assert r1cs.num_witness > 40000 and r1cs.is_satisfied(z)
    pc = _product_circuit(shape, oc)
    circ = P.L2Circuit(pc)
    assert (circ.num_constraints, circ.num_witness) == (r1cs.num_constraints, r1cs.num_witness)
    zb = circ.assign(pc)
    assert unpack32(zb) == z
    assert circ.is_satisfied(zb) == (True, None)
    # the same library, one thread (the environment variable is read once per process)
    code = ("import sys, pickle; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from zelana_b200 import l2_circuit as P\n"
            "pc = pickle.load(open(sys.argv[1], 'rb'))\n"
            "sys.stdout.buffer.write(P.L2Circuit(pc).assign(pc))\n") % (ROOT, os.path.join(ROOT, "tests"))
    import pickle
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        pickle.dump(pc, f)
    try:
        outs = {}
        for threads in ("1", "5"):
            env = dict(os.environ, ZKB_L2_ASSIGN_THREADS=threads)
            outs[threads] = subprocess.run([sys.executable, "-c", code, f.name], env=env, check=True, capture_output=True).stdout
        assert outs["1"] == zb and outs["5"] == zb
    finally:
        os.unlink(f.name)


def test_batch_prover_without_a_gpu_is_an_error_not_a_fallback():
    import torch
    import zelana_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = zelana_b200.load_library()
    h = C.c_void_p()
    assert lib.zkb_l2_batch_create(0, 4, C.byref(h)) == -1          # ZKB_ERR_NO_DEVICE from zkb_ctx_create
    assert not h.value
    assert lib.zkb_l2_batch_create(0, 0, C.byref(h)) == -3          # lanes outside 1..64
    assert lib.zkb_l2_batch_lanes(None) == 0
    out = (C.c_uint8 * 256)()
    assert lib.zkb_l2_prove(None, None, None, None, None, None, out) == -3


def test_poseidon_generator_reproduces_arkworks_known_answers():
    """The only golden vectors that exist for this stack: ark-crypto-primitives 0.5.0's own tests of the parameter generator
    `get_poseidon_config()` calls (Cargo.lock pins the crate; it is not vendored).  They are stated over BLS12-381's scalar
    field: src/sponge/poseidon/grain_lfsr.rs `test_grain_lfsr_consistency` (PoseidonGrainLFSR::new(false, 255, 3, 8, 31): two
    rejection-sampled, then two mod-p elements) and src/sponge/poseidon/traits.rs
    `bls12_381_fr_poseidon_default_parameters_test` (rate 2 optimised-for-constraints entry = (alpha 17, R_F 8, R_P 31,
    skip 0): ark[0][0] and mds[0][0])."""
    q = 52435875175126190479447740508185965837690552500527637822603658699938581184513      # BLS12-381 Fr
    g = O.GrainLFSR(False, 255, 3, 8, 31)
    assert g.get_field_elements_rejection_sampling(1, q) == [
        27117311055620256798560880810000042840428971800021819916023577129547249660720]
    assert g.get_field_elements_rejection_sampling(1, q) == [
        51641662388546346858987925410984003801092143452466182801674685248597955169158]
    assert g.get_field_elements_mod_p(1, q) == [
        30468495022634911716522728179277518871747767531215914044579216845399211650580]
    assert g.get_field_elements_mod_p(1, q) == [
        17250718238509906485015112994867732544602358855445377986727968022920517907825]
    ark, mds = O.find_poseidon_ark_and_mds(255, 2, 8, 31, 0, modulus=q)
    assert ark[0][0] == 27117311055620256798560880810000042840428971800021819916023577129547249660720
    assert mds[0][0] == 26017457457808754696901916760153646963713419596921330311675236858336250747575
    assert len(ark) == 39 and all(len(r) == 3 for r in ark)


def test_dummy_circuit_matches_the_committed_golden_digest():
    """tests/golden/l2_dummy_circuit.json (made by tests/golden/make_l2_golden.py) freezes the oracle's dummy circuit: counts,
    roots, SHA-256 of the matrices and of the assignment -- and the native synthesiser must hash to the same values."""
    import hashlib
    import importlib.util
    import json
    import os
    import struct
    from conftest import GOLDEN
    from zelana_b200 import l2_circuit as P
    spec = importlib.util.spec_from_file_location("make_l2_golden", os.path.join(GOLDEN, "make_l2_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    want = json.load(open(os.path.join(GOLDEN, "l2_dummy_circuit.json")))
    assert mk.build() == want
    pc = P.L2BlockCircuit.dummy()
    pc = pc.with_inputs(P.satisfying_inputs(pc))
    circ = P.L2Circuit(pc)
    h = hashlib.sha256()
    for rp, col, co in circ.matrices():
        data = co.tobytes()
        for i in range(len(rp) - 1):
            lo, hi = int(rp[i]), int(rp[i + 1])
            h.update(struct.pack("<I", hi - lo))
            for k in range(lo, hi):                      # the native rows are already sorted by column
                h.update(struct.pack("<I", int(col[k])) + data[32 * k:32 * k + 32])
    assert h.hexdigest() == want["matrices_sha256"]
    z = circ.assign(pc)
    assert hashlib.sha256(z).hexdigest() == want["assignment_sha256"]
    assert [z[32 * i:32 * i + 32].hex() for i in range(1, 8)] == want["public_inputs_le_hex"]


def test_roots_edge_cases_zero_amounts_u64_limits_and_refusals():
    from zelana_b200 import ZkbError, l2_circuit as P
    top = 2 ** 64 - 1
    ok = P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(1), _k(2), 0), P.TransactionWitness(_k(1), _k(2), top)],
                          initial_accounts={_k(1): top, _k(2): 0}, batch_id=top)
    ok = ok.with_inputs(P.satisfying_inputs(ok))
    circ = P.L2Circuit(ok)
    assert circ.is_satisfied(circ.assign(ok)) == (True, None)
    oc = O.with_satisfying_roots(O.L2BlockCircuit(transactions=[(_k(1), _k(2), 0), (_k(1), _k(2), top)],
                                                  initial_accounts={_k(1): top, _k(2): 0}, batch_id=top))
    assert (ok.pre_state_root, ok.post_state_root, ok.batch_hash) == (oc.pre_state_root, oc.post_state_root, oc.batch_hash)
    for bad in (P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(1), _k(2), 6)], initial_accounts={_k(1): 5, _k(2): 0}),
                P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(3), _k(2), 1)], initial_accounts={_k(1): 5, _k(2): 0}),
                P.L2BlockCircuit(transactions=[P.TransactionWitness(_k(1), _k(2), 2)], initial_accounts={_k(1): 5, _k(2): top})):
        with pytest.raises(ZkbError) as e:              # overspend / unknown sender / recipient beyond 64 bits
            P.satisfying_inputs(bad)
        assert e.value.code == -3


def test_exported_poseidon_parameters_equal_the_oracle_generator():
    """zkb_l2_poseidon_params (what the GPU hash kernel uploads) == oracle find_poseidon_ark_and_mds(254, 2, 8, 56)."""
    import ctypes as C
    import zelana_b200
    lib = zelana_b200.load_library()
    ark, mds = C.create_string_buffer(64 * 3 * 32), C.create_string_buffer(9 * 32)
    assert lib.zkb_l2_poseidon_params(ark, mds) == 0
    cfg = O.get_poseidon_config()
    want_ark = b"".join(int(v).to_bytes(32, "little") for row in cfg.ark for v in row)
    want_mds = b"".join(int(v).to_bytes(32, "little") for row in cfg.mds for v in row)
    assert ark.raw == want_ark and mds.raw == want_mds
    # host batch entry == one-at-a-time entry
    import random
    rnd = random.Random(4)
    vals = [rnd.randrange(1 << 256) for _ in range(3 * 50)]
    buf = b"".join(v.to_bytes(32, "little") for v in vals)
    out = C.create_string_buffer(50 * 32)
    assert lib.zkb_l2_poseidon_hash_batch_host(3, buf, 50, 4, out) == 0
    one = C.create_string_buffer(32)
    for i in (0, 17, 49):
        assert lib.zkb_l2_poseidon_hash(buf[96 * i:96 * i + 96], 3, one) == 0
        assert out.raw[32 * i:32 * i + 32] == one.raw


@pytest.mark.gpu
def test_gpu_poseidon_hash_batch_equals_host_hashes():
    """zkb_l2_poseidon_hash_batch (GPU, n independent leaf hashes) == the native host sponge, arities 0..3, including inputs
    >= r (reduced mod r as Fr::from_le_bytes_mod_order does) -- and == the oracle's PoseidonSponge for a few."""
    import ctypes as C
    import random
    import zelana_b200
    lib = zelana_b200.load_library()
    ctx = zelana_b200.Context(0)
    rnd = random.Random(6)
    for arity in (1, 2, 3):
        n = 1000
        vals = [rnd.randrange(1 << 256) if i % 7 == 0 else rnd.randrange(R) for i in range(n * arity)]
        vals[:arity] = [0] * arity
        buf = b"".join(v.to_bytes(32, "little") for v in vals)
        host = C.create_string_buffer(n * 32)
        assert lib.zkb_l2_poseidon_hash_batch_host(arity, buf, n, 4, host) == 0
        assert ctx.l2_poseidon_hash_batch(arity, buf) == host.raw
        for i in (1, 500):
            assert int.from_bytes(host.raw[32 * i:32 * i + 32], "little") == O.poseidon_hash([v % R for v in vals[arity * i:arity * i + arity]])
    assert len(ctx.l2_poseidon_hash_batch(0, b"", n=3)) == 96
    ctx.close()
