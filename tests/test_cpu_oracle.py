"""Pins the C++ CPU restatement (oracle/cpu_oracle.cpp: checker for big sizes + timed CPU baseline) against the
Python oracle, which is itself pinned by the reference's committed fixtures (tests/test_oracle_kat.py)."""
import json
import os
import random

import pytest

from conftest import REF_FIXTURES
from helpers import R, P, fr_bytes, unpack32, g1_raw, g2_raw, arithmetic_bases, mimc7_chain, pk_parts
from oracle import bn254 as bn
from oracle import cpu as orc
from oracle import groth16 as g16
from oracle import rng as orng
from oracle.rng import StdRng


@pytest.mark.parametrize("field,mod", [(0, R), (1, P)])
def test_field_ops(field, mod):
    rnd = random.Random(3 + field)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, mod >> 1, 1 << 253]
    a = [x for x in edge for _ in edge] + [rnd.randrange(mod) for _ in range(500)]
    b = [y for _ in edge for y in edge] + [rnd.randrange(mod) for _ in range(500)]
    pack = lambda v: b"".join(int(x).to_bytes(32, "little") for x in v)
    A, B = pack(a), pack(b)
    assert unpack32(orc.field_op(field, 0, A, B)) == [(x + y) % mod for x, y in zip(a, b)]
    assert unpack32(orc.field_op(field, 1, A, B)) == [(x - y) % mod for x, y in zip(a, b)]
    assert unpack32(orc.field_op(field, 2, A, B)) == [(x * y) % mod for x, y in zip(a, b)]
    assert unpack32(orc.field_op(field, 4, A)) == [(-x) % mod for x in a]
    assert unpack32(orc.field_op(field, 3, A[:32 * 100])) == [pow(x, -1, mod) if x else 0 for x in a[:100]]


def test_window_formula_matches_arkworks_table():
    # 2^16 -> 13, 2^21 -> 16, 2^24 -> 18, 2^26 -> 19 ; n < 32 -> 3
    assert [orc.msm_window(1 << k) for k in (16, 21, 24, 26)] == [13, 16, 18, 19]
    assert orc.msm_window(31) == 3 and orc.msm_window(32) == 5


@pytest.mark.parametrize("n", [0, 1, 5, 31, 32, 300, 2000])
def test_msm_g1(n):
    rnd = random.Random(50 + n)
    pts, ks = arithmetic_bases(bn.G1, bn.G1_GEN, n, rnd.randrange(R), rnd.randrange(R))
    sc = [rnd.randrange(R) for _ in range(n)]
    if n >= 5:
        sc[:4] = [0, 1, R - 1, 1 << 253]
        pts[4] = None
    exp = bn.G1.mul(bn.G1_GEN, sum(k * s for k, s in zip(ks, sc) if True) % R) if n else None
    if n >= 5:
        exp = bn.G1.mul(bn.G1_GEN, sum(k * s for i, (k, s) in enumerate(zip(ks, sc)) if i != 4) % R)
    assert orc.msm_g1(g1_raw(pts), fr_bytes(sc)) == bn.g1_to_raw(exp)
    if 0 < n <= 32:
        assert orc.msm_g1(g1_raw(pts), fr_bytes(sc)) == bn.g1_to_raw(bn.G1.msm_naive(pts, sc))


def test_msm_g1_cancellation_and_doubling():
    g = bn.G1_GEN
    pts = [g, g, bn.G1.neg(g), bn.G1.mul(g, 9)]
    assert orc.msm_g1(g1_raw(pts), fr_bytes([3, 3, 6, 0])) == bytes(64)
    assert orc.msm_g1(g1_raw(pts), fr_bytes([5, 5, 0, 1])) == bn.g1_to_raw(bn.G1.mul(g, 19))


@pytest.mark.parametrize("n", [1, 40, 400])
def test_msm_g2(n):
    rnd = random.Random(70 + n)
    pts, ks = arithmetic_bases(bn.G2, bn.G2_GEN, n, rnd.randrange(R), rnd.randrange(R))
    sc = [rnd.randrange(R) for _ in range(n)]
    exp = bn.G2.mul(bn.G2_GEN, sum(k * s for k, s in zip(ks, sc)) % R)
    assert orc.msm_g2(g2_raw(pts), fr_bytes(sc)) == bn.g2_to_raw(exp)


def test_arithmetic_bases_and_preparsed_msm():
    n = 5000
    b = orc.G1Bases.arithmetic(7, n)
    raw = b.read(0, 3) + b.read(n - 1, 1)
    assert raw == g1_raw([bn.G1.mul(bn.G1_GEN, k) for k in (7, 8, 9, 7 + n - 1)])
    rnd = random.Random(9)
    sc = [rnd.randrange(R) for _ in range(n)]
    exp = bn.G1.mul(bn.G1_GEN, sum((7 + i) * s for i, s in enumerate(sc)) % R)
    assert b.msm(fr_bytes(sc)) == bn.g1_to_raw(exp)
    assert b.msm(fr_bytes(sc[100:900]), off=100) == bn.g1_to_raw(
        bn.G1.mul(bn.G1_GEN, sum((7 + i) * sc[i] for i in range(100, 900)) % R))


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9])
def test_ntt_all_four_variants(log_n):
    rnd = random.Random(log_n)
    v = [rnd.randrange(R) for _ in range(1 << log_n)]
    data = fr_bytes(v)
    assert unpack32(orc.ntt(data, log_n)) == g16.fft(v)
    assert unpack32(orc.ntt(data, log_n, inverse=True)) == g16.ifft(v)
    assert unpack32(orc.ntt(data, log_n, coset=True)) == g16.coset_fft(v)
    assert unpack32(orc.ntt(data, log_n, inverse=True, coset=True)) == g16.coset_ifft(v)


def test_witness_map_and_prove_mimc():
    r1cs, z = mimc7_chain(num_perm=2, seed=42, rounds=20)
    pk = g16.circuit_specific_setup(r1cs, StdRng.seed_from_u64(0))
    m = orc.R1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    assert m.log_domain == 8
    assert unpack32(orc.witness_map(m, fr_bytes(z))) == g16.witness_map_from_matrices(r1cs, z)
    cpk = orc.ProvingKey(**pk_parts(pk))
    for r, s in ((123456789, 987654321), (0, 5), (R - 1, 0)):
        ref = g16.prove_with_rs(pk, r1cs, z, r, s)
        got = orc.prove(cpk, m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
        assert got == (bn.g1_to_raw(ref.a), bn.g2_to_raw(ref.b), bn.g1_to_raw(ref.c))


def test_prove_square_circuit_reproduces_reference_fixture():
    """onchain-programs/verifier/proof_for_onchain.json from the C++ restatement, byte for byte."""
    r1cs, z = g16.square_circuit(7)
    rng = StdRng.seed_from_u64(42)
    pk = g16.circuit_specific_setup(r1cs, rng)
    r, s = orng.rand_fr(rng), orng.rand_fr(rng)
    m = orc.R1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    a, b, c = orc.prove(orc.ProvingKey(**pk_parts(pk)), m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
    proof = g16.Proof(bn.g1_from_raw(a), bn.g2_from_raw(b), bn.g1_from_raw(c))
    pc = json.load(open(os.path.join(REF_FIXTURES, "proof_for_onchain.json")))["proof_components"]
    assert proof.serialize_uncompressed() == bytes(pc["pi_a"]) + bytes(pc["pi_b"]) + bytes(pc["pi_c"])


def test_dot_mod_r_helper_is_exact():
    """tests/helpers.py::dot_mod_r (the exact sum k_i s_i mod r behind the 2^24 known-dlog MSM check) against Python integers."""
    import numpy as np
    from helpers import dot_mod_r
    from oracle import bn254 as bn

    def rnd(n, seed):
        rs = np.random.RandomState(seed)
        a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        a[:, 7] %= 0x30644E72
        return a

    k, s = rnd(3000, 1), rnd(3000, 2)
    ints = lambda a: [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]
    assert dot_mod_r(k, s) == sum(x * y for x, y in zip(ints(k), ints(s))) % bn.R


def test_poly_eval_matches_ntt_definition():
    """oracle poly_eval (Horner, shares no code with the FFT) against the C++ FFT and the Python oracle's root of unity."""
    import random
    from oracle import bn254 as bn
    R = bn.R
    lg = 9
    n = 1 << lg
    rnd = random.Random(5)
    a = [rnd.randrange(R) for _ in range(n)]
    data = b"".join(x.to_bytes(32, "little") for x in a)
    w = orc.root_of_unity(lg)
    assert pow(w, n, R) == 1 and pow(w, n // 2, R) == R - 1
    f = orc.ntt(data, lg, coset=True)
    ks = [0, 1, 77, n - 1]
    for k, e in zip(ks, orc.poly_eval(data, [5 * pow(w, k, R) % R for k in ks])):
        assert int.from_bytes(f[32 * k:32 * k + 32], "little") == e == sum(c * pow(5 * pow(w, k, R), j, R) for j, c in enumerate(a)) % R
