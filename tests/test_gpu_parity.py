"""GPU parity tests: every C-ABI entry point against the oracle, bit-exact (integer arithmetic)."""
import random

import pytest

from helpers import (R, P, fr_bytes, fq_bytes, unpack32, g1_raw, g2_raw, arithmetic_bases, mimc7_chain, pk_parts)
from oracle import bn254 as bn
from oracle import groth16 as g16
from oracle.rng import StdRng

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import zelana_b200
    c = zelana_b200.Context(0)
    yield c
    c.close()


# ----------------------------------------------------------------------------- field arithmetic (row a9)
@pytest.mark.parametrize("field,mod", [(0, R), (1, P)])
def test_field_ops(ctx, field, mod):
    rnd = random.Random(11 + field)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, (1 << 512) % mod, mod >> 1, (mod >> 1) + 1, 1 << 253]
    a = [x for x in edge for _ in edge] + [rnd.randrange(mod) for _ in range(1 << 14)]
    b = [y for _ in edge for y in edge] + [rnd.randrange(mod) for _ in range(1 << 14)]
    pack = lambda v: b"".join(int(x).to_bytes(32, "little") for x in v)
    A, B = pack(a), pack(b)
    assert unpack32(ctx.field_op(field, 0, A, B)) == [(x + y) % mod for x, y in zip(a, b)]
    assert unpack32(ctx.field_op(field, 1, A, B)) == [(x - y) % mod for x, y in zip(a, b)]
    assert unpack32(ctx.field_op(field, 2, A, B)) == [(x * y) % mod for x, y in zip(a, b)]
    assert unpack32(ctx.field_op(field, 4, A)) == [(-x) % mod for x in a]
    inv = unpack32(ctx.field_op(field, 3, A[:32 * 600]))
    assert inv == [pow(x, -1, mod) if x else 0 for x in a[:600]]


def test_field_rejects_non_canonical(ctx):
    import zelana_b200
    bad = int(R).to_bytes(32, "little")
    with pytest.raises(zelana_b200.ZkbError) as e:
        ctx.field_op(0, 0, bad, bad)
    assert e.value.code == -5


# ----------------------------------------------------------------------------- curve arithmetic
def test_scalar_mul_g1_g2(ctx):
    rnd = random.Random(5)
    pts1 = [bn.G1.mul(bn.G1_GEN, rnd.randrange(1, R)) for _ in range(24)] + [None, bn.G1_GEN]
    ks = [rnd.randrange(R) for _ in range(24)] + [5, 0]
    ks[0], ks[1], ks[2] = 0, 1, R - 1
    out = ctx.scalar_mul(1, g1_raw(pts1), fr_bytes(ks))
    exp = g1_raw([bn.G1.mul(p, k) for p, k in zip(pts1, ks)])
    assert out == exp
    pts2 = [bn.G2.mul(bn.G2_GEN, rnd.randrange(1, R)) for _ in range(10)] + [None]
    ks2 = [rnd.randrange(R) for _ in range(10)] + [7]
    ks2[0], ks2[1] = 1, R - 1
    out = ctx.scalar_mul(2, g2_raw(pts2), fr_bytes(ks2))
    assert out == g2_raw([bn.G2.mul(p, k) for p, k in zip(pts2, ks2)])


def test_point_sum_edge_cases(ctx):
    g = bn.G1_GEN
    p5 = bn.G1.mul(g, 5)
    cases = [
        [], [None], [g], [g, g], [g, bn.G1.neg(g)], [p5, p5, p5], [g, None, p5, bn.G1.neg(p5), bn.G1.neg(g)],
        [p5, bn.G1.mul(g, 7), bn.G1.mul(g, R - 12)],
    ]
    for pts in cases:
        acc = None
        for q in pts:
            acc = bn.G1.add(acc, q)
        assert ctx.point_sum(1, g1_raw(pts)) == bn.g1_to_raw(acc), pts
    h = bn.G2_GEN
    for pts in ([h, h], [h, bn.G2.neg(h)], [h, bn.G2.mul(h, 3), None]):
        acc = None
        for q in pts:
            acc = bn.G2.add(acc, q)
        assert ctx.point_sum(2, g2_raw(pts)) == bn.g2_to_raw(acc)


def test_bases_reject_off_curve(ctx):
    import zelana_b200
    bad = (1).to_bytes(32, "little") + (3).to_bytes(32, "little")
    with pytest.raises(zelana_b200.ZkbError) as e:
        ctx.g1_bases(bad, validate=True)
    assert e.value.code == -5
    b = ctx.g1_bases(g1_raw([bn.G1_GEN, None]))
    assert b.read() == g1_raw([bn.G1_GEN, None])


# ----------------------------------------------------------------------------- MSM (rows a6, a7)
def _expected_msm(curve, gen, dlogs, scalars):
    acc = sum(k * s for k, s in zip(dlogs, scalars)) % R
    return curve.mul(gen, acc)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 257, 1000, 5000, 1 << 14])
def test_msm_g1_matches_oracle(ctx, n):
    rnd = random.Random(100 + n)
    pts, ks = arithmetic_bases(bn.G1, bn.G1_GEN, n, rnd.randrange(R), rnd.randrange(R))
    sc = [rnd.randrange(R) for _ in range(n)]
    bases = ctx.g1_bases(g1_raw(pts))
    out = ctx.msm_g1(bases, fr_bytes(sc))
    assert out == bn.g1_to_raw(_expected_msm(bn.G1, bn.G1_GEN, ks, sc))
    if 0 < n <= 33:  # also against the plain double-and-add restatement
        assert out == bn.g1_to_raw(bn.G1.msm_naive(pts, sc))


@pytest.mark.parametrize("c", [2, 5, 8, 11, 13, 16, 19, 22, 23])
def test_msm_g1_every_window_width(ctx, c):
    """The window width is fixed when the bases build their tables 2^(c j) P: force it, then load."""
    rnd = random.Random(200 + c)
    n = 3000 if c > 2 else 600
    pts, ks = arithmetic_bases(bn.G1, bn.G1_GEN, n, 17, 1)
    sc = [rnd.randrange(R) for _ in range(n)]
    sc[:6] = [0, 1, R - 1, (1 << 253), R >> 1, 2]
    ctx.set_msm_window(c)
    try:
        bases = ctx.g1_bases(g1_raw(pts))
    finally:
        ctx.set_msm_window(0)
    out = ctx.msm_g1(bases, fr_bytes(sc))
    assert out == bn.g1_to_raw(_expected_msm(bn.G1, bn.G1_GEN, ks, sc))
    # a sub-range on the same tables
    out = ctx.msm_g1(bases, fr_bytes(sc[100:500]), offset=100)
    assert out == bn.g1_to_raw(_expected_msm(bn.G1, bn.G1_GEN, ks[100:500], sc[100:500]))


def test_msm_g1_witness_like_and_degenerate(ctx):
    """Skewed digits (50% zero, 25% one), repeated bases (same bucket gets P and P: doubling path),
    cancelling pairs, infinity bases."""
    rnd = random.Random(77)
    n = 4096
    pts, ks = arithmetic_bases(bn.G1, bn.G1_GEN, n, 3, 5)
    sc = []
    for _ in range(n):
        u = rnd.random()
        sc.append(0 if u < 0.5 else (1 if u < 0.75 else rnd.randrange(R)))
    # repeated base with identical scalars -> doubling inside a bucket
    for i in range(0, 64, 2):
        pts[i + 1], ks[i + 1] = pts[i], ks[i]
        sc[i + 1] = sc[i] = 1 + (i % 3)
    # P and -P with the same scalar -> bucket sums to infinity
    for i in range(64, 128, 2):
        pts[i + 1], ks[i + 1] = bn.G1.neg(pts[i]), (-ks[i]) % R
        sc[i + 1] = sc[i] = 9
    pts[200], ks[200] = None, 0
    bases = ctx.g1_bases(g1_raw(pts))
    out = ctx.msm_g1(bases, fr_bytes(sc))
    assert out == bn.g1_to_raw(_expected_msm(bn.G1, bn.G1_GEN, ks, sc))
    # all-zero scalars -> infinity
    assert ctx.msm_g1(bases, fr_bytes([0] * n)) == bytes(64)
    # offset / sub-range
    out = ctx.msm_g1(bases, fr_bytes(sc[300:900]), offset=300)
    assert out == bn.g1_to_raw(_expected_msm(bn.G1, bn.G1_GEN, ks[300:900], sc[300:900]))


@pytest.mark.parametrize("n", [1, 40, 700, 3000])
def test_msm_g2_matches_oracle(ctx, n):
    rnd = random.Random(300 + n)
    pts, ks = arithmetic_bases(bn.G2, bn.G2_GEN, n, rnd.randrange(R), rnd.randrange(R))
    sc = [rnd.randrange(R) for _ in range(n)]
    sc[0] = 1
    bases = ctx.g2_bases(g2_raw(pts))
    out = ctx.msm_g2(bases, fr_bytes(sc))
    assert out == bn.g2_to_raw(_expected_msm(bn.G2, bn.G2_GEN, ks, sc))


def test_generated_bases_and_large_msm_known_dlog(ctx):
    """2^18 bases [k_i]G generated on the GPU; MSM must equal [sum k_i s_i]G (size-independent property)."""
    import numpy as np
    import torch
    n = 1 << 18
    rs = np.random.RandomState(1234)
    k = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    s = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    k[:, 7] &= 0x0FFFFFFF  # < 2^252 < r : canonical
    s[:, 7] &= 0x0FFFFFFF
    kd = torch.from_numpy(k.view(np.int32)).cuda()
    bases = ctx.g1_bases_generate(kd, n)
    ctx.synchronize()
    # spot-check generated points
    kin = [int.from_bytes(k[i].tobytes(), "little") for i in range(n)]
    sin = [int.from_bytes(s[i].tobytes(), "little") for i in range(n)]
    raw = bases.read(0, 4)
    assert raw == g1_raw([bn.G1.mul(bn.G1_GEN, kin[i]) for i in range(4)])
    out = ctx.msm_g1(bases, s.tobytes())
    assert out == bn.g1_to_raw(_expected_msm(bn.G1, bn.G1_GEN, kin, sin))


# ----------------------------------------------------------------------------- NTT (row a5)
@pytest.mark.parametrize("log_n", list(range(0, 13)))
def test_ntt_matches_oracle(ctx, log_n):
    rnd = random.Random(400 + log_n)
    n = 1 << log_n
    v = [rnd.randrange(R) for _ in range(n)]
    if n >= 4:
        v[0], v[1], v[2] = 0, R - 1, 1
    data = fr_bytes(v)
    assert unpack32(ctx.ntt(data, log_n)) == g16.fft(v)
    assert unpack32(ctx.ntt(data, log_n, inverse=True)) == g16.ifft(v)
    assert unpack32(ctx.ntt(data, log_n, coset=True)) == g16.coset_fft(v)
    assert unpack32(ctx.ntt(data, log_n, inverse=True, coset=True)) == g16.coset_ifft(v)


@pytest.mark.parametrize("log_n", [13, 16, 17, 20, 22])
def test_ntt_roundtrip_and_linearity_large(ctx, log_n):
    import numpy as np
    n = 1 << log_n
    rs = np.random.RandomState(log_n)
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x0FFFFFFF
    data = a.tobytes()
    f = ctx.ntt(data, log_n)
    assert ctx.ntt(f, log_n, inverse=True) == data
    cf = ctx.ntt(data, log_n, coset=True)
    assert ctx.ntt(cf, log_n, inverse=True, coset=True) == data
    # spot-check 3 output coefficients against the definition sum_j x_j w^(jk) (Horner in Python)
    w = g16.root_of_unity(n)
    xs = [int.from_bytes(a[j].tobytes(), "little") for j in range(n)] if log_n <= 16 else None
    if xs is not None:
        fo = unpack32(f)
        for k in (0, 1, n - 1, n // 2 + 3):
            wk = pow(w, k, R)
            acc = 0
            for x in reversed(xs):
                acc = (acc * wk + x) % R
            assert fo[k] == acc


# ----------------------------------------------------------------------------- witness map + prove (rows a3, a4, a8)
@pytest.fixture(scope="module")
def mimc_setup():
    r1cs, z = mimc7_chain(num_perm=2, seed=42, rounds=20)   # 161 constraints -> domain 256
    assert r1cs.is_satisfied(z)
    pk = g16.circuit_specific_setup(r1cs, StdRng.seed_from_u64(0))
    return r1cs, z, pk


def test_witness_map_matches_oracle(ctx, mimc_setup):
    r1cs, z, _ = mimc_setup
    m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    h = unpack32(ctx.witness_map(m, fr_bytes(z)))
    assert h == g16.witness_map_from_matrices(r1cs, z)
    assert h[-1] == 0


def test_prove_square_circuit_reproduces_reference_fixture(ctx):
    """The reference's own committed proof (proof_for_onchain.json) from the GPU prover, byte for byte."""
    import json, os
    from conftest import REF_FIXTURES
    from oracle import rng as orng
    r1cs, z = g16.square_circuit(7)
    rng = StdRng.seed_from_u64(42)
    pk = g16.circuit_specific_setup(r1cs, rng)
    r = orng.rand_fr(rng)
    s = orng.rand_fr(rng)
    m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    dpk = ctx.proving_key(**pk_parts(pk))
    a, b, c = ctx.prove(dpk, m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
    proof = g16.Proof(bn.g1_from_raw(a), bn.g2_from_raw(b), bn.g1_from_raw(c))
    pc = json.load(open(os.path.join(REF_FIXTURES, "proof_for_onchain.json")))["proof_components"]
    assert proof.serialize_uncompressed() == bytes(pc["pi_a"]) + bytes(pc["pi_b"]) + bytes(pc["pi_c"])


def test_prove_mimc_matches_oracle_and_verifies(ctx, mimc_setup):
    r1cs, z, pk = mimc_setup
    m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    dpk = ctx.proving_key(**pk_parts(pk))
    for seed in (0, 7):
        rng = StdRng.seed_from_u64(seed)      # prover.rs:354: seed = batch_id
        from oracle import rng as orng
        r, s = orng.rand_fr(rng), orng.rand_fr(rng)
        a, b, c = ctx.prove(dpk, m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
        ref = g16.prove_with_rs(pk, r1cs, z, r, s)
        assert (a, b, c) == (bn.g1_to_raw(ref.a), bn.g2_to_raw(ref.b), bn.g1_to_raw(ref.c))
        assert g16.verify(pk.vk, [z[1]], ref)
    # r = 0 edge (ark-groth16 skips B1) and s = 0
    a, b, c = ctx.prove(dpk, m, fr_bytes(z), fr_bytes([0]), fr_bytes([5]))
    ref = g16.prove_with_rs(pk, r1cs, z, 0, 5)
    assert (a, b, c) == (bn.g1_to_raw(ref.a), bn.g2_to_raw(ref.b), bn.g1_to_raw(ref.c))


def test_prove_shape_errors(ctx, mimc_setup):
    import zelana_b200
    r1cs, z, pk = mimc_setup
    sq, zsq = g16.square_circuit(3)
    m_sq = ctx.r1cs(sq.num_instance, sq.num_witness, sq.a, sq.b, sq.c)
    dpk = ctx.proving_key(**pk_parts(pk))
    with pytest.raises(zelana_b200.ZkbError) as e:
        ctx.prove(dpk, m_sq, fr_bytes(zsq), fr_bytes([1]), fr_bytes([1]))
    assert e.value.code == -6
