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


# ----------------------------------------------------------------------------- mid / full size against the C++ restatement
def _rand_fr_np(n, seed):
    import numpy as np
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] %= 0x30644E72
    return a


def test_msm_g1_2p16_matches_cpp_oracle_and_witness_like(ctx):
    """Random (not arithmetic-progression) bases generated on the GPU, read back, and summed by oracle/cpu_oracle.cpp."""
    import numpy as np
    import torch
    from oracle import cpu as orc
    n = 1 << 16
    k = _rand_fr_np(n, 1)
    bases = ctx.g1_bases_generate(torch.from_numpy(k.view(np.int32)).cuda(), n)
    cb = orc.G1Bases.from_raw(bases.read())
    s = _rand_fr_np(n, 2)
    assert ctx.msm_g1(bases, s) == cb.msm(s)
    # witness-like scalars: 50% zero, 25% one, a hot bucket of 2^14 equal small scalars, the rest uniform
    w = s.copy()
    r = np.random.RandomState(3).rand(n)
    w[r < 0.5] = 0
    one = np.zeros(8, dtype=np.uint32); one[0] = 1
    w[(r >= 0.5) & (r < 0.75)] = one
    w[: 1 << 14] = 0
    w[: 1 << 14, 0] = 7
    assert ctx.msm_g1(bases, w) == cb.msm(w)
    # sub-range with an offset
    assert ctx.msm_g1(bases, s[1000:30000], offset=1000) == cb.msm(s[1000:30000], off=1000)


def test_msm_g2_2p13_matches_cpp_oracle(ctx):
    import numpy as np
    import torch
    from oracle import cpu as orc
    n = 1 << 13
    k = _rand_fr_np(n, 4)
    bases = ctx.g2_bases_generate(torch.from_numpy(k.view(np.int32)).cuda(), n)
    s = _rand_fr_np(n, 5)
    assert ctx.msm_g2(bases, s) == orc.msm_g2(bases.read(), s)


def test_msm_g1_2p22_known_dlog(ctx):
    """Full-pipeline check at 2^22 (c = 21/22 tables): sum s_i [k_i]G == [sum k_i s_i]G, exact integers on the host."""
    import numpy as np
    import torch
    n = 1 << 22
    k = _rand_fr_np(n, 6)
    s = _rand_fr_np(n, 7)
    bases = ctx.g1_bases_generate(torch.from_numpy(k.view(np.int32)).cuda(), n)
    out = ctx.msm_g1(bases, s)

    def ints(a):  # [n, 8] u32 -> python ints
        b = a.astype(np.uint64)
        lo = b[:, 0] | (b[:, 1] << np.uint64(32))
        parts = [b[:, 2 * j] | (b[:, 2 * j + 1] << np.uint64(32)) for j in range(4)]
        return [int(p0) | (int(p1) << 64) | (int(p2) << 128) | (int(p3) << 192)
                for p0, p1, p2, p3 in zip(*[p.tolist() for p in parts])]

    acc = 0
    for a, b in zip(ints(k), ints(s)):
        acc += a * b
    assert out == bn.g1_to_raw(bn.G1.mul(bn.G1_GEN, acc % R))


@pytest.mark.parametrize("log_n", [14, 18])
def test_ntt_matches_cpp_oracle(ctx, log_n):
    from oracle import cpu as orc
    a = _rand_fr_np(1 << log_n, 10 + log_n)
    for inv in (False, True):
        for coset in (False, True):
            assert ctx.ntt(a, log_n, inverse=inv, coset=coset) == orc.ntt(a, log_n, inverse=inv, coset=coset)


def test_ntt_roundtrip_2p24(ctx):
    import numpy as np
    a = _rand_fr_np(1 << 24, 99)
    data = a.tobytes()
    f = ctx.ntt(data, 24, coset=True)
    assert f != data
    assert ctx.ntt(f, 24, inverse=True, coset=True) == data


def test_prove_2p13_matches_cpp_oracle(ctx):
    """Full prove on a 2^13-constraint MiMC circuit (the L2-dummy domain size): the key is a set of random
    curve points generated on the GPU (no trusted setup needed to compare two provers), loaded into BOTH provers."""
    import importlib.util
    import os
    import numpy as np
    import torch
    from conftest import ROOT
    from oracle import cpu as orc
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ni, nw, (A, B, Cm), z = bench.mimc_r1cs_numpy(np, num_perm=22, seed=77)       # 8009 constraints -> domain 2^13
    nv, n = ni + nw, 1 << 13
    m = ctx.r1cs(ni, nw, A, B, Cm)
    assert m.log_domain == 13

    def g1(cnt, seed):
        b = ctx.g1_bases_generate(torch.from_numpy(_rand_fr_np(cnt, seed).view(np.int32)).cuda(), cnt)
        return b.read()

    def g2(cnt, seed):
        b = ctx.g2_bases_generate(torch.from_numpy(_rand_fr_np(cnt, seed).view(np.int32)).cuda(), cnt)
        return b.read()

    parts = dict(alpha_g1=g1(1, 20), beta_g1=g1(1, 21), beta_g2=g2(1, 22), delta_g1=g1(1, 23), delta_g2=g2(1, 24),
                 a_query=g1(nv, 25), b_g1_query=g1(nv, 26), b_g2_query=g2(nv, 27), h_query=g1(n - 1, 28), l_query=g1(nw, 29))
    dpk = ctx.proving_key(**parts)
    cpk = orc.ProvingKey(**parts)
    cm = orc.R1cs(ni, nw, csr=(A, B, Cm))
    zb = z.reshape(-1)
    h_gpu = ctx.witness_map(m, zb)
    assert h_gpu == orc.witness_map(cm, zb)
    assert h_gpu[-32:] == bytes(32)
    for r, s in ((0x1234567, 0x7654321), (R - 5, 3)):
        got = ctx.prove(dpk, m, zb, fr_bytes([r]), fr_bytes([s]))
        assert got == orc.prove(cpk, cm, zb, fr_bytes([r]), fr_bytes([s]))


def test_msm_g1_host_sliced_path_2p21(ctx):
    """zkb_msm_g1 with >= 2^20 host scalars takes the sliced path (upload overlapped with accumulation, one bucket array per
    slice, merged before the reduction): same bytes as the device-resident path and as the C++ restatement."""
    import numpy as np
    import torch
    from oracle import cpu as orc
    n = (1 << 21) + 12345          # ragged: the last slice is shorter
    k = _rand_fr_np(n, 31)
    s = _rand_fr_np(n, 32)
    s[: 1 << 12] = 0
    bases = ctx.g1_bases_generate(torch.from_numpy(k.view(np.int32)).cuda(), n)
    out = ctx.msm_g1(bases, s)
    sd = torch.from_numpy(s.view(np.int32)).cuda()
    od = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ctx.msm_g1_dev(bases, sd, n, out_affine_dev=od)
    ctx.synchronize()
    assert out == bytes(od.cpu().numpy())
    m = 1 << 18
    assert ctx.msm_g1(bases, s[:m]) == orc.G1Bases.from_raw(bases.read(0, m)).msm(s[:m])
    # twice in a row on the same context (slice buffers and events are reused)
    assert ctx.msm_g1(bases, s) == out


# ----------------------------------------------------------------------------- compressed key load (row a11)
def test_pk_load_compressed_matches_uncompressed_and_oracle(ctx, mimc_setup):
    """Groth16Prover::from_bytes path: the ark-serialize COMPRESSED ProvingKey is decompressed + validated on the GPU;
    proofs made with it equal the oracle's.  The key holds infinity points (variables absent from B) and both y signs."""
    from oracle import rng as orng
    r1cs, z, pk = mimc_setup
    blob = pk.serialize_compressed()
    assert g16.ProvingKey.deserialize_compressed(blob).serialize_compressed() == blob
    assert any(p is None for p in pk.b_g1_query)
    m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    dpk = ctx.proving_key_compressed(blob)
    rng = StdRng.seed_from_u64(3)
    r, s = orng.rand_fr(rng), orng.rand_fr(rng)
    got = ctx.prove(dpk, m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
    ref = g16.prove_with_rs(pk, r1cs, z, r, s)
    assert got == (bn.g1_to_raw(ref.a), bn.g2_to_raw(ref.b), bn.g1_to_raw(ref.c))
    assert g16.verify(pk.vk, [z[1]], ref)


def test_pk_load_compressed_square_circuit_fixture(ctx):
    """keygen -> serialize_compressed -> from_bytes -> prove: the reference's committed proof again, via the compressed key."""
    import json, os
    from conftest import REF_FIXTURES
    from oracle import rng as orng
    r1cs, z = g16.square_circuit(7)
    rng = StdRng.seed_from_u64(42)
    pk = g16.circuit_specific_setup(r1cs, rng)
    r, s = orng.rand_fr(rng), orng.rand_fr(rng)
    m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    a, b, c = ctx.prove(ctx.proving_key_compressed(pk.serialize_compressed()), m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
    proof = g16.Proof(bn.g1_from_raw(a), bn.g2_from_raw(b), bn.g1_from_raw(c))
    pc = json.load(open(os.path.join(REF_FIXTURES, "proof_for_onchain.json")))["proof_components"]
    assert proof.serialize_uncompressed() == bytes(pc["pi_a"]) + bytes(pc["pi_b"]) + bytes(pc["pi_c"])


def test_pk_load_compressed_rejects_what_arkworks_rejects(ctx, mimc_setup):
    import zelana_b200
    _, _, pk = mimc_setup
    blob = bytearray(pk.serialize_compressed())
    vk_len = 32 + 64 * 3 + 8 + 32 * len(pk.vk.gamma_abc_g1)
    a_off = vk_len + 64 + 8                      # first a_query point

    def expect(buf, code, *, validate=True):
        with pytest.raises(zelana_b200.ZkbError) as e:
            ctx.proving_key_compressed(bytes(buf), validate=validate)
        assert e.value.code == code, str(e.value)

    expect(blob[:-1], -6)                        # truncated
    expect(blob + b"\x00", -6)                   # trailing byte
    bad = bytearray(blob); bad[a_off + 31] |= 0xC0
    expect(bad, -5)                              # both flags set
    bad = bytearray(blob); bad[a_off:a_off + 32] = (bn.P + 1).to_bytes(32, "little")
    expect(bad, -5)                              # x >= p
    x = 1
    while bn.fq_sqrt((x ** 3 + 3) % bn.P) is not None:
        x += 1
    bad = bytearray(blob); bad[a_off:a_off + 32] = x.to_bytes(32, "little")
    expect(bad, -5)                              # x^3 + 3 is not a square
    # a G2 point on the twist but outside the prime-order subgroup (in b_g2_query[0])
    b2_off = vk_len + 64 + 8 + 32 * len(pk.a_query) + 8 + 32 * len(pk.b_g1_query) + 8
    xx = 1
    while True:
        cand = (xx, 0)
        y = bn.f2_sqrt(bn.f2_add(bn.f2_mul(bn.f2_sqr(cand), cand), bn.B_G2))
        if y is not None and not bn.g2_in_subgroup((cand, y)):
            break
        xx += 1
    bad = bytearray(blob); bad[b2_off:b2_off + 64] = bn.g2_serialize((cand, y))
    expect(bad, -5)                              # not in the subgroup
    ctx.proving_key_compressed(bytes(bad), validate=False)   # Validate::No skips exactly that check
    # gamma_abc is not used by the prover but is still validated
    bad = bytearray(blob); bad[32 + 64 * 3 + 8:32 + 64 * 3 + 8 + 32] = x.to_bytes(32, "little")
    expect(bad, -5)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_sharded_prove_equals_prove(ctx, mimc_setup, world):
    """One proof over `world` key shards (emulated sequentially on one GPU): partial records combined == zkb_prove == oracle.
    The last shards hold the appended constant points (alpha, beta, delta); ragged ranges when world does not divide."""
    import torch
    from zelana_b200.api import PROVE_PARTIAL_BYTES
    r1cs, z, pk = mimc_setup
    m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    parts_host = pk_parts(pk)
    r, s = 0x1234567890ABCDEF1234567, 0x7654321FEDCBA987654321
    ref = g16.prove_with_rs(pk, r1cs, z, r, s)
    expect = (bn.g1_to_raw(ref.a), bn.g2_to_raw(ref.b), bn.g1_to_raw(ref.c))
    parts = torch.zeros(world * PROVE_PARTIAL_BYTES, dtype=torch.uint8, device="cuda")
    for shard in range(world):
        dpk = ctx.proving_key_shard(shard, world, **parts_host)
        ctx.prove_partial(dpk, m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]), parts[shard * PROVE_PARTIAL_BYTES:])
        if world > 1:
            import zelana_b200
            with pytest.raises(zelana_b200.ZkbError):
                ctx.prove(dpk, m, fr_bytes(z), fr_bytes([r]), fr_bytes([s]))   # a shard cannot make a whole proof
        dpk.free()
    assert ctx.prove_combine(parts, world, fr_bytes([r]), fr_bytes([s])) == expect
    # r = 0 (arkworks skips B1) through the sharded path
    parts.zero_()
    for shard in range(world):
        dpk = ctx.proving_key_shard(shard, world, **parts_host)
        ctx.prove_partial(dpk, m, fr_bytes(z), fr_bytes([0]), fr_bytes([9]), parts[shard * PROVE_PARTIAL_BYTES:])
        dpk.free()
    ref0 = g16.prove_with_rs(pk, r1cs, z, 0, 9)
    assert ctx.prove_combine(parts, world, fr_bytes([0]), fr_bytes([9])) == (bn.g1_to_raw(ref0.a), bn.g2_to_raw(ref0.b), bn.g1_to_raw(ref0.c))


def test_l2_sized_keygen_prove_verify_end_to_end(ctx):
    """The whole reference flow at the L2BlockCircuit::dummy() size (domain 2^13) with nothing but this library on the GPU:
    keygen (seed 0, keygen.rs:87) -> compressed key bytes -> from_bytes -> prove(batch_id) -> the oracle's pairing check."""
    import importlib.util
    import os
    import numpy as np
    from conftest import ROOT
    from oracle import cpu as orc
    from zelana_b200 import keygen as kg
    from zelana_b200 import prover as zp
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ni, nw, (A, B, Cm), z = bench.mimc_r1cs_numpy(np, num_perm=22, seed=78)
    pk_bytes, vk_bytes, raw = kg.circuit_specific_setup(ctx, ni, nw, A, B, Cm, zp.StdRng.seed_from_u64(0))
    vk = g16.VerifyingKey.deserialize_compressed(vk_bytes)
    dpk = ctx.proving_key_compressed(pk_bytes)
    m = ctx.r1cs(ni, nw, A, B, Cm)
    zb = z.reshape(-1)
    public = [int.from_bytes(z[1].tobytes(), "little")]
    for batch_id in (0, 3):
        rng = zp.StdRng.seed_from_u64(batch_id)
        r, s = zp.fr_rand(rng), zp.fr_rand(rng)
        a, b, c = ctx.prove(dpk, m, zb, fr_bytes([r]), fr_bytes([s]))
        proof = g16.Proof(bn.g1_from_raw(a), bn.g2_from_raw(b), bn.g1_from_raw(c))
        assert g16.verify(vk, public, proof)
        assert not g16.verify(vk, [public[0] + 1], proof)
        # and the C++ restatement proves the same bytes from the same (raw) key
        cpk = orc.ProvingKey(**{k: raw[k] for k in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "a_query",
                                                    "b_g1_query", "b_g2_query", "h_query", "l_query")})
        assert orc.prove(cpk, orc.R1cs(ni, nw, csr=(A, B, Cm)), zb, fr_bytes([r]), fr_bytes([s])) == (a, b, c)


# ----------------------------------------------------------------------------- the L2 batch circuit itself (rows a1, a2; 8f.3)
def _solana_to_proof(proof_bytes):
    """Inverse of proof_to_solana_bytes (prover.rs:304-334): -A || B || C  ->  oracle Proof with A un-negated."""
    a = bn.g1_from_raw(proof_bytes[:64])
    return g16.Proof(bn.G1.neg(a), bn.g2_from_raw(proof_bytes[64:192]), bn.g1_from_raw(proof_bytes[192:]))


@pytest.fixture(scope="module")
def l2_setup(ctx):
    """keygen.rs:81-131 on the GPU: dummy circuit, StdRng::seed_from_u64(0)."""
    from zelana_b200 import l2_circuit as P2
    circ, pk_bytes, vk_bytes, raw = P2.keygen(ctx)
    return circ, pk_bytes, vk_bytes, raw


def test_l2_circuit_keygen_prove_verify(ctx, l2_setup):
    """`Groth16Prover::from_bytes` + `BatchProver::prove` for the reference's own L2BlockCircuit, end to end through the C ABI
    (zkb_l2_prove): the proof verifies by pairing under the key's VK with the seven public inputs, and equals -- byte for
    byte -- the proof the C++ restatement of arkworks makes from the ORACLE's matrices, assignment and (r, s)."""
    from oracle import cpu as orc
    from oracle import l2_circuit as O
    from oracle.rng import rand_fr
    from zelana_b200 import l2_circuit as P2
    circ, pk_bytes, vk_bytes, raw = l2_setup
    vk = g16.VerifyingKey.deserialize_compressed(vk_bytes)
    assert len(vk.gamma_abc_g1) == 8
    dpk = ctx.proving_key_compressed(pk_bytes)
    prover = P2.L2Prover(ctx, circ, dpk, vk_bytes)
    cpk = orc.ProvingKey(**{k: raw[k] for k in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "a_query",
                                                "b_g1_query", "b_g2_query", "h_query", "l_query")})
    for batch_id, amount in ((0, 100), (7, 1000), (2 ** 64 - 1, 0)):
        oc = O.with_satisfying_roots(O.L2BlockCircuit(
            transactions=[(bytes([1] * 32), bytes([2] * 32), amount)],
            initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): 0}, batch_id=batch_id))
        r1cs, z = O.synthesize(oc)
        pc = P2.L2BlockCircuit(transactions=[P2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), amount)],
                               initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): 0}, batch_id=batch_id)
        inputs = P2.satisfying_inputs(pc)
        assert inputs.post_state_root == oc.post_state_root
        proof = prover.prove(inputs, pc)
        assert prover.verify(proof) and len(proof.proof_bytes) == 256
        assert g16.verify(vk, z[1:8], _solana_to_proof(proof.proof_bytes))
        assert not g16.verify(vk, z[1:7] + [(z[7] + 1) % R], _solana_to_proof(proof.proof_bytes))
        rng = StdRng.seed_from_u64(batch_id)
        r, s = rand_fr(rng), rand_fr(rng)
        a, b, c = orc.prove(cpk, orc.R1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c), fr_bytes(z),
                            fr_bytes([r]), fr_bytes([s]))
        from zelana_b200 import proof_to_solana_bytes
        assert proof.proof_bytes == proof_to_solana_bytes(a, b, c)
    # a batch whose roots are wrong still "proves" (release-mode arkworks only debug_asserts satisfaction) -- and is rejected
    bad = prover.prove(P2.BatchPublicInputs(batch_id=1), P2.L2BlockCircuit.dummy())
    zbad = unpack32(circ.assign(P2.L2BlockCircuit.dummy().with_inputs(P2.BatchPublicInputs(batch_id=1))))
    assert not g16.verify(vk, zbad[1:8], _solana_to_proof(bad.proof_bytes))
    prover.m.free()
    dpk.free()


def test_l2_circuit_key_equals_oracle_setup_on_a_smaller_shape(ctx):
    """keygen for the no-transfer shape (the smallest L2BlockCircuit): GPU setup from the native matrices == the oracle's
    circuit_specific_setup from the oracle's matrices, same seed; VK bytes identical."""
    from oracle import l2_circuit as O
    from zelana_b200 import l2_circuit as P2
    shape = P2.L2BlockCircuit(initial_accounts={bytes([9] * 32): 5})
    circ, pk_bytes, vk_bytes, raw = P2.keygen(ctx, shape, seed=0)
    r1cs, _ = O.synthesize(O.L2BlockCircuit(initial_accounts={bytes([9] * 32): 5}))
    assert (circ.num_constraints, circ.num_witness) == (r1cs.num_constraints, r1cs.num_witness)
    opk = g16.circuit_specific_setup(r1cs, StdRng.seed_from_u64(0))
    assert vk_bytes == opk.vk.serialize_compressed()
    assert raw["a_query"][:64 * 16] == g1_raw(opk.a_query[:16])
    assert raw["l_query"][-64 * 8:] == g1_raw(opk.l_query[-8:])
    assert raw["h_query"][:64 * 4] == g1_raw(opk.h_query[:4])
    circ.free()


def test_l2_prove_shape_error_is_an_error_not_a_crash(ctx, l2_setup):
    from zelana_b200 import ZkbError
    from zelana_b200 import l2_circuit as P2
    circ, pk_bytes, vk_bytes, raw = l2_setup
    dpk = ctx.proving_key_compressed(pk_bytes, validate=False)
    prover = P2.L2Prover(ctx, circ, dpk, vk_bytes)
    two = P2.L2BlockCircuit(transactions=[P2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 1)] * 2,
                            initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): 0})
    with pytest.raises(ZkbError) as e:
        prover.prove(P2.BatchPublicInputs(), two)
    assert e.value.code == -6
    prover.m.free()
    dpk.free()


# ----------------------------------------------------------------------------- CUDA-graph replay of the prove
def test_prove_graph_replay_is_bit_identical_and_survives_reallocation(mimc_setup):
    """The device part of a prove is captured as a CUDA graph on the second call for a (key, matrices) pair and replayed
    afterwards: every replay must give the bytes the direct launches give, also with two keys interleaved on one context,
    and after buffers were reallocated (a bigger circuit proved in between) the stale graph must not be replayed."""
    import zelana_b200
    r1cs, z, pk = mimc_setup
    sq, zsq = g16.square_circuit(7)
    pk_sq = g16.circuit_specific_setup(sq, StdRng.seed_from_u64(42))
    big, zbig = mimc7_chain(num_perm=6, seed=9, rounds=91)          # 2185 constraints -> domain 4096: grows every buffer
    rnd = random.Random(77)
    cases = []
    for i in range(6):                                              # different witness-independent randomness each time
        cases.append((rnd.randrange(R), rnd.randrange(R)))

    def run(graphs):
        c = zelana_b200.Context(0)
        c.set_graphs(graphs)
        m1 = c.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
        m2 = c.r1cs(sq.num_instance, sq.num_witness, sq.a, sq.b, sq.c)
        k1, k2 = c.proving_key(**pk_parts(pk)), c.proving_key(**pk_parts(pk_sq))
        out = []
        for r, s in cases:
            out.append(c.prove(k1, m1, fr_bytes(z), fr_bytes([r]), fr_bytes([s])))
            out.append(c.prove(k2, m2, fr_bytes(zsq), fr_bytes([s]), fr_bytes([r])))
        stats_before = c.graph_stats()
        # a larger circuit reallocates the prove buffers: the cached graphs are stale from here on
        m3 = c.r1cs(big.num_instance, big.num_witness, big.a, big.b, big.c)
        nv, n = big.num_instance + big.num_witness, 4096
        import torch
        import numpy as np
        k = torch.from_numpy(_rand_fr_np(n + 16, 5).view(np.int32)).cuda()
        k3 = c.proving_key_synthetic(nv, big.num_witness, n - 1, k, n + 16)
        out.append(c.prove(k3, m3, fr_bytes(zbig), fr_bytes([3]), fr_bytes([4])))
        for r, s in cases[:4]:
            out.append(c.prove(k1, m1, fr_bytes(z), fr_bytes([s]), fr_bytes([r])))
        stats = c.graph_stats()
        launches = c.launch_count()
        for h in (m1, m2, m3, k1, k2, k3):
            h.free()
        c.close()
        return out, stats_before, stats, launches

    direct, _, s_off, launches_off = run(False)
    graphed, s_mid, s_on, launches_on = run(True)
    assert s_off == (0, 0)
    assert graphed == direct
    assert s_mid[0] == 2 and s_mid[1] == 2 * (len(cases) - 1)       # each key: 1 direct warm-up, then captured + replayed
    assert s_on[0] == 3 and s_on[1] > s_mid[1]                      # recaptured once after the reallocation
    assert launches_on == launches_off                              # a replay counts the kernels it stands for
    ref = g16.prove_with_rs(pk, r1cs, z, cases[0][0], cases[0][1])
    assert graphed[0] == (bn.g1_to_raw(ref.a), bn.g2_to_raw(ref.b), bn.g1_to_raw(ref.c))


def test_l2_batch_prove_equals_single_proofs(ctx, l2_setup):
    """zkb_l2_batch_prove (lanes of contexts + host threads inside the library) returns, in order, exactly the proofs
    zkb_l2_prove makes one at a time; a batch smaller than the lane count and an empty batch work; a proof of the wrong
    shape fails alone."""
    from zelana_b200 import ZkbError
    from zelana_b200 import l2_circuit as P2
    circ, pk_bytes, vk_bytes, raw = l2_setup
    dpk = ctx.proving_key_compressed(pk_bytes, validate=False)
    single = P2.L2Prover(ctx, circ, dpk, vk_bytes)

    def batch(bid):
        c = P2.L2BlockCircuit(transactions=[P2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), bid)],
                              initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): 7 * bid}, batch_id=bid)
        return c.with_inputs(P2.satisfying_inputs(c))

    circuits = [batch(i + 1) for i in range(21)]
    expect = [single.prove_circuit(c).proof_bytes for c in circuits]
    bp = P2.L2BatchProver(ctx, circ, dpk, lanes=8)
    assert bp.lanes == 8
    for _ in range(3):                       # direct launches, graph capture, graph replay
        assert bp.prove(circuits) == expect
    assert bp.prove(circuits[:3]) == expect[:3]
    assert bp.prove([]) == []
    wrong = P2.L2BlockCircuit(transactions=circuits[0].transactions * 2, initial_accounts=circuits[0].initial_accounts)
    pack = bp.marshal(circuits[:4] + [wrong] + circuits[4:6])
    with pytest.raises(ZkbError) as e:
        bp.prove_marshalled(pack)
    assert e.value.code == -6
    n, _xs, _ws, _keep, out, status = pack
    assert list(status)[:7] == [0, 0, 0, 0, -6, 0, 0]
    assert [out.raw[256 * i:256 * i + 256] for i in (0, 1, 2, 3, 5, 6)] == expect[:6]
    # calls of <= 8 proofs take the per-proof path (one context each), larger ones the batched kernels: both sides of the switch
    assert bp.prove(circuits[:8]) == expect[:8]
    assert bp.prove(circuits[:9]) == expect[:9]
    pack = bp.marshal(circuits[:10] + [wrong] + circuits[10:12])
    with pytest.raises(ZkbError) as e:
        bp.prove_marshalled(pack)
    assert e.value.code == -6
    n, _xs, _ws, _keep, out, status = pack
    assert list(status) == [0] * 10 + [-6, 0, 0]
    assert [out.raw[256 * i:256 * i + 256] for i in list(range(10)) + [11, 12]] == expect[:12]
    bp.close()
    single.m.free()
    dpk.free()


def test_prove_batch_equals_single_proves(ctx, l2_setup):
    """zkb_prove_batch (batched mat-vecs / NTTs / MSMs: K scalar vectors against one window table, per-proof bucket arrays,
    s A and r B1 by double-and-add instead of extra MSMs) returns exactly the K proofs zkb_prove makes one at a time -- for
    K = 1, an odd K, r = 0 and s = 0 among the randomness, and a second call on the same context (buffers reused)."""
    from zelana_b200 import l2_circuit as P2
    circ, pk_bytes, vk_bytes, raw = l2_setup
    dpk = ctx.proving_key_compressed(pk_bytes, validate=False)
    a, b, c = circ.matrices()
    m = ctx.r1cs(circ.num_instance, circ.num_witness, a, b, c)

    def assignment(bid):
        ck = P2.L2BlockCircuit(transactions=[P2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 3 * bid + 1)],
                               initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): bid}, batch_id=bid)
        return circ.assign(ck.with_inputs(P2.satisfying_inputs(ck)))

    K = 11
    zs = [assignment(i + 1) for i in range(K)]
    rs = [P2.prover_randomness(i + 1) for i in range(K)]
    rs[3] = (bytes(32), rs[3][1])            # r = 0
    rs[5] = (rs[5][0], bytes(32))            # s = 0
    rs[7] = (fr_bytes([R - 1]), fr_bytes([1]))
    expect = [ctx.prove(dpk, m, z, r, s) for z, (r, s) in zip(zs, rs)]
    got = ctx.prove_batch(dpk, m, b"".join(zs), b"".join(r + s for r, s in rs))
    assert got == expect
    assert ctx.prove_batch(dpk, m, zs[2], rs[2][0] + rs[2][1]) == [expect[2]]
    assert ctx.prove_batch(dpk, m, b"".join(zs[:5]), b"".join(r + s for r, s in rs[:5])) == expect[:5]
    # 70 proofs: the sequential record folds (batches of >= 64 vectors) and five warps of the packed finish kernel
    reps = [i % K for i in range(70)]
    assert ctx.prove_batch(dpk, m, b"".join(zs[i] for i in reps), b"".join(rs[i][0] + rs[i][1] for i in reps)) == [expect[i] for i in reps]
    # the same assignments handed over as Montgomery limbs (ZKB_BATCH_Z_MONTGOMERY: what the native batch prover's host threads write)
    def mont(zb):
        return b"".join(((int.from_bytes(zb[i:i + 32], "little") << 256) % R).to_bytes(32, "little") for i in range(0, len(zb), 32))
    assert ctx.prove_batch(dpk, m, b"".join(mont(z) for z in zs[:5]), b"".join(r + s for r, s in rs[:5]), montgomery=True) == expect[:5]
    with pytest.raises(Exception):                                          # a limb pattern >= r is refused in either form
        ctx.prove_batch(dpk, m, mont(zs[0])[:-32] + R.to_bytes(32, "little"), rs[0][0] + rs[0][1], montgomery=True)
    with pytest.raises(Exception):
        ctx.prove_batch(dpk, m, b"".join(zs[:2]), rs[0][0] + rs[0][1])      # two assignments, one (r, s)
    m.free()
    dpk.free()


def test_msm_batch_equals_single_msms(ctx):
    """msm_run_batch through zkb_debug_msm_batch: K scalar vectors against one table == K separate MSMs (G1 and G2), with
    zero / one / maximal scalars and a vector of all zeros among them.  K = 9 folds the segment records with the warp levels,
    K = 67 (>= MSM_SEQ_FOLD_BATCH) with the sequential levels."""
    import numpy as np
    import torch
    for n, K in ((3000, 9), (700, 67)):
        k = _rand_fr_np(n, 81)
        for group in (1, 2):
            gen = ctx.g1_bases_generate if group == 1 else ctx.g2_bases_generate
            bases = gen(torch.from_numpy(k.view(np.int32)).cuda(), n)
            sc = _rand_fr_np(n * K, 82 + group).reshape(K, n, 8)
            sc[1] = 0
            sc[2, ::2] = 0
            sc[3, :, 1:] = 0
            sc[3, :, 0] = 1
            rm1 = np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint32)
            sc[4, :100] = rm1
            single = [(ctx.msm_g1 if group == 1 else ctx.msm_g2)(bases, sc[p]) for p in range(K)]
            got = ctx.debug_msm_batch(group, bases, torch.from_numpy(sc.view(np.int32).copy()).cuda(), n, n, K)
            assert got == single
            # a sub-range with a stride larger than the vector length
            cut = n // 6
            got = ctx.debug_msm_batch(group, bases, torch.from_numpy(sc.view(np.int32).copy()).cuda()[:, 5:], n - cut, n, K, offset=5)
            assert got == [(ctx.msm_g1 if group == 1 else ctx.msm_g2)(bases, sc[p, 5:n - cut + 5], offset=5) for p in range(K)]
            bases.free()


def test_sharded_msm_engine_orders_a_default_stream_context(ctx):
    """zelana_b200.multi.GpuMsmEngine / ShardedMsm with the DEFAULT Context (its own non-blocking stream, not torch's): the
    scalars are produced by torch kernels on torch's stream immediately before the MSM and the result is read by torch right
    after the combine -- the engine has to order the two streams itself (round-1 advisor finding).  Both groups, run twice."""
    import numpy as np
    import torch
    from zelana_b200.multi import GpuMsmEngine, ShardedMsm
    assert ctx.stream_handle != torch.cuda.current_stream().cuda_stream
    n = 200_000
    k = torch.from_numpy(_rand_fr_np(n, 71).view(np.int32)).cuda()
    for group, cnt in ((1, n), (2, 20_000)):
        bases = (ctx.g1_bases_generate if group == 1 else ctx.g2_bases_generate)(k, cnt)
        sm = ShardedMsm(GpuMsmEngine(ctx, bases, group=group))
        host = _rand_fr_np(cnt, 72 + group)
        want = (ctx.msm_g1 if group == 1 else ctx.msm_g2)(bases, host)
        for rep in range(2):
            big = torch.from_numpy(host.view(np.int32)).cuda()
            sc = (big ^ 0x55555555) ^ 0x55555555            # a torch kernel writes the scalars just before the MSM reads them
            got = sm.run(sc, cnt).cpu().numpy().tobytes()
            assert got == want, (group, rep)
        bases.free()


def test_entry_points_on_a_second_gpu_from_a_thread_on_the_first(l2_setup):
    """A host thread's current CUDA device is 0 unless it says otherwise; every entry point must put ITS context's device in
    place for everything it does, including the wait at the end (a blocking-sync event created on the wrong device is an
    'invalid resource handle': found by the 8-GPU bench, where ranks 1..7 prove small batches through zkb_prove).
    Needs two GPUs; the single-GPU box skips it."""
    import torch
    import zelana_b200
    from zelana_b200 import l2_circuit as P2
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    circ, pk_bytes, vk_bytes, raw = l2_setup
    torch.cuda.set_device(0)
    c0, c1 = zelana_b200.Context(0), zelana_b200.Context(1)
    for c in (c0, c1):
        assert c.lib.zkb_ctx_set_blocking_sync(c.h, 1) == 0
    proofs = []
    for c in (c0, c1):
        dpk = c.proving_key_compressed(pk_bytes, validate=False)
        single = P2.L2Prover(c, circ, dpk, vk_bytes)
        ck = P2.L2BlockCircuit(transactions=[P2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 5)],
                               initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): 7}, batch_id=9)
        ck = ck.with_inputs(P2.satisfying_inputs(ck))
        one = single.prove_circuit(ck).proof_bytes
        bp = P2.L2BatchProver(c, circ, dpk, lanes=4)
        assert bp.prove([ck] * 3) == [one] * 3            # per-proof path (<= 8), worker threads of the pool
        assert bp.prove([ck] * 12) == [one] * 12          # batched path
        bp.close()
        proofs.append(one)
        single.m.free()
        dpk.free()
        assert torch.cuda.current_device() == 0           # the caller's device is left alone
    assert proofs[0] == proofs[1]
    c0.close()
    c1.close()


def test_msm_multi_c_abi_equals_single_msm(ctx):
    """zkb_msm_g1_multi / zkb_msm_g2_multi (one process, one context per GPU, range-sharded bases, host
    gather of the partial sums): same bytes as one MSM over all points.  Contexts sit on distinct GPUs when the box has them,
    otherwise on the same one (the entry point does not care); also the >= 2^20 sliced-upload path per shard."""
    import numpy as np
    import torch
    import zelana_b200
    from zelana_b200.api import msm_multi
    ndev = torch.cuda.device_count()
    for group, n, world in ((1, 10001, 3), (2, 3001, 2), (1, (1 << 21) + 77, 2)):
        k = _rand_fr_np(n, 90 + group)
        s = _rand_fr_np(n, 92 + group)
        s[5:50] = 0
        gen_all = ctx.g1_bases_generate if group == 1 else ctx.g2_bases_generate
        whole = gen_all(torch.from_numpy(k.view(np.int32)).cuda(), n)
        want = (ctx.msm_g1 if group == 1 else ctx.msm_g2)(whole, s)
        raw = whole.read()
        whole.free()
        sz = 64 if group == 1 else 128
        ctxs = [zelana_b200.Context(i % ndev) for i in range(world)]
        bounds = [n * i // world for i in range(world + 1)]
        shards = [(c.g1_bases if group == 1 else c.g2_bases)(raw[sz * bounds[i]:sz * bounds[i + 1]])
                  for i, c in enumerate(ctxs)]
        assert msm_multi(ctxs, shards, s, group=group) == want
        assert msm_multi(ctxs, shards, s, group=group) == want          # buffers reused
        for b in shards:
            b.free()
        for c in ctxs:
            c.close()


def test_prove_multi_c_abi_equals_prove(ctx, mimc_setup):
    """zkb_prove_multi: ONE proof over the contexts of one process (key sharded by range, matrices replicated, partial records
    gathered through host memory) == zkb_prove."""
    import torch
    import zelana_b200
    from zelana_b200.api import prove_multi
    r1cs, z, pk = mimc_setup
    parts = pk_parts(pk)
    m = ctx.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
    dpk = ctx.proving_key(**parts)
    r, s = fr_bytes([123456789]), fr_bytes([R - 2])
    want = ctx.prove(dpk, m, fr_bytes(z), r, s)
    ndev = torch.cuda.device_count()
    for world in (1, 2, 3):
        ctxs = [zelana_b200.Context(i % ndev) for i in range(world)]
        pks = [c.proving_key_shard(i, world, **parts) for i, c in enumerate(ctxs)]
        ms = [c.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c) for c in ctxs]
        assert prove_multi(ctxs, pks, ms, fr_bytes(z), r, s) == want
        for x in pks + ms:
            x.free()
        for c in ctxs:
            c.close()
    m.free()
    dpk.free()


@pytest.mark.parametrize("c", [3, 8, 11])
def test_msm_comb_tables_equal_single_msms(ctx, c):
    """The comb-table MSM of a batched prove (every digit multiple d * 2^(c w) * P_i resident in HBM, an MSM = a sum of
    gathered points: no buckets, no sort): K scalar vectors == K separate bucket-method MSMs, G1 and G2, with zero / one /
    r - 1 scalars, an all-zero vector, infinity bases, and a sub-range with a stride."""
    import numpy as np
    import torch
    n, K = 700, 5
    k = _rand_fr_np(n, 95)
    k[10:20] = 0                     # [0] G = infinity: bases the key marks as absent
    for group in (1, 2):
        gen = ctx.g1_bases_generate if group == 1 else ctx.g2_bases_generate
        msm = ctx.msm_g1 if group == 1 else ctx.msm_g2
        bases = gen(torch.from_numpy(k.view(np.int32)).cuda(), n)
        sc = _rand_fr_np(n * K, 96 + group).reshape(K, n, 8)
        sc[1] = 0
        sc[2, ::2] = 0
        sc[3, :, 1:] = 0
        sc[3, :, 0] = 1
        sc[4, :100] = np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint32)
        sd = torch.from_numpy(sc.view(np.int32).copy()).cuda()
        assert ctx.debug_msm_comb(group, bases, sd, n, n, K, c) == [msm(bases, sc[p]) for p in range(K)]
        assert ctx.debug_msm_comb(group, bases, sd[:, 7:], n - 300, n, K, c, offset=7) == \
            [msm(bases, sc[p, 7:n - 293], offset=7) for p in range(K)]
        bases.free()


def test_scratch_guards_stay_intact(l2_setup, mimc_setup):
    """compute-sanitizer is closed on this GPU pool; instead every scratch buffer of a fresh context is wrapped in 4 KiB canaries
    (ZKB_GUARD=1) and the canaries are verified after ragged / empty / degenerate MSMs of both groups, NTTs, a prove, a batched
    prove and the MiMC / Poseidon kernels: nothing writes outside its layout."""
    import os
    import numpy as np
    import torch
    import zelana_b200
    from zelana_b200 import l2_circuit as P2
    os.environ["ZKB_GUARD"] = "1"
    try:
        g = zelana_b200.Context(0)
        for n in (1, 31, 257, 5000, (1 << 16) + 3):
            k = _rand_fr_np(n, 300 + n % 7)
            b1 = g.g1_bases_generate(torch.from_numpy(k.view(np.int32)).cuda(), n)
            s = _rand_fr_np(n, 301)
            s[: n // 3] = 0
            g.msm_g1(b1, s)
            if n <= 5000:
                b2 = g.g2_bases_generate(torch.from_numpy(k.view(np.int32)).cuda(), n)
                g.msm_g2(b2, s)
                sd = torch.from_numpy(np.repeat(s[None], 3, axis=0).view(np.int32).copy()).cuda()
                g.debug_msm_batch(2, b2, sd, n, n, 3)
                g.debug_msm_batch(1, b1, sd, n, n, 3)
                b2.free()
            b1.free()
        big = (1 << 20) + 12345           # the sliced host-scalar path
        kb = _rand_fr_np(big, 310)
        bb = g.g1_bases_generate(torch.from_numpy(kb.view(np.int32)).cuda(), big)
        g.msm_g1(bb, _rand_fr_np(big, 311))
        bb.free()
        for lg in (1, 7, 13, 17):
            g.ntt(_rand_fr_np(1 << lg, 320 + lg), lg, inverse=True, coset=True)
        r1cs, z, pk = mimc_setup
        m = g.r1cs(r1cs.num_instance, r1cs.num_witness, r1cs.a, r1cs.b, r1cs.c)
        dpk = g.proving_key(**pk_parts(pk))
        g.prove(dpk, m, fr_bytes(z), fr_bytes([5]), fr_bytes([7]))
        g.prove_batch(dpk, m, fr_bytes(z) * 3, (fr_bytes([5]) + fr_bytes([7])) * 3)
        g.mimc_hash(2, _rand_fr_np(2 * 77, 330).tobytes())
        g.l2_poseidon_hash_batch(3, _rand_fr_np(3 * 77, 331).tobytes())
        g.debug_check_guards()
        m.free()
        dpk.free()
        g.close()
    finally:
        os.environ.pop("ZKB_GUARD", None)
