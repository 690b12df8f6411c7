"""Thin object layer over the C ABI: contexts, device-resident bases, MSM, NTT, witness map, prove."""
import ctypes as C

import numpy as np

from ._lib import ZkbError, load_library, R1csDesc, PkDesc, Csr, SetupParams, SetupOut

FR, FQ = 0, 1
OP_ADD, OP_SUB, OP_MUL, OP_INV, OP_NEG = range(5)
G1_PARTIAL_BYTES, G2_PARTIAL_BYTES = 128, 256
PROVE_PARTIAL_BYTES = 768


def _buf(x):
    """bytes / bytearray / numpy uint8 array -> (ctypes pointer, keepalive)."""
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x).view(np.uint8).reshape(-1)
        return a.ctypes.data_as(C.c_void_p), a
    if isinstance(x, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(x), dtype=np.uint8)
        return a.ctypes.data_as(C.c_void_p), a
    raise TypeError("expected bytes or numpy array, got %r" % type(x))


def _devptr(t):
    """torch CUDA tensor or int device address -> c_void_p."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class Context:
    """One CUDA device + stream.  Not thread-safe; make one per thread / per GPU."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.zkb_ctx_create(device, C.byref(h))
        if rc != 0:
            raise ZkbError(rc, "zkb_ctx_create(device=%d) failed (a CUDA device is required; no CPU fallback)" % device)
        self.h = h
        self.device = device
        self.stream_handle = None      # None: the context's own non-blocking stream
        if stream is not None:
            self.set_stream(stream)

    def _check(self, rc):
        if rc != 0:
            raise ZkbError(rc, self.lib.zkb_last_error(self.h).decode())

    def set_stream(self, stream):
        self._check(self.lib.zkb_ctx_set_stream(self.h, C.c_void_p(int(stream) if stream else 0)))
        self.stream_handle = int(stream) if stream else None

    def synchronize(self):
        self._check(self.lib.zkb_ctx_synchronize(self.h))

    def debug_check_guards(self):
        """Verify the canaries around every scratch buffer (allocated with ZKB_GUARD=1 in the environment)."""
        self._check(self.lib.zkb_debug_check_guards(self.h))

    def launch_count(self):
        return int(self.lib.zkb_launch_count(self.h))

    def set_graphs(self, on=True):
        """CUDA-graph replay of a prove's device part (default on)."""
        self._check(self.lib.zkb_ctx_set_graphs(self.h, int(bool(on))))

    def graph_stats(self):
        """(captures, replays) of prove graphs on this context."""
        cap, rep = C.c_ulonglong(0), C.c_ulonglong(0)
        self._check(self.lib.zkb_graph_stats(self.h, C.byref(cap), C.byref(rep)))
        return int(cap.value), int(rep.value)

    def set_msm_window(self, c):
        self._check(self.lib.zkb_ctx_set_msm_window(self.h, c))

    def profile(self, on=True):
        """Bracket every phase (MSM digits/sort/accumulate/reduce, NTT, ...) with CUDA events on the context's stream."""
        self._check(self.lib.zkb_prof_enable(self.h, int(on)))

    def profile_reset(self):
        self._check(self.lib.zkb_prof_reset(self.h))

    def profile_read(self):
        """-> {phase name: (total device ms, spans)} since the last reset (synchronises the stream)."""
        out = {}
        for ph in range(self.lib.zkb_prof_phase_count()):
            ms, cnt = C.c_double(), C.c_ulonglong()
            self._check(self.lib.zkb_prof_read(self.h, ph, C.byref(ms), C.byref(cnt)))
            if cnt.value:
                out[self.lib.zkb_prof_phase_name(ph).decode()] = (ms.value, int(cnt.value))
        return out

    def int32_peak(self, variant=0, iters=2000):
        """Measured INT32 multiply-pipe peak (mul32/s) of this GPU; variant 0 = IMAD.WIDE, 1 = lo/hi pairs."""
        rate, ms = C.c_double(), C.c_double()
        self._check(self.lib.zkb_bench_int32_peak(self.h, variant, iters, C.byref(rate), C.byref(ms)))
        return rate.value, ms.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.zkb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parity hooks
    def field_op(self, field, op, a, b=None):
        pa, ka = _buf(a)
        n = len(ka) // 32
        out = np.empty(n * 32, dtype=np.uint8)
        if b is not None:
            pb, kb = _buf(b)
        else:
            pb, kb = C.c_void_p(0), None
        self._check(self.lib.zkb_field_op(self.h, field, op, pa, pb, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def scalar_mul(self, group, points, scalars):
        pp, kp = _buf(points)
        ps, ks = _buf(scalars)
        sz = 64 if group == 1 else 128
        n = len(kp) // sz
        out = np.empty(n * sz, dtype=np.uint8)
        self._check(self.lib.zkb_scalar_mul(self.h, group, pp, ps, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def point_sum(self, group, points):
        sz = 64 if group == 1 else 128
        if len(points) == 0:
            pp, kp, n = C.c_void_p(0), None, 0
        else:
            pp, kp = _buf(points)
            n = len(kp) // sz
        out = np.empty(sz, dtype=np.uint8)
        self._check(self.lib.zkb_point_sum(self.h, group, pp, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    # ---- bases
    def g1_bases(self, affine_bytes, validate=True):
        return G1Bases._load(self, affine_bytes, validate)

    def g2_bases(self, affine_bytes, validate=True):
        return G2Bases._load(self, affine_bytes, validate)

    def g1_bases_generate(self, k_dev, n):
        h = C.c_void_p()
        self._check(self.lib.zkb_g1_bases_generate(self.h, _devptr(k_dev), n, C.byref(h)))
        return G1Bases(self, h)

    def g2_bases_generate(self, k_dev, n):
        h = C.c_void_p()
        self._check(self.lib.zkb_g2_bases_generate(self.h, _devptr(k_dev), n, C.byref(h)))
        return G2Bases(self, h)

    # ---- MSM
    def msm_g1(self, bases, scalars, offset=0):
        ps, ks = _buf(scalars) if len(scalars) else (C.c_void_p(0), None)
        n = len(ks) // 32 if ks is not None else 0
        out = np.empty(64, dtype=np.uint8)
        self._check(self.lib.zkb_msm_g1(self.h, bases.h, offset, ps, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def msm_g2(self, bases, scalars, offset=0):
        ps, ks = _buf(scalars) if len(scalars) else (C.c_void_p(0), None)
        n = len(ks) // 32 if ks is not None else 0
        out = np.empty(128, dtype=np.uint8)
        self._check(self.lib.zkb_msm_g2(self.h, bases.h, offset, ps, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def msm_g1_dev(self, bases, scalars_dev, n, out_affine_dev=None, out_partial_dev=None, offset=0):
        self._check(self.lib.zkb_msm_g1_dev(self.h, bases.h, offset, _devptr(scalars_dev), n,
                                            _devptr(out_affine_dev), _devptr(out_partial_dev)))

    def msm_g2_dev(self, bases, scalars_dev, n, out_affine_dev=None, out_partial_dev=None, offset=0):
        self._check(self.lib.zkb_msm_g2_dev(self.h, bases.h, offset, _devptr(scalars_dev), n,
                                            _devptr(out_affine_dev), _devptr(out_partial_dev)))

    def msm_g1_partial(self, bases, scalars, out_partial_dev, offset=0):
        """Host scalars -> this rank's projective partial sum in device memory (zkb_msm_g1_partial; asynchronous)."""
        ps, ks = _buf(scalars) if len(scalars) else (C.c_void_p(0), None)
        n = len(ks) // 32 if ks is not None else 0
        self._check(self.lib.zkb_msm_g1_partial(self.h, bases.h, offset, ps, n, _devptr(out_partial_dev)))
        return ks   # keep the host buffer alive until the stream has consumed it

    def msm_g2_partial(self, bases, scalars, out_partial_dev, offset=0):
        ps, ks = _buf(scalars) if len(scalars) else (C.c_void_p(0), None)
        n = len(ks) // 32 if ks is not None else 0
        self._check(self.lib.zkb_msm_g2_partial(self.h, bases.h, offset, ps, n, _devptr(out_partial_dev)))
        return ks

    def msm_g1_combine(self, partials_dev, k, out_affine_dev):
        self._check(self.lib.zkb_msm_g1_combine(self.h, _devptr(partials_dev), k, _devptr(out_affine_dev)))

    def msm_g2_combine(self, partials_dev, k, out_affine_dev):
        self._check(self.lib.zkb_msm_g2_combine(self.h, _devptr(partials_dev), k, _devptr(out_affine_dev)))

    def debug_msm_batch(self, group, bases, scalars_dev, n, stride, batch, offset=0):
        """Parity hook: `batch` MSMs over the same bases in one batched pass -> list of canonical affine byte strings."""
        import torch
        sz = 64 if group == 1 else 128
        out = torch.zeros(batch * sz, dtype=torch.uint8, device=scalars_dev.device)
        torch.cuda.synchronize()       # the library runs on its own non-blocking stream: torch's fill must be done first
        self._check(self.lib.zkb_debug_msm_batch(self.h, group, bases.h, offset, _devptr(scalars_dev), n, stride, batch, _devptr(out)))
        self.synchronize()
        o = bytes(out.cpu().numpy())
        return [o[sz * i:sz * (i + 1)] for i in range(batch)]

    def debug_msm_comb(self, group, bases, scalars_dev, n, stride, batch, c, offset=0):
        """Parity hook: `batch` MSMs through the comb table (every digit multiple resident) -> canonical affine byte strings."""
        import torch
        sz = 64 if group == 1 else 128
        out = torch.zeros(batch * sz, dtype=torch.uint8, device=scalars_dev.device)
        torch.cuda.synchronize()
        self._check(self.lib.zkb_debug_msm_comb(self.h, group, bases.h, offset, _devptr(scalars_dev), n, stride, batch, c, _devptr(out)))
        self.synchronize()
        o = bytes(out.cpu().numpy())
        return [o[sz * i:sz * (i + 1)] for i in range(batch)]

    def debug_msm_entries(self, bases, scalars_dev, n, stride, batch, offset=0):
        """Parity hook: the sorted (key, value) entry lists of the MSM front end -> (keys, vals) numpy uint32 arrays."""
        import torch
        c, nwin = bases.window()
        cap = nwin * n * batch
        keys = torch.zeros(cap, dtype=torch.int32, device=scalars_dev.device)
        vals = torch.zeros(cap, dtype=torch.int32, device=scalars_dev.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=scalars_dev.device)
        torch.cuda.synchronize()
        self._check(self.lib.zkb_debug_msm_entries(self.h, bases.h, offset, _devptr(scalars_dev), n, stride, batch, _devptr(keys),
                                                   _devptr(vals), _devptr(cnt)))
        m = int(cnt.cpu().numpy().view(np.uint32)[0])
        return keys.cpu().numpy().view(np.uint32)[:m].copy(), vals.cpu().numpy().view(np.uint32)[:m].copy()

    # ---- NTT
    def ntt(self, data, log_n, inverse=False, coset=False):
        pi, ki = _buf(data)
        assert len(ki) == 32 << log_n
        out = np.empty(len(ki), dtype=np.uint8)
        self._check(self.lib.zkb_ntt(self.h, pi, out.ctypes.data_as(C.c_void_p), log_n, int(inverse), int(coset)))
        return out.tobytes()

    def ntt_dev(self, in_dev, out_dev, log_n, inverse=False, coset=False):
        self._check(self.lib.zkb_ntt_dev(self.h, _devptr(in_dev), _devptr(out_dev), log_n, int(inverse), int(coset)))

    # ---- the L2 circuit's Poseidon, batched
    def l2_poseidon_hash_batch(self, arity, data, n=None):
        """n independent Poseidon hashes of `arity` (0..3) elements each with the L2 circuit's parameters -> n x 32 B."""
        if arity:
            p, k = _buf(data)
            n = len(k) // (32 * arity)
        else:
            p, k = C.c_void_p(0), None
        if not n:
            return b""
        out = np.empty(n * 32, dtype=np.uint8)
        self._check(self.lib.zkb_l2_poseidon_hash_batch(self.h, arity, p, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def l2_poseidon_hash_batch_dev(self, arity, in_dev, n, out_dev):
        self._check(self.lib.zkb_l2_poseidon_hash_batch_dev(self.h, arity, _devptr(in_dev), n, _devptr(out_dev)))

    # ---- MiMC-7 / account Merkle tree (forge stack)
    def mimc_hash(self, arity, data):
        """n independent hash_arity(...) of the forge stack's MiMC-7 (poseidon.nr:62-94): data = n x arity x 32 B LE -> n x 32 B."""
        if len(data) == 0:
            return b""
        p, k = _buf(data)
        if arity < 1 or len(k) % (32 * arity):
            raise ZkbError(-3, "mimc_hash: %d bytes is not a whole number of %d-element inputs" % (len(k), arity))
        n = len(k) // (32 * arity)
        out = np.empty(n * 32, dtype=np.uint8)
        self._check(self.lib.zkb_mimc_hash(self.h, arity, p, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def mimc_hash_dev(self, arity, in_dev, n, out_dev):
        self._check(self.lib.zkb_mimc_hash_dev(self.h, arity, _devptr(in_dev), n, _devptr(out_dev)))

    def mimc_merkle_roots(self, leaves, siblings, bits, depth=32):
        """n roots: leaves n x 32 B, siblings n x depth x 32 B, bits n x depth bytes (1 = running node is the right child)."""
        pl, kl = _buf(leaves)
        n = len(kl) // 32
        ps, ks = _buf(siblings) if depth else (C.c_void_p(0), b"")
        pb, kb = _buf(bits) if depth else (C.c_void_p(0), b"")
        if len(ks) != n * depth * 32 or len(kb) != n * depth:
            raise ZkbError(-6, "mimc_merkle_roots: %d leaves need %d sibling bytes and %d index bytes" % (n, n * depth * 32, n * depth))
        out = np.empty(n * 32, dtype=np.uint8)
        self._check(self.lib.zkb_mimc_merkle_roots(self.h, pl, ps, pb, n, depth, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def mimc_merkle_roots_dev(self, leaves_dev, siblings_dev, bits_dev, n, depth, out_dev):
        self._check(self.lib.zkb_mimc_merkle_roots_dev(self.h, _devptr(leaves_dev), _devptr(siblings_dev), _devptr(bits_dev), n, depth,
                                                       _devptr(out_dev)))

    # ---- Groth16
    def r1cs(self, num_instance, num_witness, a, b, c):
        return R1csMatrices(self, num_instance, num_witness, a, b, c)

    def proving_key(self, **parts):
        return ProvingKeyDev(self, **parts)

    def proving_key_compressed(self, ark_bytes, validate=True):
        """ProvingKey::deserialize_compressed on the GPU: the bytes keygen wrote (prover/src/bin/keygen.rs:100)."""
        p, keep = _buf(ark_bytes)
        h = C.c_void_p()
        self._check(self.lib.zkb_pk_load_compressed(self.h, p, len(keep), int(validate), C.byref(h)))
        pk = ProvingKeyDev.__new__(ProvingKeyDev)
        pk.ctx, pk.h = self, h
        return pk

    def proving_key_synthetic(self, num_vars, num_witness, h_len, k_dev, k_len, shard=0, world=1):
        """Benchmark-only key (or key shard) of the given shape (query points [k_i] G from device scalars); proofs do not verify."""
        h = C.c_void_p()
        self._check(self.lib.zkb_pk_synthetic_shard(self.h, num_vars, num_witness, h_len, _devptr(k_dev), k_len, shard, world,
                                                    C.byref(h)))
        pk = ProvingKeyDev.__new__(ProvingKeyDev)
        pk.ctx, pk.h = self, h
        return pk

    def setup(self, num_instance, num_witness, a, b, c, *, alpha, beta, gamma, delta, tau, g1_generator, g2_generator):
        """Groth16 parameter generation on the GPU (zkb_setup) from explicit toxic waste -> dict of raw affine byte strings."""
        csr = [_checked_csr(m, num_instance + num_witness) for m in (a, b, c)]
        if not (len(csr[0][0]) == len(csr[1][0]) == len(csr[2][0])):
            raise ZkbError(-6, "A, B and C must have the same number of rows")
        keep = []
        d = R1csDesc()
        d.num_constraints, d.num_instance, d.num_witness = len(csr[0][0]) - 1, num_instance, num_witness
        for name, (rp, col, co) in zip(("a", "b", "c"), csr):
            keep += [rp, col, co]
            setattr(d, name, Csr(rp.ctypes.data, col.ctypes.data, co.ctypes.data))
        prm = SetupParams()
        for name, v in (("alpha", alpha), ("beta", beta), ("gamma", gamma), ("delta", delta), ("tau", tau)):
            getattr(prm, name)[:] = list(int(v).to_bytes(32, "little"))
        prm.g1_generator[:] = list(g1_generator)
        prm.g2_generator[:] = list(g2_generator)
        nv = num_instance + num_witness
        n = 1
        while n < d.num_constraints + num_instance:
            n <<= 1
        sizes = {"alpha_g1": 64, "beta_g1": 64, "delta_g1": 64, "beta_g2": 128, "gamma_g2": 128, "delta_g2": 128,
                 "gamma_abc_g1": 64 * num_instance, "a_query": 64 * nv, "b_g1_query": 64 * nv, "b_g2_query": 128 * nv,
                 "h_query": 64 * (n - 1), "l_query": 64 * num_witness}
        bufs = {k: np.zeros(max(v, 1), dtype=np.uint8) for k, v in sizes.items()}
        out = SetupOut()
        for k, arr in bufs.items():
            setattr(out, k, arr.ctypes.data)
        self._check(self.lib.zkb_setup(self.h, C.byref(d), C.byref(prm), C.byref(out)))
        return {k: bufs[k][:sizes[k]].tobytes() for k in sizes}

    @staticmethod
    def _check_z(r1cs, kz):
        """The C ABI takes a bare pointer: a short assignment would be an out-of-bounds host read, so check it here."""
        want = (r1cs.num_instance + r1cs.num_witness) * 32
        if len(kz) != want:
            raise ZkbError(-6, "assignment has %d bytes, the matrices need %d (num_instance + num_witness) x 32" % (len(kz), want))

    def witness_map(self, r1cs, z_bytes):
        pz, kz = _buf(z_bytes)
        self._check_z(r1cs, kz)
        n = 1 << r1cs.log_domain
        out = np.empty(n * 32, dtype=np.uint8)
        self._check(self.lib.zkb_witness_map(self.h, r1cs.h, pz, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def proving_key_shard(self, shard, world, **parts):
        """Rank `shard` of `world`: only its contiguous range of every query vector is uploaded (zkb_pk_load_shard)."""
        return ProvingKeyDev(self, shard=shard, world=world, **parts)

    def prove_partial(self, pk, r1cs, z_bytes, r_bytes, s_bytes, out_partial_dev):
        """This rank's share of a sharded proof -> PROVE_PARTIAL_BYTES at out_partial_dev (device memory)."""
        pz, kz = _buf(z_bytes)
        self._check_z(r1cs, kz)
        pr, kr = _buf(r_bytes)
        ps, ks = _buf(s_bytes)
        self._check(self.lib.zkb_prove_partial(self.h, pk.h, r1cs.h, pz, pr, ps, _devptr(out_partial_dev)))

    def prove_combine(self, partials_dev, world, r_bytes, s_bytes):
        pr, kr = _buf(r_bytes)
        ps, ks = _buf(s_bytes)
        oa = np.empty(64, dtype=np.uint8)
        ob = np.empty(128, dtype=np.uint8)
        oc = np.empty(64, dtype=np.uint8)
        self._check(self.lib.zkb_prove_combine(self.h, _devptr(partials_dev), world, pr, ps, oa.ctypes.data_as(C.c_void_p),
                                               ob.ctypes.data_as(C.c_void_p), oc.ctypes.data_as(C.c_void_p)))
        return oa.tobytes(), ob.tobytes(), oc.tobytes()

    def prove_batch(self, pk, r1cs, z_bytes, rs_bytes, montgomery=False):
        """K proofs of one circuit/key in one set of batched launches (zkb_prove_batch).  z_bytes: K assignments back to back;
        rs_bytes: K x (r || s).  -> list of (A, B, C) byte triples, identical to K calls of prove().
        montgomery=True: z_bytes holds Montgomery limbs (x * 2^256 mod r), zkb_prove_batch_begin_ex / ZKB_BATCH_Z_MONTGOMERY."""
        pz, kz = _buf(z_bytes)
        prs, krs = _buf(rs_bytes)
        per = (r1cs.num_instance + r1cs.num_witness) * 32
        k = len(krs) // 64
        if k < 1 or len(krs) != k * 64 or len(kz) != k * per:
            raise ZkbError(-6, "prove_batch: %d bytes of assignments and %d bytes of (r, s) do not describe the same number of proofs"
                           % (len(kz), len(krs)))
        out = np.empty(k * 256, dtype=np.uint8)
        if montgomery:
            self._check(self.lib.zkb_prove_batch_begin_ex(self.h, pk.h, r1cs.h, pz, prs, k, 1))
            self._check(self.lib.zkb_prove_batch_end(self.h, k, out.ctypes.data_as(C.c_void_p)))
        else:
            self._check(self.lib.zkb_prove_batch(self.h, pk.h, r1cs.h, pz, prs, k, out.ctypes.data_as(C.c_void_p)))
        o = out.tobytes()
        return [(o[256 * i:256 * i + 64], o[256 * i + 64:256 * i + 192], o[256 * i + 192:256 * i + 256]) for i in range(k)]

    def prove(self, pk, r1cs, z_bytes, r_bytes, s_bytes):
        pz, kz = _buf(z_bytes)
        self._check_z(r1cs, kz)
        pr, kr = _buf(r_bytes)
        ps, ks = _buf(s_bytes)
        if len(kr) != 32 or len(ks) != 32:
            raise ZkbError(-3, "r and s are 32-byte canonical Fr elements")
        oa = np.empty(64, dtype=np.uint8)
        ob = np.empty(128, dtype=np.uint8)
        oc = np.empty(64, dtype=np.uint8)
        self._check(self.lib.zkb_prove(self.h, pk.h, r1cs.h, pz, pr, ps, oa.ctypes.data_as(C.c_void_p),
                                       ob.ctypes.data_as(C.c_void_p), oc.ctypes.data_as(C.c_void_p)))
        return oa.tobytes(), ob.tobytes(), oc.tobytes()


def msm_multi(ctxs, bases, scalars, group=1):
    """zkb_msm_g{1,2}_multi: one process, one Context per GPU, bases[i] = range i of the points on ctxs[i]; host scalars."""
    lib = ctxs[0].lib
    ps, ks = _buf(scalars) if len(scalars) else (C.c_void_p(0), None)
    n = len(ks) // 32 if ks is not None else 0
    k = len(ctxs)
    ca = (C.c_void_p * k)(*[c.h for c in ctxs])
    ba = (C.c_void_p * k)(*[b.h for b in bases])
    out = np.empty(64 if group == 1 else 128, dtype=np.uint8)
    fn = lib.zkb_msm_g1_multi if group == 1 else lib.zkb_msm_g2_multi
    ctxs[0]._check(fn(ca, ba, k, ps, n, out.ctypes.data_as(C.c_void_p)))
    return out.tobytes()


def prove_multi(ctxs, pk_shards, r1cs_list, z_bytes, r_bytes, s_bytes):
    """zkb_prove_multi: ONE proof over the GPUs of this process (one Context, key shard and matrices copy per GPU)."""
    lib = ctxs[0].lib
    k = len(ctxs)
    pz, kz = _buf(z_bytes)
    Context._check_z(r1cs_list[0], kz)
    pr, kr = _buf(r_bytes)
    ps, ks = _buf(s_bytes)
    ca = (C.c_void_p * k)(*[c.h for c in ctxs])
    pa = (C.c_void_p * k)(*[p.h for p in pk_shards])
    ma = (C.c_void_p * k)(*[m.h for m in r1cs_list])
    oa, ob, oc = np.empty(64, dtype=np.uint8), np.empty(128, dtype=np.uint8), np.empty(64, dtype=np.uint8)
    ctxs[0]._check(lib.zkb_prove_multi(ca, pa, ma, k, pz, pr, ps, oa.ctypes.data_as(C.c_void_p), ob.ctypes.data_as(C.c_void_p),
                                       oc.ctypes.data_as(C.c_void_p)))
    return oa.tobytes(), ob.tobytes(), oc.tobytes()


class _Bases:
    _free = None
    _len = None
    _read = None
    _size = 0

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __len__(self):
        return int(getattr(self.ctx.lib, self._len)(self.h))

    def window(self):
        """(c, nwin): window width and number of resident window tables (nwin x len points in HBM)."""
        c, nwin = C.c_int(0), C.c_int(0)
        self.ctx._check(getattr(self.ctx.lib, self._window)(self.h, C.byref(c), C.byref(nwin)))
        return int(c.value), int(nwin.value)

    def read(self, offset=0, n=None):
        n = len(self) - offset if n is None else n
        out = np.empty(n * self._size, dtype=np.uint8)
        self.ctx._check(getattr(self.ctx.lib, self._read)(self.ctx.h, self.h, offset, n, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def free(self):
        if self.h:
            getattr(self.ctx.lib, self._free)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class G1Bases(_Bases):
    _free, _len, _read, _size = "zkb_g1_bases_free", "zkb_g1_bases_len", "zkb_g1_bases_read", 64
    _window = "zkb_g1_bases_window"

    @classmethod
    def _load(cls, ctx, data, validate):
        p, k = _buf(data) if len(data) else (C.c_void_p(0), None)
        h = C.c_void_p()
        ctx._check(ctx.lib.zkb_g1_bases_load(ctx.h, p, len(data) // 64, int(validate), C.byref(h)))
        return cls(ctx, h)


class G2Bases(_Bases):
    _free, _len, _read, _size = "zkb_g2_bases_free", "zkb_g2_bases_len", "zkb_g2_bases_read", 128
    _window = "zkb_g2_bases_window"

    @classmethod
    def _load(cls, ctx, data, validate):
        p, k = _buf(data) if len(data) else (C.c_void_p(0), None)
        h = C.c_void_p()
        ctx._check(ctx.lib.zkb_g2_bases_load(ctx.h, p, len(data) // 128, int(validate), C.byref(h)))
        return cls(ctx, h)


def _csr_arrays(rows):
    """rows: list of [(coeff:int, var:int), ...] -> (row_ptr u64, col u32, coeff bytes)."""
    row_ptr = np.zeros(len(rows) + 1, dtype=np.uint64)
    cols, coeffs = [], bytearray()
    k = 0
    for i, row in enumerate(rows):
        for co, v in row:
            cols.append(v)
            coeffs += int(co).to_bytes(32, "little")
            k += 1
        row_ptr[i + 1] = k
    col = np.asarray(cols, dtype=np.uint32) if cols else np.zeros(1, dtype=np.uint32)
    coeff = np.frombuffer(bytes(coeffs), dtype=np.uint8) if coeffs else np.zeros(32, dtype=np.uint8)
    return row_ptr, col, coeff


def _checked_csr(m, nvars):
    """rows of (coeff, var) or a prebuilt (row_ptr, col, coeff) triple -> contiguous arrays whose lengths agree with row_ptr[-1]
    (the C side trusts row_ptr to index col / coeff)."""
    rp, col, co = m if isinstance(m, tuple) else _csr_arrays(m)
    rp = np.ascontiguousarray(rp, dtype=np.uint64)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    co = np.ascontiguousarray(co, dtype=np.uint8).reshape(-1)
    nnz = int(rp[-1]) if len(rp) else 0
    if len(rp) < 1 or int(rp[0]) != 0:
        raise ZkbError(-6, "CSR row_ptr must start at 0")
    if len(col) < nnz or len(co) < nnz * 32:
        raise ZkbError(-6, "CSR arrays shorter than row_ptr[-1] = %d entries (col %d, coeff %d bytes)" % (nnz, len(col), len(co)))
    if nnz and int(col[:nnz].max()) >= nvars:
        raise ZkbError(-6, "CSR column index beyond the %d variables" % nvars)
    return rp, col, co


class R1csMatrices:
    """Device-resident ConstraintMatrices (ark-relations): rows of (coeff, variable)."""

    def __init__(self, ctx, num_instance, num_witness, a, b, c):
        """a, b, c: rows of (coeff, variable) pairs, or prebuilt CSR triples (row_ptr u64, col u32, coeff u8[nnz*32])."""
        self.ctx = ctx
        keep = []
        d = R1csDesc()
        csr = [_checked_csr(m, num_instance + num_witness) for m in (a, b, c)]
        if not (len(csr[0][0]) == len(csr[1][0]) == len(csr[2][0])):
            raise ZkbError(-6, "A, B and C must have the same number of rows")
        d.num_constraints, d.num_instance, d.num_witness = len(csr[0][0]) - 1, num_instance, num_witness
        for name, (rp, col, co) in zip(("a", "b", "c"), csr):
            keep += [rp, col, co]
            setattr(d, name, Csr(rp.ctypes.data, col.ctypes.data, co.ctypes.data))
        h = C.c_void_p()
        ctx._check(ctx.lib.zkb_r1cs_load(ctx.h, C.byref(d), C.byref(h)))
        self.h = h
        self.log_domain = int(ctx.lib.zkb_r1cs_log_domain(h))
        self.num_instance, self.num_witness = num_instance, num_witness

    def free(self):
        if self.h:
            self.ctx.lib.zkb_r1cs_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ProvingKeyDev:
    """Device-resident Groth16 proving key built from raw affine byte strings."""

    def __init__(self, ctx, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query, b_g2_query,
                 h_query, l_query, validate=True, shard=0, world=1):
        self.ctx = ctx
        keep = []

        def ptr(x):
            if len(x) == 0:
                return None
            p, k = _buf(x)
            keep.append(k)
            return p.value

        d = PkDesc()
        d.alpha_g1, d.beta_g1, d.beta_g2 = ptr(alpha_g1), ptr(beta_g1), ptr(beta_g2)
        d.delta_g1, d.delta_g2 = ptr(delta_g1), ptr(delta_g2)
        d.a_query, d.a_len = ptr(a_query), len(a_query) // 64
        d.b_g1_query, d.b_g1_len = ptr(b_g1_query), len(b_g1_query) // 64
        d.b_g2_query, d.b_g2_len = ptr(b_g2_query), len(b_g2_query) // 128
        d.h_query, d.h_len = ptr(h_query), len(h_query) // 64
        d.l_query, d.l_len = ptr(l_query), len(l_query) // 64
        h = C.c_void_p()
        ctx._check(ctx.lib.zkb_pk_load_shard(ctx.h, C.byref(d), int(validate), shard, world, C.byref(h)))
        self.h = h

    def free(self):
        if self.h:
            self.ctx.lib.zkb_pk_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
