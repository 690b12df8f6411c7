"""ORACLE (test infrastructure + CPU baseline only) -- ctypes binding of oracle/cpu_oracle.cpp.

The C++ file restates arkworks' CPU algorithms (msm_bigint_wnaf, radix-2 FFT, witness map,
create_proof_with_assignment) so that (a) parity tests can check sizes the pure-Python oracle cannot
reach in seconds and (b) bench.py has a multithreaded CPU baseline to time on the GPU box's host cores.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libzkoracle.so")
_lib = None
_P, _SZ, _I, _U64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64


def build(force=False):
    src = os.path.join(HERE, "cpu_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_max_threads.restype = _I
        L.orc_msm_window.restype = _I
        L.orc_msm_window.argtypes = [_SZ]
        L.orc_field_op.argtypes = [_I, _I, _P, _P, _SZ, _P]
        L.orc_msm_g1.argtypes = [_P, _P, _SZ, _I, _P]
        L.orc_msm_g2.argtypes = [_P, _P, _SZ, _I, _P]
        L.orc_g1_bases_new.restype = _P
        L.orc_g1_bases_new.argtypes = [_P, _SZ, _I]
        L.orc_g1_bases_arith.restype = _P
        L.orc_g1_bases_arith.argtypes = [_U64, _SZ, _I]
        L.orc_g1_bases_read.argtypes = [_P, _SZ, _SZ, _P]
        L.orc_g1_bases_free.argtypes = [_P]
        L.orc_msm_g1_pre.argtypes = [_P, _SZ, _P, _SZ, _I, _P]
        L.orc_ntt.argtypes = [_P, _I, _I, _I, _I]
        L.orc_mimc_hash.argtypes = [_I, _P, _SZ, _I, _P]
        L.orc_mimc_merkle_roots.argtypes = [_P, _P, _P, _SZ, _I, _I, _P]
        L.orc_poly_eval.argtypes = [_P, _SZ, _P, _SZ, _I, _P]
        L.orc_root_of_unity.argtypes = [_I, _P]
        L.orc_witness_map.restype = _I
        L.orc_witness_map.argtypes = [_U64, _U64, _U64] + [_P] * 9 + [_P, _P, _I]
        L.orc_pk_new.restype = _P
        L.orc_pk_new.argtypes = [_P] * 5 + [_P, _SZ, _P, _P, _P, _SZ, _P, _SZ, _I]
        L.orc_pk_free.argtypes = [_P]
        L.orc_prove.restype = _I
        L.orc_prove.argtypes = [_P, _U64, _U64, _U64] + [_P] * 9 + [_P, _P, _P, _P, _P, _P, _I]
        _lib = L
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def _np(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(b), dtype=np.uint8) if len(b) else np.zeros(1, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def field_op(field, op, a, b=None):
    A = _np(a)
    n = len(a) // 32
    out = np.empty(max(n, 1) * 32, dtype=np.uint8)
    B = _np(b) if b is not None else None
    lib().orc_field_op(field, op, _ptr(A), _ptr(B) if B is not None else None, n, _ptr(out))
    return out[:n * 32].tobytes()


def msm_window(n):
    return int(lib().orc_msm_window(n))


def msm_g1(bases_raw, scalars, threads=0):
    B, S = _np(bases_raw), _np(scalars)
    n = len(S) // 32
    out = np.empty(64, dtype=np.uint8)
    lib().orc_msm_g1(_ptr(B), _ptr(S), n, threads or max_threads(), _ptr(out))
    return out.tobytes()


def msm_g2(bases_raw, scalars, threads=0):
    B, S = _np(bases_raw), _np(scalars)
    n = len(S) // 32
    out = np.empty(128, dtype=np.uint8)
    lib().orc_msm_g2(_ptr(B), _ptr(S), n, threads or max_threads(), _ptr(out))
    return out.tobytes()


class G1Bases:
    """Pre-parsed (Montgomery, in RAM) bases, as arkworks holds a proving key."""

    def __init__(self, handle, n):
        self.h, self.n = handle, n

    @classmethod
    def from_raw(cls, raw, threads=0):
        B = _np(raw)
        n = len(raw) // 64
        return cls(lib().orc_g1_bases_new(_ptr(B), n, threads or max_threads()), n)

    @classmethod
    def arithmetic(cls, k0, n, threads=0):
        """bases[i] = (k0 + i) G."""
        return cls(lib().orc_g1_bases_arith(k0, n, threads or max_threads()), n)

    def read(self, off=0, n=None):
        n = self.n - off if n is None else n
        out = np.empty(max(n, 1) * 64, dtype=np.uint8)
        lib().orc_g1_bases_read(self.h, off, n, _ptr(out))
        return out[:n * 64].tobytes()

    def msm(self, scalars, off=0, threads=0):
        S = _np(scalars)
        n = len(S) // 32
        assert off + n <= self.n
        out = np.empty(64, dtype=np.uint8)
        lib().orc_msm_g1_pre(self.h, off, _ptr(S), n, threads or max_threads(), _ptr(out))
        return out.tobytes()

    def free(self):
        if self.h:
            lib().orc_g1_bases_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ntt(data, log_n, inverse=False, coset=False, threads=0):
    a = np.array(_np(data), copy=True)
    assert len(a) == 32 << log_n
    lib().orc_ntt(_ptr(a), log_n, int(inverse), int(coset), threads or max_threads())
    return a.tobytes()


def poly_eval(coeffs, points, threads=0):
    """[P(x) for x in points], P given by n x 32 B canonical coefficients; by definition (blocked Horner), no FFT."""
    cf = _np(coeffs)
    pts = b"".join(int(x).to_bytes(32, "little") for x in points)
    P = _np(pts)
    out = np.empty(32 * len(points), dtype=np.uint8)
    lib().orc_poly_eval(_ptr(cf), len(cf) // 32, _ptr(P), len(points), threads or max_threads(), _ptr(out))
    return [int.from_bytes(out[32 * i:32 * i + 32].tobytes(), "little") for i in range(len(points))]


def root_of_unity(log_n):
    out = np.empty(32, dtype=np.uint8)
    lib().orc_root_of_unity(log_n, _ptr(out))
    return int.from_bytes(out.tobytes(), "little")


def mimc_hash(arity, data, threads=0):
    """n x hash_arity of the forge stack's MiMC-7 (n x arity x 32 B LE in, n x 32 B out)."""
    a = _np(data)
    n = len(a) // (32 * arity)
    out = np.empty(max(n, 1) * 32, dtype=np.uint8)
    lib().orc_mimc_hash(arity, _ptr(a), n, threads or max_threads(), _ptr(out))
    return out[:n * 32].tobytes()


def mimc_merkle_roots(leaves, siblings, bits, depth=32, threads=0):
    L, S, B = _np(leaves), _np(siblings), _np(bits)
    n = len(L) // 32
    out = np.empty(max(n, 1) * 32, dtype=np.uint8)
    lib().orc_mimc_merkle_roots(_ptr(L), _ptr(S), _ptr(B), n, depth, threads or max_threads(), _ptr(out))
    return out[:n * 32].tobytes()


def csr_arrays(rows):
    """rows: [[(coeff, var), ...], ...] -> (row_ptr u64, col u32, coeff bytes u8)."""
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


class R1cs:
    def __init__(self, num_instance, num_witness, a=None, b=None, c=None, csr=None):
        """Either rows (a, b, c) or csr = ((rp, col, coeff) x 3) numpy arrays."""
        self.ni, self.nw = num_instance, num_witness
        self.m = csr if csr is not None else tuple(csr_arrays(r) for r in (a, b, c))
        self.nc = len(self.m[0][0]) - 1
        self.log_domain = max(0, (self.nc + self.ni - 1).bit_length())

    def _args(self):
        out = []
        for rp, col, co in self.m:
            out += [_ptr(rp), _ptr(col), _ptr(co)]
        return out


def witness_map(r1cs, z_bytes, threads=0):
    Z = _np(z_bytes)
    out = np.empty(32 << r1cs.log_domain, dtype=np.uint8)
    lg = lib().orc_witness_map(r1cs.nc, r1cs.ni, r1cs.nw, *r1cs._args(), _ptr(Z), _ptr(out), threads or max_threads())
    assert lg == r1cs.log_domain
    return out.tobytes()


class ProvingKey:
    def __init__(self, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query, b_g2_query, h_query,
                 l_query, threads=0):
        arrs = [_np(x) for x in (alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query, b_g2_query,
                                 h_query, l_query)]
        p = [_ptr(a) for a in arrs]
        self.h = lib().orc_pk_new(p[0], p[1], p[2], p[3], p[4], p[5], len(a_query) // 64, p[6], p[7], p[8],
                                  len(h_query) // 64, p[9], len(l_query) // 64, threads or max_threads())

    def free(self):
        if self.h:
            lib().orc_pk_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def prove(pk, r1cs, z_bytes, r_bytes, s_bytes, threads=0):
    Z, Rb, Sb = _np(z_bytes), _np(r_bytes), _np(s_bytes)
    oa, ob, oc = (np.empty(k, dtype=np.uint8) for k in (64, 128, 64))
    rc = lib().orc_prove(pk.h, r1cs.nc, r1cs.ni, r1cs.nw, *r1cs._args(), _ptr(Z), _ptr(Rb), _ptr(Sb), _ptr(oa), _ptr(ob),
                         _ptr(oc), threads or max_threads())
    if rc != 0:
        raise ValueError("orc_prove failed: %d" % rc)
    return oa.tobytes(), ob.tobytes(), oc.tobytes()
