"""ctypes loader for libzkb200.so (built in-tree by __graft_entry__.build())."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libzkb200.so")


class ZkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("zkb200 error %d: %s" % (code, msg))
        self.code = code


STATUS = {0: "OK", -1: "NO_DEVICE", -2: "CUDA", -3: "INVALID_ARG", -4: "OOM", -5: "NOT_CANONICAL", -6: "SHAPE"}


class Csr(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("coeff", C.c_void_p)]


class R1csDesc(C.Structure):
    _fields_ = [("num_constraints", C.c_uint64), ("num_instance", C.c_uint64), ("num_witness", C.c_uint64),
                ("a", Csr), ("b", Csr), ("c", Csr)]


class PkDesc(C.Structure):
    _fields_ = [("alpha_g1", C.c_void_p), ("beta_g1", C.c_void_p), ("beta_g2", C.c_void_p),
                ("delta_g1", C.c_void_p), ("delta_g2", C.c_void_p),
                ("a_query", C.c_void_p), ("a_len", C.c_size_t),
                ("b_g1_query", C.c_void_p), ("b_g1_len", C.c_size_t),
                ("b_g2_query", C.c_void_p), ("b_g2_len", C.c_size_t),
                ("h_query", C.c_void_p), ("h_len", C.c_size_t),
                ("l_query", C.c_void_p), ("l_len", C.c_size_t)]


class SetupParams(C.Structure):
    _fields_ = [("alpha", C.c_uint8 * 32), ("beta", C.c_uint8 * 32), ("gamma", C.c_uint8 * 32), ("delta", C.c_uint8 * 32),
                ("tau", C.c_uint8 * 32), ("g1_generator", C.c_uint8 * 64), ("g2_generator", C.c_uint8 * 128)]


class SetupOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "gamma_g2", "delta_g2", "gamma_abc_g1",
                                          "a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")]


class L2PublicInputs(C.Structure):
    """zkb_l2_public_inputs = BatchPublicInputs (prover.rs:48-63)"""
    _fields_ = [(n, C.c_uint8 * 32) for n in ("pre_state_root", "post_state_root", "pre_shielded_root", "post_shielded_root",
                                              "withdrawal_root", "batch_hash")] + [("batch_id", C.c_uint64)]


class L2Witness(C.Structure):
    _fields_ = [("account_pks", C.c_char_p), ("account_balances", C.POINTER(C.c_uint64)), ("n_accounts", C.c_size_t),
                ("tx_senders", C.c_char_p), ("tx_recipients", C.c_char_p), ("tx_amounts", C.POINTER(C.c_uint64)),
                ("n_txs", C.c_size_t),
                ("commitments", C.c_char_p), ("n_commitments", C.c_size_t),
                ("wd_recipients", C.c_char_p), ("wd_amounts", C.POINTER(C.c_uint64)), ("n_withdrawals", C.c_size_t)]


_P = C.c_void_p
_SZ = C.c_size_t
_I = C.c_int

# every symbol include/zkb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "zkb_version": (C.c_char_p, []),
    "zkb_device_count": (_I, []),
    "zkb_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "zkb_ctx_set_stream": (_I, [_P, _P]),
    "zkb_ctx_synchronize": (_I, [_P]),
    "zkb_ctx_destroy": (None, [_P]),
    "zkb_last_error": (C.c_char_p, [_P]),
    "zkb_host_alloc_pinned": (_I, [_SZ, C.POINTER(_P)]),
    "zkb_host_free_pinned": (None, [_P]),
    "zkb_debug_check_guards": (_I, [_P]),
    "zkb_launch_count": (C.c_ulonglong, [_P]),
    "zkb_ctx_set_msm_window": (_I, [_P, _I]),
    "zkb_ctx_set_graphs": (_I, [_P, _I]),
    "zkb_ctx_set_blocking_sync": (_I, [_P, _I]),
    "zkb_graph_stats": (_I, [_P, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "zkb_prof_phase_count": (_I, []),
    "zkb_prof_phase_name": (C.c_char_p, [_I]),
    "zkb_prof_enable": (_I, [_P, _I]),
    "zkb_prof_reset": (_I, [_P]),
    "zkb_prof_read": (_I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "zkb_bench_int32_peak": (_I, [_P, _I, _I, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "zkb_field_op": (_I, [_P, _I, _I, _P, _P, _SZ, _P]),
    "zkb_scalar_mul": (_I, [_P, _I, _P, _P, _SZ, _P]),
    "zkb_point_sum": (_I, [_P, _I, _P, _SZ, _P]),
    "zkb_g1_bases_load": (_I, [_P, _P, _SZ, _I, C.POINTER(_P)]),
    "zkb_g2_bases_load": (_I, [_P, _P, _SZ, _I, C.POINTER(_P)]),
    "zkb_g1_bases_generate": (_I, [_P, _P, _SZ, C.POINTER(_P)]),
    "zkb_g2_bases_generate": (_I, [_P, _P, _SZ, C.POINTER(_P)]),
    "zkb_g1_bases_window": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "zkb_g2_bases_window": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "zkb_g1_bases_len": (_SZ, [_P]),
    "zkb_g2_bases_len": (_SZ, [_P]),
    "zkb_g1_bases_read": (_I, [_P, _P, _SZ, _SZ, _P]),
    "zkb_g2_bases_read": (_I, [_P, _P, _SZ, _SZ, _P]),
    "zkb_g1_bases_free": (None, [_P]),
    "zkb_g2_bases_free": (None, [_P]),
    "zkb_msm_g1": (_I, [_P, _P, _SZ, _P, _SZ, _P]),
    "zkb_msm_g2": (_I, [_P, _P, _SZ, _P, _SZ, _P]),
    "zkb_msm_g1_dev": (_I, [_P, _P, _SZ, _P, _SZ, _P, _P]),
    "zkb_msm_g2_dev": (_I, [_P, _P, _SZ, _P, _SZ, _P, _P]),
    "zkb_msm_g1_partial": (_I, [_P, _P, _SZ, _P, _SZ, _P]),
    "zkb_msm_g2_partial": (_I, [_P, _P, _SZ, _P, _SZ, _P]),
    "zkb_msm_g1_multi": (_I, [C.POINTER(_P), C.POINTER(_P), _I, _P, _SZ, _P]),
    "zkb_msm_g2_multi": (_I, [C.POINTER(_P), C.POINTER(_P), _I, _P, _SZ, _P]),
    "zkb_msm_g1_combine": (_I, [_P, _P, _I, _P]),
    "zkb_msm_g2_combine": (_I, [_P, _P, _I, _P]),
    "zkb_debug_msm_batch": (_I, [_P, _I, _P, _SZ, _P, _SZ, _SZ, _I, _P]),
    "zkb_debug_msm_comb": (_I, [_P, _I, _P, _SZ, _P, _SZ, _SZ, _I, _I, _P]),
    "zkb_debug_msm_entries": (_I, [_P, _P, _SZ, _P, _SZ, _SZ, _I, _P, _P, _P]),
    "zkb_ntt": (_I, [_P, _P, _P, _I, _I, _I]),
    "zkb_ntt_dev": (_I, [_P, _P, _P, _I, _I, _I]),
    "zkb_mimc_hash": (_I, [_P, _I, _P, _SZ, _P]),
    "zkb_mimc_hash_dev": (_I, [_P, _I, _P, _SZ, _P]),
    "zkb_mimc_merkle_roots": (_I, [_P, _P, _P, _P, _SZ, _I, _P]),
    "zkb_mimc_merkle_roots_dev": (_I, [_P, _P, _P, _P, _SZ, _I, _P]),
    "zkb_r1cs_load": (_I, [_P, C.POINTER(R1csDesc), C.POINTER(_P)]),
    "zkb_r1cs_free": (None, [_P]),
    "zkb_r1cs_log_domain": (_I, [_P]),
    "zkb_r1cs_num_variables": (C.c_uint64, [_P]),
    "zkb_r1cs_num_constraints": (C.c_uint64, [_P]),
    "zkb_witness_map": (_I, [_P, _P, _P, _P]),
    "zkb_pk_load": (_I, [_P, C.POINTER(PkDesc), _I, C.POINTER(_P)]),
    "zkb_pk_load_compressed": (_I, [_P, _P, _SZ, _I, C.POINTER(_P)]),
    "zkb_pk_free": (None, [_P]),
    "zkb_pk_synthetic": (_I, [_P, _SZ, _SZ, _SZ, _P, _SZ, C.POINTER(_P)]),
    "zkb_prove": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "zkb_prove_batch_begin": (_I, [_P, _P, _P, _P, _P, _SZ]),
    "zkb_prove_batch_begin_ex": (_I, [_P, _P, _P, _P, _P, _SZ, C.c_uint]),
    "zkb_prove_batch_end": (_I, [_P, _SZ, _P]),
    "zkb_prove_batch": (_I, [_P, _P, _P, _P, _P, _SZ, _P]),
    "zkb_setup": (_I, [_P, C.POINTER(R1csDesc), C.POINTER(SetupParams), C.POINTER(SetupOut)]),
    "zkb_pk_load_shard": (_I, [_P, C.POINTER(PkDesc), _I, _I, _I, C.POINTER(_P)]),
    "zkb_pk_synthetic_shard": (_I, [_P, _SZ, _SZ, _SZ, _P, _SZ, _I, _I, C.POINTER(_P)]),
    "zkb_prove_partial": (_I, [_P, _P, _P, _P, _P, _P, _P]),
    "zkb_prove_combine": (_I, [_P, _P, _I, _P, _P, _P, _P, _P]),
    "zkb_prove_multi": (_I, [C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _I, _P, _P, _P, _P, _P, _P]),
    "zkb_l2_last_error": (C.c_char_p, []),
    "zkb_l2_circuit_create": (_I, [C.POINTER(L2Witness), C.POINTER(_P)]),
    "zkb_l2_circuit_free": (None, [_P]),
    "zkb_l2_circuit_desc": (_I, [_P, C.POINTER(R1csDesc)]),
    "zkb_l2_circuit_assign": (_I, [_P, C.POINTER(L2PublicInputs), C.POINTER(L2Witness), _P]),
    "zkb_l2_circuit_is_satisfied": (_I, [_P, _P, C.POINTER(_I), C.POINTER(C.c_uint64)]),
    "zkb_l2_roots": (_I, [C.POINTER(L2Witness), C.c_uint64, _P, C.POINTER(L2PublicInputs)]),
    "zkb_l2_poseidon_hash": (_I, [_P, _SZ, _P]),
    "zkb_l2_poseidon_hash_batch": (_I, [_P, _I, _P, _SZ, _P]),
    "zkb_l2_poseidon_hash_batch_dev": (_I, [_P, _I, _P, _SZ, _P]),
    "zkb_l2_poseidon_hash_batch_host": (_I, [_I, _P, _SZ, _I, _P]),
    "zkb_l2_poseidon_params": (_I, [_P, _P]),
    "zkb_l2_prover_randomness": (_I, [C.c_uint64, _P, _P]),
    "zkb_l2_prove": (_I, [_P, _P, _P, _P, C.POINTER(L2PublicInputs), C.POINTER(L2Witness), _P]),
    "zkb_l2_batch_create": (_I, [_I, _I, C.POINTER(_P)]),
    "zkb_l2_batch_destroy": (None, [_P]),
    "zkb_l2_batch_lanes": (_I, [_P]),
    "zkb_l2_batch_prove": (_I, [_P, _P, _P, _P, C.POINTER(L2PublicInputs), C.POINTER(L2Witness), _SZ, _P, C.POINTER(_I)]),
}

_lib = None


def load_library():
    """Load libzkb200.so and bind every declared symbol.  Fails loudly if the CUDA extension is missing:
    there is deliberately no CPU implementation to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(no CPU fallback exists)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
