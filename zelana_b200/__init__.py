"""zelana_b200 -- B200-native Groth16/BN254 proving backend for the Zelana prover hot path.

Host-side mirror (Python, ctypes) of the C ABI in include/zkb200.h.  All arithmetic happens in the CUDA
library `libzkb200.so`; there is no CPU fallback -- importing works without a GPU (so that symbols can be
checked), but creating a Context without a CUDA device raises.
"""
import os as _os

# several prove contexts x six streams each: ask for the maximum number of hardware queues before CUDA initialises
# (see zkb_device_count in csrc/zkb200.cu); a value set by the user wins
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._lib import LIB_PATH, ZkbError, load_library  # noqa: F401
from .api import Context, G1Bases, G2Bases, R1csMatrices, ProvingKeyDev  # noqa: F401
from .prover import Groth16Prover, BatchPublicInputs, BatchProof, StdRng, proof_to_solana_bytes  # noqa: F401
from .l2_circuit import (L2BlockCircuit, L2Circuit, L2Prover, TransactionWitness, ShieldedCommitmentWitness,  # noqa: F401
                         WithdrawalWitness)
