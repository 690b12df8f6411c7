"""Host mirror of the sequencer's account state tree (core/src/sequencer/storage/account_tree.rs:203-455) with the MiMC-7 node
hashes on the GPU (zkb_mimc_hash): same names, same 32-byte big-endian node encoding, same position rule (the first four bytes
of the account id), same empty-subtree roots.

The reference hashes with BigUint, one node at a time, 32 levels per insert.  Here a BATCH of account updates is applied level
by level: all leaves first (one hash_4 launch), then for each of the 32 levels the parents of the nodes that changed (one
hash_2 launch per level over all dirty pairs).  The tree ends in exactly the state sequential inserts would leave it in --
nodes, root and paths are compared with the restated reference in tests/test_mimc.py.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

from .api import Context

TREE_DEPTH = 32
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
DOMAIN_ACCOUNT = 1


def bytes_to_field(b: bytes) -> int:
    """account_tree.rs:187-192"""
    return int.from_bytes(b, "big") % R


def field_to_bytes(x: int) -> bytes:
    """account_tree.rs:194-201"""
    return int(x).to_bytes(32, "big")


def _le(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


@dataclass
class AccountMerklePath:
    """account_tree.rs:205-260"""
    siblings: List[bytes]
    path_indices: List[int]
    position: int
    _ctx: Optional[Context] = None

    def compute_root(self, leaf: bytes) -> bytes:
        out = self._ctx.mimc_merkle_roots(_le(bytes_to_field(leaf)), b"".join(_le(bytes_to_field(s)) for s in self.siblings),
                                          bytes(self.path_indices), TREE_DEPTH)
        return field_to_bytes(int.from_bytes(out, "little"))

    def verify(self, leaf: bytes, root: bytes) -> bool:
        return self.compute_root(leaf) == root

    def siblings_hex(self) -> List[str]:
        return [s.hex() for s in self.siblings]


class AccountTree:
    """account_tree.rs:270-455"""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.nodes: Dict[Tuple[int, int], bytes] = {}
        self.positions: Dict[bytes, int] = {}
        self.empty_roots = [bytes(32)]                     # compute_empty_roots (:298-311)
        for _ in range(TREE_DEPTH):
            p = _le(bytes_to_field(self.empty_roots[-1]))
            self.empty_roots.append(field_to_bytes(int.from_bytes(ctx.mimc_hash(2, p + p), "little")))
        self._root = self.empty_roots[TREE_DEPTH]

    def root(self) -> bytes:
        return self._root

    def _position(self, account_id: bytes) -> int:
        """get_or_create_position (:319-337)"""
        if account_id not in self.positions:
            self.positions[account_id] = int.from_bytes(account_id[:4], "big")
        return self.positions[account_id]

    def get_position(self, account_id: bytes) -> Optional[int]:
        return self.positions.get(account_id)

    def insert(self, account_id: bytes, balance: int, nonce: int) -> int:
        """insert (:345-358)"""
        return self.insert_batch([(account_id, balance, nonce)])[0]

    def insert_batch(self, updates: List[Tuple[bytes, int, int]]) -> List[int]:
        """Applies the updates in order (a later update of the same account wins, as with sequential inserts)."""
        if not updates:
            return []
        pos = [self._position(a) for a, _, _ in updates]
        leaf_in = b"".join(_le(DOMAIN_ACCOUNT) + _le(bytes_to_field(a)) + _le(bal) + _le(nonce) for a, bal, nonce in updates)
        leaves = self.ctx.mimc_hash(4, leaf_in)            # compute_account_leaf (:109-125)
        for k, p in enumerate(pos):
            self.nodes[(0, p)] = field_to_bytes(int.from_bytes(leaves[32 * k:32 * k + 32], "little"))
        dirty = sorted(set(pos))
        for level in range(TREE_DEPTH):                    # insert_leaf_at (:361-397), all dirty nodes of a level at once
            parents = sorted(set(i >> 1 for i in dirty))
            buf = bytearray()
            for q in parents:
                left = self.nodes.get((level, 2 * q), self.empty_roots[level])
                right = self.nodes.get((level, 2 * q + 1), self.empty_roots[level])
                buf += _le(bytes_to_field(left)) + _le(bytes_to_field(right))
            out = self.ctx.mimc_hash(2, bytes(buf))
            for k, q in enumerate(parents):
                self.nodes[(level + 1, q)] = field_to_bytes(int.from_bytes(out[32 * k:32 * k + 32], "little"))
            dirty = parents
        self._root = self.nodes[(TREE_DEPTH, 0)]
        return pos

    def path(self, account_id: bytes) -> Optional[AccountMerklePath]:
        pos = self.positions.get(account_id)
        return None if pos is None else self.path_at_position(pos)

    def path_at_position(self, position: int) -> AccountMerklePath:
        """path_at_position (:405-431)"""
        sibs, bits, idx = [], [], position
        for level in range(TREE_DEPTH):
            right = idx & 1 == 1
            bits.append(1 if right else 0)
            sibs.append(self.nodes.get((level, idx - 1 if right else idx + 1), self.empty_roots[level]))
            idx //= 2
        return AccountMerklePath(sibs, bits, position, self.ctx)

    def leaf(self, account_id: bytes) -> Optional[bytes]:
        pos = self.positions.get(account_id)
        return None if pos is None else self.nodes.get((0, pos))

    def contains(self, account_id: bytes) -> bool:
        return account_id in self.positions

    def __len__(self) -> int:
        return len(self.positions)

    def is_empty(self) -> bool:
        return not self.positions
