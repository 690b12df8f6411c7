"""ORACLE (test infrastructure only) -- MiMC-7 hashing and the depth-32 account Merkle tree of the forge stack, restated.

Follows, function by function:
  forge/circuits/zelana_lib/src/poseidon.nr:15-110   round_constant, mimc_round, mimc_permute, mimc_sponge_absorb, hash_2..hash_6,
                                                     domain separators
  forge/circuits/zelana_lib/src/merkle.nr:29-98      compute_merkle_root, verify_merkle_path, update_merkle_root
  forge/circuits/zelana_lib/src/account.nr:68-71     compute_account_leaf
  core/src/sequencer/storage/account_tree.rs:56-155  the same hash in Rust BigUint (compute_account_leaf, withdrawal root, batch hash)
  core/src/sequencer/storage/account_tree.rs:203-431 AccountMerklePath::compute_root, AccountTree (empty roots, insert_leaf_at, path)
  forge/circuits/zelana_batch/src/main.nr:136-343    the transfer part of the batch circuit (state-root replay, batch hash)
Pinned on the reference's own fixture forge/circuits/zelana_batch/Prover.toml (tests/test_mimc.py): pre_state_root is recomputed
from the first transfer's leaf and path, post_state_root by replaying the five transfers, plus batch_hash and withdrawal_root;
and on the constants the reference's tests assert (account_tree.rs:461-470, forge/crates/prover-worker/src/mimc.rs:322-330).
"""
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MIMC_ROUNDS = 91
TREE_DEPTH = 32
DOMAIN_ACCOUNT, DOMAIN_MERKLE, DOMAIN_NULLIFIER, DOMAIN_BATCH, DOMAIN_WITHDRAWAL, DOMAIN_NOTE = 1, 2, 3, 4, 5, 6


def round_constant(i):
    """poseidon.nr:19-27 / account_tree.rs:48-53: (i+1)^3 + (i+1)"""
    k = i + 1
    return (k * k * k + k) % R


ROUND_CONSTANTS = [round_constant(i) for i in range(MIMC_ROUNDS)]


def mimc_permute(x, k=0):
    """poseidon.nr:38-46: 91 rounds of x -> (x + k + c_i)^7, then + k"""
    for c in ROUND_CONSTANTS:
        t = (x + k + c) % R
        t2 = t * t % R
        t4 = t2 * t2 % R
        x = t4 * t2 % R * t % R
    return (x + k) % R


def mimc_sponge_absorb(inputs, capacity=0):
    """poseidon.nr:50-59"""
    state = capacity
    for v in inputs:
        state = mimc_permute((state + v) % R, 0)
    return state


def hash_n(*vals):
    """hash_2 .. hash_6 (poseidon.nr:62-94): the arity is the domain tag absorbed first"""
    return mimc_sponge_absorb([len(vals)] + [v % R for v in vals])


def hash_2(a, b):
    return hash_n(a, b)


def compute_account_leaf(pubkey, balance, nonce):
    """account.nr:68-71 = account_tree.rs:109-125: hash_4(domain_account, pubkey, balance, nonce)"""
    return hash_n(DOMAIN_ACCOUNT, pubkey, balance, nonce)


def compute_merkle_root(leaf, path, path_indices):
    """merkle.nr:29-52 / account_tree.rs:222-237: index bit 1 = the current node is the RIGHT child"""
    cur = leaf
    for sib, bit in zip(path, path_indices):
        cur = hash_2(sib, cur) if bit == 1 else hash_2(cur, sib)
    return cur


def update_merkle_root(old_leaf, new_leaf, path, path_indices, old_root):
    """merkle.nr:85-98"""
    assert compute_merkle_root(old_leaf, path, path_indices) == old_root, "Old leaf not in tree"
    return compute_merkle_root(new_leaf, path, path_indices)


def withdrawal_root(batch_id, withdrawals=()):
    """main.nr:144,259-260,337 / account_tree.rs:139-155: withdrawals = [(l1_recipient, amount, sender_pubkey), ...]"""
    acc = hash_2(DOMAIN_WITHDRAWAL, batch_id)
    for l1, amount, sender in withdrawals:
        acc = hash_2(acc, hash_n(l1, amount, sender))
    return hash_2(acc, len(withdrawals))


def batch_hash(batch_id, transfers=(), num_withdrawals=0, num_shielded=0):
    """main.nr:141,169-175,215,329-335 (transfers only): transfers = [(sender_pubkey, receiver_pubkey, amount, sender_nonce), ...]"""
    acc = hash_2(DOMAIN_BATCH, batch_id)
    for s, r, amount, nonce in transfers:
        acc = hash_n(acc, hash_n(s, r, amount, nonce), amount)
    return hash_n(acc, len(transfers), num_withdrawals, num_shielded)


def field_to_bytes_be(x):
    return int(x).to_bytes(32, "big")


def bytes_to_field_be(b):
    return int.from_bytes(b, "big") % R


class AccountTree:
    """account_tree.rs:280-431: sparse depth-32 tree, nodes as 32-byte big-endian strings, empty leaf = 32 zero bytes."""

    def __init__(self):
        self.nodes, self.positions = {}, {}
        self.empty_roots = [bytes(32)]
        for _ in range(TREE_DEPTH):
            p = bytes_to_field_be(self.empty_roots[-1])
            self.empty_roots.append(field_to_bytes_be(hash_2(p, p)))
        self.root = self.empty_roots[TREE_DEPTH]

    def position(self, account_id):
        if account_id not in self.positions:
            self.positions[account_id] = int.from_bytes(account_id[:4], "big")
        return self.positions[account_id]

    def insert(self, account_id, balance, nonce):
        pos = self.position(account_id)
        leaf = field_to_bytes_be(compute_account_leaf(bytes_to_field_be(account_id), balance, nonce))
        self.insert_leaf_at(pos, leaf)
        return pos

    def insert_leaf_at(self, position, leaf):
        self.nodes[(0, position)] = leaf
        idx, cur = position, leaf
        for level in range(TREE_DEPTH):
            right = idx & 1 == 1
            sib = self.nodes.get((level, idx - 1 if right else idx + 1), self.empty_roots[level])
            c, s = bytes_to_field_be(cur), bytes_to_field_be(sib)
            cur = field_to_bytes_be(hash_2(s, c) if right else hash_2(c, s))
            idx //= 2
            self.nodes[(level + 1, idx)] = cur
        self.root = cur

    def path(self, account_id):
        pos = self.positions.get(account_id)
        if pos is None:
            return None
        sibs, bits, idx = [], [], pos
        for level in range(TREE_DEPTH):
            right = idx & 1 == 1
            bits.append(1 if right else 0)
            sibs.append(self.nodes.get((level, idx - 1 if right else idx + 1), self.empty_roots[level]))
            idx //= 2
        return sibs, bits, pos

    def leaf(self, account_id):
        pos = self.positions.get(account_id)
        return None if pos is None else self.nodes.get((0, pos))


def parse_prover_toml(text):
    """Just enough TOML for forge/circuits/zelana_batch/Prover.toml: top-level scalars, [[tables]] of scalars and string arrays."""
    top, tables, cur, key, arr = {}, {}, None, None, None
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        if arr is not None:
            if line.startswith("]"):
                (cur if cur is not None else top)[key] = arr
                arr = None
            else:
                arr.append(line.strip(",").strip('"'))
            continue
        if line.startswith("[["):
            name = line.strip("[]")
            cur = {}
            tables.setdefault(name, []).append(cur)
            continue
        k, v = [x.strip() for x in line.split("=", 1)]
        if v == "[":
            key, arr = k, []
            continue
        if v.startswith("[") and v.endswith("]"):
            val = [x.strip().strip('"') for x in v[1:-1].split(",") if x.strip()]
        elif v in ("true", "false"):
            val = v == "true"
        else:
            val = v.strip('"')
        (cur if cur is not None else top)[k] = val
    return top, tables
