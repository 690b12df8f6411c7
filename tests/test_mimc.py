"""MiMC-7 / account Merkle tree: the oracle restatement pinned on the reference's own fixture
forge/circuits/zelana_batch/Prover.toml (a byte-identical copy under tests/golden/reference_fixtures/), and -- with a GPU --
the batched Fr kernels behind zkb_mimc_hash / zkb_mimc_merkle_roots against that oracle."""
import os
import random

import pytest

from conftest import REF_FIXTURES
from oracle import mimc as om

R = om.R


def _le(vals):
    return b"".join(int(v % R).to_bytes(32, "little") for v in vals)


@pytest.fixture(scope="module")
def prover_toml():
    top, tables = om.parse_prover_toml(open(os.path.join(REF_FIXTURES, "zelana_batch_Prover.toml")).read())
    return top, tables


def test_round_constants_match_the_reference_tests():
    # account_tree.rs:461-470 (RC[0] = 2, RC[1] = 10); forge/crates/prover-worker/src/mimc.rs:322-330 (2, 10, 30)
    assert [om.round_constant(i) for i in range(3)] == [2, 10, 30]
    assert om.hash_2(123, 456) != om.hash_2(456, 123)


def test_prover_toml_state_roots_batch_hash_and_withdrawal_root(prover_toml):
    """Replays forge/circuits/zelana_batch/src/main.nr:136-343 on the committed witness for batch 70: every value the circuit
    asserts is reproduced -- pre_state_root from the first leaf and path, the root after each of the ten leaf updates,
    post_state_root, batch_hash, withdrawal_root; the shielded root passes through."""
    top, tables = prover_toml
    transfers = [t for t in tables["transfers"] if t["is_valid"]]
    assert len(transfers) == int(top["num_transfers"]) == 5
    root = int(top["pre_state_root"])
    hashed = []
    for t in transfers:
        g = lambda k: int(t[k])
        sp, si = [int(x) for x in t["sender_path"]], [int(x) for x in t["sender_path_indices"]]
        rp, ri = [int(x) for x in t["receiver_path"]], [int(x) for x in t["receiver_path_indices"]]
        assert len(sp) == len(si) == len(rp) == len(ri) == 32
        s_leaf = om.compute_account_leaf(g("sender_pubkey"), g("sender_balance"), g("sender_nonce"))
        assert om.compute_merkle_root(s_leaf, sp, si) == root                      # verify_merkle_path(sender)
        assert g("sender_balance") >= g("amount") and g("signature") != 0
        s_new = om.compute_account_leaf(g("sender_pubkey"), g("sender_balance") - g("amount"), g("sender_nonce") + 1)
        root = om.update_merkle_root(s_leaf, s_new, sp, si, root)
        r_leaf = om.compute_account_leaf(g("receiver_pubkey"), g("receiver_balance"), g("receiver_nonce"))
        r_new = om.compute_account_leaf(g("receiver_pubkey"), g("receiver_balance") + g("amount"), g("receiver_nonce"))
        root = om.update_merkle_root(r_leaf, r_new, rp, ri, root)
        hashed.append((g("sender_pubkey"), g("receiver_pubkey"), g("amount"), g("sender_nonce")))
    assert root == int(top["post_state_root"])
    assert om.batch_hash(int(top["batch_id"]), hashed, int(top["num_withdrawals"]), int(top["num_shielded"])) == int(top["batch_hash"])
    assert om.withdrawal_root(int(top["batch_id"])) == int(top["withdrawal_root"])
    assert top["pre_shielded_root"] == top["post_shielded_root"]


def test_account_tree_restatement_paths_verify():
    """account_tree.rs tests (test_insert_and_path, test_root_changes_on_update, test_multiple_accounts) on the restatement."""
    t = om.AccountTree()
    assert t.root == om.AccountTree().root and t.root == t.empty_roots[32]
    roots = [t.root]
    for k, bal in ((1, 1000), (2, 2000), (3, 3000), (1, 2500)):
        t.insert(bytes([k] * 32), bal, 0 if bal != 2500 else 1)
        assert t.root not in roots
        roots.append(t.root)
    for k in (1, 2, 3):
        sibs, bits, pos = t.path(bytes([k] * 32))
        leaf = t.leaf(bytes([k] * 32))
        assert pos == int.from_bytes(bytes([k] * 4), "big")
        got = om.compute_merkle_root(om.bytes_to_field_be(leaf), [om.bytes_to_field_be(s) for s in sibs], bits)
        assert om.field_to_bytes_be(got) == t.root


def test_cpp_mimc_matches_python_restatement():
    from oracle import cpu as orc
    rnd = random.Random(21)
    rows = [[rnd.randrange(R) for _ in range(3)] for _ in range(20)]
    assert orc.mimc_hash(3, _le([v for r in rows for v in r])) == _le([om.hash_n(*r) for r in rows])
    leaves = [rnd.randrange(R) for _ in range(5)]
    sibs = [[rnd.randrange(R) for _ in range(32)] for _ in range(5)]
    bits = [[rnd.randrange(2) for _ in range(32)] for _ in range(5)]
    assert orc.mimc_merkle_roots(_le(leaves), _le([v for s in sibs for v in s]), bytes(b for r in bits for b in r)) == \
        _le([om.compute_merkle_root(l, s, b) for l, s, b in zip(leaves, sibs, bits)])


# ------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_gpu_mimc_hash_matches_oracle_and_the_reference_fixture(prover_toml):
    import zelana_b200
    ctx = zelana_b200.Context(0)
    rnd = random.Random(11)
    for arity in (2, 3, 4, 5, 6):
        rows = [[rnd.randrange(R) for _ in range(arity)] for _ in range(257)]
        rows[0] = [0] * arity
        rows[1] = [R - 1] * arity
        got = ctx.mimc_hash(arity, _le([v for r in rows for v in r]))
        want = _le([om.hash_n(*r) for r in rows])
        assert got == want, arity
    assert ctx.mimc_hash(2, b"") == b""
    # the fixture's leaves: hash_4(domain_account, pubkey, balance, nonce)
    top, tables = prover_toml
    t0 = tables["transfers"][0]
    leaf = ctx.mimc_hash(4, _le([1, int(t0["sender_pubkey"]), int(t0["sender_balance"]), int(t0["sender_nonce"])]))
    sibs = [int(x) for x in t0["sender_path"]]
    bits = bytes(int(x) for x in t0["sender_path_indices"])
    root = ctx.mimc_merkle_roots(leaf, _le(sibs), bits, 32)
    assert int.from_bytes(root, "little") == int(top["pre_state_root"])
    with pytest.raises(zelana_b200.ZkbError):
        ctx.mimc_hash(2, (R).to_bytes(32, "little") * 2)           # not canonical
    with pytest.raises(zelana_b200.ZkbError):
        ctx.mimc_hash(7, bytes(32 * 7))
    ctx.close()


@pytest.mark.gpu
def test_gpu_merkle_roots_batch_and_account_tree_batch_insert():
    """n independent (leaf, 32 siblings, 32 index bits) -> n roots == AccountMerklePath::compute_root (account_tree.rs:222-237);
    zelana_b200.AccountTree.insert_batch (level-by-level hashing of the dirty nodes on the GPU) == sequential inserts of the
    restated AccountTree: same root, same nodes, same paths."""
    import zelana_b200
    from zelana_b200.account_tree import AccountTree
    ctx = zelana_b200.Context(0)
    rnd = random.Random(12)
    n, depth = 100, 32
    leaves = [rnd.randrange(R) for _ in range(n)]
    sibs = [[rnd.randrange(R) for _ in range(depth)] for _ in range(n)]
    bits = [[rnd.randrange(2) for _ in range(depth)] for _ in range(n)]
    got = ctx.mimc_merkle_roots(_le(leaves), _le([v for s in sibs for v in s]), bytes(b for r in bits for b in r), depth)
    assert got == _le([om.compute_merkle_root(l, s, b) for l, s, b in zip(leaves, sibs, bits)])
    # tree: 40 accounts, some sharing their first bytes (neighbouring positions), two updated twice
    ids = [bytes([rnd.randrange(256) for _ in range(32)]) for _ in range(36)]
    ids += [ids[0][:3] + bytes([ids[0][3] ^ 1]) + ids[0][4:], ids[1][:3] + bytes([ids[1][3] ^ 1]) + ids[1][4:]]
    updates = [(a, rnd.randrange(1 << 40), rnd.randrange(100)) for a in ids] + [(ids[5], 7, 8), (ids[6], 9, 10)]
    ref = om.AccountTree()
    for a, bal, nonce in updates:
        ref.insert(a, bal, nonce)
    tree = AccountTree(ctx)
    assert tree.root() == om.AccountTree().root
    tree.insert_batch(updates[:20])
    tree.insert_batch(updates[20:])
    assert tree.root() == ref.root
    assert tree.nodes == ref.nodes
    for a in ids[:5]:
        p = tree.path(a)
        assert (p.siblings, p.path_indices, p.position) == ref.path(a)
        assert p.verify(tree.leaf(a), tree.root())
    one = AccountTree(ctx)
    one.insert(ids[0], 5, 6)
    r1 = om.AccountTree()
    r1.insert(ids[0], 5, 6)
    assert one.root() == r1.root
    ctx.close()


@pytest.mark.gpu
def test_gpu_mimc_at_scale_matches_cpp_restatement():
    """2^16 hash_2 (one level of a large batched tree update) and 2^12 depth-32 path roots against oracle/cpu_oracle.cpp."""
    import numpy as np
    import zelana_b200
    from oracle import cpu as orc
    ctx = zelana_b200.Context(0)
    rs = np.random.RandomState(5)

    def rnd(n):
        a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        a[:, 7] %= 0x30644E72
        return a.tobytes()

    pairs = rnd(2 << 16)
    assert ctx.mimc_hash(2, pairs) == orc.mimc_hash(2, pairs)
    n = 1 << 12
    leaves, sibs = rnd(n), rnd(n * 32)
    bits = rs.randint(0, 2, size=n * 32, dtype=np.uint8).tobytes()
    assert ctx.mimc_merkle_roots(leaves, sibs, bits, 32) == orc.mimc_merkle_roots(leaves, sibs, bits, 32)
    ctx.close()
