"""Tiny interpreter for the carry-chain PTX inside zelana_b200/csrc/fp.cuh.

No GPU in the dev container: this executes the literal asm strings (operand numbering included) of
reduce_once / add / sub / mont_step / final-merge on Python ints and compares with big-int arithmetic,
so operand-index mistakes are caught before a GPU call is spent.  Used by tests/test_ptx_sim.py.
"""
import re

M32 = 0xFFFFFFFF


def extract_asm_blocks(src):
    """Return list of (template, [operand exprs]) for each asm(...) statement, in file order."""
    blocks = []
    i = 0
    while True:
        i = src.find("asm(", i)
        if i < 0:
            break
        depth, j = 0, i + 3
        while True:
            ch = src[j]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    break
            elif ch == '"':
                j = src.find('"', j + 1)
                while src[j - 1] == "\\" and src[j - 2] != "\\":
                    j = src.find('"', j + 1)
            j += 1
        body = src[i + 4:j]
        # split template strings from operand lists at top-level ':'
        parts, cur, k, depth = [], "", 0, 0
        while k < len(body):
            ch = body[k]
            if ch == '"':
                e = k + 1
                while body[e] != '"' or body[e - 1] == "\\":
                    e += 1
                cur += body[k:e + 1]
                k = e + 1
                continue
            if ch == "(":
                depth += 1
            if ch == ")":
                depth -= 1
            if ch == ":" and depth == 0 and body[k:k + 2] != "::" and body[k - 1] != ":":
                parts.append(cur)
                cur = ""
            else:
                cur += ch
            k += 1
        parts.append(cur)
        tmpl = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', parts[0]))
        tmpl = tmpl.replace("\\n", "\n").replace("\\t", " ")
        ops = []
        for lst in parts[1:]:
            for m in re.finditer(r'"([^"]+)"\s*\(((?:[^()]|\([^()]*\))*)\)', lst):
                ops.append((m.group(1), m.group(2).strip()))
        blocks.append((tmpl, ops))
        i = j
    return blocks


def run_asm(tmpl, ops, env):
    """env: dict expr -> int for inputs ('r'/'+r'/'n'); returns dict expr -> int for outputs."""
    regs = {}
    for idx, (cons, expr) in enumerate(ops):
        if cons in ("r", "+r", "n"):
            regs[idx] = env[expr] & M32
        else:
            regs[idx] = None
    cc = 0

    def val(tok):
        tok = tok.strip()
        if tok.startswith("%"):
            v = regs[int(tok[1:])]
            assert v is not None, "read of unwritten output " + tok
            return v
        return int(tok, 0) & M32

    for line in re.split(r"[;\n]", tmpl):
        line = line.strip()
        if not line:
            continue
        op, rest = line.split(None, 1)
        args = [a.strip() for a in rest.split(",")]
        d = int(args[0][1:])
        base = op.split(".")
        name = base[0]
        use_c = name in ("addc", "subc", "madc")
        set_c = ".cc" in op
        if name in ("add", "addc"):
            t = val(args[1]) + val(args[2]) + (cc if use_c else 0)
            regs[d] = t & M32
            if set_c:
                cc = t >> 32
        elif name in ("sub", "subc"):
            t = val(args[1]) - val(args[2]) - (cc if use_c else 0)
            regs[d] = t & M32
            if set_c:
                cc = 1 if t < 0 else 0
        elif name in ("mul",):
            p = val(args[1]) * val(args[2])
            regs[d] = (p >> 32) & M32 if ".hi" in op else p & M32
        elif name in ("mad", "madc"):
            p = val(args[1]) * val(args[2])
            part = (p >> 32) & M32 if ".hi" in op else p & M32
            t = part + val(args[3]) + (cc if use_c else 0)
            regs[d] = t & M32
            if set_c:
                cc = t >> 32
        else:
            raise ValueError("unhandled PTX op " + op)
    return {expr: regs[idx] for idx, (cons, expr) in enumerate(ops) if cons.startswith("=") or cons.startswith("+")}
