"""Evaluates the reference's shape functions straight from its source, at authoring time only.

`/root/reference/src/shapes/<kind>.rs` defines `calc_interp` / `calc_deriv` as plain arithmetic on `ksi` (let-bindings,
`interp[i] = expr;`, `deriv.set(m, j, expr);`).  This module parses those statements and evaluates them with Python floats
-- IEEE double arithmetic in the order written, i.e. bit-identical to what the Rust compiler produces -- so that the
reference-element TABLES of kinds the oracle does not restate node by node can still be generated exactly
(tools/gen_shape_tables.py).  Nothing from the reference is copied into the repository: only the numbers it evaluates to.
The reference tree exists in the authoring container only; the generated tables are committed.
"""
import re
from pathlib import Path

REF = Path("/root/reference/src/shapes")


def _body(text, fname):
    i = text.index(f"pub fn {fname}(")
    i = text.index("{", i) + 1
    depth, j = 1, i
    while depth:
        c = text[j]
        depth += c == "{"
        depth -= c == "}"
        j += 1
    body = text[i:j - 1]
    return re.sub(r"//[^\n]*", "", body)


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        if ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a.strip() for a in out]


def _run(body, ksi, nnode, gd, want):
    env = {"ksi": list(ksi)}
    interp = [0.0] * nnode
    deriv = [[0.0] * gd for _ in range(nnode)]

    def ev(expr):
        return eval(" ".join(expr.split()), {"__builtins__": {}}, env)  # noqa: S307 (arithmetic on floats only)

    for stmt in body.split(";"):
        s = " ".join(stmt.split())
        if not s:
            continue
        m = re.fullmatch(r"let \((.*?)\) = \((.*)\)", s)
        if m:
            names, vals = _split_args(m.group(1)), _split_args(m.group(2))
            for n, v in zip(names, vals):
                env[n] = ev(v)
            continue
        m = re.fullmatch(r"let (\w+) = (.*)", s)
        if m:
            env[m.group(1)] = ev(m.group(2))
            continue
        m = re.fullmatch(r"interp\[(\d+)\] = (.*)", s)
        if m:
            interp[int(m.group(1))] = float(ev(m.group(2)))
            continue
        m = re.fullmatch(r"deriv\.set\((.*)\)", s)
        if m:
            a = _split_args(m.group(1))
            deriv[int(a[0])][int(a[1])] = float(ev(a[2]))
            continue
        raise ValueError(f"unexpected statement in reference shape function: {s!r}")
    return interp if want == "interp" else deriv


class RefKind:
    def __init__(self, name):
        self.name = name
        text = (REF / f"{name}.rs").read_text()
        self.nnode = int(re.search(r"pub const NNODE: usize = (\d+);", text).group(1))
        self.gd = int(re.search(r"pub const GEO_NDIM: usize = (\d+);", text).group(1))
        self._interp = _body(text, "calc_interp")
        self._deriv = _body(text, "calc_deriv")
        m = re.search(r"pub const NODE_REFERENCE_COORDS:[^=]*=\s*\[(.*?)\n    \];", text, re.S)
        rows = re.findall(r"\[([^\[\]]*)\]", re.sub(r"//[^\n]*", "", m.group(1)))
        consts = {"__builtins__": {}, "ONE_BY_3": 1.0 / 3.0, "TWO_BY_3": 2.0 / 3.0}  # russell_lab::math constants
        self.ref_coords = [[float(eval(" ".join(v.split()), consts)) for v in _split_args(r)] for r in rows]  # noqa: S307
        assert len(self.ref_coords) == self.nnode and all(len(r) == self.gd for r in self.ref_coords), name

    def interp(self, ksi):
        return _run(self._interp, ksi, self.nnode, self.gd, "interp")

    def deriv(self, ksi):
        return _run(self._deriv, ksi, self.nnode, self.gd, "deriv")


if __name__ == "__main__":
    import numpy as np

    for name in ["lin4", "lin5", "tri10", "tri15", "qua12", "qua16", "qua17", "tet20", "hex32", "hex8", "qua8", "tet10", "hex20"]:
        k = RefKind(name)
        # the reference's own shape tests: N_m(x_n) = delta_mn, sum N = 1, sum dN = 0
        worst = 0.0
        for n, x in enumerate(k.ref_coords):
            nn = np.array(k.interp(x + [0.0] * (3 - k.gd)))
            e = np.zeros(k.nnode)
            e[n] = 1.0
            worst = max(worst, float(np.max(np.abs(nn - e))))
        ksi = [0.13, -0.21, 0.37] if name[:3] in ("qua", "hex", "lin") else [0.2, 0.3, 0.1]
        nn, ll = np.array(k.interp(ksi)), np.array(k.deriv(ksi))
        # derivative by central differences
        fd = np.zeros_like(ll)
        for j in range(k.gd):
            a, b = list(ksi), list(ksi)
            a[j] += 1e-6
            b[j] -= 1e-6
            fd[:, j] = (np.array(k.interp(a)) - np.array(k.interp(b))) / 2e-6
        print(f"{name:6s} nnode {k.nnode:2d} kronecker {worst:.1e} sumN-1 {abs(nn.sum() - 1):.1e} sumL {np.max(np.abs(ll.sum(axis=0))):.1e} "
              f"L-vs-FD {np.max(np.abs(ll - fd)):.1e}")
