"""Executes short statement ranges of the reference's Fortran sources (TEST INFRASTRUCTURE ONLY).

Why: the image has no Fortran compiler, so the parts of the hot path that are plain Fortran loops - the
pairing loops, the occupation rules, the +-G unpack with the kinetic term, the density accumulation,
``kin_energy`` and ``dotp`` - could only be *restated* in the oracle.  This module reads the reference's own
statements from ``/root/reference/src/<file>`` (line ranges given by the caller), translates the small
Fortran subset they use into Python one statement at a time and executes them on NumPy data.  The result is
the reference's text doing the arithmetic, which is what ``tests/test_fsnip_pin.py`` compares the oracle with
and what ``tools/make_golden_fsnip.py`` freezes into ``tests/golden/fsnip/`` for machines without the
reference tree.

Subset: assignments, ``CALL name(args)`` (the caller supplies ``name``), ``DO v=a,b[,s]`` / ``ENDDO``, block and one-line ``IF``, ``ELSE`` / ``ELSEIF``, ``&``
continuations, ``!`` comments (OpenMP / compiler directives are comments), ``#if``/``#ifdef``/``#else``/``#endif``
(no macro is defined: ``#ifdef X`` branches are skipped, their ``#else`` taken), ``__NVTX_*`` lines (skipped), the
intrinsics CMPLX / REAL / AIMAG / CONJG / ABS, ``_real_8`` literals, ``.EQ.`` & co., ``a%b`` components.
Arrays are :class:`FArr` objects (1-based, column-major, ``A(i,j)`` reads an element, ``A(:,j)`` a column);
array assignments go through ``FArr.set``.  Nothing else of Fortran is understood; an unknown statement raises.
Nothing under ``cpmd_b200/`` imports this."""
from __future__ import annotations

import os
import re
import types

import numpy as np

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(REF_SRC)


class FArr:
    """1-based, column-major view of a NumPy array: ``A(i, j)`` like Fortran."""

    def __init__(self, a):
        self.a = np.asarray(a)

    def _idx(self, idx):
        out = []
        for i in idx:
            if isinstance(i, slice):
                out.append(i)
            else:
                i = int(i)
                if i < 1:
                    raise IndexError("Fortran index below 1")
                out.append(i - 1)
        return tuple(out)

    def __call__(self, *idx):
        v = self.a[self._idx(idx)]
        return FArr(v) if isinstance(v, np.ndarray) and v.ndim else v

    def set(self, idx, value):
        self.a[self._idx(idx)] = value

    def __len__(self):
        return self.a.shape[0]


def ns(**kw):
    """A derived-type variable (``parm%tpiba2`` -> ``parm.tpiba2``)."""
    return types.SimpleNamespace(**kw)


def _ddot_seq(n, x, kx, y, ky):
    """``ddot(n, x(kx...), 1, y(ky...), 1)`` with COMPLEX arrays passed by sequence association: the dot product of
    the n REAL words that start at element x(kx...) / y(ky...) in Fortran (column-major) memory order."""
    def words(arr, k):
        a = np.asfortranarray(arr.a)
        k = k if isinstance(k, tuple) else (k,)
        off = int(np.ravel_multi_index(tuple(int(i) - 1 for i in k), a.shape, order="F"))
        flat = a.reshape(-1, order="F")
        return flat.view(np.float64)[2 * off:] if np.iscomplexobj(flat) else flat[off:]
    n = int(n)
    return float(np.dot(words(x, kx)[:n], words(y, ky)[:n]))


_BUILTINS = {
    "CMPLX": lambda re_, im_=0.0, kind=None: complex(re_, im_),
    "REAL": lambda x, kind=None, KIND=None: float(np.real(x)),
    "AIMAG": lambda x: float(np.imag(x)),
    "CONJG": lambda x: np.conj(x),
    "ABS": abs,
    "MOD": lambda a, b: int(a) % int(b),
    "MIN": min,
    "MAX": max,
    "SIZE": lambda a, dim=None: (a.a.size if dim is None else a.a.shape[int(dim) - 1]),
    "ddot_seq": _ddot_seq,
    "real_8": None,
}


def _strip_comment(line):
    out, q = "", None
    for ch in line:
        if q:
            out += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out += ch
        elif ch == "!":
            break
        else:
            out += ch
    return out.rstrip()


def read_statements(fname, lo, hi):
    """Logical statements of lines lo..hi (1-based, inclusive) of a reference source file."""
    with open(os.path.join(REF_SRC, fname)) as fh:
        lines = fh.read().split("\n")[lo - 1:hi]
    stmts, cur, skip = [], "", []
    for raw in lines:
        s = raw.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            if d.startswith("ifdef") or d.startswith("if "):
                skip.append(True)          # no macro is defined
            elif d.startswith("ifndef"):
                skip.append(False)
            elif d.startswith("else"):
                skip[-1] = not skip[-1]
            elif d.startswith("endif"):
                skip.pop()
            continue
        if any(skip):
            continue
        s = _strip_comment(raw).strip()
        if not s or s.startswith("__NVTX"):
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1]
            continue
        stmts.append(cur + s)
        cur = ""
    if cur:
        stmts.append(cur)
    return stmts


_OPS = [(r"\.EQ\.", "=="), (r"\.NE\.", "!="), (r"\.GT\.", ">"), (r"\.GE\.", ">="), (r"\.LT\.", "<"), (r"\.LE\.", "<="),
        (r"\.AND\.", " and "), (r"\.OR\.", " or "), (r"\.NOT\.", " not "), (r"\.TRUE\.", "True"), (r"\.FALSE\.", "False")]


_PYKW = re.compile(r"\b(is|in|as|def|del|from|global|lambda|pass|try|with|yield|class|for|while|import)\b")


def _expr(e):
    e = _PYKW.sub(lambda m: m.group(1) + "_", e)                        # Fortran names that are Python keywords
    e = re.sub(r"(\d+\.?\d*(?:[eEdD][+-]?\d+)?)_real_8", lambda m: m.group(1).replace("d", "e").replace("D", "e"), e)
    e = re.sub(r"(\d)\.(?=[^\d\w]|$)", r"\1.0", e)                      # "1." -> "1.0"
    for pat, rep in _OPS:
        e = re.sub(pat, rep, e, flags=re.I)
    e = e.replace("%", ".")
    # complex literal (re,im): a parenthesised pair of numbers that is not an argument list
    e = re.sub(r"(?<![\w)])\(\s*([-+]?\d[\d.]*(?:[eE][-+]?\d+)?)\s*,\s*([-+]?\d[\d.]*(?:[eE][-+]?\d+)?)\s*\)", r"complex(\1,\2)", e)
    e = re.sub(r"\bkind\s*=\s*real_8", "kind=None", e, flags=re.I)
    # ddot(n, a(k), 1, b(k), 1): sequence association of complex arrays -> ddot_seq(n, a, k, b, k)
    e = re.sub(r"\bddot\(\s*([^,]+),\s*(\w+)\(([^)]+)\)\s*,\s*1\s*,\s*(\w+)\(([^)]+)\)\s*,\s*1\s*\)", r"ddot_seq(\1,\2,(\3,),\4,(\5,))", e)
    # a bare ':' subscript -> slice(None)
    e = re.sub(r"(?<=[(,])\s*:\s*(?=[,)])", "slice(None)", e)
    return e


def _split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def _match_paren(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses: " + s)


def _assign(s):
    """``lhs = rhs`` -> Python; array element targets go through FArr.set."""
    depth = 0
    for i, ch in enumerate(s):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0 and s[i + 1:i + 2] != "=" and s[i - 1:i] not in "<>=!/":
            lhs, rhs = s[:i].strip(), s[i + 1:].strip()
            m = re.match(r"^([\w%]+)\((.*)\)$", lhs)
            if m:
                idx = ", ".join(_expr(x) for x in _split_top(m.group(2)))
                return f"{_expr(m.group(1))}.set(({idx},), {_expr(rhs)})"
            return f"{_expr(lhs)} = {_expr(rhs)}"
    raise ValueError("not an assignment: " + s)


def translate(stmts):
    """Fortran statements -> Python source."""
    py, ind = [], 0
    for s in stmts:
        u = s.upper()
        pad = "    " * ind
        if re.match(r"^DO\s+\w+\s*=", u):
            m = re.match(r"^DO\s+(\w+)\s*=\s*(.*)$", s, flags=re.I)
            parts = [_expr(x) for x in _split_top(m.group(2))]
            step = parts[2] if len(parts) > 2 else "1"
            py.append(f"{pad}for {_expr(m.group(1))} in range(int({parts[0]}), int({parts[1]}) + 1, int({step})):")
            ind += 1
        elif u in ("ENDDO", "END DO"):
            ind -= 1
        elif u.startswith("IF") and re.match(r"^IF\s*\(", u):
            j = _match_paren(s, s.index("("))
            cond, rest = s[s.index("(") + 1:j], s[j + 1:].strip()
            if rest.upper() == "THEN":
                py.append(f"{pad}if {_expr(cond)}:")
                ind += 1
            else:
                py.append(f"{pad}if {_expr(cond)}:")
                py.append(f"{pad}    {_assign(rest)}")
        elif re.match(r"^ELSE\s*IF\s*\(", u):
            j = _match_paren(s, s.index("("))
            py.append(f"{'    ' * (ind - 1)}elif {_expr(s[s.index('(') + 1:j])}:")
        elif u == "ELSE":
            py.append(f"{'    ' * (ind - 1)}else:")
        elif u in ("ENDIF", "END IF"):
            ind -= 1
        elif u == "RETURN":
            py.append(f"{pad}pass")
        elif u.startswith("CALL "):
            py.append(pad + _expr(s[5:].strip()))          # CALL name(args): env supplies the callable
        else:
            py.append(pad + _assign(s))
        if py and py[-1].rstrip().endswith(":") and False:
            pass
    if ind != 0:
        raise ValueError("unbalanced block structure in the statement range")
    # empty bodies (e.g. an IF whose statements were all skipped) need a pass
    out = []
    for i, line in enumerate(py):
        out.append(line)
        if line.rstrip().endswith(":"):
            nxt = py[i + 1] if i + 1 < len(py) else ""
            if len(nxt) - len(nxt.lstrip()) <= len(line) - len(line.lstrip()):
                out.append(" " * (len(line) - len(line.lstrip()) + 4) + "pass")
    return "\n".join(out)


def run(fname, lo, hi, env, tail=()):
    """Execute lines lo..hi of reference file ``fname`` in ``env`` (dict of Fortran names -> Python objects;
    scalars the statements assign end up in it).  ``tail``: Fortran statements appended by the caller (e.g. a
    ``CALL record(...)`` and the ``ENDDO`` that closes a loop whose body is only partly inside the range).
    Returns env."""
    src = translate(read_statements(fname, lo, hi) + list(tail))
    g = dict(_BUILTINS)
    g.update(env)
    exec(compile(src, f"{fname}:{lo}-{hi}", "exec"), g)     # noqa: S102 - the reference's own statements
    for k, v in g.items():
        if k not in _BUILTINS and not k.startswith("__"):
            env[k] = v
    return env
