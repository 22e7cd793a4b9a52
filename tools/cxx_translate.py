"""C++ cell loops of the reference -> Python, just large enough to EXECUTE the `BoxIterator` loop bodies of AmrHydro.cpp on this path.

Companion of tools/chf_translate.py.  Part of the Picard-body arithmetic of the reference is not Fortran but plain C++ inside
`for (bit.begin(); bit.ok(); ++bit) { IntVect iv = bit(); ... }` loops over one FArrayBox: Calc_meltingRate
(src/AmrHydro.cpp:2175-2252), CalcRHS_gapHeightFAS (:2070-2171), the right-hand side of the head equation (:3044-3077), the
explicit gap-height update (:3394-3408), the moulin quadrature (:1867-2069), VCAMRNonLinearPoissonOp::getFlux
(src/VCAMRNonLinearPoissonOp.cpp:792-841) and HydroIBC::setup_iceMask_EC (src/HydroIBC.cpp:138-184).  This tool reads those loop bodies where they lie under /root/reference, turns each into a
Python function of one cell index -- same statements, same order, IEEE doubles (Python floats: no contraction, no reassociation;
std::pow / std::abs / std::max / std::min map to the same libm calls) -- and runs it over a box.  tests/golden/make_cxx_golden.py
drives it; tests/test_oracle_cxx_golden.py holds the C oracle to the outputs bit for bit.  Nothing in the product imports this file
and no reference source is copied into the repository: the translation happens in memory.

Supported subset: `//` and `/* */` comments; `Real x = e;` declarations; assignments with = += -= *= /= (and the reference's `=-`);
`FAB(iv, c)` element access of the named arrays through any IntVect variable (`IntVect ivlo = iv - shiftiv;`, `ivm1[dir] -= 1;`,
`BASISV(dir)`); `CH_assert` (dropped); if / else if / else with braces; `&&`, `||`, `!`, the alternative token `and`;
`m_suhmoParm->m_x` and `m_suhmoParm->m_ub[0]`; `iv[0]`, `iv[1]`; `for (int m = 0; m < N; m++) { }` with `FAB(iv, m)`; std::abs, std::pow,
std::max, std::min, std::sqrt, std::exp, std::sin, isnan; `pout() << ...;` (dropped) and `MayDay::Error(...)` (raises).
"""
import math
import re


class IntVect(list):
    """Chombo IntVect in 2-D: [x, y], elementwise + and -, components assignable"""

    def __add__(self, o): return IntVect([self[0] + o[0], self[1] + o[1]])
    def __sub__(self, o): return IntVect([self[0] - o[0], self[1] - o[1]])


def BASISV(d):
    return IntVect([1 if d == 0 else 0, 1 if d == 1 else 0])


class Fab:
    """FArrayBox: data a[comp][j][i] whose element (0, 0) is index-space cell `lo` = (x, y)"""

    def __init__(self, a, lo=(0, 0)):
        self.a, self.lo = a, lo

    def idx(self, iv):
        return (iv[1] - self.lo[1], iv[0] - self.lo[0])


def function_text(src, qualified_name, start_line=1):
    """text of the first definition of `qualified_name(` at or after start_line (1-based), through its closing brace"""
    lines = src.split("\n")
    off = sum(len(l) + 1 for l in lines[:start_line - 1])
    m = re.compile(r"^" + re.escape(qualified_name) + r"\s*\(", re.M).search(src, off)
    if not m:
        raise KeyError(qualified_name)
    i = src.index("{", m.end())
    return src[m.start():_match(src, i, "{", "}") + 1]


def _match(s, i, op, cl):
    """index of the bracket closing the one at s[i]"""
    depth = 0
    for k in range(i, len(s)):
        if s[k] == op:
            depth += 1
        elif s[k] == cl:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced " + op)


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def box_loops(text):
    """bodies of every `for (bit.begin(); bit.ok(); ++bit) { ... }` in text, in order (comments already stripped or not)"""
    out = []
    for m in re.finditer(r"for\s*\(\s*bit\.begin\(\)\s*;\s*bit\.ok\(\)\s*;\s*(?:\+\+bit|bit\.next\(\))\s*\)\s*\{", text):
        i = m.end() - 1
        out.append(text[i + 1:_match(text, i, "{", "}")])
    return out


def real_decls(text):
    """{name: value} of the `Real name = <constant expression>;` declarations in text (comments stripped by the caller): the quadrature
    weights and nodes of Calc_moulin_integral, taken from the source rather than retyped"""
    out = {}
    for m in re.finditer(r"\bReal\s+(\w+)\s*=\s*([-+*/. 0-9eE()]+);", text):
        out[m.group(1)] = float(eval(m.group(2), {"__builtins__": {}}))
    return out


def _expr(e, arrays):
    e = e.strip()
    e = re.sub(r"m_suhmoParm\s*->\s*", "P.", e)
    for fn in ("abs", "pow", "max", "min", "sqrt", "exp", "sin", "cos", "tanh"):
        e = re.sub(r"\bstd::" + fn + r"\b", "f_" + fn, e)
    e = re.sub(r"\bisnan\b", "f_isnan", e)
    e = e.replace("&&", " and ").replace("||", " or ")
    e = re.sub(r"!(?!=)", " not ", e)
    names = "|".join(sorted(arrays, key=len, reverse=True))
    e = re.sub(r"\b(" + names + r")\s*\.\s*nComp\s*\(\s*\)", r'len(A["\1"].a)', e)
    e = re.sub(r"\b(" + names + r")\s*\(\s*(\w+)\s*,\s*(\w+)\s*\)", r'A["\1"].a[\3][A["\1"].idx(\2)]', e)
    return e


def _statements(body, arrays, ind, out):
    """translate a brace-free-at-top-level sequence of statements"""
    k, n = 0, len(body)
    while k < n:
        if body[k].isspace():
            k += 1
            continue
        m = re.compile(r"(else\s+if|if|else)\b").match(body, k)
        if m:
            kw = m.group(1)
            k = m.end()
            cond = None
            if kw != "else":
                while body[k].isspace():
                    k += 1
                assert body[k] == "(", body[k:k + 40]
                c1 = _match(body, k, "(", ")")
                cond = body[k + 1:c1]
                k = c1 + 1
            while body[k].isspace():
                k += 1
            assert body[k] == "{", "braces required after if/else: " + body[k:k + 40]
            b1 = _match(body, k, "{", "}")
            head = {"if": "if", "else if": "elif", "else": "else"}[" ".join(kw.split())]
            out.append(" " * ind + (head + (" " + _expr(cond, arrays) if cond is not None else "") + ":"))
            sub = []
            _statements(body[k + 1:b1], arrays, ind + 4, sub)
            out.extend(sub if sub else [" " * (ind + 4) + "pass"])
            k = b1 + 1
            continue
        if body[k] == "}":
            k += 1
            continue
        m = re.compile(r"for\s*\(\s*int\s+(\w+)\s*=\s*0\s*;\s*\1\s*<\s*([^;]+?)\s*;\s*\1\+\+\s*\)\s*\{").match(body, k)
        if m:   # for (int m = 0; m < N; m++) { ... }
            b0 = m.end() - 1
            b1 = _match(body, b0, "{", "}")
            out.append(" " * ind + "for %s in range(%s):" % (m.group(1), _expr(m.group(2), arrays)))
            sub = []
            _statements(body[b0 + 1:b1], arrays, ind + 4, sub)
            out.extend(sub if sub else [" " * (ind + 4) + "pass"])
            k = b1 + 1
            continue
        semi = body.index(";", k)
        st = " ".join(body[k:semi].split())
        k = semi + 1
        if not st or re.match(r"IntVect\s+iv\s*=\s*bit\(\)$", st) or st.startswith("CH_assert"):
            continue
        m = re.match(r"IntVect\s+(\w+)\s*=\s*(.+)$", st)
        if m:   # IntVect ivm1 = bit();  IntVect ivlo = iv - shiftiv;  a copy, as in C++
            rhs = "iv" if m.group(2).replace(" ", "") == "bit()" else _expr(m.group(2), arrays)
            out.append(" " * ind + "%s = IntVect(%s)" % (m.group(1), rhs))
            continue
        if st.startswith("pout()"):
            out.append(" " * ind + "pass")
            continue
        if st.startswith("MayDay::Error"):
            out.append(" " * ind + "raise RuntimeError(" + repr(st) + ")")
            continue
        st = re.sub(r"^(const\s+)?Real\s+", "", st)
        m = re.match(r"(.+?)\s*(\+=|-=|\*=|/=|=-|=)\s*(?!=)(.*)$", st)
        assert m, st
        lhs, op, rhs = m.group(1), m.group(2), m.group(3)
        if op == "=-":   # `x =- a * b` is `x = (-a) * b` in C++ and in Python alike
            op, rhs = "=", "-" + rhs
        out.append(" " * ind + _expr(lhs, arrays) + " " + op + " " + _expr(rhs, arrays))


def compile_cell(body, arrays, name="cell"):
    """Python function cell(iv, A, P, S) of one loop body (iv: IntVect, A: name -> Fab, P: suhmo_params, S: the other names it reads).  arrays: the FArrayBox names the body indexes with (iv, c)"""
    lines = []
    _statements(strip_comments(body), set(arrays), 4, lines)
    src = ("def %s(iv, A, P, S):\n    globals().update(S)\n" % name) + "\n".join(lines) + "\n"
    env = {"f_abs": abs, "f_pow": math.pow, "f_max": max, "f_min": min, "f_sqrt": math.sqrt, "f_exp": math.exp, "f_isnan": math.isnan,
           "f_sin": math.sin, "f_cos": math.cos, "f_tanh": math.tanh, "IntVect": IntVect, "BASISV": BASISV}
    exec(compile(src, "<cxx:%s>" % name, "exec"), env)
    fn = env[name]
    fn.source = src
    return fn


def run_box(fn, A, P, S, nj, ni, lo=(0, 0)):
    """BoxIterator over the box [lo, lo + (ni, nj)), i fastest.  A: name -> Fab, or a bare array [ncomp][nj][ni] that covers exactly
    this box"""
    A = {k: v if isinstance(v, Fab) else Fab(v, lo) for k, v in A.items()}
    for j in range(nj):
        for i in range(ni):
            fn(IntVect([lo[0] + i, lo[1] + j]), A, P, S)
