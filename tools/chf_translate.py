"""ChomboFortran (.ChF) -> Python translator, just large enough to EXECUTE the reference's own 2-D kernels on this path.

Why: the reference cannot be built here (no gfortran, no Chombo), so `oracle/` is a hand restatement.  This tool removes the hand
from the kernel arithmetic: it reads the .ChF sources where they lie under /root/reference, expands the CHF_ macros for
CH_SPACEDIM = 2, turns each subroutine into a Python function that performs the same statements in the same order on IEEE doubles
(Python floats: no FMA contraction, no reassociation) and runs them.  tests/golden/make_chf_golden.py drives it to produce golden
vectors of the reference kernels; tests/test_oracle_chf_golden.py holds the C oracle to them bit for bit.  Nothing in the product
or in the GPU tests imports this file, and no reference source is copied into the repository: the translation happens in memory.

Semantics assumed (stated in DESIGN.md section 3): real literals are double precision (Chombo's DOUBLE build compiles Fortran with
-fdefault-real-8; `1000.0 * 9.8` is 9800 in double), integer division truncates, MOD takes the sign of the dividend, expressions
evaluate left to right by Fortran precedence (no -ffast-math), array layout and bounds come from the CHF_FRA / CHF_BOX arguments.

Supported subset: fixed-form continuation lines, C/c/! comments, cpp #if/#elif/#else/#endif on CH_SPACEDIM, CHF_{FRA,FRA1,CONST_FRA,
CONST_FRA1,FIA,BOX,REAL,CONST_REAL,INT,CONST_INT,CONST_REALVECT} arguments, CHF_{DDECL,AUTODECL,DTERM,IX,AUTOIX,OFFSETIX,ID,AUTOID,
MULTIDO,AUTOMULTIDO,ENDDO,LBOUND,UBOUND,NCOMP}, declarations, assignments, do / enddo, if / else if / else / endif, one-line if,
call MAYDAYERROR, return, end; intrinsics abs mod sqrt max min exp sign dble int.
"""
import keyword
import math
import re

SPACEDIM = 2
CONSTANTS = {"zero": 0.0, "one": 1.0, "two": 2.0, "three": 3.0, "four": 4.0, "five": 5.0, "six": 6.0, "seven": 7.0, "eight": 8.0,
             "nine": 9.0, "ten": 10.0, "twelve": 12.0, "half": 0.5, "third": 1.0 / 3.0, "fourth": 0.25, "sixth": 1.0 / 6.0,
             "eighth": 0.125, "tenth": 0.1}


# ------------------------------------------------------------------------------------------------------------------------
# run-time support
# ------------------------------------------------------------------------------------------------------------------------
class FInt(int):
    """Fortran INTEGER: arithmetic stays integer, division truncates toward zero"""

    def _w(self, other, op):
        if isinstance(other, int):
            return FInt(op(int(self), int(other)))
        return op(float(self), other)

    def __add__(self, o): return self._w(o, lambda a, b: a + b)
    def __radd__(self, o): return self._w(o, lambda a, b: b + a)
    def __sub__(self, o): return self._w(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._w(o, lambda a, b: b - a)
    def __mul__(self, o): return self._w(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._w(o, lambda a, b: b * a)
    def __neg__(self): return FInt(-int(self))
    def __pos__(self): return self

    def __truediv__(self, o):
        if isinstance(o, int):
            q = abs(int(self)) // abs(int(o))
            return FInt(q if (int(self) >= 0) == (int(o) >= 0) else -q)
        return float(self) / o

    def __rtruediv__(self, o):
        if isinstance(o, int):
            return FInt(o).__truediv__(self)
        return o / float(self)

    def __pow__(self, o, m=None):
        if isinstance(o, int):
            return FInt(int(self) ** int(o))
        return float(self) ** o


def f_range(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    i = a
    while (c > 0 and i <= b) or (c < 0 and i >= b):
        yield FInt(i)
        i += c


def f_abs(x): return FInt(abs(int(x))) if isinstance(x, int) else abs(x)


def f_mod(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return FInt(int(math.fmod(int(a), int(b))))
    return math.fmod(a, b)


def f_max(*a): return max(a)
def f_min(*a): return min(a)
def f_sign(a, b): return abs(a) if b >= 0 else -abs(a)
def f_pow(a, b):
    """a ** b; integer exponents by repeated multiplication left to right (what gfortran emits for small constants)"""
    if isinstance(b, int) and not isinstance(a, int):
        r = 1.0
        for _ in range(abs(int(b))):
            r = r * a
        return r if b >= 0 else 1.0 / r
    if isinstance(a, int) and isinstance(b, int):
        return FInt(int(a) ** int(b))
    return a ** b


def chf_id(a, b): return FInt(1 if int(a) == int(b) else 0)


class MayDay(Exception):
    pass


def maydayerror():
    raise MayDay("MAYDAYERROR")


class Fab:
    """a Fortran array argument phi(lo0:hi0, lo1:hi1, 0:ncomp-1) over a numpy array data[ncomp, nj, ni] (or [nj, ni] for FRA1)"""

    def __init__(self, data, lo, one=False):
        self.d, self.lo, self.one = data, (int(lo[0]), int(lo[1])), one
        self.ncomp = 1 if one else data.shape[0]

    def _ix(self, k):
        if self.one:
            i, j = k
            return (int(j) - self.lo[1], int(i) - self.lo[0])
        i, j, n = k
        return (int(n), int(j) - self.lo[1], int(i) - self.lo[0])

    def __getitem__(self, k):
        ix = self._ix(k)
        if min(ix) < 0:
            raise IndexError(f"below the lower bound: {k} (lo {self.lo})")
        v = self.d[ix]
        return FInt(v) if self.d.dtype.kind == "i" else float(v)

    def __setitem__(self, k, v):
        ix = self._ix(k)
        if min(ix) < 0:
            raise IndexError(f"below the lower bound: {k} (lo {self.lo})")
        self.d[ix] = v


class LArr:
    """a local or REALVECT array a(lo:hi)"""

    def __init__(self, lo, hi, init=None, integer=False):
        self.lo = int(lo)
        self.v = [FInt(0) if integer else 0.0] * (int(hi) - int(lo) + 1)
        if init is not None:
            self.v = list(init)

    def __getitem__(self, k):
        k = k[0] if isinstance(k, tuple) else k
        return self.v[int(k) - self.lo]

    def __setitem__(self, k, val):
        k = k[0] if isinstance(k, tuple) else k
        self.v[int(k) - self.lo] = val


class Box:
    def __init__(self, lo, hi):
        self.lo = [FInt(lo[0]), FInt(lo[1])]
        self.hi = [FInt(hi[0]), FInt(hi[1])]


RUNTIME = dict(FInt=FInt, f_range=f_range, f_abs=f_abs, f_mod=f_mod, f_max=f_max, f_min=f_min, f_sign=f_sign, f_pow=f_pow, chf_id=chf_id,
               maydayerror=maydayerror, LArr=LArr, math=math, SPACEDIM=SPACEDIM, **CONSTANTS)


# ------------------------------------------------------------------------------------------------------------------------
# source -> logical lines
# ------------------------------------------------------------------------------------------------------------------------
def cpp(lines):
    """#if / #elif / #else / #endif on CH_SPACEDIM (the only conditionals in these files); other # lines are dropped"""
    out, stack = [], []   # stack of [taken_before, active_now]

    def cond(expr):
        e = expr.replace("CH_SPACEDIM", str(SPACEDIM)).replace("&&", " and ").replace("||", " or ")
        e = re.sub(r"defined\s*\(?\s*\w+\s*\)?", "0", e)
        return bool(eval(e, {}, {}))

    for ln in lines:
        s = ln.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            if d.startswith("ifdef") or d.startswith("ifndef"):
                stack.append([True, d.startswith("ifndef")])
            elif d.startswith("if"):
                c = cond(d[2:])
                stack.append([c, c])
            elif d.startswith("elif"):
                t = stack[-1]
                c = (not t[0]) and cond(d[4:])
                t[1] = c
                t[0] = t[0] or c
            elif d.startswith("else"):
                t = stack[-1]
                t[1] = not t[0]
                t[0] = True
            elif d.startswith("endif"):
                stack.pop()
            continue
        if all(t[1] for t in stack):
            out.append(ln)
    return out


def logical_lines(text):
    """fixed-form Fortran: drop comments, join continuation lines (a non-blank, non-zero character in column 6)"""
    lines = cpp(text.replace("\t", "        ").split("\n"))
    out = []
    for ln in lines:
        if not ln.strip():
            continue
        if ln[0] in "Cc*!":
            continue
        if ln.lstrip().startswith("!"):
            continue
        # trailing ! comment (no string literals with ! in these kernels)
        if "!" in ln:
            ln = ln[:ln.index("!")]
        if len(ln) > 5 and ln[:5].strip() == "" and ln[5] not in " 0":
            out[-1] += " " + ln[6:].strip()
        else:
            out.append(ln.strip())
    return out


# ------------------------------------------------------------------------------------------------------------------------
# macro expansion
# ------------------------------------------------------------------------------------------------------------------------
def _split_top(s, sep=";"):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def _find_macro(s, name):
    """first occurrence of NAME[ ... ] with balanced brackets -> (start, end, inner) or None"""
    m = re.search(r"\b" + name + r"\s*\[", s, re.I)
    if not m:
        return None
    i, depth = m.end(), 1
    while depth:
        depth += {"[": 1, "]": -1}.get(s[i], 0)
        i += 1
    return m.start(), i, s[m.end():i - 1]


def expand_macros(text):
    """CHF_ macros for SPACEDIM = 2 on the text of one subroutine body (logical lines joined by newlines)"""
    def repl(name, fn):
        nonlocal text
        while True:
            f = _find_macro(text, name)
            if not f:
                return
            a, b, inner = f
            text = text[:a] + fn(inner) + text[b:]

    repl("CHF_DTERM", lambda s: "".join(_split_top(s)[:SPACEDIM]))
    repl("CHF_DINVTERM", lambda s: "".join(reversed(_split_top(s)[:SPACEDIM])))
    repl("CHF_DDECL", lambda s: ",".join(p.strip() for p in _split_top(s)[:SPACEDIM]))
    repl("CHF_AUTODECL", lambda s: ",".join(f"{s.strip()}{d}" for d in range(SPACEDIM)))
    repl("CHF_OFFSETIX", lambda s: ",".join(f"{_split_top(s)[0].strip()}{d}{_split_top(s)[1].strip()}{d}" for d in range(SPACEDIM)))
    repl("CHF_AUTOIX", lambda s: ",".join(f"{s.strip()}{d}" for d in range(SPACEDIM)))
    repl("CHF_IX", lambda s: ",".join(p.strip() for p in _split_top(s)[:SPACEDIM]))
    def autoid(s):   # CHF_AUTOID[v; dir; s]: v_d = s * delta(d, dir), s = 1 when omitted
        p = [q.strip() for q in _split_top(s)]
        mul = f"{p[2]}*" if len(p) > 2 and p[2] else ""
        return "\n".join(f"{p[0]}{d} = {mul}chf_id({d},{p[1]})" for d in range(SPACEDIM))

    repl("CHF_AUTOID", autoid)

    def multido(s):
        p = [q.strip() for q in _split_top(s)]
        box, ivs = p[0], p[1:1 + SPACEDIM]
        return "\n".join(f"do {ivs[d]} = {box}__lo({d}), {box}__hi({d})" for d in reversed(range(SPACEDIM)))

    def automultido(s):
        p = [q.strip() for q in _split_top(s)]
        return "\n".join(f"do {p[1]}{d} = {p[0]}__lo({d}), {p[0]}__hi({d})" for d in reversed(range(SPACEDIM)))

    repl("CHF_AUTOMULTIDO", automultido)
    repl("CHF_MULTIDO", multido)
    text = re.sub(r"\bCHF_ENDDO\b", "\n".join(["enddo"] * SPACEDIM), text, flags=re.I)
    repl("CHF_LBOUND", lambda s: f"{_split_top(s)[0].strip()}__lo({_split_top(s)[1].strip()})")
    repl("CHF_UBOUND", lambda s: f"{_split_top(s)[0].strip()}__hi({_split_top(s)[1].strip()})")
    repl("CHF_NCOMP", lambda s: f"{s.strip()}__ncomp")
    text = re.sub(r"\bCHF_ID\s*\(", "chf_id(", text, flags=re.I)
    # D_TERM(a, b, c) of SPACE.H: the first SPACEDIM arguments, concatenated
    while True:
        m = re.search(r"\bD_TERM\s*\(", text, re.I)
        if not m:
            break
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        text = text[:m.start()] + "".join(_split_top(text[m.end():i - 1], ",")[:SPACEDIM]) + text[i:]
    return text


# ------------------------------------------------------------------------------------------------------------------------
# statements -> Python
# ------------------------------------------------------------------------------------------------------------------------
_KW = set(keyword.kwlist) | {"print", "max", "min", "abs", "int", "float", "math", "re", "id"}
_INTRINSIC = {"abs": "f_abs", "mod": "f_mod", "sqrt": "math.sqrt", "max": "f_max", "min": "f_min", "exp": "math.exp", "sign": "f_sign",
              "dble": "float", "int": "FInt", "sin": "math.sin", "cos": "math.cos", "atan": "math.atan", "log": "math.log",
              "chf_id": "chf_id"}


def pyname(n):
    return n + "_v" if n in _KW else n


def convert_expr(e, arrays):
    """a Fortran expression (lower-case) -> Python: operators, literals, array references name(...) -> name[...]"""
    e = re.sub(r"\.ne\.", " != ", e)
    e = re.sub(r"\.eq\.", " == ", e)
    e = re.sub(r"\.le\.", " <= ", e)
    e = re.sub(r"\.lt\.", " < ", e)
    e = re.sub(r"\.ge\.", " >= ", e)
    e = re.sub(r"\.gt\.", " > ", e)
    e = re.sub(r"\.and\.", " and ", e)
    e = re.sub(r"\.or\.", " or ", e)
    e = re.sub(r"\.not\.", " not ", e)
    e = re.sub(r"\.true\.", " True ", e)
    e = re.sub(r"\.false\.", " False ", e)
    e = e.replace("/=", "!=")
    # numeric literals: 1.0d-16 -> 1.0e-16, 1.d0 -> 1.e0; integer literals stay integers
    e = re.sub(r"(?<![\w.])(\d+\.?\d*|\.\d+)[dD]([+-]?\d+)", r"\1e\2", e)
    out, i, n = "", 0, len(e)
    while i < n:
        m = re.match(r"[a-z_]\w*", e[i:])
        if m and (i == 0 or not (e[i - 1].isalnum() or e[i - 1] in "._")):
            name = m.group(0)
            j = i + len(name)
            if name in ("and", "or", "not", "True", "False"):
                out += name
                i = j
                continue
            k = j
            while k < n and e[k] == " ":
                k += 1
            if k < n and e[k] == "(":
                depth, t = 1, k + 1
                while depth:
                    depth += {"(": 1, ")": -1}.get(e[t], 0)
                    t += 1
                inner = convert_expr(e[k + 1:t - 1], arrays)
                if name in arrays:
                    out += f"{pyname(name)}[{inner}]"
                elif name.endswith("__lo") or name.endswith("__hi"):
                    out += f"{name[:-4]}.{name[-2:]}[{inner}]"
                elif name in _INTRINSIC:
                    out += f"{_INTRINSIC[name]}({inner})"
                else:
                    raise ValueError(f"unknown function or array '{name}' in: {e}")
                i = t
                continue
            if name in ("and", "or", "not", "True", "False"):
                out += name
            elif re.fullmatch(r"\d*e[+-]?\d+", name):   # tail of a literal like 1.0e-16 (cannot start an identifier here)
                out += name
            else:
                out += pyname(name)
            i = j
            continue
        out += e[i]
        i += 1
    # ** -> f_pow for the few places it occurs (constants): a**b with simple operands
    out = re.sub(r"([\w.\]\)]+)\s*\*\*\s*([\w.]+)", r"f_pow(\1,\2)", out)
    return out.strip()


class Sub:
    def __init__(self, name, args, body):
        self.name, self.args, self.body = name, args, body   # args: list of (kind, name)


def parse_file(path):
    """-> {NAME: Sub} for every subroutine of a .ChF file"""
    lines = logical_lines(open(path).read())
    subs, i = {}, 0
    while i < len(lines):
        m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)\s*$", lines[i], re.I | re.S)
        if not m:
            i += 1
            continue
        name, arglist = m.group(1).upper(), m.group(2)
        args = []
        for a in _split_top(arglist, ","):
            am = re.match(r"\s*(CHF_\w+)\s*\[\s*(\w+)\s*\]\s*$", a, re.I)
            if not am:
                raise ValueError(f"{name}: argument not understood: {a!r}")
            args.append((am.group(1).upper(), am.group(2).lower()))
        body = []
        i += 1
        while i < len(lines) and not re.match(r"end\s*$", lines[i], re.I):
            body.append(lines[i])
            i += 1
        subs[name] = Sub(name, args, body)
        i += 1
    return subs


def translate(sub):
    """Sub -> Python source of a function `name(**kwargs)`; arguments are passed by keyword under their (lower-case) Fortran names:
    Fab for CHF_FRA*, Box for CHF_BOX, float / int for scalars, a 2-list for CHF_CONST_REALVECT"""
    arrays, pre, types = set(), [], {}
    for kind, a in sub.args:
        p = pyname(a)
        if kind in ("CHF_FRA", "CHF_CONST_FRA", "CHF_FRA1", "CHF_CONST_FRA1", "CHF_FIA", "CHF_CONST_FIA", "CHF_FIA1", "CHF_CONST_FIA1"):
            arrays.add(a)
            pre.append(f"{p} = kw['{a}']")
            pre.append(f"{a}__ncomp = FInt({p}.ncomp)")
        elif kind == "CHF_BOX":
            pre.append(f"{p} = kw['{a}']")
        elif kind in ("CHF_CONST_REALVECT", "CHF_REALVECT"):
            arrays.add(a)
            pre.append(f"{p} = LArr(0, SPACEDIM - 1, [float(x) for x in kw['{a}']])")
        elif kind in ("CHF_CONST_INTVECT", "CHF_INTVECT"):
            arrays.add(a)
            pre.append(f"{p} = LArr(0, SPACEDIM - 1, [FInt(x) for x in kw['{a}']])")
        elif kind in ("CHF_REAL", "CHF_CONST_REAL"):
            types[a] = "real"
            pre.append(f"{p} = float(kw['{a}'])")
        elif kind in ("CHF_INT", "CHF_CONST_INT"):
            types[a] = "int"
            pre.append(f"{p} = FInt(kw['{a}'])")
        else:
            raise ValueError(f"{sub.name}: argument kind {kind} not supported")
    text = expand_macros("\n".join(sub.body)).lower()
    text = text.replace("ch_spacedim", str(SPACEDIM))
    src, ind = [f"def {sub.name.lower()}(**kw):"] + ["    " + p for p in pre], 1

    def emit(s):
        src.append("    " * ind + s)

    def stmt(s):
        """a simple statement: assignment, call, return"""
        if re.match(r"call\s+maydayerror", s):
            return "maydayerror()"
        if s == "return":
            return "return"
        if s == "continue":
            return "pass"
        m = re.match(r"(.+?)(?<![=<>/!])=(?!=)(.+)$", s)
        if m:
            lhs, rhs = m.group(1).strip(), convert_expr(m.group(2).strip(), arrays)
            base = re.match(r"\w+", lhs).group(0)
            t = types.get(base)
            if t == "real":      # assignment converts to the declared type of the left-hand side
                rhs = f"float({rhs})"
            elif t == "int":
                rhs = f"FInt({rhs})"
            return f"{convert_expr(lhs, arrays)} = {rhs}"
        raise ValueError(f"{sub.name}: statement not understood: {s!r}")

    for raw in text.split("\n"):
        s = raw.strip()
        if not s or s.startswith("ch_flops"):   # Chombo's flop counter
            continue
        s = s.rstrip(";").strip()
        dm = re.match(r"(real_t|integer|real\*8|double precision|logical)\s+(.*)$", s)
        if dm:
            isint = dm.group(1) == "integer"
            for item in _split_top(dm.group(2), ","):
                item = item.strip()
                am = re.match(r"(\w+)\s*\(\s*(.+?)\s*\)$", item)
                if am:
                    dims = _split_top(am.group(2), ":")
                    lo, hi = (dims[0], dims[1]) if len(dims) == 2 else ("1", dims[0])
                    arrays.add(am.group(1))
                    types[am.group(1)] = "int" if isint else "real"
                    emit(f"{pyname(am.group(1))} = LArr({convert_expr(lo, arrays)}, {convert_expr(hi, arrays)}, integer={isint})")
                elif re.fullmatch(r"\w+", item):
                    types[item] = "int" if isint else "real"
            continue
        if re.match(r"implicit\s", s):
            continue
        m = re.match(r"do\s+(\w+)\s*=\s*(.+)$", s)
        if m:
            parts = _split_top(m.group(2), ",")
            emit(f"for {pyname(m.group(1))} in f_range({', '.join(convert_expr(p.strip(), arrays) for p in parts)}):")
            ind += 1
            continue
        if re.match(r"end\s*do$", s):
            emit("pass")
            ind -= 1
            continue
        m = re.match(r"(else\s*if|if)\s*\((.*)\)\s*then$", s)
        if m:
            if m.group(1) != "if":
                emit("pass")
                ind -= 1
            emit(f"{'if' if m.group(1) == 'if' else 'elif'} {convert_expr(m.group(2), arrays)}:")
            ind += 1
            continue
        if s == "else":
            emit("pass")
            ind -= 1
            emit("else:")
            ind += 1
            continue
        if re.match(r"end\s*if$", s):
            emit("pass")
            ind -= 1
            continue
        m = re.match(r"if\s*\(", s)
        if m:   # one-line if
            depth, t = 1, m.end()
            while depth:
                depth += {"(": 1, ")": -1}.get(s[t], 0)
                t += 1
            emit(f"if {convert_expr(s[m.end():t - 1], arrays)}:")
            src.append("    " * (ind + 1) + stmt(s[t:].strip()))
            continue
        emit(stmt(s))
    return "\n".join(src) + "\n"


def load(path, names=None):
    """compile the subroutines of a .ChF file -> {NAME: python function}"""
    out = {}
    for name, sub in parse_file(path).items():
        if names is not None and name not in names:
            continue
        ns = dict(RUNTIME)
        exec(compile(translate(sub), f"<{path}:{name}>", "exec"), ns)
        out[name] = ns[name.lower()]
    return out


if __name__ == "__main__":
    import sys
    for nm, sub in parse_file(sys.argv[1]).items():
        if len(sys.argv) < 3 or nm in sys.argv[2:]:
            print(translate(sub))
