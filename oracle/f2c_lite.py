"""f2c_lite -- a small fixed-form Fortran-77 -> C translator (TEST INFRASTRUCTURE ONLY).

Purpose: the reference's hot path (cem_maxwell_op_rk and everything below it) is Fortran and
this image has no Fortran compiler, so `oracle/_ref` cannot be built the usual way.  This tool
reads the reference's OWN source files where they lie (`/root/reference/src/*.F`, the COMMON
include files and a case's `SIZE`), translates the selected subroutines statement by statement
into C, and `oracle/build_ref.py` compiles the result into `oracle/_ref/libnekcem_ref.so`.
Nothing of the reference is copied into the repository: the generated C lives only under the
git-ignored `oracle/_ref/`.

The translation is mechanical and keeps the reference's arithmetic exactly: every expression is
emitted with the same operand order and explicit parentheses for the Fortran precedence /
left-to-right association, `real` is `double` (the reference builds with -fdefault-real-8 /
-r8), integer division truncates, `x**n` with integer n uses repeated multiplication, and the
C is compiled with -ffp-contract=off.  All arguments are passed by reference and procedure
names get the gfortran `_` suffix, so the translated code links against the reference's C
library `src/jl` (gs_setup_/gs_op_fields_) unchanged.

Supported subset (what the Maxwell path uses): subroutine / function units, include,
implicit none and implicit typing, real/integer/logical/character declarations with
dimensions (lower bounds, adjustable and assumed size), parameter, common, dimension,
external/save/data (ignored unless used), assignments, call, do/enddo and labelled do, do
while, block/logical if, goto/continue, return, stop; write/print/format are dropped.
`parameter` constants become run-time globals so that one library serves every SIZE: the
values of a SIZE file can be overridden through `ref_set_param` before `ref_alloc`.
"""
import re

DOT_OPS = {"eq": "==", "ne": "!=", "lt": "<", "le": "<=", "gt": ">", "ge": ">=",
           "and": "&&", "or": "||", "not": "!", "eqv": "==", "neqv": "!=",
           "true": "1", "false": "0"}
C_RESERVED_PREFIX = "v_"


class F2CError(Exception):
    pass


# ------------------------------------------------------------------ preprocessing
def preprocess(text, defines):
    """Minimal cpp: #ifdef/#ifndef/#if defined()/#elif/#else/#endif; other directives dropped."""
    out, stack = [], []  # stack of [active_before, taken_any, active_now]

    def active():
        return all(s[2] for s in stack)

    def evalcond(expr):
        expr = re.sub(r"defined\s*\(\s*(\w+)\s*\)", lambda m: "1" if m.group(1) in defines else "0", expr)
        expr = re.sub(r"defined\s+(\w+)", lambda m: "1" if m.group(1) in defines else "0", expr)
        expr = expr.replace("&&", " and ").replace("||", " or ").replace("!", " not ")
        expr = re.sub(r"\b([A-Za-z_]\w*)\b", lambda m: m.group(1) if m.group(1) in ("and", "or", "not")
                      else ("1" if m.group(1) in defines else "0"), expr)
        return bool(eval(expr))

    for line in text.split("\n"):
        s = line.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            if d.startswith("ifdef"):
                stack.append([active(), False, d.split()[1] in defines])
                stack[-1][1] = stack[-1][2]
            elif d.startswith("ifndef"):
                stack.append([active(), False, d.split()[1] not in defines])
                stack[-1][1] = stack[-1][2]
            elif d.startswith("if"):
                v = evalcond(d[2:])
                stack.append([active(), v, v])
            elif d.startswith("elif"):
                if stack[-1][1]:
                    stack[-1][2] = False
                else:
                    v = evalcond(d[4:])
                    stack[-1][1] = stack[-1][2] = v
            elif d.startswith("else"):
                stack[-1][2] = not stack[-1][1]
                stack[-1][1] = True
            elif d.startswith("endif"):
                stack.pop()
            continue
        if active():
            out.append(line)
    return "\n".join(out)


def strip_inline_comment(s):
    q = False
    for i, ch in enumerate(s):
        if ch == "'":
            q = not q
        elif ch == "!" and not q:
            return s[:i]
    return s


def lower_outside_strings(s):
    out, q = [], False
    for ch in s:
        if ch == "'":
            q = not q
        out.append(ch if q else ch.lower())
    return "".join(out)


def logical_lines(text):
    """fixed form: comment lines dropped, continuations joined.  -> [(label, statement)]"""
    stmts = []
    for raw in text.split("\n"):
        line = raw.replace("\t", "      ").rstrip()
        if not line.strip():
            continue
        if line[0] in "cC*!":
            continue
        if line.lstrip().startswith("!"):
            continue
        line = strip_inline_comment(line).rstrip()
        if not line.strip():
            continue
        if len(line) > 5 and line[5] not in " 0" and line[:5].strip() == "":
            if not stmts:
                raise F2CError("continuation without statement: " + raw)
            stmts[-1][1] += line[6:72]
            continue
        label = line[:5].strip()
        stmts.append([label, line[6:72]])
    return [(lab, lower_outside_strings(s).strip()) for lab, s in stmts if s.strip()]


# ------------------------------------------------------------------ tokens and expressions
TOK_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<str>'(?:[^']|'')*')
  | (?P<num>(?:\d+\.?\d*|\.\d+)(?:[ed][+-]?\d+)?)
  | (?P<dot>\.[a-z]+\.)
  | (?P<id>[a-z_$][a-z0-9_$]*)
  | (?P<op>\*\*|==|/=|<=|>=|//|[-+*/(),=<>:])
""", re.X)


def tokenize(s):
    toks, i = [], 0
    while i < len(s):
        m = TOK_RE.match(s, i)
        if not m:
            raise F2CError("cannot tokenize: %r at %r" % (s, s[i:]))
        k = m.lastgroup
        t = m.group(k)
        if k == "num":
            # "1.eq.2" : the dot belongs to the operator; "1.e5" is a number
            mm = re.match(r"(\d+)\.([a-z]+)\.", s[i:])
            if mm and mm.group(2) in DOT_OPS:
                t = mm.group(1)
                toks.append(("num", t))
                i += len(t)
                continue
            # "2.d0" ok; "1.d" not expected
        i = m.end()
        if k == "ws":
            continue
        toks.append((k, t))
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", "")

    def next(self):
        tk = self.peek()
        self.i += 1
        return tk

    def accept(self, kind, val=None):
        k, v = self.peek()
        if k == kind and (val is None or v == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        if not self.accept(kind, val):
            raise F2CError("expected %s %s, got %s in %s" % (kind, val, self.peek(), self.t))

    def done(self):
        return self.i >= len(self.t)

    # precedence: .eqv. < .or. < .and. < .not. < relational < +,- < *,/ < **
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        l = self.p_or()
        while self.peek() in (("dot", ".eqv."), ("dot", ".neqv.")):
            op = self.next()[1]
            l = ("bin", "==" if op == ".eqv." else "!=", l, self.p_or())
        return l

    def p_or(self):
        l = self.p_and()
        while self.peek() == ("dot", ".or."):
            self.next()
            l = ("bin", "||", l, self.p_and())
        return l

    def p_and(self):
        l = self.p_not()
        while self.peek() == ("dot", ".and."):
            self.next()
            l = ("bin", "&&", l, self.p_not())
        return l

    def p_not(self):
        if self.peek() == ("dot", ".not."):
            self.next()
            return ("un", "!", self.p_not())
        return self.p_rel()

    def p_rel(self):
        l = self.p_add()
        k, v = self.peek()
        rel = None
        if k == "dot" and v.strip(".") in ("eq", "ne", "lt", "le", "gt", "ge"):
            rel = DOT_OPS[v.strip(".")]
        elif k == "op" and v in ("==", "/=", "<", "<=", ">", ">="):
            rel = "!=" if v == "/=" else v
        if rel:
            self.next()
            return ("bin", rel, l, self.p_add())
        return l

    def p_add(self):
        k, v = self.peek()
        if k == "op" and v in "+-":
            self.next()
            l = self.p_mul()
            if v == "-":
                l = ("un", "-", l)
        else:
            l = self.p_mul()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("+", "-"):
                self.next()
                l = ("bin", v, l, self.p_mul())
            else:
                return l

    def p_mul(self):
        l = self.p_pow()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("*", "/"):
                self.next()
                l = ("bin", v, l, self.p_pow())
            else:
                return l

    def p_pow(self):
        b = self.p_prim()
        if self.peek() == ("op", "**"):
            self.next()
            # right associative; exponent may carry a sign
            k, v = self.peek()
            if k == "op" and v in "+-":
                self.next()
                e = self.p_pow()
                if v == "-":
                    e = ("un", "-", e)
            else:
                e = self.p_pow()
            return ("pow", b, e)
        return b

    def p_prim(self):
        k, v = self.next()
        if k == "num":
            isint = re.fullmatch(r"\d+", v) is not None
            return ("num", v, isint)
        if k == "str":
            return ("str", v[1:-1].replace("''", "'"))
        if k == "dot" and v in (".true.", ".false."):
            return ("logical", 1 if v == ".true." else 0)
        if k == "op" and v == "(":
            e = self.expr()
            if self.accept("op", ","):          # (re, im): complex constant
                im = self.expr()
                self.expect("op", ")")
                return ("cplx", e, im)
            self.expect("op", ")")
            return ("paren", e)
        if k == "op" and v in "+-":
            e = self.p_prim()
            return ("un", "-", e) if v == "-" else e
        if k == "id":
            if self.accept("op", "("):
                args = []
                if not self.accept("op", ")"):
                    while True:
                        args.append(self.arg())
                        if self.accept("op", ")"):
                            break
                        self.expect("op", ",")
                return ("call", v, args)
            return ("var", v)
        raise F2CError("unexpected token %s %s in %s" % (k, v, self.t))

    def arg(self):
        # expression or lo:hi (dimension declarators); '*' alone = assumed size
        if self.peek() == ("op", "*"):
            nk, nv = self.t[self.i + 1] if self.i + 1 < len(self.t) else ("eof", "")
            if nk == "op" and nv in (",", ")"):
                self.next()
                return ("star",)
        if self.peek() == ("op", ":"):
            self.next()
            return ("range", None, self.expr() if self.peek()[1] not in (",", ")") else None)
        e = self.expr()
        if self.accept("op", ":"):
            if self.peek() == ("op", "*"):
                self.next()
                return ("range", e, ("star",))
            if self.peek()[1] in (",", ")"):
                return ("range", e, None)
            return ("range", e, self.expr())
        return e


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.done():
        raise F2CError("trailing tokens in expression %r" % s)
    return e


def split_top(s, sep=","):
    """split at top-level separators (outside parentheses and strings)"""
    parts, depth, q, cur = [], 0, False, []
    for ch in s:
        if ch == "'":
            q = not q
        if not q:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == sep and depth == 0:
                parts.append("".join(cur))
                cur = []
                continue
        cur.append(ch)
    parts.append("".join(cur))
    return [p.strip() for p in parts]


def match_paren(s, i):
    """s[i] == '(' -> index of the matching ')'"""
    depth, q = 0, False
    for j in range(i, len(s)):
        ch = s[j]
        if ch == "'":
            q = not q
        if q:
            continue
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return j
    raise F2CError("unbalanced parentheses: " + s)


def top_level_eq(s):
    """index of the assignment '=' at depth 0 (not ==, /=, <=, >=), or -1"""
    depth, q = 0, False
    for j, ch in enumerate(s):
        if ch == "'":
            q = not q
        if q:
            continue
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if s[j + 1:j + 2] == "=" or s[j - 1:j] in ("=", "/", "<", ">"):
                continue
            return j
    return -1


# ------------------------------------------------------------------ symbols
class Sym:
    def __init__(self, name):
        self.name = name
        self.ftype = None      # 'real' | 'integer' | 'logical' | 'character'
        self.dims = None       # list of (lo_ast, hi_ast|('star',)|None)
        self.kind = "local"    # local | dummy | common | param | func
        self.block = None
        self.charlen = 1
        self.value = None      # parameter expression (ast)
        self.external = False
        self.static = False    # SAVE / DATA
        self.init = None
        self.data_unsupported = False
        self.init_list = None
        self.kind8 = False     # integer*8
        self.gname = name      # name of the C global (COMMON variables)

    def ctype(self):
        if self.ftype == "integer" and self.kind8:
            return "long long"
        return {"real": "double", "integer": "int", "logical": "int", "character": "char",
                "complex": "double _Complex"}[self.ftype]


INTRINSICS = {"sqrt", "abs", "sin", "cos", "tan", "exp", "log", "log10", "atan", "atan2", "asin",
              "acos", "sinh", "cosh", "tanh", "max", "min", "mod", "real", "dble", "int", "nint",
              "float", "sign", "iand", "ior", "dsqrt", "dabs", "dsin", "dcos", "dexp", "dlog",
              "datan", "datan2", "dmax1", "dmin1", "amax1", "amin1", "max0", "min0", "iabs",
              "modulo", "ishft", "ifix", "dfloat", "sngl", "aint", "anint", "floor", "ceiling",
              "cexp", "csqrt", "clog", "cmplx", "dcmplx", "conjg", "aimag", "dimag", "imag",
              "dreal", "cabs"}

TYPE_RE = re.compile(r"^(real|integer|logical|character|double\s*precision|complex)\s*(\*\s*(\d+|\(\s*\*\s*\)))?\s*(.*)$")


class Unit:
    def __init__(self, kind, name, args, ftype=None):
        self.kind, self.name, self.args, self.ftype = kind, name, args, ftype
        self.syms = {}
        self.implicit_none = False
        self.body = []         # (label, stmt) executable statements
        self.decl_order = []   # parameter names in order of definition
        self.common_blocks = {}
        self.suffix = ""

    def sym(self, name):
        if name not in self.syms:
            self.syms[name] = Sym(name)
        return self.syms[name]


def implicit_type(name):
    return "integer" if name[0] in "ijklmn" else "real"


class Translator:
    def __init__(self, include_dirs, defines=()):
        self.include_dirs = include_dirs
        self.defines = set(defines)
        self.units = {}        # name -> Unit
        self.order = []
        self.globals = {}      # common variable name -> (Sym, block)
        self.params = []       # ordered (name, ast) of every parameter seen (first definition wins)
        self.param_names = set()
        self.externals = {}    # called procedure name -> return ftype or None
        self._inc_cache = {}

    # ---- reading
    def read_include(self, name):
        if name in self._inc_cache:
            return self._inc_cache[name]
        import os
        for d in self.include_dirs:
            p = os.path.join(d, name)
            if os.path.exists(p):
                txt = preprocess(open(p, errors="replace").read(), self.defines)
                ll = self.expand_includes(logical_lines(txt))
                self._inc_cache[name] = ll
                return ll
        raise F2CError("include file not found: " + name)

    def expand_includes(self, lines):
        out = []
        for lab, s in lines:
            m = re.match(r"^include\s*'([^']+)'", s)
            if m:
                out.extend([("@" + l.lstrip("@"), t) for l, t in self.read_include(m.group(1))])
            else:
                out.append((lab, s))
        return out

    def load(self, path, wanted=None, suffix=""):
        """parse a source file; keep the units named in `wanted` (all if None).  With a
        suffix the kept units are emitted as <name><suffix>_ (several .usr files define the
        same callback names) and calls between them follow the renaming."""
        txt = preprocess(open(path, errors="replace").read(), self.defines)
        lines = logical_lines(txt)
        cur = None
        found = []
        for lab, s in lines:
            m = re.match(r"^(?:(real|integer|logical|double\s*precision)\s*(?:\*\s*\d+)?\s+)?(subroutine|function|program)\s+([a-z_0-9]+)\s*(\((.*)\))?\s*$", s)
            if m and top_level_eq(s) < 0:
                ftype = m.group(1)
                if ftype and ftype.startswith("double"):
                    ftype = "real"
                args = [a.strip() for a in (m.group(5) or "").split(",") if a.strip()]
                cur = Unit(m.group(2), m.group(3), args, ftype)
                cur.raw = []
                continue
            if cur is None:
                continue
            if s == "end" or re.match(r"^end\s*(subroutine|function|program)(\s+\w+)?$", s):
                if wanted is None or cur.name in wanted:
                    cur.fname = cur.name          # Fortran name (function result variable)
                    cur.rename = {}
                    cur.suffix = suffix
                    key = cur.name + suffix
                    self.units[key] = cur
                    self.order.append(key)
                    found.append(key)
                cur = None
                continue
            cur.raw.append((lab, s))
        if suffix:
            ren = {self.units[k].fname: k for k in found}
            for k in found:
                self.units[k].rename = ren
        for name in found:
            u = self.units[name]
            u.raw = self.expand_includes(u.raw)
            self.declare(u)
            u.name = name if not suffix else u.name
            u.cname = name
        return found

    # ---- declarations
    def parse_declarator(self, u, text, ftype=None, charlen=None, kind=None, block=None):
        text = text.strip()
        m = re.match(r"^([a-z_$][a-z0-9_$]*)\s*(\(.*\))?\s*(\*\s*(\d+|\(\s*\*\s*\)))?$", text)
        if not m:
            raise F2CError("bad declarator %r in %s" % (text, u.name))
        name = m.group(1)
        sy = u.sym(name)
        if m.group(2):
            p = Parser(tokenize(m.group(2)))
            p.expect("op", "(")
            dims = []
            while True:
                a = p.arg()
                if a[0] == "range":
                    dims.append((a[1] if a[1] is not None else ("num", "1", True), a[2]))
                elif a[0] == "star":
                    dims.append((("num", "1", True), ("star",)))
                else:
                    dims.append((("num", "1", True), a))
                if p.accept("op", ")"):
                    break
                p.expect("op", ",")
            sy.dims = dims
        if ftype:
            sy.ftype = ftype
            if ftype == "integer" and str(charlen) == "8":
                sy.kind8 = True
            if ftype == "character":
                cl = m.group(4) if m.group(3) else charlen
                sy.charlen = int(cl) if cl and str(cl).isdigit() else 1
        if kind and sy.kind in ("local",):
            sy.kind = kind
        if block:
            sy.block = block
        return sy

    def parse_data(self, u, s):
        """DATA a,b /1,2/ [, c /3/] for scalars: static storage with an initial value"""
        rest = s[len("data"):].strip()
        for m in re.finditer(r"([^/]+)/([^/]*)/\s*,?", rest):
            names = split_top(m.group(1))
            vals = []
            for v in split_top(m.group(2)):
                mm = re.match(r"^(\d+)\s*\*\s*(.*)$", v)
                if mm:
                    vals.extend([mm.group(2)] * int(mm.group(1)))
                else:
                    vals.append(v)
            names = [n for n in names if n]
            if len(names) == 1 and "(" not in names[0] and len(vals) > 1:
                sy = u.sym(names[0])     # whole array: DATA a /v1,v2,.../
                sy.static = True
                sy.init_list = [parse_expr(v) for v in vals]
                continue
            if len(names) != len(vals) or any("(" in n for n in names):
                for n in names:
                    u.sym(re.match(r"[a-z_0-9$]+", n).group(0)).data_unsupported = True
                continue
            for n, v in zip(names, vals):
                sy = u.sym(n)
                sy.static = True
                sy.init = parse_expr(v)

    def declare(self, u):
        for a in u.args:
            u.sym(a).kind = "dummy"
        body_started = False
        for lab, s in u.raw:
            from_inc = lab.startswith("@")
            lab = lab.lstrip("@")
            if not body_started:
                if s.startswith("implicit"):
                    if "none" in s:
                        u.implicit_none = True
                    continue
                m = TYPE_RE.match(s)
                if m and top_level_eq(s) < 0 and not re.match(r"^(real|int)\s*\(", s):
                    ftype = m.group(1)
                    if ftype.startswith("double"):
                        ftype = "real"
                    rest = m.group(4)
                    charlen = m.group(3)
                    if rest.startswith("function"):
                        continue
                    for d in split_top(rest):
                        if d:
                            self.parse_declarator(u, d, ftype, charlen)
                    continue
                if s.startswith("dimension"):
                    for d in split_top(s[len("dimension"):]):
                        self.parse_declarator(u, d)
                    continue
                if s.startswith("parameter"):
                    inner = s[s.index("(") + 1:match_paren(s, s.index("("))]
                    for d in split_top(inner):
                        j = top_level_eq(d)
                        name, val = d[:j].strip(), parse_expr(d[j + 1:])
                        sy = u.sym(name)
                        sy.value = val
                        if not from_inc:
                            sy.kind = "lparam"  # a constant of this unit only
                            u.decl_order.append(name)
                            continue
                        sy.kind = "param"
                        if name not in self.param_names:
                            self.param_names.add(name)
                            self.params.append((name, val, u))
                    continue
                if s.startswith("common"):
                    rest = s[len("common"):].strip()
                    # common /a/ x,y /b/ z   (blank common not supported)
                    pieces = re.split(r"/\s*([a-z_0-9]*)\s*/", rest)
                    # pieces: ['', blk1, vars1, blk2, vars2...]
                    for bi in range(1, len(pieces), 2):
                        blk = pieces[bi]
                        names = []
                        for d in split_top(pieces[bi + 1]):
                            if d:
                                sy = self.parse_declarator(u, d, kind="common", block=blk)
                                sy.kind, sy.block = "common", blk
                                if not from_inc and u.suffix:
                                    # a COMMON of this .usr file: other .usr files reuse the
                                    # block names with different layouts
                                    sy.gname = sy.name + u.suffix
                                names.append(sy.name)
                        u.common_blocks.setdefault(blk, []).extend(names)
                    continue
                if re.match(r"^(external|save|intrinsic|equivalence|data)\b", s) and top_level_eq(s) < 0 or s.startswith("data "):
                    if s.startswith("external"):
                        for d in split_top(s[len("external"):]):
                            u.sym(d).external = True
                    if s.startswith("save"):
                        for d in split_top(s[len("save"):]):
                            if d and not d.startswith("/"):
                                u.sym(d).static = True
                    if s.startswith("data"):
                        self.parse_data(u, s)
                    continue
                if s.startswith("structure") or s.startswith("record"):
                    raise F2CError("unsupported declaration in %s: %s" % (u.name, s))
                body_started = True
            u.body.append((lab, s))
        # default types
        for sy in u.syms.values():
            if sy.ftype is None:
                sy.ftype = implicit_type(sy.name)
        if u.kind == "function" and u.ftype is None:
            r = u.syms.get(u.name)
            u.ftype = r.ftype if r else implicit_type(u.name)
        if u.kind == "function" and u.name not in u.syms:
            # `real function f()` under implicit none: the header types the result variable
            u.sym(u.name).ftype = u.ftype


# ------------------------------------------------------------------ emission
class Emitter:
    def __init__(self, tr):
        self.tr = tr
        self.used_globals = {}   # name -> Sym (from the first unit that used it)
        self.used_params = set()
        self.called = {}         # procedure -> ftype|None

    # ---- names
    def vname(self, n):
        return C_RESERVED_PREFIX + n.replace("$", "_")

    # ---- types
    def typ(self, u, e):
        k = e[0]
        if k == "num":
            return "int" if e[2] else "double"
        if k == "logical":
            return "int"
        if k == "str":
            return "char"
        if k == "paren":
            return self.typ(u, e[1])
        if k == "cplx":
            return "double _Complex"
        if k == "un":
            return "int" if e[1] == "!" else self.typ(u, e[2])
        if k == "pow":
            if "double _Complex" in (self.typ(u, e[1]), self.typ(u, e[2])):
                return "double _Complex"
            return self.typ(u, e[1]) if self.typ(u, e[2]) == "int" else "double"
        if k == "bin":
            if e[1] in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return "int"
            a, b = self.typ(u, e[2]), self.typ(u, e[3])
            if "double _Complex" in (a, b):
                return "double _Complex"
            return "double" if "double" in (a, b) else ("long long" if "long long" in (a, b) else "int")
        if k == "var":
            sy = self.lookup(u, e[1])
            return sy.ctype()
        if k == "call":
            name = e[1]
            sy = u.syms.get(name)
            if sy is not None and sy.dims is not None:
                return sy.ctype()
            if name in INTRINSICS and not (sy and sy.external):
                if name in ("int", "nint", "iand", "ior", "max0", "min0", "iabs", "ishft", "ifix",
                            "floor", "ceiling"):
                    return "int"
                ts = [self.typ(u, a) for a in e[2]]
                if name in ("cexp", "csqrt", "cmplx", "dcmplx", "conjg", "clog", "ccos", "csin"):
                    return "double _Complex"
                if name in ("exp", "sqrt", "log", "sin", "cos") and "double _Complex" in ts:
                    return "double _Complex"
                if name in ("max", "min", "mod", "abs", "sign", "modulo"):
                    if "double _Complex" in ts:
                        return "double"
                    return "double" if "double" in ts else "int"
                return "double"
            ft = (sy.ftype if sy and sy.ftype else implicit_type(name))
            return {"real": "double", "integer": "int", "logical": "int",
                    "complex": "double _Complex"}.get(ft, "double")
        raise F2CError("typ: " + str(e))

    def lookup(self, u, name):
        if name not in u.syms:
            if u.implicit_none:
                raise F2CError("undeclared %s in %s" % (name, u.name))
            sy = u.sym(name)
            sy.ftype = implicit_type(name)
        sy = u.syms[name]
        if sy.ftype is None:
            sy.ftype = implicit_type(name)
        return sy

    # ---- expressions
    def ref(self, u, name):
        """C lvalue of a scalar variable"""
        sy = self.lookup(u, name)
        if sy.kind == "dummy":
            return "(*%s)" % self.vname(name)
        if sy.kind == "common":
            self.note_global(u, sy)
            return self.vname(sy.gname)
        if sy.kind == "param":
            self.used_params.add(name)
        if u.kind == "function" and name == u.name:
            return "f_result"
        return self.vname(name)

    def note_global(self, u, sy):
        g = self.used_globals.get(sy.gname)
        if g is None:
            self.used_globals[sy.gname] = (sy, u)
            # dims may reference parameters
            if sy.dims:
                for lo, hi in sy.dims:
                    for x in (lo, hi):
                        if x is not None and x[0] != "star":
                            self.scan_params(u, x)
        else:
            g0 = g[0]
            if g0.block != sy.block or (g0.dims is None) != (sy.dims is None) or g0.ftype != sy.ftype:
                raise F2CError("common variable %s declared inconsistently (%s/%s vs %s/%s)" %
                               (sy.name, g0.block, g[1].name, sy.block, u.name))

    def scan_params(self, u, e):
        if e is None:
            return
        if e[0] == "var":
            sy = u.syms.get(e[1])
            if sy is not None and sy.kind == "param":
                if e[1] not in self.used_params:
                    self.used_params.add(e[1])
                    self.scan_params(u, sy.value)
        elif e[0] in ("bin",):
            self.scan_params(u, e[2]); self.scan_params(u, e[3])
        elif e[0] in ("un",):
            self.scan_params(u, e[2])
        elif e[0] == "paren":
            self.scan_params(u, e[1])
        elif e[0] == "pow":
            self.scan_params(u, e[1]); self.scan_params(u, e[2])
        elif e[0] == "call":
            for a in e[2]:
                self.scan_params(u, a)

    def scan_all_params(self, u, e):
        """add every parameter named in expression e"""
        if e is None or not isinstance(e, tuple):
            return
        if e[0] == "var":
            sy = u.syms.get(e[1])
            if sy is not None and sy.kind == "param":
                self.used_params.add(e[1])
            return
        for x in e[1:]:
            if isinstance(x, tuple):
                self.scan_all_params(u, x)
            elif isinstance(x, list):
                for y in x:
                    self.scan_all_params(u, y)

    def index(self, u, sy, args):
        """flat C index of array element sy(args): column-major with the declared bounds"""
        if len(args) != len(sy.dims):
            raise F2CError("rank mismatch for %s in %s" % (sy.name, u.name))
        terms = None
        # ((a3-lo3)*e2 + (a2-lo2))*e1 + (a1-lo1)
        for d in range(len(args) - 1, -1, -1):
            lo, hi = sy.dims[d]
            off = "((%s)-(%s))" % (self.ex(u, args[d]), self.ex(u, lo))
            if terms is None:
                terms = off
            else:
                ext = self.extent(u, sy, d)
                terms = "(%s*(long)(%s)+%s)" % (terms, ext, off)
        return terms

    def extent(self, u, sy, d):
        lo, hi = sy.dims[d]
        if hi is None or hi[0] == "star":
            raise F2CError("extent of assumed-size dimension used: %s in %s" % (sy.name, u.name))
        if lo[0] == "num" and lo[1] == "1":
            return self.ex(u, hi)
        return "((%s)-(%s)+1)" % (self.ex(u, hi), self.ex(u, lo))

    def base(self, u, sy):
        if sy.kind == "common":
            self.note_global(u, sy)
            return self.vname(sy.gname)
        return self.vname(sy.name)

    def ex(self, u, e):
        k = e[0]
        if k == "num":
            v = e[1]
            if e[2]:
                return v
            v = v.replace("d", "e")
            if "." not in v and "e" not in v:
                v += ".0"
            if v.endswith("."):
                v += "0"
            if v.startswith("."):
                v = "0" + v
            v = v.replace(".e", ".0e")
            return v
        if k == "logical":
            return str(e[1])
        if k == "str":
            return '"%s"' % e[1].replace("\\", "\\\\").replace('"', '\\"')
        if k == "paren":
            return "(%s)" % self.ex(u, e[1])
        if k == "cplx":
            return "CMPLX(%s,%s)" % (self.ex(u, e[1]), self.ex(u, e[2]))
        if k == "un":
            return "(%s%s)" % (e[1], self.ex(u, e[2]))
        if k == "pow":
            b, x = e[1], e[2]
            if "double _Complex" in (self.typ(u, b), self.typ(u, x)):
                return "cpow(%s,%s)" % (self.ex(u, b), self.ex(u, x))
            if self.typ(u, x) == "int":
                if self.typ(u, b) == "int":
                    return "f_ipow(%s,%s)" % (self.ex(u, b), self.ex(u, x))
                return "f_powi(%s,%s)" % (self.ex(u, b), self.ex(u, x))
            return "pow(%s,%s)" % (self.ex(u, b), self.ex(u, x))
        if k == "bin":
            if e[1] in ("==", "!=") and "char" in (self.typ(u, e[2]), self.typ(u, e[3])):
                (pa, la), (pb, lb) = self.chref(u, e[2]), self.chref(u, e[3])
                return "(f_chcmp(%s,%d,%s,%d)%s0)" % (pa, la, pb, lb, e[1])
            la, lb = self.ex(u, e[2]), self.ex(u, e[3])
            ta, tb = self.typ(u, e[2]), self.typ(u, e[3])
            if e[1] in ("+", "-", "*", "/") and (ta == "double _Complex") != (tb == "double _Complex"):
                # Fortran converts the real operand to COMPLEX (x, 0.0) first and then applies the
                # complex operation; C's mixed real/complex arithmetic skips the imaginary part of
                # the real operand, which changes signed zeros (1.0 - (4.0,+0.0) is (-3.0,+0.0) in
                # Fortran but (-3.0,-0.0) in C: the other side of csqrt's branch cut)
                if ta != "double _Complex":
                    la = "CMPLX((double)(%s),0.0)" % la
                else:
                    lb = "CMPLX((double)(%s),0.0)" % lb
            return "(%s%s%s)" % (la, e[1], lb)
        if k == "var":
            sy = self.lookup(u, e[1])
            if sy.dims is not None:
                raise F2CError("whole-array use of %s in expression (%s)" % (e[1], u.name))
            return self.ref(u, e[1])
        if k == "call":
            name, args = e[1], e[2]
            sy = u.syms.get(name)
            if sy is not None and sy.dims is not None:
                return "%s[%s]" % (self.base(u, sy), self.index(u, sy, args))
            if name in INTRINSICS and not (sy and sy.external):
                return self.intrinsic(u, name, args)
            # external function
            ft = (sy.ftype if sy and sy.ftype else implicit_type(name))
            name = u.rename.get(name, name)
            self.called[name] = ft
            return "%s_(%s)" % (name, ", ".join(self.argref(u, a) for a in args))
        raise F2CError("ex: " + str(e))

    def chref(self, u, e):
        """character operand -> (C pointer expression, length) ; literals have length -1"""
        if e[0] == "paren":
            return self.chref(u, e[1])
        if e[0] == "str":
            return self.ex(u, e), -1
        if e[0] == "var":
            sy = self.lookup(u, e[1])
            if sy.ftype != "character" or sy.dims is not None:
                raise F2CError("character operand expected: %s in %s" % (e[1], u.name))
            if sy.kind == "common":
                self.note_global(u, sy)
                return self.vname(sy.gname), sy.charlen
            return self.vname(e[1]), sy.charlen
        if e[0] == "call":
            sy = u.syms.get(e[1])
            if sy is None or sy.dims is None or sy.ftype != "character":
                raise F2CError("character operand expected: %s in %s" % (e[1], u.name))
            return "(%s+(long)%d*%s)" % (self.base(u, sy), sy.charlen, self.index(u, sy, e[2])), sy.charlen
        raise F2CError("unsupported character expression in %s: %s" % (u.name, e))

    def intrinsic(self, u, name, args):
        a = [self.ex(u, x) for x in args]
        ts = [self.typ(u, x) for x in args]
        isd = "double" in ts
        isc = "double _Complex" in ts
        if name in ("cexp",) or (name == "exp" and isc):
            return "cexp(%s)" % a[0]
        if name in ("csqrt",) or (name == "sqrt" and isc):
            return "csqrt(%s)" % a[0]
        if name in ("clog",) or (name == "log" and isc):
            return "clog(%s)" % a[0]
        if name in ("cmplx", "dcmplx"):
            return "CMPLX(%s,%s)" % (a[0], a[1] if len(a) > 1 else "0.0")
        if name == "conjg":
            return "conj(%s)" % a[0]
        if name in ("aimag", "dimag", "imag"):
            return "cimag(%s)" % a[0]
        if isc and name in ("real", "dble", "dreal"):
            return "creal(%s)" % a[0]
        if isc and name in ("abs", "cabs"):
            return "cabs(%s)" % a[0]
        one = {"sqrt": "sqrt", "dsqrt": "sqrt", "sin": "sin", "dsin": "sin", "cos": "cos", "dcos": "cos",
               "tan": "tan", "exp": "exp", "dexp": "exp", "log": "log", "dlog": "log", "log10": "log10",
               "atan": "atan", "datan": "atan", "asin": "asin", "acos": "acos", "sinh": "sinh",
               "cosh": "cosh", "tanh": "tanh", "aint": "trunc", "anint": "round"}
        if name in one:
            return "%s(%s)" % (one[name], a[0])
        if name in ("atan2", "datan2"):
            return "atan2(%s,%s)" % (a[0], a[1])
        if name in ("abs", "dabs", "iabs"):
            return ("fabs(%s)" if isd else "abs(%s)") % a[0]
        if name in ("real", "dble", "float", "dfloat", "sngl"):
            return "((double)(%s))" % a[0]
        if name in ("int", "ifix"):
            return "((int)(%s))" % a[0]
        if name == "nint":
            return "((int)lround(%s))" % a[0]
        if name == "floor":
            return "((int)floor(%s))" % a[0]
        if name == "ceiling":
            return "((int)ceil(%s))" % a[0]
        if name in ("max", "dmax1", "amax1", "max0", "min", "dmin1", "amin1", "min0"):
            ismax = "max" in name
            f = ("f_dmax" if ismax else "f_dmin") if isd else ("f_imax" if ismax else "f_imin")
            r = a[0]
            for x in a[1:]:
                r = "%s(%s,%s)" % (f, r, x)
            return r
        if name == "mod":
            return ("fmod(%s,%s)" if isd else "((%s)%%(%s))") % (a[0], a[1])
        if name == "modulo":
            if isd:
                raise F2CError("real modulo")
            return "f_imodulo(%s,%s)" % (a[0], a[1])
        if name == "sign":
            return ("f_dsign(%s,%s)" if isd else "f_isign(%s,%s)") % (a[0], a[1])
        if name == "iand":
            return "((%s)&(%s))" % (a[0], a[1])
        if name == "ishft":
            return "f_ishft(%s,%s)" % (a[0], a[1])
        if name == "ior":
            return "((%s)|(%s))" % (a[0], a[1])
        raise F2CError("intrinsic not supported: " + name)

    def argref(self, u, e):
        """actual argument -> C pointer expression (Fortran passes everything by reference)"""
        if e[0] == "var":
            sy = self.lookup(u, e[1])
            if sy.dims is not None:
                return self.base(u, sy)
            if sy.kind == "dummy":
                if sy.external:
                    return self.vname(e[1])
                return self.vname(e[1])
            if sy.kind == "lparam":
                return "&(%s){%s}" % (sy.ctype(), self.vname(e[1]))
            if sy.kind == "param":
                self.used_params.add(e[1])
                return "&(int){%s}" % self.vname(e[1]) if sy.ftype != "real" else "&(double){%s}" % self.vname(e[1])
            if sy.external or (e[1] in self.tr.units and sy.kind == "local" and sy.dims is None
                               and not self.assigned_anywhere(u, e[1])):
                self.called.setdefault(e[1], None)
                return "(void*)%s_" % e[1]
            return "&" + self.ref(u, e[1])
        if e[0] == "call":
            sy = u.syms.get(e[1])
            if sy is not None and sy.dims is not None:
                return "&%s[%s]" % (self.base(u, sy), self.index(u, sy, e[2]))
        if e[0] == "str":
            return self.ex(u, e)
        t = self.typ(u, e)
        return "&(%s){%s}" % (t, self.ex(u, e))

    def assigned_anywhere(self, u, name):
        pat = re.compile(r"(^|[^a-z0-9_])%s\s*=" % re.escape(name))
        return any(pat.search(s) for _, s in u.body)

    # ---- statements
    def emit_unit(self, u):
        out = []
        self.u = u
        body = self.emit_body(u)
        # signature
        rt = "void" if u.kind != "function" else {"real": "double", "integer": "int", "logical": "int"}[u.ftype]
        params = []
        for a in u.args:
            sy = u.syms[a]
            if sy.external or (sy.dims is None and self.is_called_as_proc(u, a)):
                params.append("void (*%s)()" % self.vname(a))
            else:
                params.append("%s *%s" % (sy.ctype(), self.vname(a)))
        out.append("%s %s_(%s)\n{" % (rt, u.cname, ", ".join(params) if params else "void"))
        # constants of this unit (PARAMETER in the unit's own text), in definition order
        for name in u.decl_order:
            sy = u.syms[name]
            out.append("    const %s %s = %s;" % (sy.ctype(), self.vname(name), self.ex(u, sy.value)))
        # locals
        for name, sy in sorted(u.syms.items()):
            if sy.kind != "local" or name == u.name:
                continue
            if not self.name_used(u, name):
                continue
            if sy.ftype == "character":
                out.append("    char %s[%d] = {0};" % (self.vname(name), sy.charlen + 1))
                continue
            if sy.data_unsupported:
                raise F2CError("unsupported DATA for %s in %s" % (name, u.name))
            if sy.dims is not None and sy.init_list is not None:
                out.append("    static %s %s[] = {%s};" % (sy.ctype(), self.vname(name),
                           ", ".join(self.ex(u, v) for v in sy.init_list)))
            elif sy.dims is not None:
                n = "*".join("(long)(%s)" % self.extent(u, sy, d) for d in range(len(sy.dims)))
                out.append("    %s *%s = (%s*)f_scratch(sizeof(%s)*(%s));" %
                           (sy.ctype(), self.vname(name), sy.ctype(), sy.ctype(), n))
                self.needs_scratch = True
            elif name in self.tr.units or name in self.called and name not in self.assigned_names(u):
                continue
            else:
                if sy.data_unsupported:
                    raise F2CError("unsupported DATA for %s in %s" % (name, u.name))
                out.append("    %s%s %s = %s;" % ("static " if sy.static else "", sy.ctype(), self.vname(name),
                                                   self.ex(u, sy.init) if sy.init is not None else "0"))
        if u.kind == "function":
            out.append("    %s f_result = 0;" % rt)
        out.extend(body)
        out.append("f_return: ;")
        out.append("    f_scratch_release(f_mark);" if self.needs_scratch else "")
        if u.kind == "function":
            out.append("    return f_result;")
        out.append("}\n")
        if self.needs_scratch:
            # insert the mark right after the opening brace
            out.insert(1, "    long f_mark = f_scratch_mark();")
        return "\n".join(x for x in out if x != "")

    def assigned_names(self, u):
        names = set()
        for _, s in u.body:
            j = top_level_eq(s)
            if j > 0:
                m = re.match(r"^([a-z_$][a-z0-9_$]*)", s[:j].strip())
                if m:
                    names.add(m.group(1))
        return names

    def is_called_as_proc(self, u, a):
        pat = re.compile(r"\bcall\s+%s\b" % re.escape(a))
        return any(pat.search(s) for _, s in u.body)

    def name_used(self, u, name):
        pat = re.compile(r"(^|[^a-z0-9_$])%s($|[^a-z0-9_$])" % re.escape(name))
        return any(pat.search(s) for _, s in u.body)

    def emit_body(self, u):
        self.needs_scratch = False
        out = []
        ind = 1
        do_labels = []  # stack of labels closing labelled do loops (None for enddo loops)
        labels_used = set()
        for _, s in u.body:
            for m in re.finditer(r"\bgo\s*to\s+(\d+)", s):
                labels_used.add(m.group(1))

        def w(line):
            out.append("    " * ind + line)

        for lab, s in u.body:
            closes = 0
            if lab:
                if lab in labels_used:
                    out.append("L%s: ;" % lab)
                while do_labels and do_labels[-1] == lab:
                    closes += 1
                    do_labels.pop()
            ind_delta_after = 0
            stmts = self.stmt(u, s, do_labels)
            for line in stmts:
                if line.startswith("}"):
                    ind -= 1
                w(line)
                if line.endswith("{"):
                    ind += 1
            for _ in range(closes):
                ind -= 1
                w("}}")
        return out

    def simple_stmt(self, u, s):
        """statement allowed after a logical IF"""
        r = self.stmt(u, s, None)
        return r

    def stmt(self, u, s, do_labels):
        if s in ("continue",):
            return [";"]
        if s == "return":
            return ["goto f_return;"]
        if s == "exit":
            return ["break;"]
        if s == "cycle":
            return ["continue;"]
        if s.startswith("stop"):
            return ["exit(1);"]
        if re.match(r"^(write|print|format|read|open|close|rewind)\b", s) and top_level_eq(s.split(")")[0] if "(" in s else s) < 0:
            return ["/* io dropped */;"]
        m = re.match(r"^go\s*to\s+(\d+)$", s)
        if m:
            return ["goto L%s;" % m.group(1)]
        if re.match(r"^end\s*do$", s):
            if do_labels is not None:
                if not do_labels or do_labels[-1] is not None:
                    raise F2CError("enddo mismatch in " + u.name)
                do_labels.pop()
            return ["}}"]
        if re.match(r"^end\s*if$", s):
            return ["}"]
        if s == "else":
            return ["} else {"]
        m = re.match(r"^else\s*if\s*\(", s)
        if m:
            i0 = s.index("(")
            i1 = match_paren(s, i0)
            if s[i1 + 1:].strip() != "then":
                raise F2CError("bad elseif: " + s)
            return ["} else if (%s) {" % self.ex(u, parse_expr(s[i0 + 1:i1]))]
        m = re.match(r"^if\s*\(", s)
        if m:
            i0 = s.index("(")
            i1 = match_paren(s, i0)
            cond = self.ex(u, parse_expr(s[i0 + 1:i1]))
            rest = s[i1 + 1:].strip()
            if rest == "then":
                return ["if (%s) {" % cond]
            if re.match(r"^\d+\s*,", rest):
                raise F2CError("arithmetic IF not supported: " + s)
            inner = self.stmt(u, rest, None)
            return ["if (%s) {" % cond] + inner + ["}"]
        m = re.match(r"^do\s+while\s*\(", s)
        if m:
            i0 = s.index("(")
            i1 = match_paren(s, i0)
            do_labels.append(None)
            return ["{while (%s) {" % self.ex(u, parse_expr(s[i0 + 1:i1]))]
        m = re.match(r"^do\s*(\d+)?\s*,?\s*([a-z_][a-z0-9_]*)\s*=(.*)$", s)
        if m and len(split_top(m.group(3))) >= 2:
            lab, var, rng = m.group(1), m.group(2), split_top(m.group(3))
            v = self.ref(u, var)
            a = self.ex(u, parse_expr(rng[0]))
            b = self.ex(u, parse_expr(rng[1]))
            do_labels.append(lab)
            if len(rng) == 3:
                c = self.ex(u, parse_expr(rng[2]))
                return ["{int f_e=%s, f_s=%s; for (%s=%s; f_s>0 ? %s<=f_e : %s>=f_e; %s+=f_s) {" %
                        (b, c, v, a, v, v, v)]
            return ["{int f_e=%s; for (%s=%s; %s<=f_e; %s++) {" % (b, v, a, v, v)]
        m = re.match(r"^call\s+([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", s)
        if m:
            name = m.group(1)
            args = []
            if m.group(2):
                p = Parser(tokenize(m.group(2)))
                p.expect("op", "(")
                if not p.accept("op", ")"):
                    while True:
                        args.append(p.arg())
                        if p.accept("op", ")"):
                            break
                        p.expect("op", ",")
            sy = u.syms.get(name)
            if sy is not None and sy.kind == "dummy":
                return ["%s(%s);" % (self.vname(name), ", ".join(self.argref(u, a) for a in args))]
            name = u.rename.get(name, name)
            self.called.setdefault(name, None)
            return ["%s_(%s);" % (name, ", ".join(self.argref(u, a) for a in args))]
        j = top_level_eq(s)
        if j > 0:
            lhs, rhs = s[:j].strip(), s[j + 1:].strip()
            le = parse_expr(lhs)
            re_ = parse_expr(rhs)
            if (le[0] == "var" and self.lookup(u, le[1]).ftype == "character") or \
               (le[0] == "call" and u.syms.get(le[1]) is not None and u.syms[le[1]].dims is not None
                    and u.syms[le[1]].ftype == "character"):
                (pa, la), (pb, lb) = self.chref(u, le), self.chref(u, re_)
                return ["f_chassign(%s,%d,%s,%d);" % (pa, la, pb, lb)]
            if le[0] == "var":
                sy = self.lookup(u, le[1])
                if sy.dims is not None:
                    raise F2CError("whole-array assignment in %s: %s" % (u.name, s))
                if sy.kind in ("param", "lparam"):
                    raise F2CError("assignment to parameter: " + s)
                return ["%s = %s;" % (self.ref(u, le[1]), self.ex(u, re_))]
            if le[0] == "call":
                sy = u.syms.get(le[1])
                if sy is None or sy.dims is None:
                    raise F2CError("statement function or undeclared array in %s: %s" % (u.name, s))
                return ["%s[%s] = %s;" % (self.base(u, sy), self.index(u, sy, le[2]), self.ex(u, re_))]
        raise F2CError("unsupported statement in %s: %s" % (u.name, s))

    # ---- whole file
    def emit_all(self, unit_names, param_overridable=()):
        bodies = []
        for n in unit_names:
            bodies.append(self.emit_unit(self.tr.units[n]))
        out = [PRELUDE]
        # closure of the parameters used (a parameter's value may name other parameters)
        changed = True
        while changed:
            before = len(self.used_params)
            for name, val, u in self.tr.params:
                if name in self.used_params:
                    self.scan_params(u, val)
                    self.scan_all_params(u, val)
            changed = len(self.used_params) != before
        # parameters (run-time globals), in definition order
        pnames = [p for p in self.tr.params if p[0] in self.used_params]
        for name, val, u in pnames:
            sy = u.syms[name]
            out.append("%s %s;" % (sy.ctype(), self.vname(name)))
        # commons
        for name, (sy, u) in sorted(self.used_globals.items()):
            if sy.ftype == "character" and sy.dims is None:
                out.append("char %s[%d];" % (self.vname(name), sy.charlen + 1))
            elif sy.dims is None:
                out.append("%s %s;" % (sy.ctype(), self.vname(name)))
            else:
                out.append("%s *%s;" % (sy.ctype(), self.vname(name)))
        # prototypes
        for name, ft in sorted(self.called.items()):
            if name in self.tr.units and name in unit_names:
                continue
            rt = "void" if ft is None else {"real": "double", "integer": "int", "logical": "int"}[ft]
            out.append("extern %s %s_();" % (rt, name))
        for n in unit_names:
            u = self.tr.units[n]
            rt = "void" if u.kind != "function" else {"real": "double", "integer": "int", "logical": "int"}[u.ftype]
            out.append("%s %s_();" % (rt, n))
        out.append("")
        # registry + parameter evaluation + allocation
        out.append("static struct { const char *name; int isint; } f_ovr_dummy;")
        out.append("#define F_MAXOVR 64\nstatic const char *f_ovr_name[F_MAXOVR]; static long f_ovr_val[F_MAXOVR]; static int f_novr;")
        out.append("EXPORT void ref_set_param(const char *name, long val)\n{\n    for (int i = 0; i < f_novr; i++) if (!strcmp(f_ovr_name[i], name)) { f_ovr_val[i] = val; return; }\n"
                   "    if (f_novr < F_MAXOVR) { f_ovr_name[f_novr] = strdup(name); f_ovr_val[f_novr++] = val; }\n}")
        out.append("static int f_ovr(const char *name, long *v)\n{\n    for (int i = 0; i < f_novr; i++) if (!strcmp(f_ovr_name[i], name)) { *v = f_ovr_val[i]; return 1; }\n    return 0;\n}")
        out.append("static void f_eval_params(void)\n{\n    long ov;")
        for name, val, u in pnames:
            out.append("    %s = f_ovr(\"%s\", &ov) ? ov : (%s);" % (self.vname(name), name, self.ex(u, val)))
        out.append("}")
        out.append("EXPORT void ref_alloc(void)\n{\n    f_eval_params();")
        for name, (sy, u) in sorted(self.used_globals.items()):
            if sy.dims is not None:
                n = "*".join("(long)(%s)" % self.extent(u, sy, d) for d in range(len(sy.dims)))
                if sy.ftype == "character":
                    n = "(%s)*%d" % (n, sy.charlen)
                # 2x slack: the reference zeroes some face arrays past their end in 2D (SETAREA
                # clears 6 faces of arrays dimensioned 2*ldim, src/nek5_coef.F:1003-1007), which
                # in the Fortran image lands in the following members of the COMMON block
                out.append("    free(%s); %s = (%s*)calloc(2 * (size_t)(%s) + 64, sizeof(%s));" %
                           (self.vname(name), self.vname(name), sy.ctype(), n, sy.ctype()))
                if sy.ftype == "character":
                    out.append("    memset(%s, ' ', (size_t)(%s));" % (self.vname(name), n))
        out.append("}")
        out.append("EXPORT void *ref_sym(const char *name, int *isint, long *count)\n{")
        for name, val, u in pnames:
            out.append("    if (!strcmp(name, \"%s\")) { *isint = 1; *count = 1; return &%s; }" % (name, self.vname(name)))
        for name, (sy, u) in sorted(self.used_globals.items()):
            isint = 0 if sy.ftype == "real" else (2 if sy.ftype == "character" else (
                4 if sy.ftype == "complex" else (3 if sy.kind8 else 1)))
            if sy.ftype == "character":
                n = "*".join("(long)(%s)" % self.extent(u, sy, d) for d in range(len(sy.dims))) if sy.dims else "1"
                out.append("    if (!strcmp(name, \"%s\")) { *isint = 2; *count = (%s)*%d; return %s; }" %
                           (name, n, sy.charlen, self.vname(name)))
            elif sy.dims is None:
                out.append("    if (!strcmp(name, \"%s\")) { *isint = %d; *count = 1; return &%s; }" %
                           (name, isint, self.vname(name)))
            else:
                n = "*".join("(long)(%s)" % self.extent(u, sy, d) for d in range(len(sy.dims)))
                out.append("    if (!strcmp(name, \"%s\")) { *isint = %d; *count = %s; return %s; }" %
                           (name, isint, n, self.vname(name)))
        out.append("    return 0;\n}")
        out.append("EXPORT const char *ref_units(void) { return \"%s\"; }" % " ".join(unit_names))
        out.append("")
        out.extend(bodies)
        # export the translated entry points under their Fortran names too
        return "\n".join(out)


PRELUDE = r"""/* GENERATED by oracle/f2c_lite.py from the reference's Fortran sources -- do not commit */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define EXPORT __attribute__((visibility("default")))
static inline double f_powi(double x, int n)
{   /* x**n by repeated multiplication (gfortran expands small integer powers the same way) */
    int m = n < 0 ? -n : n;
    double r = 1.0, b = x;
    if (m == 2) return x * x;
    if (m == 3) return x * x * x;
    while (m) { if (m & 1) r *= b; m >>= 1; if (m) b *= b; }
    return n < 0 ? 1.0 / r : r;
}
static inline int f_ipow(int x, int n) { int r = 1; for (int i = 0; i < n; i++) r *= x; return r; }
static inline double f_dmax(double a, double b) { return a > b ? a : b; }
static inline double f_dmin(double a, double b) { return a < b ? a : b; }
static inline int f_imax(int a, int b) { return a > b ? a : b; }
static inline int f_imin(int a, int b) { return a < b ? a : b; }
static inline int f_imodulo(int a, int p) { int r = a % p; return (r != 0 && ((r < 0) != (p < 0))) ? r + p : r; }
static inline int f_ishft(int a, int s) { return s >= 0 ? (int)((unsigned)a << s) : (int)((unsigned)a >> (-s)); }
static inline double f_dsign(double a, double b) { return b >= 0 ? fabs(a) : -fabs(a); }
static inline int f_isign(int a, int b) { return b >= 0 ? abs(a) : -abs(a); }
/* character assignment / comparison with blank padding (Fortran semantics) */
static void f_chassign(char *d, int ld, const char *s, int ls)
{
    if (ls < 0) ls = (int)strlen(s);
    for (int i = 0; i < ld; i++) d[i] = i < ls ? s[i] : ' ';
}
static int f_chcmp(const char *a, int la, const char *b, int lb)
{
    if (la < 0) la = (int)strlen(a);
    if (lb < 0) lb = (int)strlen(b);
    int n = la > lb ? la : lb;
    for (int i = 0; i < n; i++) {
        char ca = i < la ? a[i] : ' ', cb = i < lb ? b[i] : ' ';
        if (ca != cb) return ca < cb ? -1 : 1;
    }
    return 0;
}
/* scratch stack for local (non-COMMON) arrays */
static char *f_scr; static long f_scr_cap, f_scr_top;
static long f_scratch_mark(void) { return f_scr_top; }
static void f_scratch_release(long m) { f_scr_top = m; }
static void *f_scratch(size_t n)
{
    n = (n + 63) & ~(size_t)63;
    if (f_scr_top + (long)n > f_scr_cap) {
        if (f_scr_top != 0) { fprintf(stderr, "f2c_lite: scratch overflow\n"); abort(); }
        f_scr_cap = (long)n * 4 + (64L << 20); free(f_scr); f_scr = (char*)malloc(f_scr_cap);
    }
    void *p = f_scr + f_scr_top; f_scr_top += (long)n; memset(p, 0, n); return p;
}
"""


def translate(sources, include_dirs, defines=()):
    """sources: [(path, [unit names])] -> C text"""
    tr = Translator(include_dirs, defines)
    names = []
    for src in sources:
        path, wanted = src[0], src[1]
        suffix = src[2] if len(src) > 2 else ""
        got = tr.load(path, set(wanted), suffix)
        missing = set(w + suffix for w in wanted) - set(got)
        if missing:
            raise F2CError("units not found in %s: %s" % (path, sorted(missing)))
        names.extend([w + suffix for w in wanted])
    em = Emitter(tr)
    return em.emit_all(names), em
