"""A small SystemVerilog-subset evaluator -- TEST INFRASTRUCTURE ONLY (like the rest of oracle/).

Why it exists: the fixed-point mode mirrors the reference RTL's integer datapath
(``rtl/unopt/gradient_compute.sv:108-141``, ``window_accumulator.sv:112-167``,
``flow_solver.sv:45,83-149``), there is no SystemVerilog simulator in the build image, and the
reference ships no reproducible known-answer data for it.  ``lk_fixed_oracle.py``
restates those modules by *reading* them.  This file removes the reading: it parses the modules' own
source text (declarations, ``always_comb`` / ``always_ff`` bodies) and executes the procedural
blocks with the expression sizing and signedness rules of IEEE 1800-2017 clause 11 (11.6 expression
bit lengths, 11.7 signed expressions, 11.8 expression evaluation rules), so quirks such as the
sign-extended 9-bit frame average or the truncation of the 64-bit products to their low 32 bits fall
out of the text mechanically.  ``tests/test_fixed_oracle_vs_rtl_text.py`` diffs it against the oracle.

What it is not: a simulator.  No event scheduling, no 4-state logic (x / z never arise on the paths
evaluated), no hierarchy -- module instantiations are skipped and their outputs are treated as inputs
the caller sets.  One call = one evaluation of a procedural block on the current variable values;
``always_ff`` bodies are executed as one clock edge (right-hand sides read the old values, the
non-blocking updates are committed together).

Supported subset (everything the three modules use): ``parameter int``, ``localparam``, ``logic
[signed] [msb:lsb] name[dims]`` declarations (ports, module items, block-local), ``int`` loop
variables, blocking / non-blocking assignments to scalars and array elements, ``begin/end``,
``if / else``, ``for (int i = a; i < b; i++)``, and in expressions: sized / unsized / fill literals,
identifiers, array indexing, bit and part selects, concatenation, ``$signed``, ``$unsigned``,
``$clog2``, unary ``- ~ !``, binary ``* / % + - << >> <<< >>> < <= > >= == != & ^ | && ||``, ``?:``.
"""

from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# ----------------------------------------------------------------------------------------------
# lexer
# ----------------------------------------------------------------------------------------------
_TOKEN = re.compile(
    r"""
    (?P<ws>\s+)
  | (?P<sized>\d*\s*'[sS]?[bBdDhHoO]\s*[0-9a-fA-F_xXzZ]+)
  | (?P<fill>'[01])
  | (?P<num>\d[\d_]*)
  | (?P<ident>[$A-Za-z_][A-Za-z0-9_$]*)
  | (?P<string>"[^"]*")
  | (?P<op><<<|>>>|<<|>>|<=|>=|==|!=|&&|\|\||\+\+|--|[-+*/%<>=!~&|^?:;,.(){}\[\]@\#])
    """,
    re.VERBOSE,
)


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"\(\*.*?\*\)", " ", text, flags=re.S)  # attributes: (* use_dsp = "yes" *)
    text = re.sub(r"`timescale[^\n]*", " ", text)
    return text


def tokenize(text: str) -> List[Tuple[str, str]]:
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at: {text[pos:pos + 40]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group()))
    return out


# ----------------------------------------------------------------------------------------------
# values and variables
# ----------------------------------------------------------------------------------------------
def _mask(w: int) -> int:
    return (1 << w) - 1


def to_signed(bits: int, w: int) -> int:
    bits &= _mask(w)
    return bits - (1 << w) if (bits >> (w - 1)) & 1 else bits


@dataclass
class Var:
    width: int
    signed: bool
    dims: Tuple[int, ...] = ()
    data: Dict[Tuple[int, ...], int] = field(default_factory=dict)  # index tuple -> bit pattern

    def get(self, idx: Tuple[int, ...]) -> int:
        return self.data.get(idx, 0)

    def set(self, idx: Tuple[int, ...], bits: int) -> None:
        if len(idx) != len(self.dims) or any(not (0 <= i < d) for i, d in zip(idx, self.dims)):
            raise IndexError(f"index {idx} out of range for dims {self.dims}")
        self.data[idx] = bits & _mask(self.width)


# ----------------------------------------------------------------------------------------------
# expression AST: every node knows its self-determined width / signedness (IEEE 1800-2017 11.6.1,
# 11.8.1) and evaluates in a context (width, signed) handed down by its parent (11.8.2)
# ----------------------------------------------------------------------------------------------
class Node:
    def width(self, env) -> int:
        raise NotImplementedError

    def signed(self, env) -> bool:
        raise NotImplementedError

    def eval(self, env, w: int, s: bool) -> int:
        """bit pattern of the expression evaluated at context width w and context type s"""
        raise NotImplementedError

    def self_eval(self, env) -> Tuple[int, int, bool]:
        w, s = self.width(env), self.signed(env)
        return self.eval(env, w, s), w, s


def _extend(bits: int, w_from: int, w_to: int, as_signed: bool) -> int:
    """11.8.2: an operand is converted to the propagated type and size; sign-extended only if the
    propagated type is signed"""
    bits &= _mask(w_from)
    if w_to <= w_from:
        return bits & _mask(w_to)
    if as_signed and (bits >> (w_from - 1)) & 1:
        bits |= _mask(w_to) & ~_mask(w_from)
    return bits


class Num(Node):
    def __init__(self, bits: int, w: int, s: bool, fill: bool = False):
        self.bits, self.w, self.s, self.fill = bits, w, s, fill

    def width(self, env):
        return self.w

    def signed(self, env):
        return self.s

    def eval(self, env, w, s):
        if self.fill:  # '0 / '1: every bit of the context width
            return _mask(w) if self.bits else 0
        return _extend(self.bits, self.w, w, s)


class Ref(Node):
    """identifier with optional unpacked-array indices, then an optional bit / part select"""

    def __init__(self, name: str, indices: List[Node], select: Optional[Tuple[Node, Optional[Node]]]):
        self.name, self.indices, self.select = name, indices, select

    def _var(self, env) -> Var:
        return env.lookup(self.name)

    def _split(self, env):
        v = self._var(env)
        nd = len(v.dims)
        idx = self.indices[:nd]
        rest = self.indices[nd:]
        sel = self.select
        if rest:  # name[i][j][bit]: a trailing single index beyond the unpacked dims is a bit select
            if len(rest) != 1 or sel is not None:
                raise SyntaxError(f"too many indices on {self.name}")
            sel = (rest[0], None)
        return v, idx, sel

    def width(self, env):
        v, _, sel = self._split(env)
        if sel is None:
            return v.width
        if sel[1] is None:
            return 1
        return env.const(sel[0]) - env.const(sel[1]) + 1

    def signed(self, env):
        v, _, sel = self._split(env)
        return v.signed if sel is None else False  # selects are unsigned (11.8.1)

    def eval(self, env, w, s):
        v, idx, sel = self._split(env)
        if len(idx) != len(v.dims):
            raise SyntaxError(f"{self.name}: whole-array reference in an expression")
        key = tuple(env.index_value(i) for i in idx)
        if any(not (0 <= k < d) for k, d in zip(key, v.dims)):
            raise IndexError(f"{self.name}{list(key)} out of range")
        bits = v.get(key)
        if sel is None:
            return _extend(bits, v.width, w, s)
        if sel[1] is None:
            b = env.index_value(sel[0])
            return _extend((bits >> b) & 1, 1, w, s)
        hi, lo = env.const(sel[0]), env.const(sel[1])
        return _extend((bits >> lo) & _mask(hi - lo + 1), hi - lo + 1, w, s)


class Concat(Node):
    def __init__(self, parts: List[Node]):
        self.parts = parts

    def width(self, env):
        return sum(p.width(env) for p in self.parts)

    def signed(self, env):
        return False

    def eval(self, env, w, s):
        bits, total = 0, 0
        for p in self.parts:  # operands are self-determined
            b, pw, _ = p.self_eval(env)
            bits = (bits << pw) | (b & _mask(pw))
            total += pw
        return _extend(bits, total, w, s)


class Cast(Node):
    """$signed / $unsigned: the argument is self-determined, the result keeps its width"""

    def __init__(self, arg: Node, to_signed_: bool):
        self.arg, self.s = arg, to_signed_

    def width(self, env):
        return self.arg.width(env)

    def signed(self, env):
        return self.s

    def eval(self, env, w, s):
        b, aw, _ = self.arg.self_eval(env)
        return _extend(b, aw, w, s)


class Unary(Node):
    def __init__(self, op: str, a: Node):
        self.op, self.a = op, a

    def width(self, env):
        return 1 if self.op == "!" else self.a.width(env)

    def signed(self, env):
        return False if self.op == "!" else self.a.signed(env)

    def eval(self, env, w, s):
        if self.op == "!":
            b, _, _ = self.a.self_eval(env)
            return _extend(0 if b else 1, 1, w, s)
        a = self.a.eval(env, w, s)
        if self.op == "-":
            return (-a) & _mask(w)
        if self.op == "~":
            return (~a) & _mask(w)
        if self.op == "+":
            return a
        raise SyntaxError(self.op)


_ARITH = {"+", "-", "*", "/", "%", "&", "|", "^"}
_SHIFT = {"<<", ">>", "<<<", ">>>"}
_REL = {"<", "<=", ">", ">=", "==", "!="}


class Binary(Node):
    def __init__(self, op: str, a: Node, b: Node):
        self.op, self.a, self.b = op, a, b

    def width(self, env):
        if self.op in _ARITH:
            return max(self.a.width(env), self.b.width(env))
        if self.op in _SHIFT:
            return self.a.width(env)
        return 1

    def signed(self, env):
        if self.op in _ARITH:
            return self.a.signed(env) and self.b.signed(env)
        if self.op in _SHIFT:
            return self.a.signed(env)
        return False

    def eval(self, env, w, s):
        op = self.op
        if op in _ARITH:
            a, b = self.a.eval(env, w, s), self.b.eval(env, w, s)
            if op == "+":
                r = a + b
            elif op == "-":
                r = a - b
            elif op == "*":
                r = a * b  # low w bits of the product are the same for either interpretation
            elif op in ("/", "%"):
                if b == 0:
                    raise ZeroDivisionError("division by zero gives x in SystemVerilog")
                if s:
                    sa, sb = to_signed(a, w), to_signed(b, w)
                    q = abs(sa) // abs(sb)  # truncation toward zero (11.4.2)
                    q = -q if (sa < 0) != (sb < 0) else q
                    r = q if op == "/" else sa - q * sb
                else:
                    r = a // b if op == "/" else a % b
            elif op == "&":
                r = a & b
            elif op == "|":
                r = a | b
            else:
                r = a ^ b
            return r & _mask(w)
        if op in _SHIFT:
            a = self.a.eval(env, w, s)
            n, _, _ = self.b.self_eval(env)  # the right operand is self-determined and unsigned in effect
            if op in ("<<", "<<<"):
                return (a << n) & _mask(w)
            if op == ">>>" and s:  # arithmetic only when the expression type is signed (11.4.10)
                return (to_signed(a, w) >> n) & _mask(w)
            return a >> n
        if op in _REL:
            # the two operands size each other; signed comparison only if both are signed (11.4.4, 11.8.1)
            cw = max(self.a.width(env), self.b.width(env))
            cs = self.a.signed(env) and self.b.signed(env)
            a, b = self.a.eval(env, cw, cs), self.b.eval(env, cw, cs)
            if cs:
                a, b = to_signed(a, cw), to_signed(b, cw)
            r = {"<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b, "==": a == b, "!=": a != b}[op]
            return _extend(1 if r else 0, 1, w, s)
        if op in ("&&", "||"):
            a, _, _ = self.a.self_eval(env)
            b, _, _ = self.b.self_eval(env)
            r = (a != 0 and b != 0) if op == "&&" else (a != 0 or b != 0)
            return _extend(1 if r else 0, 1, w, s)
        raise SyntaxError(op)


class Ternary(Node):
    def __init__(self, c: Node, a: Node, b: Node):
        self.c, self.a, self.b = c, a, b

    def width(self, env):
        return max(self.a.width(env), self.b.width(env))

    def signed(self, env):
        return self.a.signed(env) and self.b.signed(env)

    def eval(self, env, w, s):
        c, _, _ = self.c.self_eval(env)
        return (self.a if c else self.b).eval(env, w, s)


# ----------------------------------------------------------------------------------------------
# statements
# ----------------------------------------------------------------------------------------------
@dataclass
class Assign:
    target: Ref
    rhs: Node
    blocking: bool


@dataclass
class Block:
    stmts: list


@dataclass
class If:
    cond: Node
    then: object
    other: object


@dataclass
class For:
    var: str
    init: Node
    cond: Node
    step: int
    body: object


@dataclass
class Decl:
    name: str
    width_hi: Optional[Node]
    width_lo: Optional[Node]
    signed: bool
    dims: List[Node]
    is_int: bool = False


# ----------------------------------------------------------------------------------------------
# parser
# ----------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k][1] if self.i + k < len(self.t) else None

    def kind(self, k=0):
        return self.t[self.i + k][0] if self.i + k < len(self.t) else None

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok[1]

    def expect(self, s):
        got = self.next()
        if got != s:
            ctx = " ".join(x[1] for x in self.t[max(0, self.i - 8): self.i + 4])
            raise SyntaxError(f"expected {s!r}, got {got!r} near: {ctx}")

    def accept(self, s):
        if self.peek() == s:
            self.i += 1
            return True
        return False

    def skip_balanced(self, open_, close_):
        self.expect(open_)
        depth = 1
        while depth:
            tok = self.next()
            depth += tok == open_
            depth -= tok == close_

    # ---- expressions, lowest precedence first (11.3.2 operator precedence) ----------------------
    def expr(self) -> Node:
        c = self.lor()
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return Ternary(c, a, b)
        return c

    def _left(self, sub, ops):
        a = sub()
        while self.peek() in ops and self.kind() == "op":
            op = self.next()
            a = Binary(op, a, sub())
        return a

    def lor(self):
        return self._left(self.land, {"||"})

    def land(self):
        return self._left(self.bor, {"&&"})

    def bor(self):
        return self._left(self.bxor, {"|"})

    def bxor(self):
        return self._left(self.band, {"^"})

    def band(self):
        return self._left(self.equality, {"&"})

    def equality(self):
        return self._left(self.relational, {"==", "!="})

    def relational(self):
        return self._left(self.shift, {"<", "<=", ">", ">="})

    def shift(self):
        return self._left(self.additive, {"<<", ">>", "<<<", ">>>"})

    def additive(self):
        return self._left(self.multiplicative, {"+", "-"})

    def multiplicative(self):
        return self._left(self.unary, {"*", "/", "%"})

    def unary(self) -> Node:
        if self.kind() == "op" and self.peek() in ("-", "+", "~", "!"):
            op = self.next()
            return Unary(op, self.unary())
        return self.primary()

    def primary(self) -> Node:
        k, v = self.kind(), self.peek()
        if k == "sized":
            self.next()
            return parse_sized(v)
        if k == "fill":
            self.next()
            return Num(int(v[1]), 1, False, fill=True)
        if k == "num":
            self.next()
            return Num(int(v.replace("_", "")), 32, True)  # unsized decimal: a signed 32-bit integer (5.7.1)
        if v == "(":
            self.next()
            e = self.expr()
            self.expect(")")
            return e
        if v == "{":
            self.next()
            parts = [self.expr()]
            while self.accept(","):
                parts.append(self.expr())
            self.expect("}")
            return Concat(parts)
        if k == "ident":
            name = self.next()
            if name in ("$signed", "$unsigned"):
                self.expect("(")
                e = self.expr()
                self.expect(")")
                return Cast(e, name == "$signed")
            if name == "$clog2":
                self.expect("(")
                e = self.expr()
                self.expect(")")
                return Clog2(e)
            return self.ref_tail(name)
        raise SyntaxError(f"unexpected token {v!r} in expression")

    def ref_tail(self, name) -> Ref:
        indices, select = [], None
        while self.peek() == "[":
            self.next()
            a = self.expr()
            if self.accept(":"):
                b = self.expr()
                self.expect("]")
                select = (a, b)
                break
            self.expect("]")
            indices.append(a)
        return Ref(name, indices, select)

    # ---- declarations ---------------------------------------------------------------------------
    def decl_after_logic(self) -> List[Decl]:
        """after the keyword `logic`: [signed] [msb:lsb] name[dims] {, name[dims]}"""
        signed_ = self.accept("signed")
        self.accept("unsigned")
        hi = lo = None
        if self.peek() == "[":
            self.next()
            hi = self.expr()
            self.expect(":")
            lo = self.expr()
            self.expect("]")
        out = []
        while True:
            name = self.next()
            dims = []
            while self.peek() == "[":
                self.next()
                dims.append(self.expr())
                self.expect("]")
            out.append(Decl(name, hi, lo, signed_, dims))
            # a following `, name` continues this declaration; `, input ...` (port list) does not
            if self.peek() == "," and self.kind(1) == "ident" and self.peek(1) not in ("input", "output", "inout", "logic"):
                self.next()
                continue
            break
        return out

    # ---- statements -------------------------------------------------------------------------------
    def statement(self):
        v = self.peek()
        if v == "begin":
            self.next()
            if self.accept(":"):
                self.next()
            stmts = []
            while self.peek() != "end":
                st = self.statement()
                stmts.extend(st) if isinstance(st, list) else stmts.append(st)  # a declaration: Decl entries of this block
            self.next()
            if self.accept(":"):
                self.next()
            return Block(stmts)
        if v == "logic":
            self.next()
            d = self.decl_after_logic()
            self.expect(";")
            return d
        if v == "if":
            self.next()
            self.expect("(")
            c = self.expr()
            self.expect(")")
            then = self.statement()
            other = None
            if self.accept("else"):
                other = self.statement()
            return If(c, then, other)
        if v == "for":
            self.next()
            self.expect("(")
            self.accept("int")
            var = self.next()
            self.expect("=")
            init = self.expr()
            self.expect(";")
            cond = self.expr()
            self.expect(";")
            var2 = self.next()
            if var2 != var:
                raise SyntaxError("for-loop step on another variable")
            stepop = self.next()
            if stepop not in ("++", "--"):
                raise SyntaxError("only i++ / i-- steps are supported")
            self.expect(")")
            body = self.statement()
            return For(var, init, cond, 1 if stepop == "++" else -1, body)
        if self.kind() == "ident" and v.startswith("$"):  # $display(...) etc.: no effect on values
            self.next()
            if self.peek() == "(":
                self.skip_balanced("(", ")")
            self.expect(";")
            return Block([])
        # assignment
        name = self.next()
        target = self.ref_tail(name)
        op = self.next()
        if op not in ("=", "<="):
            raise SyntaxError(f"expected an assignment, got {op!r} after {name}")
        rhs = self.expr()
        self.expect(";")
        return Assign(target, rhs, op == "=")


class Clog2(Node):
    def __init__(self, a):
        self.a = a

    def width(self, env):
        return 32

    def signed(self, env):
        return True

    def eval(self, env, w, s):
        n = env.const(self.a)
        return _extend(max(0, (n - 1).bit_length()), 32, w, s)


def parse_sized(text: str) -> Num:
    m = re.match(r"(\d*)\s*'([sS]?)([bBdDhHoO])\s*([0-9a-fA-F_xXzZ]+)", text)
    size, s, base, digits = m.groups()
    digits = digits.replace("_", "")
    if re.search(r"[xXzZ]", digits):
        raise SyntaxError("x / z literals are outside the evaluated subset")
    val = int(digits, {"b": 2, "d": 10, "h": 16, "o": 8}[base.lower()])
    w = int(size) if size else 32
    return Num(val & _mask(w), w, bool(s))


# ----------------------------------------------------------------------------------------------
# module: declarations + procedural blocks, and the environment they execute in
# ----------------------------------------------------------------------------------------------
class Module:
    """Parsed module: `params` (name -> int), `vars` (name -> Var), `blocks` (list of (kind, statement))
    in source order, kind in {"always_comb", "always_ff", "assign"}."""

    def __init__(self, text: str, param_overrides: Optional[Dict[str, int]] = None):
        self.params: Dict[str, int] = {}
        self.param_types: Dict[str, Tuple[int, bool]] = {}
        self.vars: Dict[str, Var] = {}
        self.blocks: List[Tuple[str, object]] = []
        self.instances: List[str] = []
        self._scopes: List[Dict[str, Var]] = []
        self._pending: List[Tuple[Var, Tuple[int, ...], int]] = []
        self.last_locals: Dict[str, Var] = {}
        self._parse(tokenize(strip_comments(text)), dict(param_overrides or {}))

    # ---- environment interface used by the AST -------------------------------------------------
    def lookup(self, name: str) -> Var:
        for sc in reversed(self._scopes):
            if name in sc:
                return sc[name]
        if name in self.vars:
            return self.vars[name]
        if name in self.params:
            w, s = self.param_types.get(name, (32, True))
            v = Var(w, s)
            v.data[()] = self.params[name] & _mask(w)
            return v
        raise NameError(f"unknown identifier {name!r}")

    def const(self, node: Node) -> int:
        b, w, s = node.self_eval(self)
        return to_signed(b, w) if s else b

    def index_value(self, node: Node) -> int:
        return self.const(node)

    # ---- parsing ----------------------------------------------------------------------------------
    def _declare(self, d: Decl, scope: Optional[Dict[str, Var]] = None):
        if d.is_int:
            v = Var(32, True)
        else:
            w = 1 if d.width_hi is None else self.const(d.width_hi) - self.const(d.width_lo) + 1
            v = Var(w, d.signed, tuple(self.const(x) for x in d.dims))
        (self.vars if scope is None else scope)[d.name] = v
        return v

    def _parse(self, toks, overrides):
        p = Parser(toks)
        p.expect("module")
        self.name = p.next()
        if p.accept("#"):
            p.expect("(")
            while p.peek() != ")":
                p.expect("parameter")
                p.accept("int")
                name = p.next()
                p.expect("=")
                val = self.const(p.expr())
                self.params[name] = overrides.get(name, val)
                p.accept(",")
            p.expect(")")
        p.expect("(")
        while p.peek() != ")":
            direction = p.next()
            if direction not in ("input", "output", "inout"):
                raise SyntaxError(f"port direction expected, got {direction!r}")
            p.expect("logic")
            for d in p.decl_after_logic():
                self._declare(d)
            p.accept(",")
        p.expect(")")
        p.expect(";")
        while p.peek() is not None and p.peek() != "endmodule":
            v = p.peek()
            if v == "localparam":
                p.next()
                w, s = 32, True
                if p.accept("int"):
                    pass
                elif p.accept("logic"):
                    s = p.accept("signed")
                    w = 1
                    if p.accept("["):
                        hi = self.const(p.expr())
                        p.expect(":")
                        lo = self.const(p.expr())
                        p.expect("]")
                        w = hi - lo + 1
                name = p.next()
                p.expect("=")
                e = p.expr()
                p.expect(";")
                ew, es = e.width(self), e.signed(self)
                bits = e.eval(self, max(w, ew), es) & _mask(w)  # assignment-like context
                self.params[name] = to_signed(bits, w) if s else bits
                self.param_types[name] = (w, s)
            elif v == "logic":
                p.next()
                for d in p.decl_after_logic():
                    self._declare(d)
                p.expect(";")
            elif v == "always_comb":
                p.next()
                self.blocks.append(("always_comb", p.statement()))
            elif v == "always_ff":
                p.next()
                p.expect("@")
                p.skip_balanced("(", ")")
                self.blocks.append(("always_ff", p.statement()))
            elif v == "assign":
                p.next()
                name = p.next()
                target = p.ref_tail(name)
                p.expect("=")
                rhs = p.expr()
                p.expect(";")
                self.blocks.append(("assign", Assign(target, rhs, True)))
            elif p.kind() == "ident":  # module instantiation: name [#(...)] inst (...);
                inst_of = p.next()
                if p.accept("#"):
                    p.skip_balanced("(", ")")
                inst = p.next()
                p.skip_balanced("(", ")")
                p.expect(";")
                self.instances.append(f"{inst_of} {inst}")
            else:
                raise SyntaxError(f"unsupported module item starting at {v!r}")

    # ---- execution --------------------------------------------------------------------------------
    def set(self, name: str, value: int, *idx: int) -> None:
        """store the low `width` bits of value (two's complement for negative Python ints)"""
        self.vars[name].set(tuple(idx), value)

    def get(self, name: str, *idx: int, signed: Optional[bool] = None) -> int:
        v = self.vars[name]
        bits = v.get(tuple(idx))
        s = v.signed if signed is None else signed
        return to_signed(bits, v.width) if s else bits

    def _assign(self, a: Assign):
        v, idx_nodes, sel = a.target._split(self)
        if sel is not None:
            raise SyntaxError("assignments to bit / part selects are outside the evaluated subset")
        key = tuple(self.index_value(i) for i in idx_nodes)
        # 11.8.2 / 10.7: the right-hand side is evaluated at max(width of lhs, self-determined width of rhs)
        # with the rhs's own signedness, then truncated to the lhs
        rw, rs = a.rhs.width(self), a.rhs.signed(self)
        bits = a.rhs.eval(self, max(v.width, rw), rs) & _mask(v.width)
        if a.blocking:
            v.set(key, bits)
        else:
            self._pending.append((v, key, bits))

    def _exec(self, st):
        if isinstance(st, Assign):
            self._assign(st)
        elif isinstance(st, Decl):
            self._declare(st, self._scopes[-1])
        elif isinstance(st, Block):
            self._scopes.append({})
            try:
                for s_ in st.stmts:
                    self._exec(s_)
            finally:
                scope = self._scopes.pop()
                if not self._scopes:
                    self.last_locals = scope  # the outermost block's local variables, for inspection
        elif isinstance(st, If):
            c, _, _ = st.cond.self_eval(self)
            if c:
                self._exec(st.then)
            elif st.other is not None:
                self._exec(st.other)
        elif isinstance(st, For):
            loop = Var(32, True)
            self._scopes.append({st.var: loop})
            try:
                loop.data[()] = st.init.eval(self, 32, True)
                guard = 0
                while st.cond.self_eval(self)[0]:
                    self._exec(st.body)
                    loop.data[()] = (loop.data[()] + st.step) & _mask(32)
                    guard += 1
                    if guard > 1 << 20:
                        raise RuntimeError("runaway for loop")
            finally:
                self._scopes.pop()
        else:
            raise TypeError(st)

    def run(self, kind: str, nth: int = 0) -> None:
        """Execute the nth procedural block of `kind` once.  Block-local declarations made inside a
        `begin ... end` live for that execution.  always_ff: one clock edge."""
        blocks = [b for k, b in self.blocks if k == kind]
        self._pending = []
        # a block's local declarations are Decl entries inside its outer Block
        self._exec(blocks[nth])
        for v, key, bits in self._pending:
            v.set(key, bits)
        self._pending = []
