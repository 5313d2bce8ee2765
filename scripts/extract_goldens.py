"""Transliterates the reference's own backend tests into golden fixtures (run in the build container only).

Reads crates/burn-backend-tests/tests/**.rs under /root/reference (read-only, absent on the GPU box), interprets
the subset of Rust those tests are written in — tensor literals, method chains, operators, `for` loops over
literal lists, `assert_eq` / `assert_approx_eq` on `into_data()` — and records every assertion as
    {"name", "cite": "<file>:<line>", "expr": <op tree over literal tensors>, "expected": <literal>, "tol": ...}
into tests/golden/burn_backend_tests_expr.json.  Nothing is copied but the numeric literals the assertions pin
and the order of the public Tensor-API calls that produce them.  A test using anything outside the subset is
skipped (listed at the end); a case the CPU oracle cannot evaluate is skipped too; a case where the oracle
DISAGREES with the reference literal is reported loudly and not written — that is an oracle bug to fix.

usage: python scripts/extract_goldens.py [--verbose]
"""
from __future__ import annotations

import json
import math
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/crates/burn-backend-tests/tests"
FILES = [
    "tensor/float/ops/maxmin.rs", "tensor/float/ops/comparison.rs", "tensor/float/ops/mask.rs",
    "tensor/float/ops/aggregation.rs", "tensor/float/ops/matmul.rs", "tensor/float/ops/arg.rs",
    "tensor/float/ops/add.rs", "tensor/float/ops/sub.rs", "tensor/float/ops/mul.rs", "tensor/float/ops/div.rs",
    "tensor/float/ops/exp.rs", "tensor/float/ops/log.rs", "tensor/float/ops/log1p.rs", "tensor/float/ops/sqrt.rs",
    "tensor/float/ops/erf.rs", "tensor/float/ops/abs.rs", "tensor/float/ops/neg.rs", "tensor/float/ops/recip.rs",
    "tensor/float/ops/powf.rs", "tensor/float/ops/powf_scalar.rs", "tensor/float/ops/clamp.rs",
    "tensor/float/ops/floor.rs", "tensor/float/ops/ceil.rs", "tensor/float/ops/round.rs", "tensor/float/ops/trunc.rs",
    "tensor/float/ops/sign.rs", "tensor/float/ops/trig.rs", "tensor/float/ops/remainder.rs", "tensor/float/ops/square.rs",
    "tensor/float/ops/gather_scatter.rs", "tensor/float/ops/select.rs", "tensor/float/ops/transpose.rs",
    "tensor/float/ops/reshape.rs", "tensor/float/ops/permute.rs", "tensor/float/ops/prod.rs",
    "tensor/float/ops/cat.rs", "tensor/float/ops/flip.rs", "tensor/float/ops/repeat_dim.rs",
    "tensor/float/ops/slice.rs", "tensor/float/ops/slice_assign.rs", "tensor/float/ops/expand.rs",
    "tensor/float/ops/nan.rs", "tensor/float/ops/inf.rs", "tensor/float/ops/negative_dims.rs",
    "tensor/float/activation/gelu.rs", "tensor/float/activation/relu.rs", "tensor/float/activation/sigmoid.rs",
    "tensor/float/activation/softmax.rs", "tensor/float/activation/log_softmax.rs",
    "fusion/reduce_broadcasted.rs", "fusion/inplace.rs", "fusion/reduce_logical.rs", "fusion/fusion_shape.rs",
]


class Unsupported(Exception):
    pass


# ------------------------------------------------------------------ tokenizer
TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<num>(?:\d[\d_]*\.\d[\d_]*|\d[\d_]*\.(?![.\w])|\d[\d_]*)(?:[eE][+-]?\d+)?(?:_?(?:f32|f64|i32|i64|u8|usize|isize|u32|i8|i16|u16|u64))?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*!?)
  | (?P<str>"(?:[^"\\]|\\[\s\S])*")
  | (?P<op>::|\.\.=|\.\.|->|=>|==|!=|<=|>=|&&|\|\||[-+*/%&|!<>=.,;:()\[\]{}#?'])
""", re.X)


def tokenize(src: str):
    out, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN.match(src, pos)
        if not m:
            raise Unsupported(f"cannot tokenize at line {line}: {src[pos:pos + 20]!r}")
        kind = m.lastgroup
        text = m.group()
        if kind != "ws":
            out.append((kind, text, line))
        line += text.count("\n")
        pos = m.end()
    return out


# ------------------------------------------------------------------ values
class Data:      # TensorData
    def __init__(self, lit, node=None):
        self.lit, self.node = lit, node


def sym(op, **kw):
    d = {"op": op}
    d.update(kw)
    return d


def lit_kind(lit):
    x = lit
    while isinstance(x, list):
        if not x:
            return "float"
        x = x[0]
    if isinstance(x, bool):
        return "bool"
    if isinstance(x, int):
        return "int"
    return "float"


def map_lit(lit, fn):
    return [map_lit(e, fn) for e in lit] if isinstance(lit, list) else fn(lit)


def tensor_from(lit, kind):
    if isinstance(lit, Data):
        if lit.node is not None:
            raise Unsupported("tensor from computed data")
        lit = lit.lit
    if isinstance(lit, dict):
        return lit
    if kind == "float":
        lit = map_lit(lit, lambda v: float(v))
    elif kind == "int":
        lit = map_lit(lit, lambda v: int(v))
    else:
        lit = map_lit(lit, lambda v: bool(v))
    return sym("lit", kind=kind, value=lit)


def kind_of(node):
    if node["op"] == "lit":
        return node["kind"]
    if node["op"] in ("int", "argmax", "argmin"):
        return "int"
    if node["op"] == "float":
        return "float"
    if node["op"] in ("equal", "not_equal", "greater", "greater_equal", "lower", "lower_equal", "equal_elem", "not_equal_elem",
                      "greater_elem", "greater_equal_elem", "lower_elem", "lower_equal_elem", "bool", "is_nan", "is_inf"):
        return "bool"
    if node["op"] in ("max_dim_with_indices", "min_dim_with_indices") and node.get("out") == 1:
        return "int"
    return kind_of(node["x"]) if "x" in node else "float"


VIEW_OPS = {"reshape", "transpose", "swap_dims", "permute", "flatten", "unsqueeze", "squeeze", "expand", "slice", "flip", "t"}
TENSOR_METHODS = {
    # elementwise
    "add", "sub", "mul", "div", "remainder", "powf", "add_scalar", "sub_scalar", "mul_scalar", "div_scalar", "remainder_scalar",
    "powf_scalar", "powi_scalar", "exp", "log", "log1p", "sqrt", "abs", "neg", "recip", "tanh", "erf", "sin", "cos", "tan", "floor",
    "ceil", "round", "trunc", "sign", "square", "clamp", "clamp_min", "clamp_max", "is_nan", "is_inf",
    # compare / mask
    "equal", "not_equal", "greater", "greater_equal", "lower", "lower_equal", "equal_elem", "not_equal_elem", "greater_elem",
    "greater_equal_elem", "lower_elem", "lower_equal_elem", "mask_fill", "mask_where",
    # reduce
    "sum", "mean", "prod", "sum_dim", "mean_dim", "prod_dim", "max", "min", "max_dim", "min_dim", "argmax", "argmin",
    "max_dim_with_indices", "min_dim_with_indices", "max_abs", "max_abs_dim",
    # contraction / index / movement
    "matmul", "gather", "scatter", "select", "select_assign", "repeat_dim", "slice_assign", "slice_fill", "narrow",
    # casts
    "int", "float", "bool",
}


# ------------------------------------------------------------------ parser / evaluator
class Interp:
    def __init__(self, toks, file, fn_name):
        self.t, self.i, self.env = toks, 0, {}
        self.file, self.fn = file, fn_name
        self.cases = []

    # -- token helpers
    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "", -1)

    def at(self, text):
        return self.peek()[1] == text

    def eat(self, text=None):
        tok = self.peek()
        if text is not None and tok[1] != text:
            raise Unsupported(f"expected {text!r}, got {tok[1]!r} (line {tok[2]})")
        self.i += 1
        return tok

    def skip_generic(self):     # at '<'
        depth = 0
        while True:
            tok = self.eat()
            if tok[1] == "<":
                depth += 1
            elif tok[1] == ">":
                depth -= 1
                if depth == 0:
                    return
            elif tok[0] == "eof":
                raise Unsupported("unterminated generic")

    # -- statements
    def block(self):
        self.eat("{")
        while not self.at("}"):
            self.statement()
        self.eat("}")

    def statement(self):
        tok = self.peek()
        if tok[1] == "let":
            self.eat()
            if self.at("mut"):
                self.eat()
            if self.at("("):
                self.eat()
                names = []
                while not self.at(")"):
                    names.append(self.eat()[1])
                    if self.at(","):
                        self.eat()
                self.eat(")")
            else:
                names = self.eat()[1]
            if self.at(":"):    # type annotation
                self.eat()
                while not self.at("="):
                    if self.at("<"):
                        self.skip_generic()
                    else:
                        self.eat()
            self.eat("=")
            val = self.expr()
            self.eat(";")
            if isinstance(names, list):
                if not isinstance(val, tuple) or len(val) != len(names):
                    raise Unsupported("tuple destructuring of a non-tuple")
                for n, v in zip(names, val):
                    self.env[n] = v
            else:
                self.env[names] = val
            return
        if tok[1] == "for":
            self.eat()
            var = self.eat()[1]
            self.eat("in")
            seq = self.expr()
            if isinstance(seq, tuple) and seq and seq[0] == "range":
                seq = list(range(seq[1], seq[2]))
            if not isinstance(seq, list):
                raise Unsupported("for over a non-literal")
            start = self.i
            for v in seq:
                self.i = start
                self.env[var] = v
                self.block()
            if not seq:
                raise Unsupported("empty for")
            return
        if tok[1] in ("if", "while", "match", "loop", "fn", "use", "#"):
            raise Unsupported(f"statement {tok[1]}")
        self.expr()
        if self.at(";"):
            self.eat()

    # -- expressions
    def expr(self):
        lhs = self.term()
        while self.peek()[1] in ("+", "-"):
            op = self.eat()[1]
            rhs = self.term()
            lhs = self.binop(op, lhs, rhs)
        if self.at(".."):       # range a..b
            self.eat()
            hi = self.term()
            return ("range", lhs, hi)
        return lhs

    def term(self):
        lhs = self.unary()
        while self.peek()[1] in ("*", "/", "%"):
            op = self.eat()[1]
            rhs = self.unary()
            lhs = self.binop(op, lhs, rhs)
        return lhs

    def binop(self, op, a, b):
        num = (int, float)
        if isinstance(a, num) and isinstance(b, num) and not isinstance(a, bool):
            if op == "/" and isinstance(a, int) and isinstance(b, int):
                return a // b
            return {"+": a + b, "-": a - b, "*": a * b, "/": a / b if b else math.nan, "%": math.fmod(a, b) if b else math.nan}[op]
        name = {"+": "add", "-": "sub", "*": "mul", "/": "div", "%": "remainder"}[op]
        if isinstance(a, dict) and isinstance(b, dict):
            return sym(name, x=a, args=[b])
        if isinstance(a, dict) and isinstance(b, num):
            return sym(name + "_scalar", x=a, args=[b])
        raise Unsupported(f"operator {op} on {type(a).__name__}, {type(b).__name__}")

    def unary(self):
        if self.at("-"):
            self.eat()
            v = self.unary()
            if isinstance(v, (int, float)):
                return -v
            if isinstance(v, dict):
                return sym("neg", x=v, args=[])
            raise Unsupported("negation")
        if self.at("&"):
            self.eat()
            if self.at("mut"):
                self.eat()
            return self.unary()
        if self.at("!"):
            raise Unsupported("not")
        return self.postfix()

    def args(self):
        self.eat("(")
        out = []
        while not self.at(")"):
            out.append(self.expr())
            if self.at(","):
                self.eat()
        self.eat(")")
        return out

    def postfix(self):
        v = self.primary()
        while True:
            if self.at("."):
                nxt = self.peek(1)
                if nxt[0] == "num":      # tuple field .0
                    self.eat()
                    idx = int(self.eat()[1])
                    if not isinstance(v, tuple):
                        raise Unsupported("field of non-tuple")
                    v = v[idx]
                    continue
                self.eat()
                name = self.eat()[1]
                if self.at("::"):
                    self.eat()
                    self.skip_generic()
                if not self.at("("):
                    raise Unsupported(f"field access .{name}")
                line = self.peek()[2]
                a = self.args()
                v = self.method(v, name, a, line)
            elif self.at("as"):
                self.eat()
                self.eat()
            elif self.at("?"):
                self.eat()
            else:
                return v

    def primary(self):
        kind, text, line = self.peek()
        if kind == "num":
            self.eat()
            t = re.sub(r"_?(f32|f64|i32|i64|u8|usize|isize|u32|i8|i16|u16|u64)$", "", text).replace("_", "")
            is_float = "." in t or "e" in t.lower() or text.endswith(("f32", "f64"))
            return float(t) if is_float else int(t)
        if text in ("true", "false"):
            self.eat()
            return text == "true"
        if text == "[":
            self.eat()
            items = []
            while not self.at("]"):
                items.append(self.expr())
                if self.at(";"):     # [v; n]
                    self.eat()
                    n = self.expr()
                    self.eat("]")
                    return [items[0]] * int(n)
                if self.at(","):
                    self.eat()
            self.eat("]")
            return items
        if text == "(":
            self.eat()
            items = []
            while not self.at(")"):
                items.append(self.expr())
                if self.at(","):
                    self.eat()
            self.eat(")")
            return items[0] if len(items) == 1 else tuple(items)
        if kind == "str":
            self.eat()
            return text
        if kind == "id" and text == "vec!":
            self.eat()
            return self.primary()          # vec![a, b] reads as the list literal
        if kind == "id":
            return self.path()
        raise Unsupported(f"unexpected token {text!r} (line {line})")

    def path(self):
        segs = [self.eat()[1]]
        while self.at("::"):
            self.eat()
            if self.at("<"):
                self.skip_generic()
                continue
            segs.append(self.eat()[1])
        name = "::".join(segs)
        last = segs[-1]
        if name in self.env and not self.at("("):
            return self.env[name]
        consts = {"NAN": math.nan, "INFINITY": math.inf, "NEG_INFINITY": -math.inf, "MAX": 3.4028234663852886e38,
                  "MIN": -3.4028234663852886e38, "EPSILON": 1.1920928955078125e-07, "MIN_POSITIVE": 1.1754943508222875e-38}
        if len(segs) == 2 and segs[0] in ("f32", "f64", "FloatElem") and last in consts and not self.at("("):
            return consts[last]
        if name in ("core::f32::consts::PI", "std::f32::consts::PI", "f32::consts::PI"):
            return math.pi
        if name in ("core::f32::consts::E", "std::f32::consts::E", "f32::consts::E"):
            return math.e
        if not self.at("("):
            if name in ("Default::default",):
                return None
            if last in ("Add", "Assign", "Mul", "Max", "Min") and "IndexingUpdateOp" in name:
                return ("update", last)
            raise Unsupported(f"unknown name {name}")
        line = self.peek()[2]
        a = self.args()
        return self.call(segs, a, line)

    # -- calls
    def call(self, segs, a, line):
        head, last = segs[0], segs[-1]
        name = "::".join(segs)
        if name == "Default::default":
            return None
        if name in ("TensorData::from", "TensorData::new"):
            if name.endswith("new"):
                raise Unsupported("TensorData::new")
            return Data(a[0])
        ttype = {"TestTensor": "float", "TestTensorInt": "int", "TestTensorBool": "bool", "Tensor": "float"}.get(head)
        if ttype is not None:
            if last in ("from", "from_data", "from_floats", "from_ints", "from_bool"):
                k = {"from_floats": "float", "from_ints": "int", "from_bool": "bool"}.get(last, ttype)
                return tensor_from(a[0], k)
            if last in ("ones", "zeros", "full"):
                shape = a[0]
                fill = {"ones": 1, "zeros": 0}.get(last, a[1] if last == "full" else 0)
                lit = fill
                for n in reversed(shape):
                    lit = [lit] * int(n)
                return tensor_from(json.loads(json.dumps(lit)), ttype)
            if last == "cat":
                parts = a[0]
                if not (isinstance(parts, list) and parts and all(isinstance(t, dict) for t in parts)):
                    raise Unsupported("cat of non-tensors")
                return sym("cat", x=parts[0], args=[parts[1:], int(a[1])])
            if last == "arange":
                r = a[0]
                if not (isinstance(r, tuple) and r[0] == "range"):
                    raise Unsupported("arange arg")
                return tensor_from(list(range(int(r[1]), int(r[2]))), "int")
            raise Unsupported(f"constructor {name}")
        if last in ("softmax", "log_softmax", "relu", "gelu", "sigmoid") and "activation" in segs[:-1] + [""]:
            return sym(last, x=a[0], args=a[1:])
        if last in ("softmax", "log_softmax", "relu", "gelu", "sigmoid") and len(segs) == 1:
            return sym(last, x=a[0], args=a[1:])
        if name in ("Tolerance::default", "Tolerance::permissive", "Tolerance::strict", "Tolerance::balanced"):
            return ("tol", {"default": (5e-3, 1e-5), "permissive": (1e-2, 1e-2), "strict": (1e-5, 1e-8), "balanced": (1e-3, 1e-5)}[last])
        if name == "Tolerance::rel_abs":
            return ("tol", (float(a[0]), float(a[1])))
        if name == "Tolerance::relative":
            return ("tol", (float(a[0]), 1e-5))
        if name == "Tolerance::absolute":
            return ("tol", (5e-3, float(a[0])))
        raise Unsupported(f"call {name}")

    def method(self, recv, name, a, line):
        if recv is None and name in ("sync", "unwrap", "clone"):
            return None              # device.sync().unwrap(): a stream flush, no value
        if isinstance(recv, tuple) and recv and recv[0] == "tol":
            return recv              # .set_half_precision_* etc: irrelevant for f32
        if isinstance(recv, Data):
            if name in ("assert_eq", "assert_approx_eq"):
                other = a[0]
                tol = "exact"
                if name == "assert_approx_eq":
                    t = a[1]
                    if not (isinstance(t, tuple) and t[0] == "tol"):
                        raise Unsupported("tolerance expr")
                    tol = list(t[1])
                if not isinstance(other, Data):
                    raise Unsupported("assert against non-data")
                actual, expected = (recv, other) if recv.node is not None else (other, recv)
                if actual.node is None or expected.node is not None:
                    raise Unsupported("assert between two literals / two computed")
                self.cases.append({"line": line, "expr": actual.node, "expected": expected.lit, "tol": tol})
                return None
            if name in ("clone",):
                return recv
            if name == "convert":
                return recv
            raise Unsupported(f"TensorData.{name}")
        if isinstance(recv, list) and name in ("clone", "into", "to_vec"):
            return recv
        if not isinstance(recv, dict):
            raise Unsupported(f"method {name} on {type(recv).__name__}")
        if name in ("clone", "detach", "require_grad", "to_device"):
            return recv
        if name in ("into_data", "to_data"):
            return Data(None, recv)
        if name in ("into_scalar",):
            raise Unsupported("into_scalar")
        if name == "dims" or name == "shape":
            raise Unsupported("shape query")
        if name in VIEW_OPS:
            if name in ("transpose", "t"):
                return sym("transpose", x=recv, args=[])
            return sym(name, x=recv, args=[self.plain(v) for v in a])
        if name in TENSOR_METHODS:
            if name in ("max_dim_with_indices", "min_dim_with_indices"):
                return (sym(name, x=recv, args=[self.plain(a[0])], out=0), sym(name, x=recv, args=[self.plain(a[0])], out=1))
            if name == "square":
                return sym("mul", x=recv, args=[recv])
            if name in ("scatter", "select_assign"):
                upd = a[3] if len(a) > 3 else ("update", "Add")
                if not (isinstance(upd, tuple) and upd[0] == "update" and upd[1] == "Add"):
                    raise Unsupported("indexing update op other than Add")
                return sym({"scatter": "scatter_add", "select_assign": "select_add"}[name], x=recv, args=[self.plain(v) for v in a[:3]])
            if name in ("add", "sub", "mul", "div", "remainder", "powf") and a and isinstance(a[0], (int, float)):
                return sym(name + "_scalar", x=recv, args=[a[0]])
            return sym(name, x=recv, args=[self.plain(v) for v in a])
        raise Unsupported(f"tensor method {name}")

    @staticmethod
    def plain(v):
        if isinstance(v, Data):
            raise Unsupported("TensorData as argument")
        if isinstance(v, tuple) and len(v) == 3 and v[0] == "range" and all(isinstance(e, int) for e in v[1:]):
            return {"range": [v[1], v[2]]}
        if isinstance(v, tuple):
            raise Unsupported(f"argument {v[0] if v else v}")
        if isinstance(v, list):
            return [Interp.plain(e) for e in v]
        return v


# ------------------------------------------------------------------ driver
def split_tests(src: str):
    """Yields (name, first_line, body_tokens) of every #[test] fn."""
    toks = tokenize(src)
    i = 0
    while i < len(toks):
        if toks[i][1] == "#" and i + 3 < len(toks) and toks[i + 1][1] == "[" and toks[i + 2][1] == "test" and toks[i + 3][1] == "]":
            j = i + 4
            should_panic = False
            while toks[j][1] == "#":       # further attributes
                k = j
                while toks[k][1] != "]":
                    if toks[k][1] == "should_panic":
                        should_panic = True
                    k += 1
                j = k + 1
            if toks[j][1] != "fn":
                i += 1
                continue
            name, line = toks[j + 1][1], toks[j + 1][2]
            k = j + 2
            while toks[k][1] != "{":
                k += 1
            depth, start = 0, k
            while True:
                if toks[k][1] == "{":
                    depth += 1
                elif toks[k][1] == "}":
                    depth -= 1
                    if depth == 0:
                        break
                k += 1
            if not should_panic:
                yield name, line, toks[start:k + 1]
            i = k + 1
        else:
            i += 1


def main():
    verbose = "--verbose" in sys.argv
    from tests import golden_expr as G
    oracle_backend = G.OracleBackend()
    cases, skipped_tests, skipped_cases, disagreements = [], [], [], []
    for rel in FILES:
        path = os.path.join(REF, rel)
        if not os.path.exists(path):
            print("missing", rel)
            continue
        src = open(path).read()
        try:
            tests = list(split_tests(src))
        except Unsupported as e:
            print("cannot tokenize", rel, e)
            continue
        for name, line, toks in tests:
            it = Interp(toks, rel, name)
            try:
                it.block()
            except Unsupported as e:
                skipped_tests.append(f"{rel}:{line} {name}: {e}")
                if not it.cases:
                    continue
            except (IndexError, KeyError, TypeError, ValueError) as e:
                skipped_tests.append(f"{rel}:{line} {name}: parser: {type(e).__name__} {e}")
                if not it.cases:
                    continue
            seen = set()
            for k, c in enumerate(it.cases):
                key = json.dumps([c["expr"], c["expected"], c["tol"]], sort_keys=True, default=str)
                if key in seen:
                    continue          # the same assertion from a loop iteration / cloned-vs-inplace pair
                seen.add(key)
                case = {"name": f"{os.path.basename(rel)[:-3]}::{name}#{k}", "cite": f"{rel}:{c['line']}", "expr": c["expr"],
                        "expected": G.encode_lit(c["expected"]), "tol": c["tol"]}
                try:
                    case["expr"] = G.encode_tree(case["expr"])
                    got = G.evaluate(case["expr"], oracle_backend)
                    G.check(case, got)
                except NotImplementedError as e:
                    skipped_cases.append(f"{case['cite']} {case['name']}: oracle lacks {e}")
                    continue
                except AssertionError as e:
                    disagreements.append(f"{case['cite']} {case['name']}: {str(e)[:300]}")
                    continue
                except Exception as e:   # noqa: BLE001
                    skipped_cases.append(f"{case['cite']} {case['name']}: {type(e).__name__}: {str(e)[:200]}")
                    continue
                cases.append(case)
    out = os.path.join(ROOT, "tests", "golden", "burn_backend_tests_expr.json")
    about = ("Assertions of the reference's own backend tests (crates/burn-backend-tests/tests/**), extracted by "
             "scripts/extract_goldens.py: each case is an op tree over literal tensors plus the literal the reference asserts "
             "(cite = file:line of the assertion).  tol 'exact' = TensorData::assert_eq(strict=false); [rel, abs] = "
             "assert_approx_eq with burn's Tolerance |x-y| < max(rel*|x+y|, abs).")
    with open(out, "w") as f:
        f.write('{"_about": %s,\n "cases": [\n' % json.dumps(about))
        f.write(",\n".join("  " + json.dumps(c) for c in cases))
        f.write("\n ]\n}\n")
    print(f"{len(cases)} cases written to {out}")
    print(f"{len(skipped_tests)} tests outside the interpreted subset, {len(skipped_cases)} cases the oracle cannot run, "
          f"{len(disagreements)} ORACLE DISAGREEMENTS")
    for d in disagreements:
        print("  DISAGREE", d)
    if verbose:
        for s in skipped_tests:
            print("  skip-test", s)
        for s in skipped_cases:
            print("  skip-case", s)


if __name__ == "__main__":
    main()
