"""TEST INFRASTRUCTURE -- mechanical Julia -> Python transliteration of the reference's hot-path functions.

The reference (GeneratingAdequacy/PowerSystemAdequacy.jl) cannot run in the build image (no Julia), and its Monte Carlo
draws come from an unseeded global RNG.  To pin the oracle to the reference all the same, this module reads the
reference SOURCE TEXT at run time, cuts the named top-level functions out of it and rewrites them line by line into
Python with a fixed set of syntactic rules (below) -- no statement is re-ordered, added or dropped except `println`.
`rand()` (and, for the injected-duration form, the three `-log(rand())/rate` expressions at PSA.jl:224,243,246, replaced
by per-unit list reads exactly as tools/patched_reference.jl does for a real Julia) is supplied by the caller, so the
transliterated function replays a recorded stream.  scripts/make_reference_golden.py runs it here and commits the
input / output vectors under tests/golden/ref_*.npz; tests/test_reference_pin.py checks the C oracle and the hand-written
transcription oracle/psa_literal.py against those vectors everywhere, and re-derives them from the reference text
whenever /root/reference exists.  Nothing of the reference is stored in the repository.

Rules (Julia subset used by PSA.jl:67-269):
  function f(a::T, b::T; k::T=v) ... end   -> def f(a, b, k=v):          (type annotations dropped)
  for i in a:b / for x in xs / for (i, x) in enumerate(xs)   -> range(a, b + 1) / same / enumerate(xs, 1)
  if / elseif / else / while / end          -> if: / elif: / else: / while: / (block closed by indentation)
  `if c; stmt; end` on one line             -> if c: stmt
  x[i] (1-based)                            -> x[(i) - 1]
  c ? a : b                                 -> (a if c else b)
  &&, ||, true, false, push!(v, x), trues(n), zeros(n), Float64[], T[...]   -> and, or, True, False, v.append(x), ...
  a .* b                                    -> jl_bmul(a, b)
  g.lambda                                  -> g.lambda_               (Python keyword)
  println(...), @printf(...)                -> pass
Arithmetic is IEEE binary64 in both languages; `log` is the platform libm (compare durations, not uniforms, across
machines).  Julia's sum / cumsum of Float64 vectors are pairwise; the prelude's are sequential (differences ~1e-16
relative, far inside the 1e-9 bar of the analytical indices; the Monte Carlo functions do not use them)."""
from __future__ import annotations

import math
import re
import time as _time
from typing import Callable, Dict, Iterable, List, Tuple

PSA_REL = "GeneratingAdequacy/PowerSystemAdequacy.jl"

# the three draws of run_sequential_mc and the line numbers they must sit on
SEQ_DRAW_SUBSTITUTIONS: Tuple[Tuple[int, str, str], ...] = (
    (224, "ttf = [-log(rand())/g.lambda for g in gens]", "ttf = [PSRA_INJ_next(i) for (i, g) in enumerate(gens)]"),
    (243, "ttf[i] += -log(rand())/g.mu", "ttf[i] += PSRA_INJ_next(i)"),
    (246, "ttf[i] += -log(rand())/g.lambda", "ttf[i] += PSRA_INJ_next(i)"),
)
SEQ_RECORD_SUBSTITUTION = ("cum_eue += year_eue", "cum_eue += year_eue; PSRA_INJ_record(year_lole, year_eue)")
NONSEQ_RECORD_SUBSTITUTION = ("cum_eue += iter_eue", "cum_eue += iter_eue; PSRA_INJ_record(iter_lole, iter_eue)")


def apply_substitutions(src: str, subs: Iterable[Tuple[int, str, str]]) -> Tuple[str, List[int]]:
    """Replace each `old` by `new`, requiring exactly one occurrence, and return the 1-based line numbers hit."""
    lines_hit = []
    for line_no, old, new in subs:
        n = src.count(old)
        if n != 1:
            raise ValueError(f"reference text changed: {n} occurrences of `{old}` (expected 1)")
        at = src.index(old)
        hit = src.count("\n", 0, at) + 1
        if line_no and hit != line_no:
            raise ValueError(f"reference text changed: `{old}` is on line {hit}, expected {line_no}")
        lines_hit.append(hit)
        src = src.replace(old, new)
    return src, lines_hit


def extract_function(src: str, name: str, nth: int = 0) -> str:
    """Text of the nth top-level `function name(...) ... end` (column 0 to the matching column-0 `end`)."""
    starts = [m.start() for m in re.finditer(rf"^function {re.escape(name)}\(", src, flags=re.M)]
    if len(starts) <= nth:
        raise ValueError(f"function {name} (#{nth}) not found in the reference text")
    rest = src[starts[nth]:]
    m = re.search(r"^end\s*$", rest, flags=re.M)
    if not m:
        raise ValueError(f"no closing `end` for function {name}")
    return rest[:m.end()]


def _split_comment(line: str) -> Tuple[str, str]:
    in_str = False
    for i, ch in enumerate(line):
        if ch == '"':
            in_str = not in_str
        elif ch == "#" and not in_str:
            return line[:i], line[i:]
    return line, ""


_RANGE = re.compile(r"\bin\s+([\w.]+(?:\([^()]*\))?|\([^()]*\)):([\w.]+(?:\([^()]*\))?|\([^()]*\))")
_INDEX = re.compile(r"(?<![\w\]])([A-Za-z_][\w.]*)\[([^\[\]]+)\]")


def _expr(code: str) -> str:
    code = re.sub(r"\bFloat64\[\]", "[]", code)
    code = re.sub(r"\b(?:Float64|Int)\[", "[", code)
    code = code.replace("&&", " and ").replace("||", " or ")
    code = re.sub(r"\btrue\b", "True", code)
    code = re.sub(r"\bfalse\b", "False", code)
    code = re.sub(r"\.lambda\b", ".lambda_", code)
    code = re.sub(r"push!\(\s*([\w.]+)\s*,\s*(.+)\)\s*$", r"\1.append(\2)", code)
    code = re.sub(r"([\w.]+(?:\([^()]*\))?)\s*\.\*\s*([\w.]+(?:\([^()]*\))?)", r"jl_bmul(\1, \2)", code)
    code = re.sub(r"for\s+\((\w+)\s*,\s*(\w+)\)\s+in\s+enumerate\(([^()]+)\)", r"for (\1, \2) in enumerate(\3, 1)", code)
    code = _RANGE.sub(lambda m: f"in range({m.group(1)}, ({m.group(2)}) + 1)", code)
    # 1-based indexing (comprehensions / array literals start with `[` after an operator or `=`, never after a name)
    code = _INDEX.sub(_index_repl, code)
    # ternary (one per statement in the subset)
    m = re.match(r"^(\s*[\w.\[\]() -]+\s*=\s*)(.+?)\s\?\s(.+?)\s:\s(.+)$", code)
    if m:
        code = f"{m.group(1)}(({m.group(3)}) if ({m.group(2)}) else ({m.group(4)}))"
    return code


def _index_repl(m: "re.Match[str]") -> str:
    inner = m.group(2)
    if inner.endswith(") - 1") and inner.startswith("("):     # already rewritten
        return m.group(0)
    return f"{m.group(1)}[({inner}) - 1]"


def transliterate(jl: str) -> str:
    """Julia function text (subset above) -> Python source of the same function."""
    # join continuation lines (a statement whose code part ends with a binary operator or an opening bracket / comma)
    raw = jl.split("\n")
    joined: List[str] = []
    for line in raw:
        code, _ = _split_comment(line)
        if joined and re.search(r"[+\-*/,(]\s*$", _split_comment(joined[-1])[0]) and code.strip():
            joined[-1] = _split_comment(joined[-1])[0].rstrip() + " " + code.strip()
        else:
            joined.append(line)
    out: List[str] = []
    for line in joined:
        code, comment = _split_comment(line)
        indent = code[:len(code) - len(code.lstrip())]
        stmt = code.strip()
        if not stmt:
            out.append(indent + comment if comment else "")
            continue
        if stmt == "end":
            continue
        m = re.match(r"^function\s+(\w+)\((.*)\)\s*$", stmt)
        if m:
            args = re.sub(r"::[\w.]+(\{[^{}]*\})?", "", m.group(2)).replace(";", ",")
            out.append(f"{indent}def {m.group(1)}({_expr(args)}):")
            continue
        if re.match(r"^(println|@printf|print)\b", stmt):
            out.append(indent + "pass")
            continue
        m = re.match(r"^if\s+(.+?);\s*(.+?);\s*end$", stmt)           # one-line if
        if m:
            out.append(f"{indent}if {_expr(m.group(1))}: {_expr(m.group(2))}")
            continue
        m = re.match(r"^(if|elseif|while|for)\s+(.+)$", stmt)
        if m:
            kw = "elif" if m.group(1) == "elseif" else m.group(1)
            out.append(f"{indent}{_expr(kw + ' ' + m.group(2))}:")
            continue
        if stmt == "else":
            out.append(indent + "else:")
            continue
        # several statements on one line
        parts = [p.strip() for p in stmt.split(";") if p.strip()]
        out.append(indent + "; ".join(_expr(p) for p in parts))
    return "\n".join(out) + "\n"


# ----------------------------------------------------------------------------------------------- prelude
class _Struct:
    _fields: Tuple[str, ...] = ()

    def __init__(self, *args):
        if len(args) != len(self._fields):
            raise TypeError(f"{type(self).__name__} takes {len(self._fields)} fields, got {len(args)}")
        for k, v in zip(self._fields, args):
            setattr(self, k, v)


def prelude(rand: Callable[[], float]) -> Dict[str, object]:
    """Names the transliterated functions may use: Julia Base functions of the subset and the reference's structs
    (field lists as at PSA.jl:20-58; the derived Generator fields come from the transliterated outer constructor)."""

    class Generator(_Struct):
        _fields = ("id", "capacity", "mttf", "mttr", "lambda_", "mu", "for_rate")

    class LoadModel(_Struct):
        _fields = ("hourly_load", "peak_load")

    class ReliabilityResult(_Struct):
        _fields = ("method", "lole_hours_yr", "eue_mwh_yr", "computation_time", "convergence_history")

    class COPT(_Struct):
        _fields = ("capacity_outage", "probability")

    def cumsum(v):
        out, s = [], 0.0
        for i, x in enumerate(v):
            s = x if i == 0 else s + x
            out.append(s)
        return out

    def jl_sum(x, *rest):
        s = None
        for v in x:
            s = v if s is None else s + v
        return 0.0 if s is None else s

    return dict(
        Generator=Generator, LoadModel=LoadModel, ReliabilityResult=ReliabilityResult, COPT=COPT,
        rand=rand, log=math.log, time=_time.time, length=len, isempty=lambda v: len(v) == 0, maximum=max,
        zeros=lambda n: [0.0] * int(n), trues=lambda n: [True] * int(n), Int=int, ceil=math.ceil, floor=math.floor,
        round=round, abs=abs, sum=jl_sum, cumsum=cumsum, reverse=lambda v: list(reversed(v)),
        jl_bmul=lambda a, b: [x * y for x, y in zip(a, b)], enumerate=enumerate, range=range,
    )


def load_reference(ref_root: str = "/root/reference") -> str:
    with open(f"{ref_root}/{PSA_REL}", "r", encoding="utf-8") as f:
        return f.read()


def compile_functions(src: str, names: Iterable[Tuple[str, int]], env: Dict[str, object]) -> Dict[str, str]:
    """Transliterate and exec the named (name, nth) functions of `src` into `env`; returns the Python sources."""
    py = {}
    for name, nth in names:
        text = transliterate(extract_function(src, name, nth))
        py[f"{name}#{nth}"] = text
        exec(compile(text, f"<transliterated {name}>", "exec"), env)
    return py


class Injected:
    """Per-unit duration lists / record of the per-trial sums -- the Python twin of module PSRA_INJ in
    tools/patched_reference.jl."""

    def __init__(self, durations=None):
        self.D = durations                  # D[unit][k]
        self.used = [0] * (len(durations) if durations is not None else 0)
        self.lole: List[float] = []
        self.eue: List[float] = []

    def next(self, i: int) -> float:        # i is 1-based, as in the Julia text
        k = self.used[i - 1]
        self.used[i - 1] = k + 1
        return float(self.D[i - 1][k])

    def record(self, lole: float, eue: float) -> None:
        self.lole.append(lole)
        self.eue.append(eue)


def reference_sequential(src: str, cap, mttf, mttr, load, years: int, durations):
    """The reference's run_sequential_mc (transliterated from `src`) on injected per-unit duration lists.
    Returns (ReliabilityResult, per-year LOL hours, per-year ENS, substitution lines hit)."""
    patched, hit = apply_substitutions(src, SEQ_DRAW_SUBSTITUTIONS)
    patched, _ = apply_substitutions(patched, [(0,) + SEQ_RECORD_SUBSTITUTION])
    inj = Injected(durations)

    def no_rand():
        raise RuntimeError("run_sequential_mc consumed rand() although its three draws are injected")

    env = prelude(no_rand)
    env["PSRA_INJ_next"] = inj.next
    env["PSRA_INJ_record"] = inj.record
    compile_functions(patched, [("Generator", 0), ("run_sequential_mc", 0)], env)
    make = env["Generator"]                                   # now the transliterated outer constructor (4 arguments)
    struct7 = prelude(no_rand)["Generator"]
    env["Generator"] = struct7                                # ... which calls the 7-field struct
    gens = [make(i + 1, float(cap[i]), float(mttf[i]), float(mttr[i])) for i in range(len(cap))]
    lm = env["LoadModel"]([float(x) for x in load], max(float(x) for x in load))
    res = env["run_sequential_mc"](gens, lm, int(years))
    return res, inj.lole, inj.eue, hit


def reference_non_sequential(src: str, cap, mttf, mttr, load, iterations: int, uniforms):
    """The reference's run_non_sequential_mc (transliterated, text unchanged apart from the per-iteration record) with
    rand() replaying `uniforms` (consumed unit by unit, iteration by iteration: PSA.jl:183)."""
    patched, _ = apply_substitutions(src, [(0,) + NONSEQ_RECORD_SUBSTITUTION])
    it = iter(uniforms)
    inj = Injected()
    env = prelude(lambda: float(next(it)))
    env["PSRA_INJ_record"] = inj.record
    compile_functions(patched, [("Generator", 0), ("run_non_sequential_mc", 0)], env)
    make = env["Generator"]
    env["Generator"] = prelude(lambda: 0.0)["Generator"]
    gens = [make(i + 1, float(cap[i]), float(mttf[i]), float(mttr[i])) for i in range(len(cap))]
    lm = env["LoadModel"]([float(x) for x in load], max(float(x) for x in load))
    res = env["run_non_sequential_mc"](gens, lm, int(iterations))
    return res, inj.lole, inj.eue, [g.for_rate for g in gens]


def reference_analytical(src: str, cap, mttf, mttr, load, step_size: float):
    """The reference's add_unit_convolution + run_analytical (transliterated, text unchanged).
    Returns (ReliabilityResult, final COPT probabilities)."""
    env = prelude(lambda: 0.0)
    tables = []
    compile_functions(src, [("Generator", 0), ("add_unit_convolution", 0), ("run_analytical", 0)], env)
    inner = env["add_unit_convolution"]

    def spy(old, unit, step):
        t = inner(old, unit, step)
        tables.append(t)
        return t

    env["add_unit_convolution"] = spy
    make = env["Generator"]
    env["Generator"] = prelude(lambda: 0.0)["Generator"]
    gens = [make(i + 1, float(cap[i]), float(mttf[i]), float(mttr[i])) for i in range(len(cap))]
    lm = env["LoadModel"]([float(x) for x in load], max(float(x) for x in load))
    res = env["run_analytical"](gens, lm, step_size=float(step_size))
    return res, list(tables[-1].probability), [g.for_rate for g in gens]
