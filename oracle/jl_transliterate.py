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
  x[i], a[i, j], x[r], x[end] (1-based)     -> the same on JArr, a 1-based array class; [..] displays -> JArr([..])
  c ? a : b                                 -> (a if c else b)
  &&, ||, !x, true, false, nothing, x -> e, push!(v, x), f!(..)   -> and, or, not x, True, False, None, lambda x: e, v.append(x), f_b(..)
  a .* b, a .- b, x[r] .-= c, x ./= c       -> jl_bmul(a, b), jl_bsub(a, b), jl_bsub_at(x, r, c), jl_bdiv_all(x, c)
  a:b as a value                            -> jl_range(a, b)
  push!(v, (k = x, ...)) (named tuple)      -> v.append(dict(k=x, ...))
  collect(a:s:b), T[] (struct type), x = @sprintf(..)   -> jl_collect_range(a, s, b), JArr([]), x = ""
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
    """Text of the nth `function name(...) ... end`: from its `function` line to the `end` at the same indentation (a
    constructor inside a struct block is indented; the text is returned de-indented)."""
    ms = list(re.finditer(rf"^([ \t]*)function {re.escape(name)}\(", src, flags=re.M))
    if len(ms) <= nth:
        raise ValueError(f"function {name} (#{nth}) not found in the reference text")
    ind = ms[nth].group(1)
    rest = src[ms[nth].start():]
    m = re.search(rf"^{ind}end\s*$", rest, flags=re.M)
    if not m:
        raise ValueError(f"no closing `end` for function {name}")
    text = rest[:m.end()]
    if ind:
        text = "\n".join(l[len(ind):] if l.startswith(ind) else l for l in text.split("\n"))
    return text


def _split_comment(line: str) -> Tuple[str, str]:
    in_str = False
    for i, ch in enumerate(line):
        if ch == '"':
            in_str = not in_str
        elif ch == "#" and not in_str:
            return line[:i], line[i:]
    return line, ""


_RANGE_FOR = re.compile(r"\bin\s+([\w.]+(?:\([^()]*\))?|\([^()]*\)):([\w.]+(?:\([^()]*\))?|\([^()]*\))")
_RANGE_VAL = re.compile(r"(?<![\w\])])(\b[\w.]+|\([^()]*\)):(\([^()]*\)|[\w.]+(?:\([^()]*\))?)")


def _wrap_array_displays(code: str) -> str:
    """`[a, b]` / `[f(i) for i in r]` that is NOT an index (not preceded by a name, `]` or `)`) -> JArr([...])."""
    out = []
    stack = []                       # True for array displays, False for index brackets
    prev = ""
    for ch in code:
        if ch == "[":
            is_display = not (prev.isalnum() or prev in "_])")
            stack.append(is_display)
            out.append("JArr([" if is_display else "[")
        elif ch == "]" and stack:
            out.append("])" if stack.pop() else "]")
        else:
            out.append(ch)
        if not ch.isspace():
            prev = ch
    return "".join(out)


def _expr(code: str) -> str:
    code = re.sub(r"\b(?:Float64|Int)\[", "[", code)                           # typed array display
    code = re.sub(r"(?<![\w.])[A-Z][A-Za-z0-9]*\[\]", "[]", code)               # empty array of a struct type
    code = re.sub(r"=\s*@sprintf\(.*\)\s*$", '= ""', code)                      # strings that only feed printing
    code = re.sub(r"collect\(\s*([\w.]+)\s*:\s*([\w.]+)\s*:\s*([\w.]+)\s*\)", r"jl_collect_range(\1, \2, \3)", code)
    code = re.sub(r"push!\(\s*([\w.]+)\s*,\s*\(\s*(\w+)\s*=(?!=)", r"\1.append(dict(\2=", code)     # named tuple -> dict
    code = re.sub(r"push!\(\s*([\w.]+)\s*,\s*", r"\1.append(", code)
    code = re.sub(r"\b(\w+)!\(", r"\1_b(", code)                                # fill!(..) -> fill_b(..), popfirst! ...
    code = code.replace("&&", " and ").replace("||", " or ")
    code = re.sub(r"===\s*nothing", " is None", code)
    code = re.sub(r"(?<![\w=!<>])!(?!=)", " not ", code)
    code = re.sub(r"\btrue\b", "True", code)
    code = re.sub(r"\bfalse\b", "False", code)
    code = re.sub(r"\bnothing\b", "None", code)
    code = re.sub(r"\.lambda\b", ".lambda_", code)
    code = re.sub(r"\b(\w+)\s*->\s*", r"lambda \1: ", code)
    code = re.sub(r"([\w.]+)\[end\]", r"\1[length(\1)]", code)
    # broadcasting
    code = re.sub(r"^([\w.]+)\[(.+)\]\s*\.-=\s*(.+)$", r"jl_bsub_at(\1, \2, \3)", code)
    code = re.sub(r"^([\w.]+)\s*\./=\s*(.+)$", r"jl_bdiv_all(\1, \2)", code)
    code = re.sub(r"([\w.]+(?:\[[^\[\]]*\])?(?:\([^()]*\))?)\s*\.\*\s*([\w.]+(?:\[[^\[\]]*\])?(?:\([^()]*\))?)", r"jl_bmul(\1, \2)", code)
    code = re.sub(r"([\w.]+(?:\[[^\[\]]*\])?)\s*\.-\s*([\w.]+(?:\[[^\[\]]*\])?)", r"jl_bsub(\1, \2)", code)
    code = re.sub(r"for\s+\((\w+)\s*,\s*(\w+)\)\s+in\s+enumerate\(([^()]+)\)", r"for (\1, \2) in enumerate(\3, 1)", code)
    code = _RANGE_FOR.sub(lambda m: f"in range({m.group(1)}, ({m.group(2)}) + 1)", code)
    # a range used as a value: `window = a:(b)`, `x[a:(b)]`
    if not re.match(r"^\s*(for|if|elif|while)\b", code):
        code = _RANGE_VAL.sub(lambda m: f"jl_range({m.group(1)}, {m.group(2)})", code)
    code = _wrap_array_displays(code)
    # ternary (one per statement in the subset)
    m = re.match(r"^(\s*(?:return\s+|[\w.\[\]() -]+\s*=\s*))(.+?)\s\?\s(.+?)\s:\s(.+)$", code)
    if m and "lambda" not in m.group(1):
        code = f"{m.group(1)}(({m.group(3)}) if ({m.group(2)}) else ({m.group(4)}))"
    return code


def transliterate(jl: str) -> str:
    """Julia function text (subset above) -> Python source of the same function."""
    # join continuation lines (a statement whose code part ends with a binary operator, an opening bracket or a comma)
    raw = jl.split("\n")
    joined: List[str] = []
    for line in raw:
        code, _ = _split_comment(line)
        if joined and re.search(r"(?:[+\-*/,(=]|&&|\|\|)\s*$", _split_comment(joined[-1])[0]) and code.strip():
            joined[-1] = _split_comment(joined[-1])[0].rstrip() + " " + code.strip()
        elif joined and code.strip().startswith(")") and _split_comment(joined[-1])[0].count("(") > _split_comment(joined[-1])[0].count(")"):
            joined[-1] = _split_comment(joined[-1])[0].rstrip() + code.strip()
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
        m = re.match(r"^function\s+(\w+)(!?)\((.*)\)\s*$", stmt)
        if m:
            args = re.sub(r"::[\w.]+(\{[^{}]*\})?", "", m.group(3)).replace(";", ",")
            out.append(f"{indent}def {m.group(1)}{'_b' if m.group(2) else ''}({_expr(args)}):")
            continue
        if re.match(r"^(println|@printf|print)\b", stmt):
            out.append(indent + "pass")
            continue
        m = re.match(r"^if\s+(.+?)\s+(break|continue)\s+end$", stmt)   # `if c break end`
        if m:
            out.append(f"{indent}if {_expr(m.group(1))}: {m.group(2)}")
            continue
        m = re.match(r"^if\s+(.+?);\s*(.+?);\s*end$", stmt)           # one-line if
        if m:
            out.append(f"{indent}if {_expr(m.group(1))}: {_expr(m.group(2))}")
            continue
        m = re.match(r"^for\s+(.+?);\s*(.+?);\s*end$", stmt)          # one-line for
        if m:
            out.append(f"{indent}{_expr('for ' + m.group(1))}: {_expr(m.group(2))}")
            continue
        m = re.match(r"^(if|elseif|while|for)\s+(.+)$", stmt)
        if m:
            kw = "elif" if m.group(1) == "elseif" else m.group(1)
            out.append(f"{indent}{_expr(kw + ' ' + m.group(2))}:")
            continue
        if stmt == "else":
            out.append(indent + "else:")
            continue
        if re.match(r"^new\(", stmt):                                # inner constructor: the value of the block
            out.append(indent + "return " + _expr(stmt))
            continue
        # several statements on one line
        parts = [p.strip() for p in stmt.split(";") if p.strip()]
        out.append(indent + "; ".join(_expr(p) for p in parts))
    return "\n".join(out) + "\n"


# ----------------------------------------------------------------------------------------------- prelude
class JArr:
    """Julia Vector / Matrix: 1-based, `a[i]`, `a[i, j]`, `a[r]` with a range r (a copy, as in Julia)."""

    def __init__(self, items=()):
        self.v = list(items)

    def __len__(self):
        return len(self.v)

    def __iter__(self):
        return iter(self.v)

    def __getitem__(self, k):
        if isinstance(k, tuple):
            return self.v[k[0] - 1].v[k[1] - 1]
        if isinstance(k, range):
            return JArr(self.v[i - 1] for i in k)
        return self.v[k - 1]

    def __setitem__(self, k, x):
        if isinstance(k, tuple):
            self.v[k[0] - 1].v[k[1] - 1] = x
        else:
            self.v[k - 1] = x

    def append(self, x):
        self.v.append(x)

    def __eq__(self, o):
        return list(self) == list(o)

    def tolist(self):
        return [x.tolist() if isinstance(x, JArr) else x for x in self.v]


def jarr(x):
    """numpy / list (1-D or 2-D) -> JArr of Python floats."""
    try:
        return JArr(jarr(r) for r in x) if hasattr(x[0], "__len__") else JArr(float(t) for t in x)
    except (IndexError, TypeError):
        return JArr(x)


class _Struct:
    _fields: Tuple[str, ...] = ()

    def __init__(self, *args):
        if len(args) != len(self._fields):
            raise TypeError(f"{type(self).__name__} takes {len(self._fields)} fields, got {len(args)}")
        for k, v in zip(self._fields, args):
            setattr(self, k, v)


def _zeros(*a):
    dims = [int(x) for x in a if not isinstance(x, type)]
    zero = 0 if (a and a[0] is int) else 0.0
    if len(dims) == 2:
        return JArr(JArr([zero] * dims[1]) for _ in range(dims[0]))
    return JArr([zero] * dims[0])


def _copy(x):
    return JArr(_copy(t) if isinstance(t, JArr) else t for t in x)


def _findfirst(f, v):
    for i, x in enumerate(v, 1):
        if f(x):
            return i
    return None


def _bsub_at(x, r, c):
    for i in r:
        x[i] = x[i] - c


def _bdiv_all(x, c):
    for i in range(1, len(x) + 1):
        x[i] = x[i] / c


def _sum(x, *rest):
    s = None
    for v in x:
        s = v if s is None else s + v
    return 0.0 if s is None else s


def _cumsum(v):
    out, s = [], 0.0
    for i, x in enumerate(v):
        s = x if i == 0 else s + x
        out.append(s)
    return JArr(out)


def base_prelude(rand: Callable[[], float], randn: Callable[[], float] = None) -> Dict[str, object]:
    """Julia Base functions of the subset (arrays are JArr, 1-based)."""
    return dict(
        JArr=JArr, rand=rand, randn=randn, log=math.log, time=_time.time, length=len, isempty=lambda v: len(v) == 0,
        maximum=max, minimum=min, max=max, min=min, zeros=_zeros, trues=lambda n: JArr([True] * int(n)),
        fill=lambda v, n: JArr([v] * int(n)), fill_b=lambda x, v: [x.__setitem__(i, v) for i in range(1, len(x) + 1)] and None,
        Int=int, Float64=float, ceil=math.ceil, floor=math.floor, round=round, abs=abs, div=lambda a, b: int(a) // int(b),
        Inf=float("inf"), exp=math.exp, sum=_sum, cumsum=_cumsum, reverse=lambda v: JArr(reversed(list(v))), copy=_copy,
        findfirst=_findfirst, isnothing=lambda x: x is None, popfirst_b=lambda q: q.v.pop(0), all=lambda f, v: all(f(x) for x in v),
        sort=lambda x, by=None, rev=False: JArr(sorted(x, key=by, reverse=rev)),
        jl_bmul=lambda a, b: JArr(x * y for x, y in zip(a, b)), jl_bsub=lambda a, b: JArr(x - y for x, y in zip(a, b)),
        jl_bsub_at=_bsub_at, jl_bdiv_all=_bdiv_all, jl_range=lambda a, b: range(int(a), int(b) + 1),
        enumerate=enumerate, range=range, dict=dict, mod=lambda a, b: math.fmod(a, b) if (a >= 0) == (b >= 0) else math.fmod(math.fmod(a, b) + b, b),
        jl_collect_range=lambda a, st, b: JArr(a + i * st for i in range(int(math.floor((b - a) / st + 1e-12)) + 1)),
    )


def prelude(rand: Callable[[], float]) -> Dict[str, object]:
    """base_prelude + the structs of PowerSystemAdequacy.jl (field lists as at PSA.jl:20-58; the derived Generator
    fields come from the transliterated outer constructor)."""

    class Generator(_Struct):
        _fields = ("id", "capacity", "mttf", "mttr", "lambda_", "mu", "for_rate")

    class LoadModel(_Struct):
        _fields = ("hourly_load", "peak_load")

    class ReliabilityResult(_Struct):
        _fields = ("method", "lole_hours_yr", "eue_mwh_yr", "computation_time", "convergence_history")

    class COPT(_Struct):
        _fields = ("capacity_outage", "probability")

    env = base_prelude(rand)
    env.update(Generator=Generator, LoadModel=LoadModel, ReliabilityResult=ReliabilityResult, COPT=COPT)
    return env


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
    gens = JArr(make(i + 1, float(cap[i]), float(mttf[i]), float(mttr[i])) for i in range(len(cap)))
    lm = env["LoadModel"](jarr(load), max(float(x) for x in load))
    res = env["run_sequential_mc"](gens, lm, int(years))
    res.convergence_history = list(res.convergence_history)
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
    gens = JArr(make(i + 1, float(cap[i]), float(mttf[i]), float(mttr[i])) for i in range(len(cap)))
    lm = env["LoadModel"](jarr(load), max(float(x) for x in load))
    res = env["run_non_sequential_mc"](gens, lm, int(iterations))
    res.convergence_history = list(res.convergence_history)
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
    gens = JArr(make(i + 1, float(cap[i]), float(mttf[i]), float(mttr[i])) for i in range(len(cap)))
    lm = env["LoadModel"](jarr(load), max(float(x) for x in load))
    res = env["run_analytical"](gens, lm, step_size=float(step_size))
    return res, list(tables[-1].probability), [g.for_rate for g in gens]


# ------------------------------------------------------------------ the "next" rows: other files of GeneratingAdequacy/
MULTI_AREA_REL = "GeneratingAdequacy/AdequacyAssessmentII.jl"
TAIL_RISK_REL = "GeneratingAdequacy/tail_risk.jl"
COMPREHENSIVE_REL = "GeneratingAdequacy/generating_adequacy_comprehensive.jl"

# run_fast_sequential_simulation draws its two durations at AdequacyAssessmentII.jl:210,213 (and the initial one in the
# Generator constructor, :25); injected per generator exactly like the three draws of run_sequential_mc
MULTI_AREA_DRAW_SUBSTITUTIONS: Tuple[Tuple[int, str, str], ...] = (
    (210, "g.time_to_transition += -log(rand()) * g.mttr", "g.time_to_transition += PSRA_INJ_next(g)"),
    (213, "g.time_to_transition += -log(rand()) * g.mttf", "g.time_to_transition += PSRA_INJ_next(g)"),
)


def load_text(ref_root: str, rel: str) -> str:
    with open(f"{ref_root}/{rel}", "r", encoding="utf-8") as f:
        return f.read()


def _multi_area_env(rand):
    class Generator(_Struct):
        _fields = ("id", "capacity", "mttf", "mttr", "current_state", "time_to_transition")

    class Area(_Struct):
        _fields = ("id", "name", "generators", "hourly_load")

    class System(_Struct):
        _fields = ("areas", "tie_lines", "topology_matrix")

    env = base_prelude(rand)
    env.update(Generator=Generator, Area=Area, System=System, ISOLATED=0, INTERCONNECTED=1)
    return env


def reference_solve_curtailment(src: str, topology, margins, policy: int):
    """solve_curtailment_fast (AdequacyAssessmentII.jl:73-179), transliterated, text unchanged."""
    env = _multi_area_env(lambda: 0.0)
    compile_functions(src, [("solve_curtailment_fast", 0)], env)
    sys_ = env["System"](JArr([]), JArr([]), jarr(topology))
    return list(env["solve_curtailment_fast"](sys_, jarr(margins), int(policy)))


def reference_multi_area_year(src: str, unit_area, cap, mttf, mttr, loads, topology, policy: int, durations):
    """run_fast_sequential_simulation (AdequacyAssessmentII.jl:185-250, transliterated) for ONE year of all-up generators
    whose durations come from per-unit lists (durations[u][0] = the initial time to failure of :25).  Returns
    (hours with curtailment per area, curtailed energy per area, substitution lines hit)."""
    patched, hit = apply_substitutions(src, MULTI_AREA_DRAW_SUBSTITUTIONS)

    def no_rand():
        raise RuntimeError("run_fast_sequential_simulation consumed rand() although its draws are injected")

    env = _multi_area_env(no_rand)
    used = {}

    def nxt(g):
        u = int(g.id)
        used[u] = used.get(u, 0) + 1
        return float(durations[u][used[u]])

    env["PSRA_INJ_next"] = nxt
    compile_functions(patched, [("solve_curtailment_fast", 0), ("run_fast_sequential_simulation", 0)], env)
    n_areas = len(loads)
    areas = []
    for a in range(n_areas):
        gens = JArr(env["Generator"](u, float(cap[u]), float(mttf[u]), float(mttr[u]), True, float(durations[u][0]))
                    for u in range(len(cap)) if int(unit_area[u]) == a)
        areas.append(env["Area"](a + 1, f"area{a + 1}", gens, jarr(loads[a])))
    sys_ = env["System"](JArr(areas), JArr([]), jarr(topology))
    res = env["run_fast_sequential_simulation"](sys_, int(policy), 1)
    return [r["lole"] for r in res], [r["eue"] for r in res], hit


def _detailed_env(rand, randn):
    class Generator(_Struct):
        _fields = ("name", "capacity", "for_rate", "maintenance_weeks", "energy_limit", "effective_q", "scheduled_outage_start",
                   "history_q")

    class COPT(_Struct):
        _fields = ("capacity_outage", "probability")

    env = base_prelude(rand, randn)
    env.update(Generator=Generator, COPT=COPT)
    return env


def reference_schedule_maintenance(src_comprehensive: str, cap, maintenance_weeks, weekly_peaks):
    """schedule_maintenance! (generating_adequacy_comprehensive.jl:86-112), transliterated; returns the start weeks."""
    env = _detailed_env(lambda: 0.0, lambda: 0.0)
    compile_functions(src_comprehensive, [("schedule_maintenance!", 0)], env)
    gens = JArr(env["Generator"](f"g{i}", float(cap[i]), 0.0, int(maintenance_weeks[i]), float("inf"), 0.0, 0, JArr([]))
                for i in range(len(cap)))
    env["schedule_maintenance_b"](gens, jarr(weekly_peaks))
    return [int(g.scheduled_outage_start) for g in gens]


def reference_detailed_mc(src_tail: str, cap, for_rate, maint_start, maint_weeks, energy_limit, base_load, lfu_sigma_percent,
                          n_years: int, unif, norm):
    """run_detailed_mc (tail_risk.jl:12-91), transliterated, text unchanged.  unif[year][hour][unit] / norm[year][hour] are
    replayed in the reference's own consumption order: rand() is only called for a unit that is NOT on maintenance in that
    hour (tail_risk.jl:39-44), randn() once per hour (:60)."""
    U, H = len(cap), len(base_load)

    def uniform_stream():
        for y in range(n_years):
            for h in range(1, H + 1):
                week = (h - 1) // 168 + 1
                for u in range(U):
                    if maint_start[u] <= week < maint_start[u] + maint_weeks[u]:
                        continue
                    yield float(unif[y][h - 1][u])

    us = uniform_stream()
    ns = (float(norm[y][h]) for y in range(n_years) for h in range(H))
    env = _detailed_env(lambda: next(us), lambda: next(ns))
    compile_functions(src_tail, [("run_detailed_mc", 0)], env)
    gens = JArr(env["Generator"](f"g{i}", float(cap[i]), float(for_rate[i]), int(maint_weeks[i]), float(energy_limit[i]),
                                 float(for_rate[i]), int(maint_start[i]), JArr([])) for i in range(U))
    yl, hf = env["run_detailed_mc"](gens, jarr(base_load), float(lfu_sigma_percent), int(n_years))
    return list(yl), list(hf)


# ---- frequency & duration recursion (generating_adequacy_frequency.jl) and the stand-alone COPT demo (generating_adequacy_assessment.jl)
FREQUENCY_REL = "GeneratingAdequacy/generating_adequacy_frequency.jl"
ASSESSMENT_REL = "GeneratingAdequacy/generating_adequacy_assessment.jl"


def reference_fd(src_freq: str, cap, mtbf_h, mttr_h, peak_load: float):
    """GeneratorFD constructor (:21-32), add_unit_educational! (:53-147) and evaluate_risk (:155-186), transliterated.
    Returns (outage levels, cumulative probabilities, cumulative frequencies, (LOLE hours, LOLF, LOLD))."""

    class GeneratorFD(_Struct):
        _fields = ("name", "capacity", "mtbf", "mttr", "lambda_", "mu", "p", "q")

    class COPT(_Struct):
        _fields = ("outage_levels", "cum_prob", "cum_freq")

    env = base_prelude(lambda: 0.0)
    env.update(COPT=COPT, new=GeneratorFD)
    compile_functions(src_freq, [("GeneratorFD", 0), ("add_unit_educational!", 0), ("evaluate_risk", 0)], env)
    copt = COPT(JArr([0.0]), JArr([1.0]), JArr([0.0]))                      # run_educational_demo, :195
    total = 0.0
    for i in range(len(cap)):
        u = env["GeneratorFD"](f"Unit {i + 1}", float(cap[i]), float(mtbf_h[i]), float(mttr_h[i]))
        copt = env["add_unit_educational_b"](copt, u)
        total += u.capacity
    return list(copt.outage_levels), list(copt.cum_prob), list(copt.cum_freq), tuple(env["evaluate_risk"](copt, float(peak_load), total))


def reference_gaa(src_assess: str, cap, for_rate, step: float, ldc):
    """add_unit (:30-107) and calculate_indices (:113-146) of generating_adequacy_assessment.jl, transliterated."""

    class Generator(_Struct):
        _fields = ("capacity", "for_rate", "name")

    class COPT(_Struct):
        _fields = ("capacity_outage", "probability")

    env = base_prelude(lambda: 0.0)
    env.update(Generator=Generator, COPT=COPT)
    compile_functions(src_assess, [("add_unit", 0), ("calculate_indices", 0)], env)
    copt = COPT(JArr([0.0]), JArr([1.0]))
    for i in range(len(cap)):
        copt = env["add_unit"](copt, Generator(float(cap[i]), float(for_rate[i]), f"G{i + 1}"), float(step))
    res = env["calculate_indices"](copt, jarr(ldc))
    return list(copt.probability), tuple(res)


# ---- Markov_process.jl is a script: two of its top-level blocks, cut out by their first / last line and executed as they stand
MARKOV_REL = "GeneratingAdequacy/Markov_process.jl"


def extract_block(src: str, first: str, last: str, first_line: int = 0) -> str:
    """The lines from the one that starts with `first` to the next one that is exactly `last` (top-level script code)."""
    lines = src.split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.startswith(first))
    if first_line and i0 + 1 != first_line:
        raise ValueError(f"reference text changed: `{first}` is on line {i0 + 1}, expected {first_line}")
    i1 = next(i for i in range(i0, len(lines)) if lines[i].rstrip() == last)
    return "\n".join(lines[i0:i1 + 1])


def reference_failure_times(src_markov: str, lam: float, dt: float, max_time: float, n: int, uniforms):
    """Markov_process.jl:46-60 (the constant-hazard experiment), rand() replayed from uniforms[component][k]."""
    state = {"i": -1, "k": 0}

    def rand():
        k = state["k"]; state["k"] = k + 1
        return float(uniforms[state["i"]][k])

    env = base_prelude(rand)
    block = extract_block(src_markov, "failure_times = Float64[]", "end", 46)
    code = transliterate(block)
    # the component index is implicit in the reference (one stream); the replay needs it: count the outer loop's iterations
    code = code.replace("    t = 0.0\n", "    PSRA_INJ_component()\n    t = 0.0\n", 1)

    def nxt():
        state["i"] += 1; state["k"] = 0

    env.update({"λ": float(lam), "dt": float(dt), "max_time": max_time, "N_samples": int(n), "PSRA_INJ_component": nxt})
    exec(compile(code, "<transliterated Markov_process.jl:46-60>", "exec"), env)
    return list(env["failure_times"]), code


# Markov_process.jl:83-110 (PART 4): the 2 x 2 transition matrix and its 1 x 2 row-vector product.  Julia's matrix syntax is
# outside the transliterator's subset, so the three matrix expressions of the block are replaced textually (one hit each,
# asserted) by helpers that spell the same products out; everything else is transliterated as it stands.
MARKOV2_SUBSTITUTIONS: Tuple[Tuple[int, str, str], ...] = (
    (94, "P = [p00 p01; \n     p10 p11]", "P = jl_mat2x2(p00, p01, p10, p11)"),
    # lines 103 and 108 of the file; one less here because the two-line matrix literal above has become one line
    (102, "current_state_prob = [1.0 0.0]", "current_state_prob = jl_rowvec(1.0, 0.0)"),
    (107, "current_state_prob = current_state_prob * P", "current_state_prob = jl_row_times_mat(current_state_prob, P)"),
)


def reference_markov2(src_markov: str, mttf: float = None, mttr: float = None, steps: int = None):
    """Markov_process.jl:83-110: probability of the DOWN state after t = 1..steps hourly transitions of a two-state chain that
    starts UP.  λ, μ, dt come from the script's own definitions (:16-22, :42) unless mttf / mttr are given; `steps` overrides
    the script's 200.  Returns (prob_down list, λ, μ, dt)."""
    src, hit = apply_substitutions(src_markov, MARKOV2_SUBSTITUTIONS)
    lines = src.split("\n")
    head = [l for l in lines[:45] if re.match(r"^(MTTF_val|MTTR_val|λ|μ|dt)\s*=", l)]
    if len(head) != 5:
        raise ValueError("reference text changed: MTTF_val / MTTR_val / λ / μ / dt definitions not found in the script head")
    block = extract_block(src, "p01 = 1 - exp", "end", 89)          # ... to the `end` of the evolution loop (:110)
    block = "\n".join(l for l in block.split("\n") if not re.match(r"^\s*(display\(|global )", l))
    env = base_prelude(lambda: 0.5)
    env.update(jl_mat2x2=lambda a, b, c, d: JArr([JArr([a, b]), JArr([c, d])]), jl_rowvec=lambda a, b: JArr([a, b]),
               jl_row_times_mat=lambda v, M: JArr([v[1] * M[1, 1] + v[2] * M[2, 1], v[1] * M[1, 2] + v[2] * M[2, 2]]))
    exec(compile(transliterate("\n".join(head)), "<Markov_process.jl head>", "exec"), env)
    if mttf is not None:
        env["λ"] = 1 / float(mttf); env["μ"] = 1 / float(mttr)
    code = transliterate(block)
    if steps is not None:
        if code.count("steps = 200") != 1:
            raise ValueError("reference text changed: `steps = 200` not found once")
        code = code.replace("steps = 200", f"steps = {int(steps)}")
    exec(compile(code, "<transliterated Markov_process.jl:89-110>", "exec"), env)
    return list(env["prob_down_analytical"]), env["λ"], env["μ"], env["dt"], hit


def reference_dtmc_capacity(src_markov: str, uniforms):
    """Markov_process.jl:153-195 (five generators, hourly two-state DTMC, 1000 hours) with the script's own unit data;
    rand() replayed from uniforms[hour][generator].  Returns (capacity series, mttfs, mttrs, capacities)."""
    it = (float(x) for row in uniforms for x in row)
    env = base_prelude(lambda: next(it))
    block = extract_block(src_markov, "num_gens = 5", "end", 153)
    # the block holds two top-level loops; take everything up to the end of the simulation loop
    lines = src_markov.split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.startswith("num_gens = 5"))
    i1 = next(i for i in range(i0, len(lines)) if lines[i].startswith("    system_available_capacity[t] = current_cap"))
    code = transliterate("\n".join(lines[i0:i1 + 2]))
    exec(compile(code, "<transliterated Markov_process.jl:153-195>", "exec"), env)
    return list(env["system_available_capacity"]), list(env["gen_mttfs"]), list(env["gen_mttrs"]), [float(c) for c in env["gen_capacities"]]


def reference_detailed_analytical(src_tail: str, src_comprehensive: str, cap, for_rate, maint_start, maint_weeks, energy_limit,
                                  base_load, lfu_sigma_percent: float):
    """run_detailed_analytical (tail_risk.jl:96-141) with add_unit / get_lfu_distribution / calculate_expected_generation /
    update_elu! of generating_adequacy_comprehensive.jl, all transliterated, text unchanged.
    Returns (sum of the hourly risk, hourly risk profile, effective FOR per unit, history of the effective FOR per unit)."""
    env = _detailed_env(lambda: 0.0, lambda: 0.0)
    compile_functions(src_comprehensive, [("add_unit", 0), ("get_lfu_distribution", 0), ("calculate_expected_generation", 0),
                                          ("update_elu!", 0)], env)
    compile_functions(src_tail, [("run_detailed_analytical", 0)], env)
    gens = JArr(env["Generator"](f"g{i}", float(cap[i]), float(for_rate[i]), int(maint_weeks[i]), float(energy_limit[i]),
                                 float(for_rate[i]), int(maint_start[i]), JArr([float(for_rate[i])])) for i in range(len(cap)))
    total, profile = env["run_detailed_analytical"](gens, jarr(base_load), float(lfu_sigma_percent))
    return total, list(profile), [g.effective_q for g in gens], [list(g.history_q) for g in gens]
