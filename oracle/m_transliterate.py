"""TEST INFRASTRUCTURE -- mechanical MATLAB -> Python transliteration of the two MATLAB functions on the hot path.

Montecarlo_seq/seq_mcsampling.m (the next-event sampler with its round / ceil discretisation, SURVEY a-8) and
Montecarlo_seq/calnlc.m (number of load curtailments = entries into a loss episode, SURVEY a-9) cannot run in the build
image (no MATLAB / Octave).  As oracle/jl_transliterate.py does for the Julia files, this module reads the reference SOURCE
TEXT at run time and rewrites the two functions line by line into Python with the fixed rules below (no statement
re-ordered, added or dropped), then executes them: the two duration draws `-mttf * log(rand_val)` / `-mttr * log(rand_val)`
(seq_mcsampling.m:52,59) read from per-component duration lists -- the substitution test asserts one hit each, on exactly those
lines --, `rand(1)` itself still runs (a stub) so the statement order is untouched.  scripts/make_reference_golden.py commits
what they produce (tests/golden/ref_matlab.npz); tests/test_reference_pin.py holds the C oracle to it and re-derives it from
the reference text whenever /root/reference exists.  Nothing of the reference is stored in the repository.

Rules (MATLAB subset of the two files):
  function out = f(a, b) ... end            -> def f(a, b): ... return out
  % comment, trailing ;                     -> dropped
  for i = a : b                             -> for i in range(a, (b) + 1):
  while c / if c / elseif c / else / end    -> while c: / if c: / elif c: / else: / (block closed by indentation)
  ~x, true, false                           -> not x, True, False
  v(i), M(i, j), M(i, a:b) for the ARRAY names of the function (its matrix arguments and the matrix it builds)
                                            -> v[i], M[i, j], M[i, m_range(a, b)] on MArr, a 1-based dense array class
  sparse(r, c)                              -> MArr.zeros(r, c)
  diff(v), sum(v), v == 1                   -> element-wise helpers of the prelude (MATLAB semantics)
  round(x)                                  -> MATLAB round: half away from zero (NOT Python's banker's rounding)
Arithmetic is IEEE binary64 in both languages."""
from __future__ import annotations

import math
import re
from typing import Callable, Dict, Iterable, List, Sequence, Tuple

SAMPLING_REL = "Montecarlo_seq/seq_mcsampling.m"
CALNLC_REL = "Montecarlo_seq/calnlc.m"

# the two duration draws of seq_mcsampling and the lines they must sit on
SAMPLING_DRAW_SUBSTITUTIONS: Tuple[Tuple[int, str, str], ...] = (
    (52, "duration = -mttf * log(rand_val);", "duration = PSRA_INJ_next(i);"),
    (59, "duration = -mttr * log(rand_val);", "duration = PSRA_INJ_next(i);"),
)


def apply_substitutions(src: str, subs: Iterable[Tuple[int, str, str]]) -> Tuple[str, List[int]]:
    """Replace each `old` by `new`, requiring exactly one occurrence, and return the 1-based line numbers hit."""
    hit_lines = []
    for line_no, old, new in subs:
        n = src.count(old)
        if n != 1:
            raise ValueError(f"reference text changed: {n} occurrences of `{old}` (expected 1)")
        hit = src.count("\n", 0, src.index(old)) + 1
        if line_no and hit != line_no:
            raise ValueError(f"reference text changed: `{old}` is on line {hit}, expected {line_no}")
        hit_lines.append(hit)
        src = src.replace(old, new)
    return src, hit_lines


def _strip_comment(line: str) -> str:
    in_str = False
    for i, ch in enumerate(line):
        if ch == "'":
            in_str = not in_str
        elif ch == "%" and not in_str:
            return line[:i]
    return line


def _index_arrays(code: str, arrays: Sequence[str]) -> str:
    """name( ... ) -> name[ ... ] for the array names; a:b inside such an index -> m_range(a, b)."""
    for name in arrays:
        out, i = [], 0
        pat = re.compile(rf"(?<![\w.]){re.escape(name)}\(")
        while True:
            m = pat.search(code, i)
            if not m:
                out.append(code[i:])
                break
            out.append(code[i:m.start()] + name + "[")
            depth, j = 1, m.end()
            while depth:
                depth += {"(": 1, ")": -1}.get(code[j], 0)
                j += 1
            inner = code[m.end():j - 1]
            inner = re.sub(r"([\w.]+)\s*:\s*([\w.]+)", r"m_range(\1, \2)", inner)
            out.append(inner + "]")
            i = j
        code = "".join(out)
    return code


def _expr(code: str, arrays: Sequence[str]) -> str:
    code = _index_arrays(code, arrays)
    code = re.sub(r"\bsparse\(", "MArr.zeros(", code)
    code = re.sub(r"~(?!=)", " not ", code)
    code = re.sub(r"\btrue\b", "True", code)
    code = re.sub(r"\bfalse\b", "False", code)
    return code


def transliterate(m_src: str, arrays: Sequence[str]) -> str:
    """MATLAB function file (subset above) -> Python source of the same function."""
    out: List[str] = []
    ret = None
    for line in m_src.split("\n"):
        code = _strip_comment(line).rstrip()
        indent = code[:len(code) - len(code.lstrip())]
        stmt = code.strip().rstrip(";").rstrip()
        if not stmt or stmt == "end":
            continue
        m = re.match(r"^function\s+(\w+)\s*=\s*(\w+)\((.*)\)$", stmt)
        if m:
            ret = m.group(1)
            out.append(f"def {m.group(2)}({m.group(3)}):")
            continue
        m = re.match(r"^for\s+(\w+)\s*=\s*(.+?)\s*:\s*(.+)$", stmt)
        if m:
            out.append(f"{indent}for {m.group(1)} in range({_expr(m.group(2), arrays)}, ({_expr(m.group(3), arrays)}) + 1):")
            continue
        m = re.match(r"^(if|elseif|while)\s+(.+)$", stmt)
        if m:
            kw = "elif" if m.group(1) == "elseif" else m.group(1)
            out.append(f"{indent}{kw} {_expr(m.group(2), arrays)}:")
            continue
        if stmt == "else":
            out.append(indent + "else:")
            continue
        out.append(indent + _expr(stmt, arrays))
    if ret is None:
        raise ValueError("no `function out = name(args)` line in the reference text")
    out.append(f"    return {ret}")
    return "\n".join(out) + "\n"


# ----------------------------------------------------------------------------------------------- prelude
class MArr:
    """MATLAB matrix / row vector of doubles: 1-based, M[i, j], M[i, r] = scalar with a range r, v[i]."""

    def __init__(self, rows: int, cols: int, data=None):
        self.rows, self.cols = int(rows), int(cols)
        self.d = [0.0] * (self.rows * self.cols) if data is None else list(data)

    @staticmethod
    def zeros(rows, cols):
        return MArr(rows, cols)

    @staticmethod
    def row(values):
        v = [float(x) for x in values]
        return MArr(1, len(v), v)

    @staticmethod
    def matrix(rows_of_values):
        rows = [list(map(float, r)) for r in rows_of_values]
        return MArr(len(rows), len(rows[0]) if rows else 0, [x for r in rows for x in r])

    def _lin(self, i, j):
        if not (1 <= i <= self.rows and 1 <= j <= self.cols):
            raise IndexError(f"index ({i}, {j}) out of bounds for a {self.rows} x {self.cols} array")
        return (int(i) - 1) * self.cols + int(j) - 1

    def __getitem__(self, k):
        if isinstance(k, tuple):
            return self.d[self._lin(k[0], k[1])]
        if self.rows != 1:
            raise IndexError("linear index into a matrix is outside the subset")
        return self.d[self._lin(1, k)]

    def __setitem__(self, k, x):
        if isinstance(k, tuple) and isinstance(k[1], range):
            for j in k[1]:
                self.d[self._lin(k[0], j)] = float(x)
        elif isinstance(k, tuple):
            self.d[self._lin(k[0], k[1])] = float(x)
        else:
            self.d[self._lin(1, k)] = float(x)

    def __eq__(self, other):          # element-wise, as in MATLAB
        return MArr(self.rows, self.cols, [1.0 if x == other else 0.0 for x in self.d])

    __hash__ = None

    def tolist(self):
        return [self.d[r * self.cols:(r + 1) * self.cols] for r in range(self.rows)]


def _diff(v: MArr) -> MArr:
    if v.rows != 1:
        raise ValueError("diff of a matrix is outside the subset")
    return MArr(1, max(v.cols - 1, 0), [v.d[i + 1] - v.d[i] for i in range(v.cols - 1)])


def _sum(v: MArr) -> float:
    s = 0.0
    for x in v.d:
        s += x
    return s


def matlab_round(x: float) -> float:
    """MATLAB round: nearest integer, ties away from zero."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def prelude(rand: Callable[[], float]) -> Dict[str, object]:
    return dict(MArr=MArr, m_range=lambda a, b: range(int(a), int(b) + 1), rand=lambda *a: rand(), log=math.log,
                round=matlab_round, ceil=lambda x: float(math.ceil(x)), min=min, diff=_diff, sum=_sum, range=range)


def _load(root: str, rel: str) -> str:
    with open(f"{root}/{rel}", "r", encoding="utf-8") as f:
        return f.read()


def load_seq_mcsampling(root: str = "/root/reference"):
    """Returns (sampler, python source, lines hit): sampler(reliability_data rows [MTTF, MTTR], n_generators, n_lines, years,
    hours_per_year, durations) runs the reference's seq_mcsampling with the k-th duration draw of component i taken from
    durations[i - 1][k] and returns the dense 0 / 1 state matrix as a list of rows (1 = DOWN)."""
    src, hit = apply_substitutions(_load(root, SAMPLING_REL), SAMPLING_DRAW_SUBSTITUTIONS)
    py = transliterate(src, arrays=("reliability_data", "state_duration_matrix"))
    ns: Dict[str, object] = prelude(lambda: 0.5)

    def sampler(reliability_rows, n_generators, n_lines, years, hours_per_year, durations):
        pos = [0] * len(durations)

        def nxt(i):
            k = pos[i - 1]
            pos[i - 1] = k + 1
            return float(durations[i - 1][k])

        ns["PSRA_INJ_next"] = nxt
        out = ns["seq_mcsampling"](MArr.matrix(reliability_rows), int(n_generators), int(n_lines), int(years), int(hours_per_year))
        return out.tolist(), list(pos)

    exec(compile(py, "<seq_mcsampling.m transliterated>", "exec"), ns)
    return sampler, py, hit


def load_calnlc(root: str = "/root/reference"):
    """Returns (calnlc, python source): calnlc(flags) -> number of curtailment events of a 0 / 1 hour series."""
    py = transliterate(_load(root, CALNLC_REL), arrays=("system_status_series",))
    ns: Dict[str, object] = prelude(lambda: 0.5)
    exec(compile(py, "<calnlc.m transliterated>", "exec"), ns)
    return (lambda flags: ns["calnlc"](MArr.row(flags))), py
