"""Plan-specialised fused kernel: the word trie compiled to straight-line CUDA.

The generic kernel of ``csrc/lns.cuh`` interprets the prefix trie at run time
(one lane per node, parents passed through shared memory) and spends most of
its issue slots on that interpretation.  This module *compiles* a slice --
trie, semiring, weighting mode, sieve set -- into CUDA source in which

* one thread owns one series and one *part* of the trie (a few dozen nodes
  plus the ancestors they need); a warp is 32 series of the same part, so
  control flow never diverges;
* every running iterated sum and every sieve accumulator is a named register;
  the only memory traffic of a time step is the read of ``x[t]`` from a
  shared-memory tile that the CTA stages with ``cp.async``;
* letter products are shared between siblings: ``P*x1*x1*x2`` reuses the
  product ``P*x1*x1`` of the sibling letter ``[11]`` -- the reference
  multiplies once per letter occurrence, dimensions ascending
  (fruits/iss/semiring.py:114-120, :142-148), so every partial product of a
  letter is the full product of a shorter letter and the result is bit
  identical while one multiplication per node remains;
* thresholds sit in constant memory and are folded into the compare
  instructions.

The floating point order of every emitted value is the reference's
(fruits/iss/semiring.py:93-158 Reals, :282-338 Arctic): unweighted Reals and
all Arctic results are bit-identical, weighted Reals differ only through
``exp`` (computed on the device).  The source is compiled with NVRTC for
sm_100a by ``libfruits_b200.so`` (``fb_jit_*`` in include/fruits_b200.h) and
cached as a cubin next to the library.
"""
import ctypes
import hashlib
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field


from . import _backend as be

JIT_VERSION = 21            # bump to invalidate cached cubins
CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "jit")

# threshold table columns (include/fruits_b200.h, FB_NTHR)
_COL_U = {0: (0, 1), 1: (2, 3), 2: (4, 5)}
_COL_PPV, _COL_MAX, _COL_MIN = 6, (8, 9), (10, 11)
_COL_CPV, _COL_CUR = 7, (12, 13)


# ---------------------------------------------------------------------------
# program: trie -> parts
# ---------------------------------------------------------------------------

@dataclass
class SieveSet:
    """Which accumulators an emitted node carries (from the feature list)."""
    feats: list                     # [(FEAT_*, arg)]
    cnt: tuple = (False, False, False)     # a unit (increment depth) is used
    avg: tuple = (False, False, False)     # ... and its sum is needed
    ppv: bool = False
    mx: bool = False
    mn: bool = False
    hi: bool = False                # finite upper bounds possible
    mmb: bool = False               # MAX/MIN restricted to (lo, hi]
    xpi: tuple = (False, False, False)     # mean index of the selected increments of a unit
    lpi: tuple = (False, False, False)     # longest run of selected increments of a unit
    cur: bool = False               # sum of the squared second increments in (lo, hi]
    cpv: bool = False               # rising edges of (y >= threshold)
    cut: bool = False               # segment sieves look at [0, cut[series]) only

    @staticmethod
    def make(feats, bounded_hi, bounded_mm, cut: bool = False) -> "SieveSet":
        cnt, avg, xpi, lpi = [False] * 3, [False] * 3, [False] * 3, [False] * 3
        ppv = mx = mn = cur = cpv = False
        for kind, arg in feats:
            if kind in (be.FEAT_CNT, be.FEAT_AVG, be.FEAT_XPI, be.FEAT_LPI):
                cnt[arg] = True                  # the unit (its predicate and count) is on
                if kind == be.FEAT_AVG:
                    avg[arg] = True
                elif kind == be.FEAT_XPI:
                    xpi[arg] = True
                elif kind == be.FEAT_LPI:
                    lpi[arg] = True
            elif kind == be.FEAT_PPV:
                ppv = True
            elif kind == be.FEAT_MAX:
                mx = True
            elif kind == be.FEAT_MIN:
                mn = True
            elif kind == be.FEAT_CUR:
                cur = True
            elif kind == be.FEAT_CPV:
                cpv = True
        return SieveSet(list(feats), tuple(cnt), tuple(avg), ppv, mx, mn,
                        bool(bounded_hi), bool(bounded_mm), tuple(xpi), tuple(lpi), cur, cpv,
                        bool(cut))

    @property
    def rank2(self) -> bool:
        """Accumulators only the thread-per-series kernel knows."""
        return any(self.xpi) or any(self.lpi) or self.cur or self.cpv or self.cut

    @property
    def end(self) -> bool:
        return any(kind == be.FEAT_END for kind, _ in self.feats)

    def thr_cols(self) -> list:
        cols = []
        for k in range(3):
            if self.cnt[k]:
                cols.append(_COL_U[k][0])
                if self.hi:
                    cols.append(_COL_U[k][1])
        if self.ppv:
            cols.append(_COL_PPV)
        if self.mx and self.mmb:
            cols += list(_COL_MAX)
        if self.mn and self.mmb:
            cols += list(_COL_MIN)
        if self.cpv:
            cols.append(_COL_CPV)
        if self.cur:
            cols += list(_COL_CUR)
        return cols

    # registers (32-bit) of sieve state per emitted node
    def regs(self) -> int:
        r = 0
        ncnt = sum(self.cnt) + (1 if self.ppv else 0)
        r += (ncnt + 1) // 2
        r += 2 * sum(self.avg)
        if self.cnt[2] or self.cur:
            r += 2          # previous first increment
        r += 2 * (int(self.mx) + int(self.mn))
        r += sum(self.xpi) + 2 * sum(self.lpi) + (2 if self.cur else 0) + (2 if self.cpv else 0)
        if self.cut and self.end:
            r += 2          # the value at the end of the segment
        return r

    # rough issue slots per emitted node and step
    def cost(self) -> int:
        c = 0
        for k in range(3):
            if self.cnt[k]:
                c += 2 + (1 if self.hi else 0) + (1 if self.avg[k] else 0)
        if self.cnt[1] or self.cnt[2]:
            c += 1
        if self.cnt[2]:
            c += 2
        if self.ppv:
            c += 2
        c += 3 * (int(self.mx) + int(self.mn))
        if self.mmb:
            c += 2 * (int(self.mx) + int(self.mn))
        c += sum(self.xpi) + 3 * sum(self.lpi) + (6 if self.cur else 0) + (5 if self.cpv else 0)
        if self.cut and self.end:
            c += 2
        return c


@dataclass
class Part:
    owned: list = field(default_factory=list)     # trie node ids whose features this part writes
    snodes: list = field(default_factory=list)    # trie node ids with a running sum (owned + ancestors)
    cost: int = 0
    regs: int = 0


def _occurrences(expo, dim_index):
    """Letter -> [(used-dim index, is_division)] in the reference's order."""
    occ = []
    for d, e in enumerate(expo):
        if e != 0:
            occ += [(dim_index[d], e < 0)] * abs(e)
    return tuple(occ)


class Program:
    """Partition of a trie into parts plus everything the emitter needs."""

    def __init__(self, trie, semiring: int, weight_mode: int, sieves: SieveSet,
                 reg_budget: int = 150, parts_multiple: int = 1) -> None:
        self.trie = trie
        self.semiring = semiring
        self.weight_mode = weight_mode
        self.sieves = sieves
        self.reals = semiring == be.SEMIRING_REALS
        # (max, times): the Arctic schedule with products instead of sums
        # (fruits/iss/semiring.py:461-494); its exponential weightings stay on the scan kernel
        self.bayes = semiring == be.SEMIRING_BAYESIAN
        if self.bayes and weight_mode != be.WEIGHT_NONE:
            raise NotImplementedError("weighted Bayesian sums are not generated")
        self.used = trie.used_dims()
        self.dim_index = {d: u for u, d in enumerate(self.used)}
        nodes = trie.nodes
        weighted = weight_mode != be.WEIGHT_NONE
        self.alphas = sorted({n.alpha for n in nodes}) if weighted else []
        self.aidx = {a: i for i, a in enumerate(self.alphas)}
        self.has_children = [bool(n.children) for n in nodes]

        # state registers of one running-sum node (a combination output --
        # virtual node with a ``combo`` term list -- keeps its previous value)
        def sregs(v, owned):
            r = 2
            if weight_mode == be.WEIGHT_NONTOTAL and nodes[v].children:
                r += 2          # second accumulator C_k / arctic carry
            if weight_mode == be.WEIGHT_TOTAL and owned:
                r += 2          # previous output
            return r + (sieves.regs() if owned else 0)

        def scost(v, owned):
            combo = getattr(nodes[v], "combo", None)
            if combo is not None:
                return 1 + sum(1 + t[2] + t[4] for t in combo) + sieves.cost()
            dp = getattr(nodes[v], "dp", None)
            if dp is not None:
                return 2 + len(dp[1])
            c = 2 + (2 if weighted else 0)
            if self.bayes:
                c += 1 + max(0, sum(abs(e) for e in nodes[v].expo) - 1)
            elif not self.reals:
                c += 1 + max(0, sum(1 for e in nodes[v].expo if e != 0) - 1)
            return c + (sieves.cost() if owned else 0)

        self._sregs, self._scost = sregs, scost

        emitted = [v for v in trie.dfs() if nodes[v].emit >= 0]
        # fewest parts the register budget allows, rounded up to whole CTAs;
        # then the smallest cost cap that still needs no more parts than that
        # (greedy filling of the DFS order is optimal for a given cap)
        k = len(self._split(emitted, float("inf"), reg_budget))
        k = -(-k // parts_multiple) * parts_multiple
        lo, hi = 1.0, float(sum(self._scost(v, True) for v in emitted)) + 1.0
        while hi - lo > 0.5:
            mid = (lo + hi) / 2
            if len(self._split(emitted, mid, reg_budget)) <= k:
                hi = mid
            else:
                lo = mid
        parts = self._split(emitted, hi, reg_budget)
        # work of the duplicated ancestors relative to the work of the trie itself
        ideal = sum(self._scost(v, nodes[v].emit >= 0) for v in range(len(nodes)))
        self.overhead = sum(pt.cost for pt in parts) / max(ideal, 1)
        self.max_regs = max(pt.regs for pt in parts)
        while len(parts) % parts_multiple:
            parts.append(Part())
        self.parts = parts

    def _chain(self, v):
        """Nodes ``v`` needs, parents first, ``v`` last."""
        combo = getattr(self.trie.nodes[v], "combo", None)
        dp = getattr(self.trie.nodes[v], "dp", None)
        if combo is not None or dp is not None:
            # several predecessors: the terms of a combination / the states of
            # the previous level of a separable recurrence (iss/cos.py)
            memo = self.__dict__.setdefault("_chain_memo", {})
            if v in memo:
                return memo[v]
            preds = [term[1] for term in combo] if combo is not None else \
                [u for u, _ in dp[1] if u >= 0]
            seen, out = set(), []
            for u in preds:
                for a in self._chain(u):
                    if a not in seen:
                        seen.add(a)
                        out.append(a)
            memo[v] = out + [v]
            return memo[v]
        out = []
        while v >= 0:
            out.append(v)
            v = self.trie.nodes[v].parent
        return out[::-1]

    def _split(self, emitted, cap, reg_budget):
        parts, cur, have = [], Part(), set()
        for v in emitted:
            need = [a for a in self._chain(v) if a not in have]
            add_regs = sum(self._sregs(a, a == v) for a in need)
            add_cost = sum(self._scost(a, a == v) for a in need)
            if cur.owned and (cur.regs + add_regs > reg_budget or cur.cost + add_cost > cap):
                parts.append(cur)
                cur, have = Part(), set()
                need = self._chain(v)
                add_regs = sum(self._sregs(a, a == v) for a in need)
                add_cost = sum(self._scost(a, a == v) for a in need)
            for a in need:
                cur.snodes.append(a)
                have.add(a)
            cur.owned.append(v)
            cur.regs += add_regs
            cur.cost += add_cost
        if cur.owned:
            parts.append(cur)
        return parts


# ---------------------------------------------------------------------------
# emitter
# ---------------------------------------------------------------------------

def _dlit(x: float) -> str:
    """C++ literal of a double (exact)."""
    if x != x:
        return "__longlong_as_double(0x7ff8000000000000LL)"
    if x == float("inf"):
        return "D_INF"
    if x == float("-inf"):
        return "D_NINF"
    return float(x).hex()


class Emitter:
    """CUDA source of one slice program."""

    def __init__(self, prog: Program, dims: list, ppc: int, gpc: int, shared_extra: bool,
                 tt: int = 8, n_shared_rows: int = 0, stage: int = 1, abi: int = 3) -> None:
        """dims[u] = (raw_dim, inc) of used dimension u; ppc/gpc = parts and
        series groups per CTA; shared_extra: the weighting rows are the same
        for every series (Indices)."""
        self.p = prog
        self.ppc, self.gpc = ppc, gpc
        self.tt = tt                 # time steps per shared-memory tile (even)
        self.stage = stage
        self.abi = abi
        self.shared_extra = shared_extra
        self.sv = prog.sieves
        self.cols = self.sv.thr_cols()
        self.ntc = len(self.cols)
        self.colpos = {c: i for i, c in enumerate(self.cols)}
        # dims[u] = (raw_dim, inc) reads the series; ("row", r) reads the shared
        # extra row r (e.g. the sin / cos rows of the cosine weighted ISS)
        self.dims = [d if d[0] != "row" else None for d in dims]
        self.row_dims = {u: d[1] for u, d in enumerate(dims) if d[0] == "row"}
        dims = [d for d in dims if d[0] != "row"]
        # distinct raw rows staged per series
        self.raw_rows = sorted({r for r, _ in dims}) or [0]
        self.row_of = {r: i for i, r in enumerate(self.raw_rows)}
        self.nrow = len(self.raw_rows)
        # extra (weighting) rows: Reals: ep_a, em_a per alpha; Arctic: g
        wm = prog.weight_mode
        if n_shared_rows:
            assert wm == be.WEIGHT_NONE and shared_extra
            self.nextra = n_shared_rows
        elif wm == be.WEIGHT_NONE:
            self.nextra = 0
        elif prog.reals:
            self.nextra = 2 * len(prog.alphas)
        else:
            self.nextra = 1

    # -- names ---------------------------------------------------------------
    def th(self, emit: int, col: int) -> str:
        return f"TH[{emit * self.ntc + self.colpos[col]}]"

    # -- one step of one part ------------------------------------------------
    def step(self, part: Part) -> list:
        p, nodes = self.p, self.p.trie.nodes
        L = []
        sidx = {v: i for i, v in enumerate(part.snodes)}
        owned = set(part.owned)
        oidx = {v: i for i, v in enumerate(part.owned)}
        wm = p.weight_mode
        inset = set(part.snodes)
        if p.reals:
            # children of every parent that live in this part
            groups = {}
            for v in part.snodes:
                if getattr(nodes[v], "combo", None) is None and getattr(nodes[v], "dp", None) is None:
                    groups.setdefault(nodes[v].parent, []).append(v)
            self._step_dp(L, part, sidx)
            # deepest parents first: a node is updated after its children read it
            order = sorted(groups, key=lambda u: -(nodes[u].depth if u >= 0 else 0))
            for u in order:
                if u < 0:
                    base = None
                elif wm == be.WEIGHT_NONE:
                    base = f"S[{sidx[u]}]"
                else:
                    a = p.aidx[nodes[u].alpha]
                    acc = f"S[{sidx[u]}]" if wm == be.WEIGHT_TOTAL else f"A2[{sidx[u]}]"
                    L.append(f"const double b{u} = __dmul_rn({acc}, em{a});")
                    base = f"b{u}"
                # product tree over the occurrences of the children's letters
                prods = {}       # occ prefix -> variable

                def prod(occ):
                    if occ in prods:
                        return prods[occ]
                    (d, div) = occ[-1]
                    x = f"x{d}"
                    if len(occ) == 1:
                        if base is None:
                            # 1.0 * x is x; 1.0 / x needs the division
                            name = x if not div else None
                            if name is None:
                                name = f"v{u if u >= 0 else 'r'}_{len(prods)}"
                                L.append(f"const double {name} = __ddiv_rn(1.0, {x});")
                        else:
                            name = f"v{u}_{len(prods)}"
                            op = "__ddiv_rn" if div else "__dmul_rn"
                            L.append(f"const double {name} = {op}({base}, {x});")
                    else:
                        src = prod(occ[:-1])
                        name = f"v{u if u >= 0 else 'r'}_{len(prods)}"
                        op = "__ddiv_rn" if div else "__dmul_rn"
                        L.append(f"const double {name} = {op}({src}, {x});")
                    prods[occ] = name
                    return name

                kids = sorted(groups[u], key=lambda v: _occurrences(nodes[v].expo, p.dim_index))
                vals = {}
                for v in kids:
                    occ = _occurrences(nodes[v].expo, p.dim_index)
                    if not occ:
                        # empty letter: the product is the base itself
                        vals[v] = base if base is not None else "1.0"
                    else:
                        vals[v] = prod(occ)
                for v in kids:
                    self._update_reals(L, v, vals[v], sidx[v], v in owned, oidx.get(v))
            # combination outputs: sum of coeff * (running sum [* trailing factors])
            # in term order (fruits/iss/cos.py:43-48); the previous value is kept in S
            for v in part.snodes:
                combo = getattr(nodes[v], "combo", None)
                if combo is None:
                    continue
                L.append(f"double y{v} = 0.0;")
                for ti, (coeff, node, sp, su, cp, cu) in enumerate(combo):
                    term = f"S[{sidx[node]}]"
                    if coeff is None:
                        # separable form: the coefficient is the shared row ``su``
                        L.append(f"y{v} = fma(x{p.dim_index[su]}, {term}, y{v});")
                        continue
                    if sp or cp:
                        L.append(f"double z{v}_{ti} = {term};")
                        for _ in range(sp):
                            L.append(f"z{v}_{ti} = __dmul_rn(z{v}_{ti}, x{p.dim_index[su]});")
                        for _ in range(cp):
                            L.append(f"z{v}_{ti} = __dmul_rn(z{v}_{ti}, x{p.dim_index[cu]});")
                        term = f"z{v}_{ti}"
                    L.append(f"y{v} = fma({float(coeff)!r}, {term}, y{v});")
                si = sidx[v]
                L.append(f"const double q{v} = S[{si}]; S[{si}] = y{v};")
                if v in owned:
                    self._sieve(L, v, f"S[{si}]", f"q{v}", oidx[v])
        else:
            # arctic: parents first, children read the parent's new value
            for v in part.snodes:
                u = nodes[v].parent
                if u < 0:
                    base = "0.0"
                elif wm == be.WEIGHT_NONE:
                    base = f"S[{sidx[u]}]"
                elif wm == be.WEIGHT_TOTAL:
                    base = f"OP[{sidx[u]}]"
                else:
                    base = f"A2[{sidx[u]}]"
                expr = base
                if p.bayes:
                    # one multiplication / division per letter occurrence, dimensions
                    # ascending (semiring.py:476-482); 1.0 * x is x
                    expr = f"S[{sidx[u]}]" if u >= 0 else None
                    for d, e in enumerate(nodes[v].expo):
                        x = f"x{p.dim_index[d]}"
                        for _ in range(abs(e)):
                            if expr is None:
                                expr = x if e > 0 else f"__ddiv_rn(1.0, {x})"
                            else:
                                expr = f"{'__dmul_rn' if e > 0 else '__ddiv_rn'}({expr}, {x})"
                    expr = "1.0" if expr is None else expr
                else:
                    for d, e in enumerate(nodes[v].expo):
                        if e != 0:
                            expr = f"fma({float(e)!r}, x{p.dim_index[d]}, {expr})"
                L.append(f"const double w{v} = {expr};")
                self._update_arctic(L, v, f"w{v}", sidx[v], v in owned, oidx.get(v))
        return L

    def _step_dp(self, L, part, sidx) -> None:
        """Nodes of a separable recurrence (cosine weighted ISS, iss/cos.py):
        ``S_v[t] = S_v[t-1] + letter[t] * sum_j row_j[t] * S_{u_j}[t-1]`` -- the
        states of the previous level weighted with shared rows.  Letter products
        are formed once per step and distinct letter; deeper levels first, so a
        node is updated after the nodes that read it."""
        p, nodes = self.p, self.p.trie.nodes
        dps = [v for v in part.snodes if getattr(nodes[v], "dp", None) is not None]
        if not dps:
            return
        letters = {}
        for v in dps:
            occ = nodes[v].dp[0]
            if occ and occ not in letters:
                name = f"lt{len(letters)}"
                expr = None
                for d, div in occ:
                    x = f"x{p.dim_index[d]}"
                    if expr is None:
                        expr = x if not div else f"__ddiv_rn(1.0, {x})"
                    else:
                        expr = f"{'__ddiv_rn' if div else '__dmul_rn'}({expr}, {x})"
                L.append(f"const double {name} = {expr};")
                letters[occ] = name
        for v in sorted(dps, key=lambda v: -nodes[v].depth):
            occ, preds = nodes[v].dp
            z = None
            for u, row in preds:
                w = f"x{p.dim_index[row]}"
                if u < 0:
                    term = w                                   # first level: the row itself
                    z = term if z is None else f"__dadd_rn({z}, {term})"
                elif z is None:
                    z = f"__dmul_rn({w}, S[{sidx[u]}])"
                else:
                    z = f"fma({w}, S[{sidx[u]}], {z})"
            lt = letters.get(occ)
            if z is None:
                val = lt if lt is not None else "1.0"
            elif lt is None:
                val = z
            else:
                val = f"__dmul_rn({lt}, {z})"
            L.append(f"const double dv{v} = {val};")
        for v in dps:
            L.append(f"S[{sidx[v]}] = __dadd_rn(S[{sidx[v]}], dv{v});")

    def _update_reals(self, L, v, val, si, is_owned, oi):
        p = self.p
        node = p.trie.nodes[v]
        wm = p.weight_mode
        S = f"S[{si}]"
        if wm == be.WEIGHT_TOTAL:
            a = p.aidx[node.alpha]
            L.append(f"{S} = __dadd_rn({S}, __dmul_rn({val}, ep{a}));")
            if is_owned:
                L.append(f"const double o{v} = __dmul_rn({S}, em{a});")
                L.append(f"const double q{v} = OP[{si}]; OP[{si}] = o{v};")
                self._sieve(L, v, f"o{v}", f"q{v}", oi)
        else:
            if is_owned:
                L.append(f"const double q{v} = {S};")
            L.append(f"{S} = __dadd_rn({S}, {val});")
            if wm == be.WEIGHT_NONTOTAL and node.children:
                a = p.aidx[node.alpha]
                L.append(f"A2[{si}] = __dadd_rn(A2[{si}], __dmul_rn({val}, ep{a}));")
            if is_owned:
                self._sieve(L, v, S, f"q{v}", oi)

    def _update_arctic(self, L, v, val, si, is_owned, oi):
        p = self.p
        node = p.trie.nodes[v]
        wm = p.weight_mode
        S = f"S[{si}]"
        mx = lambda a, b: f"(({b}) > ({a}) ? ({b}) : ({a}))"   # noqa: E731  keeps a on NaN
        if wm == be.WEIGHT_TOTAL:
            al = _dlit(float(node.alpha))
            L.append(f"const double y{v} = fma(g0, {al}, {val});")
            L.append(f"{S} = {mx(S, f'y{v}')};")
            L.append(f"const double q{v} = OP[{si}];")
            L.append(f"OP[{si}] = fma(-g0, {al}, {S});")
            if is_owned:
                self._sieve(L, v, f"OP[{si}]", f"q{v}", oi)
        elif wm == be.WEIGHT_NONTOTAL:
            u = node.parent
            if u >= 0:
                alp = _dlit(float(p.trie.nodes[u].alpha))
                L.append(f"const double y{v} = fma(-g0, {alp}, {val});")
            else:
                L.append(f"const double y{v} = {val};")
            if is_owned:
                L.append(f"const double q{v} = {S};")
            L.append(f"{S} = {mx(S, f'y{v}')};")
            if node.children:
                al = _dlit(float(node.alpha))
                L.append(f"const double z{v} = fma(g0, {al}, y{v});")
                L.append(f"A2[{si}] = {mx(f'A2[{si}]', f'z{v}')};")
            if is_owned:
                self._sieve(L, v, S, f"q{v}", oi)
        else:
            if is_owned:
                L.append(f"const double q{v} = {S};")
            L.append(f"{S} = {mx(S, val)};")
            if is_owned:
                self._sieve(L, v, S, f"q{v}", oi)

    def _sieve(self, L, v, out, prev, oi):
        """Feed the new value ``out`` (previous value ``prev``) of the owned
        node with index ``oi`` to its accumulators."""
        sv = self.sv
        e = self.p.trie.nodes[v].emit
        cregs = self._cnt_layout()

        def unit(k, val):
            """Accumulators of the increment unit ``k`` fed with ``val``: one
            predicate ``lo < val [<= hi]``, then count (NPI), sum (MPI), sum of
            the time indices (XPI), current / longest run (LPI)."""
            reg, hi16 = cregs[("U", k)]
            inc = "0x10000" if hi16 else "1"
            outs = {"cn": f'"+r"(CN[{oi}][{reg}])'}
            ins = {"v": f'"d"({val})', "lo": f'"d"({self.th(e, _COL_U[k][0])})'}
            if sv.cut:
                ins["live"] = '"r"(live)'
                asm = ["{ .reg .pred p, q;", "setp.ne.b32 q, %live, 0;", "setp.gt.and.f64 p, %v, %lo, q;"]
            else:
                asm = ["{ .reg .pred p;", "setp.gt.f64 p, %v, %lo;"]
            if sv.hi:
                ins["hi"] = f'"d"({self.th(e, _COL_U[k][1])})'
                asm.append("setp.le.and.f64 p, %v, %hi, p;")
            asm.append(f"@p add.u32 %cn, %cn, {inc};")
            if sv.avg[k]:
                outs["sm"] = f'"+d"(SM{k}[{oi}])'
                asm.append("@p add.rn.f64 %sm, %sm, %v;")
            if sv.xpi[k]:
                outs["xs"] = f'"+r"(XS{k}[{oi}])'
                ins["tix"] = '"r"(tix)'
                asm.append("@p add.u32 %xs, %xs, %tix;")
            if sv.lpi[k]:
                outs["lc"] = f'"+r"(LC{k}[{oi}])'
                outs["ll"] = f'"+r"(LL{k}[{oi}])'
                asm += ["@p add.u32 %lc, %lc, 1;", "@!p mov.u32 %lc, 0;", "max.u32 %ll, %ll, %lc;"]
            asm.append("}")
            text = " ".join(asm)
            # operands are numbered outputs first, then inputs (longest names first, so
            # that %lo is not mistaken for a prefix of another placeholder)
            names = list(outs) + list(ins)
            for name in sorted(names, key=len, reverse=True):
                text = text.replace("%" + name, "%" + str(names.index(name)))
            L.append('asm("' + text + '" : ' + ", ".join(outs.values()) + " : "
                     + ", ".join(ins.values()) + ");")

        if sv.cnt[0]:
            unit(0, out)
        if sv.cnt[1] or sv.cnt[2] or sv.cur:
            # (the very first step sees prev = 0 / -inf instead of the zero
            # padding of the reference; fixup() repairs these units after it)
            L.append(f"const double d{v} = __dadd_rn({out}, -{prev});")
            if sv.cnt[1]:
                unit(1, f"d{v}")
            if sv.cnt[2] or sv.cur:
                L.append(f"const double dd{v} = __dadd_rn(d{v}, -D1[{oi}]); D1[{oi}] = d{v};")
            if sv.cnt[2]:
                unit(2, f"dd{v}")
            if sv.cur:
                # CUR: sum of dd^2 over lo < dd <= hi (fruits/sieving/segment.py:246-258)
                live = ("setp.ne.b32 q, %4, 0; setp.gt.and.f64 p, %1, %2, q; " if sv.cut
                        else "setp.gt.f64 p, %1, %2; ")
                L.append('asm("{ .reg .pred p, q; .reg .f64 s; ' + live +
                         'setp.le.and.f64 p, %1, %3, p; fma.rn.f64 s, %1, %1, %0; '
                         f'selp.f64 %0, s, %0, p; }}" : "+d"(SQ[{oi}]) : "d"(dd{v}), '
                         f'"d"({self.th(e, _COL_CUR[0])}), "d"({self.th(e, _COL_CUR[1])})'
                         + (', "r"(live)' if sv.cut else "") + ");")
        if sv.cpv:
            # CPV: rising edges of (y >= threshold); CPP = the indicator of the previous
            # step, 1 before the first (the increments of the indicator are zero padded)
            L.append('asm("{ .reg .pred p; .reg .u32 c, e; setp.ge.f64 p, %2, %3; selp.u32 c, 1, 0, p; '
                     'not.b32 e, %1; and.b32 e, e, c; add.u32 %0, %0, e; mov.u32 %1, c; }" : '
                     f'"+r"(CPC[{oi}]), "+r"(CPP[{oi}]) : "d"({out}), "d"({self.th(e, _COL_CPV)}));')
        if sv.ppv:
            reg, hi16 = cregs[("P", 0)]
            inc = "0x10000" if hi16 else "1"
            L.append('asm("{ .reg .pred p; setp.ge.f64 p, %1, %2; @p add.u32 %0, %0, ' + inc + '; }" : '
                     f'"+r"(CN[{oi}][{reg}]) : "d"({out}), "d"({self.th(e, _COL_PPV)}));')
        for on, arr, cmp_, cols in ((sv.mx, "MX", "gt", _COL_MAX), (sv.mn, "MN", "lt", _COL_MIN)):
            if not on:
                continue
            asm = ["{ .reg .pred p, q;", f"setp.{cmp_}.f64 p, %1, %0;"]
            ins = [f'"d"({out})']
            if sv.mmb:
                asm.append("setp.gt.and.f64 p, %1, %2, p;")
                asm.append("setp.le.and.f64 p, %1, %3, p;")
                ins += [f'"d"({self.th(e, cols[0])})', f'"d"({self.th(e, cols[1])})']
            if sv.cut:
                ins.append('"r"(live)')
                asm.append(f"setp.ne.and.b32 p, %{len(ins)}, 0, p;")
            asm.append("selp.f64 %0, %1, %0, p; }")
            L.append('asm("' + " ".join(asm) + f'" : "+d"({arr}[{oi}]) : ' + ", ".join(ins) + ");")
        if sv.cut and sv.end:
            # END of the segment: the value at step cut - 1
            L.append(f"EN[{oi}] = last ? {out} : EN[{oi}];")

    def fixup(self, part: Part) -> list:
        """After the step t = 0: the increments of the reference are zero
        padded (fruits/sieving/increment.py:63-71, fruits/cache.py:8-13), so
        the first value of every increment unit is 0.0, not y[0] - 0."""
        sv = self.sv
        nodes = self.p.trie.nodes
        cregs = self._cnt_layout()
        L = []
        for oi, v in enumerate(part.owned):
            e = nodes[v].emit
            for k in (1, 2):
                if not sv.cnt[k]:
                    continue
                reg, hi16 = cregs[("U", k)]
                cond = f"(0.0 > {self.th(e, _COL_U[k][0])})"
                if sv.hi:
                    cond += f" && (0.0 <= {self.th(e, _COL_U[k][1])})"
                if sv.cut:
                    cond += " && (cend > 0)"
                keep, one = ("0x0000ffffu", "0x10000u") if hi16 else ("0xffff0000u", "1u")
                L.append(f"CN[{oi}][{reg}] = (CN[{oi}][{reg}] & {keep}) | (({cond}) ? {one} : 0u);")
                if sv.avg[k]:
                    L.append(f"SM{k}[{oi}] = 0.0;")
                if sv.xpi[k]:
                    L.append(f"XS{k}[{oi}] = 0u;")           # index 0 adds nothing
                if sv.lpi[k]:
                    L.append(f"LC{k}[{oi}] = LL{k}[{oi}] = ({cond}) ? 1u : 0u;")
            if sv.cnt[2] or sv.cur:
                L.append(f"D1[{oi}] = 0.0;")
            if sv.cur:
                L.append(f"SQ[{oi}] = 0.0;")                  # the first second increment is 0
        return L

    def _cnt_layout(self):
        """16-bit counters packed two per register: {(kind, k): (reg, high half)}."""
        keys = [("U", k) for k in range(3) if self.sv.cnt[k]]
        if self.sv.ppv:
            keys.append(("P", 0))
        return {key: (i // 2, bool(i % 2)) for i, key in enumerate(keys)}

    def n_cnt_regs(self) -> int:
        return (len(self._cnt_layout()) + 1) // 2

    # -- epilogue of one part --------------------------------------------------
    def epilogue(self, part: Part) -> list:
        p, sv = self.p, self.sv
        nodes = p.trie.nodes
        L = []
        sidx = {v: i for i, v in enumerate(part.snodes)}
        cregs = self._cnt_layout()
        nf = len(sv.feats)

        def count(oi, key):
            reg, hi16 = cregs[key]
            return f"(CN[{oi}][{reg}] >> 16)" if hi16 else f"(CN[{oi}][{reg}] & 0xffffu)"

        for oi, v in enumerate(part.owned):
            e = nodes[v].emit
            endv = f"OP[{sidx[v]}]" if p.weight_mode == be.WEIGHT_TOTAL else f"S[{sidx[v]}]"
            if sv.cut:
                endv = f"EN[{oi}]"
            for f, (kind, arg) in enumerate(sv.feats):
                if kind == be.FEAT_CNT:
                    val = f"(double){count(oi, ('U', arg))}"
                elif kind == be.FEAT_AVG:
                    c = count(oi, ("U", arg))
                    val = f"({c} ? __ddiv_rn(SM{arg}[{oi}], (double){c}) : 0.0)"
                elif kind == be.FEAT_PPV:
                    val = f"__ddiv_rn((double){count(oi, ('P', 0))}, (double)T)"
                elif kind == be.FEAT_MAX:
                    val = f"(MX[{oi}] == D_NINF ? 0.0 : MX[{oi}])"
                elif kind == be.FEAT_MIN:
                    val = f"(MN[{oi}] == D_INF ? 0.0 : MN[{oi}])"
                elif kind == be.FEAT_XPI:
                    c = count(oi, ("U", arg))
                    val = f"({c} ? __ddiv_rn((double)XS{arg}[{oi}], (double){c}) : 0.0)"
                elif kind == be.FEAT_LPI:
                    val = f"(double)LL{arg}[{oi}]"
                elif kind == be.FEAT_CUR:
                    val = f"SQ[{oi}]"
                elif kind == be.FEAT_CPV:
                    # 2 * edges / length rounded up to even (fruits/sieving/implicit.py:173-176)
                    val = f"__ddiv_rn((double)(2u * CPC[{oi}]), (double)(T + (T & 1)))"
                else:
                    val = endv
                L.append((e * nf + f, f"fin({val}, a.sanitize & 1)"))
        return L

    def staging_width(self) -> int:
        """Doubles per series row of a warp's staging area in the epilogue (the
        tile buffers of its group, free once the time loop is over), or 0 if
        the features are stored directly."""
        if not self.stage or self.ppc not in (1, 2, 4):
            return 0
        row = self.nrow * self.tt + 2
        sw = row if self.ppc <= 2 else row // 2
        if sw % 2 == 0:
            sw -= 1                      # odd stride: conflict-free column writes
        return sw if sw >= 5 else 0

    def epilogue_code(self, part: Part) -> list:
        """Store the features of one part.  Staged form: every lane puts its
        values into shared memory, then the warp writes the 32 rows with
        consecutive lanes on consecutive columns -- full sectors instead of one
        8-byte piece per lane, which matters most when ``a.out`` is the
        NVSwitch multicast mapping (every store becomes a packet to all GPUs)."""
        vals = sorted(self.epilogue(part))
        sw = self.staging_width()
        if not sw:
            return (["if (ns_ < a.n) {",
                     "    double *o = a.out + (size_t)ns_ * a.out_ld + a.col0;"]
                    + [f"    put(o + {col}, {expr}, mc);" for col, expr in vals] + ["}"])
        L = []
        i = 0
        while i < len(vals):
            j = i + 1
            while j < len(vals) and j - i < sw - 1 and vals[j][0] == vals[j - 1][0] + 1:
                j += 1
            for k in range(i, j):
                L.append(f"stg[lane * {sw} + {k - i}] = {vals[k][1]};")
            L.append("__syncwarp();")
            L.append("#pragma unroll 1")
            L.append("for (int r = 0; r < 32; r++) {")
            L.append("    if (nrow0 + r >= a.n) break;")
            L.append(f"    double *orow = a.out + (size_t)(nrow0 + r) * a.out_ld + a.col0 + {vals[i][0]};")
            L.append(f"    for (int e = lane; e < {j - i}; e += 32) put(orow + e, stg[r * {sw} + e], mc);")
            L.append("}")
            L.append("__syncwarp();")
            i = j
        return L

    def _noret(self) -> str:
        return "__attribute__((noreturn)) " if self.abi >= 1 else ""

    def _param(self) -> str:
        return {0: "const Args a", 1: "const Args a", 2: "const Args &a", 3: "const Args &a_"}[self.abi]

    # -- whole kernel ------------------------------------------------------------
    def part_name(self, pi: int) -> str:
        # padding parts (the part count is rounded up to whole groups) share one
        # function that only takes part in the staging and the barriers
        return f"fb_part_{pi}" if self.p.parts[pi].owned else "fb_part_idle"

    def entry_source(self, minb: int) -> str:
        """Translation unit of the kernel itself: the threshold table and the
        dispatch of every warp to the device function of its trie part (the
        parts are compiled separately, in parallel, and linked with nvJitLink)."""
        p = self.p
        npg = len(p.parts) // self.ppc
        nthr = max(1, len(p.trie.emits) * self.ntc)
        src = [self._common(), f"__constant__ double TH[{nthr}];"]
        for name in sorted({self.part_name(pi) for pi in range(len(p.parts))}):
            src.append(f'extern "C" __device__ {self._noret()}void {name}({self._param()});')
        src.append("// (the parts never return -- no callee-saved registers go to local memory -- and")
        src.append("// read the kernel parameters in place: __grid_constant__ makes &a a device address)")
        src.append(f'extern "C" __global__ void __launch_bounds__(NT, {minb}) '
                   f'fb_jit_slice(const {"__grid_constant__ " if self.abi >= 2 else ""}Args a)')
        src.append("{")
        src.append("    // consecutive CTAs work on the same series with different parts: the")
        src.append("    // input tile is read from HBM once and hits L2 for the other parts")
        src.append(f"    const int part = (int)(blockIdx.x % {npg}u) * PPC + (threadIdx.x >> 5) % PPC;")
        src.append("    switch (part) {")
        for pi, part in enumerate(p.parts):
            if part.owned:
                src.append(f"    case {pi}: {self.part_name(pi)}(a); break;")
        if any(not part.owned for part in p.parts):
            src.append("    default: fb_part_idle(a); break;")
        else:
            src.append("    default: break;")
        src.append("    }")
        src.append("}")
        return "\n".join(src) + "\n"

    def _common(self) -> str:
        TT = self.tt
        row = self.nrow * TT + 2           # doubles per series and tile
        erow = self.nextra * TT + 2
        src = []
        A = src.append
        A("// generated by fruits_b200/_jit.py -- do not edit")
        A(f"#define TT {TT}")
        A(f"#define PPC {self.ppc}")
        A(f"#define GPC {self.gpc}")
        A("#define NT (32 * PPC * GPC)")
        A(f"#define NROW {self.nrow}")
        A(f"#define ROW {row}")
        A(f"#define NEXTRA {self.nextra}")
        A(f"#define EROW {erow}")
        A("#define D_INF __longlong_as_double(0x7ff0000000000000LL)")
        A("#define D_NINF __longlong_as_double(0xfff0000000000000LL)")
        A("struct Args { const double *X; const double *E; double *out; long long n, d, t, e_ld, out_ld, col0; int sanitize; const int *cut; };")
        return "\n".join(src)

    def source(self, pi: int) -> str:
        """Translation unit of one trie part: ``fb_part_<pi>`` runs the whole
        time loop of the part for the 32 series of the calling warp."""
        p, sv = self.p, self.sv
        parts = [p.parts[pi]]
        wm = p.weight_mode
        ns = max(1, max(len(pt.snodes) for pt in parts))
        no = max(1, max(len(pt.owned) for pt in parts))
        need_first = sv.cnt[1] or sv.cnt[2] or sv.cur
        ncr = max(1, self.n_cnt_regs())
        nthr = max(1, len(p.trie.emits) * self.ntc)
        du = len(p.used)
        TT = self.tt
        row = self.nrow * TT + 2           # doubles per series and tile
        erow = self.nextra * TT + 2
        per_series_extra = self.nextra and not self.shared_extra
        src = [self._common()]
        A = src.append
        A(f"extern __constant__ double TH[{nthr}];")
        A("__device__ __forceinline__ int raw_row(int r) { return "
          + "".join(f"r == {i} ? {raw} : " for i, raw in enumerate(self.raw_rows[:-1]))
          + f"{self.raw_rows[-1]}; }}")
        A("__device__ __forceinline__ double fin(double v, int sanitize) {")
        A("    if (!sanitize) return v;")
        A("    if (v != v) return 0.0;")
        A("    if (v == D_INF) return __longlong_as_double(0x7fefffffffffffffLL);")
        A("    if (v == D_NINF) return __longlong_as_double(0xffefffffffffffffLL);")
        A("    return v; }")
        A("// a.sanitize bit 1: a.out is an NVSwitch multicast mapping -- such addresses may")
        A("// only be accessed with multimem.* instructions (PTX ISA).  No fence here: the")
        A("// stores are complete at system scope when the grid ends, and the ranks meet in")
        A("// the device-side barrier launched behind this kernel (parallel.PeerGather.finish)")
        A("__device__ __forceinline__ void put(double *p, double v, bool mc) {")
        A('    if (mc) asm volatile("multimem.st.weak.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");')
        A("    else *p = v; }")
        A("__device__ __forceinline__ void cp16(double *dst, const double *src) {")
        A('    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory"); }')
        A("__device__ __forceinline__ void cp8(double *dst, const double *src) {")
        A('    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory"); }')
        A(f'extern "C" __device__ __noinline__ {self._noret()}void '
          f'{self.part_name(pi)}({self._param()})')
        A("{")
        if self.abi == 3:
            A("    const Args a = a_;      // into registers once: no reloads behind the cp.async fences")
        A("    extern __shared__ __align__(16) double smem[];")
        A("    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;")
        A("    const int part = 0;")
        A("    const int sg = warp / PPC;")
        A(f"    const long long nbase = (long long)(blockIdx.x / {len(p.parts) // self.ppc}u) * (GPC * 32);")
        A("    const int T = (int)a.t;")
        A("    const bool mc = (a.sanitize & 2) != 0; (void)mc;")
        A("    // tile buffers: [2][GPC*32 series][ROW] then the weighting rows")
        A("    double *xbuf = smem;")
        A("    double *ebuf = smem + 2 * GPC * 32 * ROW;")
        A(f"    double S[{ns}];")
        if wm == be.WEIGHT_NONTOTAL:
            A(f"    double A2[{ns}];")
        if wm == be.WEIGHT_TOTAL:
            A(f"    double OP[{ns}];")
        A(f"    unsigned CN[{no}][{ncr}];")
        for k in range(3):
            if sv.avg[k]:
                A(f"    double SM{k}[{no}];")
        if sv.cnt[2] or sv.cur:
            A(f"    double D1[{no}];")
        if sv.mx:
            A(f"    double MX[{no}];")
        if sv.mn:
            A(f"    double MN[{no}];")
        for k in range(3):
            if sv.xpi[k]:
                A(f"    unsigned XS{k}[{no}];")
            if sv.lpi[k]:
                A(f"    unsigned LC{k}[{no}], LL{k}[{no}];")
        if sv.cur:
            A(f"    double SQ[{no}];")
        if sv.cpv:
            A(f"    unsigned CPC[{no}], CPP[{no}];")
        if sv.cut and sv.end:
            A(f"    double EN[{no}];")
        init = "0.0" if p.reals else "D_NINF"
        A("#pragma unroll")
        A(f"    for (int i = 0; i < {ns}; i++) {{ S[i] = {init};"
          + (f" A2[i] = {init};" if wm == be.WEIGHT_NONTOTAL else "")
          + (" OP[i] = 0.0;" if wm == be.WEIGHT_TOTAL else "") + " }")
        A("#pragma unroll")
        A(f"    for (int i = 0; i < {no}; i++) {{")
        A("#pragma unroll")
        A(f"        for (int j = 0; j < {ncr}; j++) CN[i][j] = 0u;")
        for k in range(3):
            if sv.avg[k]:
                A(f"        SM{k}[i] = 0.0;")
        if sv.cnt[2] or sv.cur:
            A("        D1[i] = 0.0;")
        if sv.mx:
            A("        MX[i] = D_NINF;")
        if sv.mn:
            A("        MN[i] = D_INF;")
        for k in range(3):
            if sv.xpi[k]:
                A(f"        XS{k}[i] = 0u;")
            if sv.lpi[k]:
                A(f"        LC{k}[i] = 0u; LL{k}[i] = 0u;")
        if sv.cur:
            A("        SQ[i] = 0.0;")
        if sv.cpv:
            A("        CPC[i] = 0u; CPP[i] = 1u;")
        if sv.cut and sv.end:
            A("        EN[i] = 0.0;")
        A("    }")
        if sv.cut:
            # segment [0, cend) of this lane's series (fruits/sieving/segment.py:51-64);
            # END reads position cend - 1, wrapping like the reference's negative index
            A("    const long long nser_ = nbase + sg * 32 + lane;")
            A("    const int craw = a.cut ? a.cut[nser_ < a.n ? nser_ : a.n - 1] : T;")
            A("    const int cend = craw < 0 ? 0 : (craw > T ? T : craw);")
            A("    const int eidx = craw >= 1 ? craw - 1 : craw - 1 + T;  (void)eidx;")
        # previous raw values of the dimensions that are read as increments
        inc_rows = sorted({self.row_of[d[0]] for d in self.dims if d is not None and d[1]})
        for r in inc_rows:
            A(f"    double xp{r} = 0.0;")
        # ---- staging ----
        # Every group of 32 series is staged by its own PPC warps and synchronised
        # with its own named barrier, so the groups of a CTA drift freely.
        A("    const bool even = ((a.t & 1) == 0) && ((((unsigned long long)a.X) & 15) == 0);")
        A("    const int gt = (warp % PPC) * 32 + lane;      // thread within the group")
        A("    const long long gbase = nbase + sg * 32;       // first series of the group")
        A("    double *gx = xbuf + (size_t)sg * 32 * ROW;     // + buf * GPC * 32 * ROW")
        # slow path: any alignment, tail of the batch
        A("    auto stage_slow = [&](int buf, int t0) {")
        A("        const int nchunk = 32 * NROW * (TT / 2);")
        A("        for (int c = gt; c < nchunk; c += 32 * PPC) {")
        A("            const int k = c % (TT / 2), sr = c / (TT / 2);")
        A("            const int r = sr % NROW, s = sr / NROW;")
        A("            long long n = gbase + s; if (n >= a.n) n = a.n - 1;")
        A("            const int t = t0 + 2 * k;")
        A("            const double *g = a.X + ((size_t)n * a.d + raw_row(r)) * (size_t)T + t;")
        A("            double *d = gx + ((size_t)(buf * GPC * 32 + s)) * ROW + r * TT + 2 * k;")
        A("            if (even) { if (t < T) cp16(d, g); }")
        A("            else { if (t < T) cp8(d, g); if (t + 1 < T) cp8(d + 1, g + 1); }")
        A("        }")
        if per_series_extra:
            A("        const int echunk = 32 * NEXTRA * (TT / 2);")
            A("        const bool eeven = ((a.t & 1) == 0) && ((((unsigned long long)a.E) & 15) == 0) && ((a.e_ld & 1) == 0);")
            A("        for (int c = gt; c < echunk; c += 32 * PPC) {")
            A("            const int k = c % (TT / 2), sr = c / (TT / 2);")
            A("            const int r = sr % NEXTRA, s = sr / NEXTRA;")
            A("            const int t = t0 + 2 * k;")
            A("            long long n = gbase + s; if (n >= a.n) n = a.n - 1;")
            A("            const double *g = a.E + (size_t)n * a.e_ld + (size_t)r * T + t;")
            A("            double *d = ebuf + ((size_t)(buf * GPC * 32 + sg * 32 + s)) * EROW + r * TT + 2 * k;")
            A("            if (eeven) { if (t < T) cp16(d, g); }")
            A("            else { if (t < T) cp8(d, g); if (t + 1 < T) cp8(d + 1, g + 1); }")
            A("        }")
        A("    };")
        # fast path: whole group inside the batch, 16-byte aligned rows; every
        # thread copies the same (series, chunk) pattern each tile, so all
        # offsets are compile-time multiples of T and D*T
        ch = TT // 2
        gthreads = 32 * self.ppc
        fast_ok = gthreads % ch == 0 and 32 % (gthreads // ch) == 0
        sp = gthreads // ch if fast_ok else 1     # series covered per pass
        passes = 32 // sp if fast_ok else 0
        A(f"    const bool fast = {'true' if fast_ok else 'false'} && even && (gbase + 32 <= a.n)"
          + (" && ((((unsigned long long)a.E) & 15) == 0) && ((a.e_ld & 1) == 0)" if per_series_extra else "") + ";")
        A(f"    const int fk = (gt % {ch}) * 2, fs = gt / {ch};")
        A("    const double *fsrc = a.X + ((size_t)(gbase + fs) * a.d) * (size_t)T + fk;")
        A("    double *fdst = gx + (size_t)fs * ROW + fk;")
        A("    const size_t DT = (size_t)a.d * (size_t)T;")
        if per_series_extra:
            A("    const double *fesrc = a.E + (size_t)(gbase + fs) * a.e_ld + fk;")
            A("    double *fedst = ebuf + (size_t)(sg * 32 + fs) * EROW + fk;")
        A("    auto stage_fast = [&](int buf, int t0) {")
        A("        if (t0 + fk < T) {")
        A("            const double *g = fsrc + t0;")
        A("            double *d = fdst + (size_t)buf * (GPC * 32 * ROW);")
        for j in range(passes):
            for r, raw in enumerate(self.raw_rows):
                A(f"            cp16(d + {j * sp} * ROW + {r} * TT, g + {j * sp} * DT + {raw} * (size_t)T);")
        if per_series_extra:
            A("            const double *ge = fesrc + t0;")
            A("            double *de = fedst + (size_t)buf * (GPC * 32 * EROW);")
            for j in range(passes):
                for r in range(self.nextra):
                    A(f"            cp16(de + {j * sp} * EROW + {r} * TT, ge + {j * sp} * (size_t)a.e_ld + {r} * (size_t)T);")
        A("        }")
        A("    };")
        A("    auto stage = [&](int buf, int t0) {")
        A("        if (fast) stage_fast(buf, t0); else stage_slow(buf, t0);")
        if self.nextra and not per_series_extra:
            A("        // weighting rows shared by all series (one copy per group)")
            A("        const bool eeven = ((a.t & 1) == 0) && ((((unsigned long long)a.E) & 15) == 0);")
            A("        for (int c = gt; c < NEXTRA * (TT / 2); c += 32 * PPC) {")
            A("            const int k = c % (TT / 2), r = c / (TT / 2);")
            A("            const int t = t0 + 2 * k;")
            A("            const double *g = a.E + (size_t)r * T + t;")
            A("            double *d = ebuf + (size_t)(buf * GPC + sg) * EROW + r * TT + 2 * k;")
            A("            if (eeven) { if (t < T) cp16(d, g); }")
            A("            else { if (t < T) cp8(d, g); if (t + 1 < T) cp8(d + 1, g + 1); }")
            A("        }")
        A('        asm volatile("cp.async.commit_group;" ::: "memory");')
        A("    };")
        A("    stage(0, 0);")
        A("    int buf = 0;")
        A("    for (int t0 = 0; t0 < T; t0 += TT, buf ^= 1) {")
        A('        asm volatile("cp.async.wait_all;" ::: "memory");')
        A("        // the PPC warps of this group (each inside its own part function)")
        A('        asm volatile("bar.sync %0, %1;" :: "r"(sg + 1), "n"(32 * PPC) : "memory");')
        A("        if (t0 + TT < T) stage(buf ^ 1, t0 + TT);")
        A("        const double *xs = xbuf + ((size_t)(buf * GPC * 32 + sg * 32 + lane)) * ROW;")
        if self.nextra:
            if per_series_extra:
                A("        const double *es = ebuf + ((size_t)(buf * GPC * 32 + sg * 32 + lane)) * EROW;")
            else:
                A("        const double *es = ebuf + (size_t)(buf * GPC + sg) * EROW;")
        A("        const int tend = min(TT, T - t0);")
        A("        int tt = 0;")
        if need_first:
            A("        // the very first step runs on its own: the zero padding of the")
            A("        // increments is repaired right after it (fixup)")
            A("        bool fix = (t0 == 0);")
            A("        int stop = fix ? 1 : tend;")
        else:
            A("        const int stop = tend;")
        A("        for (;;) {")
        A("        switch (part) {")
        for pi, part in enumerate(parts):
            if not part.owned:
                continue
            A(f"        case {pi}: {{")
            A("#pragma unroll 1")
            A("            for (; tt < stop; tt++) {")
            A("                const bool is0 = (t0 + tt) == 0;")
            A("                const unsigned tix = (unsigned)(t0 + tt); (void)tix;")
            if sv.cut:
                A("                const int live = (int)tix < cend; const bool last = (int)tix == eidx;")
                A("                (void)live; (void)last;")
            for ln in self._loads() + self.step(part) + self._after():
                A("                " + ln)
            A("            }")
            A("        } break;")
        A("        default: break;")
        A("        }")
        if need_first:
            A("        if (!fix) break;")
            A("        switch (part) {")
            for pi, part in enumerate(parts):
                if not part.owned:
                    continue
                A(f"        case {pi}: {{")
                for ln in self.fixup(part):
                    A("            " + ln)
                A("        } break;")
            A("        default: break;")
            A("        }")
            A("        stop = tend;")
            A("        fix = false;")
        else:
            A("        break;")
        A("        }")
        A("    }")
        # ---- epilogue ----
        A("    const long long ns_ = nbase + sg * 32 + lane;")
        A("    const long long nrow0 = nbase + sg * 32;        // first series of this warp")
        if self.staging_width():
            # all warps of the group are done with the tile buffers: reuse them
            A('    asm volatile("bar.sync %0, %1;" :: "r"(sg + 1), "n"(32 * PPC) : "memory");')
            if self.ppc <= 2:
                A("    double *stg = xbuf + ((size_t)((warp % PPC) * GPC * 32 + sg * 32)) * ROW;")
            else:
                A("    double *stg = xbuf + ((size_t)(((warp % PPC) / 2) * GPC * 32 + sg * 32)) * ROW"
                  " + ((warp % PPC) % 2) * 16 * ROW;")
        A("    (void)ns_; (void)nrow0;")
        A("    switch (part) {")
        for pi, part in enumerate(parts):
            if not part.owned:
                continue
            A(f"    case {pi}: {{")
            for ln in self.epilogue_code(part):
                A("        " + ln)
            A("    } break;")
        A("    default: break;")
        A("    }")
        A("    // the kernel has nothing left to do after its part: end the thread here, so")
        A("    // that the compiler need not save / restore the caller's registers (measured:")
        A("    // 208 B of local-memory stores and loads per thread, 0.6x the feature bytes)")
        if self.abi >= 1:
            A('    asm volatile("exit;");')
            A("    __builtin_unreachable();")
        A("}")
        del du
        return "\n".join(src) + "\n"

    def _loads(self) -> list:
        """Values of the current step: x{u} per used dimension, ep/em/g0."""
        L = []
        for r in range(self.nrow):
            L.append(f"const double r{r} = xs[{r} * TT + tt];")
        for u, dim in enumerate(self.dims):
            if dim is None:                  # shared extra row (sin / cos of CosWISS)
                L.append(f"const double x{u} = es[{self.row_dims[u]} * TT + tt];")
                continue
            raw, inc = dim
            r = self.row_of[raw]
            if inc:
                L.append(f"const double x{u} = is0 ? 0.0 : __dadd_rn(r{r}, -xp{r});")
            else:
                L.append(f"const double x{u} = r{r};")
        p = self.p
        if p.weight_mode != be.WEIGHT_NONE:
            if p.reals:
                for a in range(len(p.alphas)):
                    L.append(f"const double ep{a} = es[{2 * a} * TT + tt];")
                    L.append(f"const double em{a} = es[{2 * a + 1} * TT + tt];")
            else:
                L.append("const double g0 = es[tt];")
        return L

    def _after(self) -> list:
        inc_rows = sorted({self.row_of[d[0]] for d in self.dims if d is not None and d[1]})
        return [f"xp{r} = r{r};" for r in inc_rows]

    def smem_bytes(self) -> int:
        row = self.nrow * self.tt + 2
        erow = self.nextra * self.tt + 2
        n = 2 * self.gpc * 32 * row
        if self.nextra:
            n += 2 * self.gpc * (erow if self.shared_extra else 32 * erow)
        return n * 8


# ---------------------------------------------------------------------------
# run time: compile (NVRTC inside libfruits_b200.so), cache, load, launch
# ---------------------------------------------------------------------------

class FbJitGeometry(ctypes.Structure):
    _fields_ = [("n_parts", ctypes.c_int32), ("parts_per_cta", ctypes.c_int32),
                ("groups_per_cta", ctypes.c_int32), ("smem_bytes", ctypes.c_int32)]


DEFAULT_OPTS = {"budget": 70, "ppc": 2, "gpc": 8, "minb": 1, "unroll": 2, "tt": 16,
                "stage": 1,        # epilogue: features through shared memory, coalesced rows
                "abi": 3}          # part functions: 0 by value + return, 1 by value + exit,
                                   # 2 by reference + exit, 3 = 2 with a local copy of the struct


def options() -> dict:
    """Generator options; ``FRUITS_B200_JIT_OPTS="budget=150,ppc=4"`` overrides."""
    opts = dict(DEFAULT_OPTS)
    env = os.environ.get("FRUITS_B200_JIT_OPTS", "")
    for item in filter(None, env.split(",")):
        key, val = item.split("=")
        if key not in opts:
            raise ValueError(f"unknown JIT option {key!r}")
        opts[key] = int(val)
    return opts


def options_key() -> tuple:
    return tuple(sorted(options().items()))


def enabled(n_series: int = None) -> bool:
    """``FRUITS_B200_JIT``: "0" never, "force" always, default: whenever the
    batch is large enough to fill the GPU with one thread per series and trie
    part (small batches are better served by the generic kernel, which runs
    one lane per trie node)."""
    mode = os.environ.get("FRUITS_B200_JIT", "1")
    if mode == "0":
        return False
    if mode == "force" or n_series is None:
        return True
    return n_series >= MIN_SERIES


MIN_SERIES = 4096          # from this batch size on a slice is compiled (seconds, cached on disk)
MIN_SERIES_CACHED = 1000   # ... and from this size on an already compiled kernel is used:
                           # it beats the generic kernel from ~1,000 series (scripts/crossover.py;
                           # C2 slice 0, 1,000 x 512: 0.60 against 0.73 ms)


class NotCompiled(NotImplementedError):
    """The generated kernel of a plan exists neither in memory nor on disk."""


def generate(trie, semiring: int, weight_mode: int, sieves: SieveSet, dims: list,
             shared_extra: bool, opts: dict, n_shared_rows: int = 0, max_state_regs: int = 0,
             small_batch: bool = False):
    """-> Generated: one translation unit per trie part plus the kernel that
    dispatches to them.  Raises NotImplementedError for plans the generated
    kernel cannot hold (the caller then uses the generic kernel).

    ``small_batch``: fewer than ``MIN_SERIES`` series -- too few CTAs for the layout
    with fewer, larger parts, so the plan keeps the many-parts layout if it fits
    (C2, 1,000 series: slice 0 0.33 -> 0.27 ms, CosWISS slices 0.34 -> 0.31 and
    0.45 -> 0.38 ms)."""
    prog = Program(trie, semiring, weight_mode, sieves, reg_budget=opts["budget"],
                   parts_multiple=opts["ppc"])
    narrow_fits = prog.overhead <= 1.6 and prog.max_regs <= (max_state_regs or 190)
    if ((prog.overhead > 1.15 or prog.max_regs > opts["budget"] + 20 or sieves.regs() > 6)
            and not (small_batch and narrow_fits)
            and "FRUITS_B200_JIT_OPTS" not in os.environ):
        # deep or sieve-heavy tries: fewer, larger parts (one CTA per SM, 255 registers);
        # measured on C4 slice 0 (depth 9, overhead 1.21 -> 1.08): 92 -> 71 ms
        # (256 threads x 255 registers; with every register in use the parts take their
        # parameters by value: measured on C4 slice 0, abi 1 / 2 / 3 = 70.4 / 75.5 / 90.2 ms)
        opts = dict(opts, budget=150, ppc=1, gpc=8, minb=1, unroll=1, abi=1)
        prog = Program(trie, semiring, weight_mode, sieves, reg_budget=opts["budget"],
                       parts_multiple=opts["ppc"])
    # (plans without another fused route -- CosWISS -- may spill a few sums to local memory)
    if prog.overhead > 1.6 or prog.max_regs > (max_state_regs or 190):
        # long chains (e.g. arctic words of 24-48 letters): every part would
        # recompute most of the chain -- the generic kernel is the better shape
        raise NotImplementedError("trie too deep for the plan-specialised kernel")
    # largest tile (then most series per CTA) whose double buffer fits
    em = None
    for gpc, tt in ((opts["gpc"], opts["tt"]), (opts["gpc"], 8), (max(1, opts["gpc"] // 2), 8),
                    (max(1, opts["gpc"] // 4), 8), (1, 4), (1, 2)):
        cand = Emitter(prog, dims, opts["ppc"], gpc, shared_extra, tt, n_shared_rows,
                       opts.get("stage", 1), opts.get("abi", 3))
        if cand.smem_bytes() <= 200 * 1024:
            em = cand
            break
    if em is None:
        raise NotImplementedError("input tile exceeds shared memory")
    if len(trie.emits) * em.ntc * 8 > 60 * 1024:
        raise NotImplementedError("threshold table exceeds the constant bank")
    if len(prog.parts) // opts["ppc"] > 65535:
        raise NotImplementedError("too many parts")
    minb = opts["minb"]
    nt = 32 * em.ppc * em.gpc
    # register file: 64K per SM, allocated per 4 warps, 8 registers granularity
    warps = -(-(nt // 32) // 4) * 4
    max_regs = min(255, (65536 // (warps * 32 * minb)) // 8 * 8)
    srcs = []
    idle = [pi for pi, part in enumerate(prog.parts) if not part.owned][:1]
    for pi, part in enumerate(prog.parts):
        if part.owned or pi in idle:
            src = em.source(pi)
            src = src.replace("#pragma unroll 1\n            for (; tt < stop; tt++)",
                              f"#pragma unroll {opts['unroll']}\n            for (; tt < stop; tt++)")
            srcs.append(src)
    return Generated(srcs, em.entry_source(minb), max_regs, em)


@dataclass
class Generated:
    parts: list          # source of one translation unit per trie part
    entry: str           # source of the kernel (dispatch + threshold table)
    max_regs: int        # register cap of the part functions
    em: Emitter

    def digest(self) -> str:
        d = self.__dict__.get("_digest")
        if d is None:
            h = hashlib.sha256(f"v{JIT_VERSION} r{self.max_regs}\n".encode())
            for src in self.parts + [self.entry]:
                h.update(src.encode())
                h.update(b"\0")
            d = self._digest = h.hexdigest()[:24]
        return d


def _nvrtc(src: str, name: str, relocatable: bool, max_regs: int) -> bytes:
    L = be.lib()
    cubin, size = ctypes.c_void_p(), ctypes.c_size_t()
    log = ctypes.create_string_buffer(1 << 16)
    rc = L.fb_jit_compile(src.encode(), name.encode(), int(relocatable), int(max_regs),
                          ctypes.byref(cubin), ctypes.byref(size), log, len(log))
    if rc != 0:
        msg = L.fb_last_error().decode(errors="replace")
        raise RuntimeError(f"JIT compilation failed: {msg}\n{log.value.decode(errors='replace')}")
    data = ctypes.string_at(cubin.value, size.value)
    L.fb_jit_free(cubin)
    return data


def build_cubin(gen: Generated) -> bytes:
    """Compile every part (NVRTC, in parallel threads: ctypes releases the
    GIL and NVRTC is thread safe), link them with the entry kernel
    (nvJitLink) and cache the loadable cubin on disk, keyed by the sources."""
    path = os.path.join(CACHE_DIR, gen.digest() + ".cubin")
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    jobs = [(src, f"fb_part_{i}.cu", True, gen.max_regs) for i, src in enumerate(gen.parts)]
    jobs.append((gen.entry, "fb_jit_slice.cu", True, 0))
    workers = max(1, min(len(jobs), os.cpu_count() or 1, 32))
    with ThreadPoolExecutor(max_workers=workers) as ex:
        objs = list(ex.map(lambda j: _nvrtc(*j), jobs))
    L = be.lib()
    n = len(objs)
    ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(ctypes.c_char_p(o), ctypes.c_void_p).value
                                   for o in objs])
    sizes = (ctypes.c_size_t * n)(*[len(o) for o in objs])
    out, size = ctypes.c_void_p(), ctypes.c_size_t()
    log = ctypes.create_string_buffer(1 << 16)
    rc = L.fb_jit_link(ptrs, sizes, n, ctypes.byref(out), ctypes.byref(size), log, len(log))
    if rc != 0:
        msg = L.fb_last_error().decode(errors="replace")
        raise RuntimeError(f"JIT link failed: {msg}\n{log.value.decode(errors='replace')}")
    data = ctypes.string_at(out.value, size.value)
    L.fb_jit_free(out)
    try:
        os.makedirs(CACHE_DIR, exist_ok=True)
        tmp = path + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
    except OSError:
        pass
    return data


class JitSlice:
    """The loaded plan-specialised kernel of one slice."""

    _loaded: dict = {}     # (source digest, device) -> JitSlice (a cudaLibrary belongs to a context)

    def __init__(self, gen: Generated) -> None:
        self.em = gen.em
        self.gen = gen
        cubin = build_cubin(gen)
        handle = ctypes.c_void_p()
        be.check(be.lib().fb_jit_load(cubin, len(cubin), ctypes.byref(handle)))
        self.handle = handle
        em = gen.em
        self.geo = FbJitGeometry(len(em.p.parts), em.ppc, em.gpc, em.smem_bytes())
        self.cols = list(em.cols)

    @classmethod
    def get(cls, trie, semiring, weight_mode, sieves, dims, shared_extra,
            n_shared_rows: int = 0, max_state_regs: int = 0) -> "JitSlice":
        gen = generate(trie, semiring, weight_mode, sieves, dims, shared_extra, options(),
                       n_shared_rows, max_state_regs)
        return cls.load(gen)

    @classmethod
    def load(cls, gen: Generated, cached_only: bool = False) -> "JitSlice":
        """The loaded kernel of ``gen``; with ``cached_only`` only if it is
        already in memory or compiled on disk (else ``NotCompiled``)."""
        import torch
        digest = gen.digest()
        key = (digest, torch.cuda.current_device() if torch.cuda.is_available() else -1)
        obj = cls._loaded.get(key)
        if obj is None:
            if cached_only and not os.path.exists(os.path.join(CACHE_DIR, digest + ".cubin")):
                raise NotCompiled(digest)
            obj = cls(gen)
            cls._loaded[key] = obj
        return obj

    def n_launches(self, n_series: int = 0, length: int = 0) -> int:
        return 1

    def launch(self, X, extra, extra_ld, thr_compact, out, col0, sanitize, multicast=None,
               cuts=None) -> None:
        """X[n, d, t] cuda float64; extra: weighting rows or None; cuts: int32 [n]
        end of the sieved segment per series (kernels generated with a cut);
        thr_compact: [n_emit * len(cols)] cuda float64.  ``multicast`` =
        (address, row stride) of the same rows in an NVSwitch multicast
        mapping: the kernel then stores there (with ``multimem.st``) instead of
        into ``out``."""
        batch = be.FbBatch()
        batch.X = X.data_ptr()
        batch.n, batch.d, batch.t = X.shape
        n_thr = 0 if thr_compact is None else thr_compact.numel()
        if self.em.sv.cut and cuts is None:
            raise ValueError("this kernel was generated for a cut: pass the cut table")
        be.check(be.lib().fb_jit_slice_features_cut(
            self.handle, ctypes.byref(self.geo), ctypes.byref(batch), be.ptr(extra),
            int(extra_ld), be.ptr(thr_compact), n_thr, be.ptr(cuts),
            out.data_ptr() if multicast is None else int(multicast[0]),
            out.stride(0) if multicast is None else int(multicast[1]), int(col0),
            int(bool(sanitize)) | (2 if multicast is not None else 0), be.stream_ptr()))
