"""Chain-pipelined generated kernel for deep Arctic tries (lane = trie node).

The plan-specialised kernel of ``_jit.py`` gives every thread one series and a
*part* of the trie; a part has to recompute the ancestors of its nodes, which
is cheap for the bushy tries of ``of_weight`` words and ruinous for the
24-48-letter alternating-sign chains of ``experiments/fruit_reduced.py:42-49``
/ ``fruit_general.py:42-51`` (every part would recompute most of its chain).
For those this module generates the opposite layout:

* one **lane per trie node**, a warp owns one series and one block of up to
  ``R x 32`` nodes in depth-first order, all running sums and sieve
  accumulators in registers, thresholds and letters in per-lane registers;
* Arctic sums read the parent at the *same* time step
  (fruits/iss/semiring.py:326-333), so the lanes run **skewed in time by
  their depth**: at step ``s`` a node of depth ``k`` works on ``t = s-(k-1)``
  and its parent -- one lane to the left in depth-first order -- holds exactly
  ``A_parent[t]`` from the step before.  Depth-first position ``i`` lives in
  lane ``i // R``, row ``i % R``: the parent of a chain node is the previous
  row of the *same lane* (a register, no instruction at all) and only row 0
  takes its parent from the last row of the lane to its left (one rotation
  shuffle per step); branch points and duplicated ancestors use an indexed
  shuffle.  Nothing goes through shared memory except the input itself;
* the whole (prepared) series sits in shared memory, staged once per CTA, so a
  lane reads ``x[t]`` at its own skewed position with one ``LDS.64``;
* the first ``depth-1`` and the last ``depth-1`` steps (pipeline fill and
  drain, zero padding of the increments) run through a masked copy of the
  step body, all steps in between through an unmasked one.

Floating point order: ``fma(e_d, x_d[t], parent)`` per dimension of the
letter, ascending, then the running maximum -- the reference's (numba contracts
``tmp + el*Z``; verified bit for bit by the oracle).  A maximum is exactly
associative and every lane walks its own time axis serially, so all results
are bit-identical to the reference.
"""
import ctypes
import hashlib
import os
from dataclasses import dataclass

from . import _backend as be
from . import _jit

CHAIN_VERSION = 6
MIN_SERIES = 512           # from this batch size on a chain kernel is compiled (cached on disk)
MIN_SERIES_CACHED = 16     # ... and from this size on an already compiled one is used


class FbJitChainGeometry(ctypes.Structure):
    _fields_ = [("blocks_per_series", ctypes.c_int32), ("series_per_cta", ctypes.c_int32),
                ("rows_staged", ctypes.c_int32), ("pad", ctypes.c_int32)]


# ---------------------------------------------------------------------------
# layout: trie -> blocks of R x 32 slots, depth-first, closed under ancestors
# ---------------------------------------------------------------------------

@dataclass
class Slot:
    node: int           # trie node id
    owned: bool         # this block writes the node's features
    parent: int         # position of the parent inside the block, -1 = root


def partition(trie, rows: int) -> list:
    """Blocks of at most ``rows * 32`` slots: the depth-first order cut into
    pieces of about equal size; ancestors a block needs but does not own are
    duplicated right before their first use (they cost one lane each)."""
    cap = 32 * rows
    if trie.max_depth > cap:
        raise NotImplementedError("word longer than one block of the chain kernel")
    dup_ok = chain_like(trie)        # (bushy tries would duplicate a node per sibling)
    order = trie.dfs()
    n_blocks = max(1, -(-len(order) // cap))
    while True:
        target = -(-len(order) // n_blocks)
        # lanes left over by the owned nodes of a block may carry duplicates
        spare = 0
        blocks, cur, pos, owned = [], [], {}, 0
        for v in order:
            chain, a = [], v
            while a >= 0:
                chain.append(a)
                a = trie.nodes[a].parent
            need = [a for a in reversed(chain) if a not in pos]
            if cur and (len(cur) + len(need) > cap or owned >= target):
                blocks.append(cur)
                cur, pos, owned = [], {}, 0
                need = list(reversed(chain))
            par = trie.nodes[v].parent
            if (dup_ok and need == [v] and par >= 0 and pos[par] != len(cur) - 1
                    and trie.nodes[par].depth <= 2
                    and len(cur) + trie.nodes[par].depth + 1 <= cap - spare):
                # a branch off a shallow node that sits far back in the block (the second
                # word of an alternating-sign pair shares only its first letter): a
                # duplicate of that node right here keeps the parent in the previous slot
                # -- no indexed shuffle in every step -- for the price of a free lane
                need = list(reversed(chain))
            for a in need:
                par = trie.nodes[a].parent
                if a in pos and a != v:
                    cur.append(Slot(a, False, len(cur) - 1 if par >= 0 else -1))
                    pos[a] = len(cur) - 1          # later nodes hang off the duplicate
                    continue
                pos[a] = len(cur)
                cur.append(Slot(a, a == v, pos[par] if par >= 0 else -1))
            owned += 1
        blocks.append(cur)
        if all(len(b) <= cap for b in blocks):
            return blocks
        n_blocks += 1


class ChainProgram:
    """Everything the emitter needs about one (trie, sieve set)."""

    def __init__(self, trie, sieves: "_jit.SieveSet", rows: int) -> None:
        self.trie, self.sieves, self.rows = trie, sieves, rows
        self.used = trie.used_dims()
        self.dim_index = {d: u for u, d in enumerate(self.used)}
        self.blocks = partition(trie, rows)
        self.max_skew = trie.max_depth - 1
        nodes = trie.nodes
        # dimension/exponent pairs of a letter, ascending dimension (the reference
        # adds el * Z[d] for every dimension in order, fruits/iss/semiring.py:326-327)
        self.pairs = {}
        for b in self.blocks:
            for sl in b:
                self.pairs[sl.node] = [(self.dim_index[d], e) for d, e in
                                       enumerate(nodes[sl.node].expo) if e != 0]
        # per row: how many pairs the code applies, which parent wirings occur
        # (position i of a block = lane i // rows, row i % rows)
        self.npairs = [0] * rows
        self.same = [False] * rows          # parent = previous row of the same lane
        self.rot = False                    # row 0: parent = last row of lane - 1
        self.roots = [False] * rows
        self.irregular = [set() for _ in range(rows)]    # source rows of indexed shuffles
        for b in self.blocks:
            for i, sl in enumerate(b):
                r = i % rows
                self.npairs[r] = max(self.npairs[r], len(self.pairs[sl.node]))
                if sl.parent < 0:
                    self.roots[r] = True
                elif sl.parent == i - 1 and r > 0:
                    self.same[r] = True
                elif sl.parent == i - 1:
                    self.rot = True
                else:
                    self.irregular[r].add(sl.parent % rows)
        self.irregular = [sorted(s) for s in self.irregular]
        self.n_slots = sum(len(b) for b in self.blocks)

    def tables(self):
        """Per-lane constants, [block][row][lane]:
        node word = emit (16 bits, 0xffff: none) | skew << 16 | root << 24,
        pair words = dim row | (exponent & 0xff) << 8,
        irregular words (one per source row of the row) = src lane | take << 8."""
        nodes = self.trie.nodes
        R = self.rows
        node_w, pair_w, irr_w = [], [], []
        for b in self.blocks:
            for r in range(R):
                for l in range(32):
                    i = l * R + r
                    if i >= len(b):
                        # padding lane: a root with letter 0 that emits nothing
                        node_w.append(0xffff | (1 << 24))
                        pair_w.append([0] * self.npairs[r])
                        irr_w.append([0] * len(self.irregular[r]))
                        continue
                    sl = b[i]
                    nd = nodes[sl.node]
                    emit = nd.emit if (sl.owned and nd.emit >= 0) else 0xffff
                    node_w.append(emit | ((nd.depth - 1) << 16) | ((1 if sl.parent < 0 else 0) << 24))
                    pw = [(u | ((e & 0xff) << 8)) for u, e in self.pairs[sl.node]]
                    pair_w.append(pw + [0] * (self.npairs[r] - len(pw)))
                    iw = []
                    regular = sl.parent < 0 or sl.parent == i - 1
                    for q in self.irregular[r]:
                        if not regular and sl.parent % R == q:
                            iw.append((sl.parent // R) | (1 << 8))
                        else:
                            iw.append(0)
                    irr_w.append(iw)
        return node_w, pair_w, irr_w


# ---------------------------------------------------------------------------
# emitter
# ---------------------------------------------------------------------------

class ChainEmitter:
    def __init__(self, prog: ChainProgram, dims: list, spc: int, unroll: int = 2) -> None:
        """dims[u] = (raw_dim, inc) of used dimension u; spc = series per CTA."""
        self.p, self.sv = prog, prog.sieves
        self.dims = dims
        self.spc = spc
        self.unroll = unroll
        self.cols = self.sv.thr_cols()
        self.ntc = len(self.cols)
        self.colpos = {c: i for i, c in enumerate(self.cols)}
        self.nb = len(prog.blocks)
        self.nrow = len(dims)                  # staged rows = used dimensions (prepared)
        self.pad = max(2, -(-(prog.max_skew + 1) // 2) * 2)

    # counters: 16 bits each, two per register (series shorter than 65,536 steps)
    def _cnt_layout(self):
        keys = [("U", k) for k in range(3) if self.sv.cnt[k]]
        if self.sv.ppv:
            keys.append(("P", 0))
        return {key: (i // 2, bool(i % 2)) for i, key in enumerate(keys)}

    def n_cnt_regs(self) -> int:
        return max(1, (len(self._cnt_layout()) + 1) // 2)

    def th(self, r: int, col: int) -> str:
        return f"th{r}_{self.colpos[col]}"

    # -- sieve accumulators of row r fed with `out` (previous value `prev`) ----
    def _sieve(self, L, r, out, prev, masked: bool):
        sv = self.sv
        cregs = self._cnt_layout()
        act = f', "r"(act{r})' if masked else ""

        def pred_open(nin):
            # predicate q = lane is active (masked body only)
            return [f"setp.ne.b32 q, %{nin}, 0;"] if masked else []

        def unit(k, val):
            reg, hi16 = cregs[("U", k)]
            inc = "0x10000" if hi16 else "1"
            outs = [f'"+r"(cn{r}_{reg})']
            if sv.avg[k]:
                outs.append(f'"+d"(sm{k}_{r})')
            iv = len(outs)
            ins = [f'"d"({val})', f'"d"({self.th(r, _jit._COL_U[k][0])})']
            if sv.hi:
                ins.append(f'"d"({self.th(r, _jit._COL_U[k][1])})')
            nin = iv + len(ins)
            asm = ["{ .reg .pred p, q;"] + pred_open(nin)
            asm.append(f"setp.gt{'.and' if masked else ''}.f64 p, %{iv}, %{iv + 1}{', q' if masked else ''};")
            if sv.hi:
                asm.append(f"setp.le.and.f64 p, %{iv}, %{iv + 2}, p;")
            asm.append(f"@p add.u32 %0, %0, {inc};")
            if sv.avg[k] and masked:
                asm.append(f"@p add.rn.f64 %1, %1, %{iv};")
            elif sv.avg[k]:
                # sum += val * (p ? 1 : 0): one select less than a predicated add
                # (ptxas turns that into an add and two selects).  Exact: every value
                # of the unmasked steps is finite (running maxima of sums of finite
                # inputs) and the sum itself is never -0.0
                asm.append("{ .reg .f64 m; selp.f64 m, 0d3FF0000000000000, 0d0000000000000000, p;")
                asm.append(f"fma.rn.f64 %1, %{iv}, m, %1; }}")
            asm.append("}")
            L.append('asm("' + " ".join(asm) + '" : ' + ", ".join(outs) + " : "
                     + ", ".join(ins) + act + ");")

        if sv.cnt[0]:
            unit(0, out)
        if sv.cnt[1] or sv.cnt[2]:
            # zero padding of the increments (fruits/cache.py:8-13): the first
            # increment of a row is 0.0, so is the first second increment
            if masked:
                L.append(f"const double d{r} = fst{r} ? 0.0 : __dadd_rn({out}, -{prev});")
            else:
                L.append(f"const double d{r} = __dadd_rn({out}, -{prev});")
            if sv.cnt[1]:
                unit(1, f"d{r}")
            if sv.cnt[2]:
                if masked:
                    L.append(f"const double dd{r} = fst{r} ? 0.0 : __dadd_rn(d{r}, -d1_{r});")
                    L.append(f"d1_{r} = act{r} ? d{r} : d1_{r};")
                else:
                    L.append(f"const double dd{r} = __dadd_rn(d{r}, -d1_{r}); d1_{r} = d{r};")
                unit(2, f"dd{r}")
        if sv.ppv:
            reg, hi16 = cregs[("P", 0)]
            inc = "0x10000" if hi16 else "1"
            asm = ["{ .reg .pred p, q;"] + pred_open(3)
            asm.append(f"setp.ge{'.and' if masked else ''}.f64 p, %1, %2{', q' if masked else ''};")
            asm.append(f"@p add.u32 %0, %0, {inc}; }}")
            L.append('asm("' + " ".join(asm) + f'" : "+r"(cn{r}_{reg}) : "d"({out}), '
                     f'"d"({self.th(r, _jit._COL_PPV)})' + act + ");")
        for on, arr, cmp_, cols in ((sv.mx, "mx", "gt", _jit._COL_MAX),
                                    (sv.mn, "mn", "lt", _jit._COL_MIN)):
            if not on:
                continue
            ins = [f'"d"({out})']
            if sv.mmb:
                ins += [f'"d"({self.th(r, cols[0])})', f'"d"({self.th(r, cols[1])})']
            nin = 1 + len(ins)
            asm = ["{ .reg .pred p, q;"] + pred_open(nin)
            asm.append(f"setp.{cmp_}{'.and' if masked else ''}.f64 p, %1, %0{', q' if masked else ''};")
            if sv.mmb:
                asm.append("setp.gt.and.f64 p, %1, %2, p;")
                asm.append("setp.le.and.f64 p, %1, %3, p;")
            asm.append("selp.f64 %0, %1, %0, p; }")
            L.append('asm("' + " ".join(asm) + f'" : "+d"({arr}{r}) : ' + ", ".join(ins) + act + ");")

    # -- one time step over all rows ---------------------------------------------
    def step(self, masked: bool) -> list:
        p = self.p
        R = p.rows
        L = []
        # phase 1: shuffles read the state before this step
        if p.rot:
            L.append(f"const double rot = __shfl_sync(0xffffffffu, S{R - 1}, lm1);")
        for r in range(R):
            for qi, q in enumerate(p.irregular[r]):
                L.append(f"const double t{r}_{qi} = __shfl_sync(0xffffffffu, S{q}, isrc{r}_{qi});")
        # phase 2, last row first: row r reads S[r-1] of its own lane before
        # row r-1 is updated
        for r in reversed(range(R)):
            if r > 0 and p.same[r]:
                base = f"S{r - 1}"
            elif r == 0 and p.rot:
                base = "rot"
            else:
                base = None
            others = bool(p.irregular[r]) or p.roots[r]
            if base is not None and not others:
                par = base
            else:
                L.append(f"double par{r} = {base or '0.0'};")
                for qi in range(len(p.irregular[r])):
                    L.append(f"par{r} = itake{r}_{qi} ? t{r}_{qi} : par{r};")
                if p.roots[r] and base is not None:
                    L.append(f"par{r} = root{r} ? 0.0 : par{r};")
                par = f"par{r}"
            if masked:
                L.append(f"const int tl{r} = s - skew{r};")
                L.append(f"const int act{r} = (unsigned)tl{r} < (unsigned)T;")
                L.append(f"const bool fst{r} = tl{r} == 0;")
            expr = par
            for j in range(p.npairs[r]):
                expr = f"fma(e{r}_{j}, *xp{r}_{j}, {expr})"
            L.append(f"const double w{r} = {expr};")
            L.append(f"const double q{r} = S{r};")
            if masked:
                L.append(f"S{r} = (act{r} && w{r} > S{r}) ? w{r} : S{r};")
            else:
                L.append(f"S{r} = (w{r} > S{r}) ? w{r} : S{r};")      # keeps S on NaN like the reference's loop
            self._sieve(L, r, f"S{r}", f"q{r}", masked)
        for r in range(R):
            for j in range(p.npairs[r]):
                L.append(f"xp{r}_{j}++;")
        return L

    def epilogue(self, r: int) -> list:
        sv = self.sv
        cregs = self._cnt_layout()
        nf = len(sv.feats)

        def count(key):
            reg, hi16 = cregs[key]
            return f"(cn{r}_{reg} >> 16)" if hi16 else f"(cn{r}_{reg} & 0xffffu)"

        L = [f"if (emit{r} != 0xffff) {{",
             f"    double *o = a.out + (size_t)n * a.out_ld + a.col0 + (size_t)emit{r} * {nf};"]
        for f, (kind, arg) in enumerate(sv.feats):
            if kind == be.FEAT_CNT:
                val = f"(double){count(('U', arg))}"
            elif kind == be.FEAT_AVG:
                c = count(("U", arg))
                val = f"({c} ? __ddiv_rn(sm{arg}_{r}, (double){c}) : 0.0)"
            elif kind == be.FEAT_PPV:
                val = f"__ddiv_rn((double){count(('P', 0))}, (double)T)"
            elif kind == be.FEAT_MAX:
                val = f"(mx{r} == D_NINF ? 0.0 : mx{r})"
            elif kind == be.FEAT_MIN:
                val = f"(mn{r} == D_INF ? 0.0 : mn{r})"
            else:
                val = f"S{r}"
            L.append(f"    put(o + {f}, fin({val}, a.sanitize & 1), mc);")
        L.append("}")
        return L

    def source(self, minb: int) -> str:
        p, sv = self.p, self.sv
        R, NB = p.rows, self.nb
        node_w, pair_w, irr_w = p.tables()
        nthr = max(1, len(p.trie.emits) * self.ntc)
        ncr = self.n_cnt_regs()
        src = []
        A = src.append
        A("// generated by fruits_b200/_jit_chain.py -- do not edit")
        A(f"#define R {R}")
        A(f"#define NB {NB}")
        A(f"#define SPC {self.spc}")
        A("#define NT (32 * NB * SPC)")
        A(f"#define NROW {self.nrow}")
        A(f"#define PAD {self.pad}")
        A("#define D_INF __longlong_as_double(0x7ff0000000000000LL)")
        A("#define D_NINF __longlong_as_double(0xfff0000000000000LL)")
        A("struct Args { const double *X; const double *E; double *out; long long n, d, t, e_ld, out_ld, col0; int sanitize; const int *cut; };")
        A(f"__constant__ double TH[{nthr}];")
        A(f"__device__ const unsigned NODE[{NB * R * 32}] = {{" + ",".join(map(str, node_w)) + "};")
        npmax = max(1, max(p.npairs))
        flat_pairs = []
        for pw in pair_w:
            flat_pairs += pw + [0] * (npmax - len(pw))
        A(f"__device__ const unsigned short PAIR[{NB * R * 32 * npmax}] = {{"
          + ",".join(map(str, flat_pairs)) + "};")
        nimax = max(1, max(len(q) for q in p.irregular))
        flat_irr = []
        for iw in irr_w:
            flat_irr += iw + [0] * (nimax - len(iw))
        A(f"__device__ const unsigned short IRR[{NB * R * 32 * nimax}] = {{"
          + ",".join(map(str, flat_irr)) + "};")
        A("// a.sanitize bit 1: a.out is an NVSwitch multicast mapping (multimem.* only)")
        A("__device__ __forceinline__ void put(double *p, double v, bool mc) {")
        A('    if (mc) asm volatile("multimem.st.weak.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");')
        A("    else *p = v; }")
        A("__device__ __forceinline__ double fin(double v, int sanitize) {")
        A("    if (!sanitize) return v;")
        A("    if (v != v) return 0.0;")
        A("    if (v == D_INF) return __longlong_as_double(0x7fefffffffffffffLL);")
        A("    if (v == D_NINF) return __longlong_as_double(0xffefffffffffffffLL);")
        A("    return v; }")
        A(f'extern "C" __global__ void __launch_bounds__(NT, {minb}) fb_jit_slice(const Args a)')
        A("{")
        A("    extern __shared__ __align__(16) double smem[];")
        A("    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;")
        A("    const int blk = warp % NB, sl = warp / NB;")
        A("    const int T = (int)a.t;")
        A("    const int XLD = T + 2 * PAD;                 // padded row of one staged dimension")
        A("    const long long n0 = (long long)blockIdx.x * SPC;")
        # ---- staging of the prepared series of this CTA ----
        A("    // whole series in shared memory: [SPC][NROW][PAD | T | PAD], pads are zero")
        A("    for (int i = threadIdx.x; i < SPC * NROW * XLD; i += NT) {")
        A("        const int sr = i / XLD, pos = i - sr * XLD;")
        A("        const int s_ = sr / NROW, r = sr - s_ * NROW;")
        A("        const int t = pos - PAD;")
        A("        const long long n = n0 + s_;")
        A("        double v = 0.0;")
        A("        if (t >= 0 && t < T && n < a.n) {")
        A("            int raw, inc;")
        A("            switch (r) {")
        for u, (raw, inc) in enumerate(self.dims):
            A(f"            case {u}: raw = {raw}; inc = {int(bool(inc))}; break;")
        A("            default: raw = 0; inc = 0; break;")
        A("            }")
        A("            const double *row = a.X + ((size_t)n * a.d + raw) * (size_t)T;")
        A("            v = row[t];")
        A("            if (inc) v = (t > 0) ? __dadd_rn(v, -row[t - 1]) : 0.0;")
        A("        }")
        A("        smem[i] = v;")
        A("    }")
        A("    __syncthreads();")
        A("    const long long n = n0 + sl;")
        A("    if (n >= a.n) return;")
        A("    const int lm1 = (lane + 31) & 31; (void)lm1;")
        # ---- per-lane constants and state ----
        for r in range(R):
            A(f"    const unsigned nw{r} = NODE[(blk * R + {r}) * 32 + lane];")
            A(f"    const unsigned emit{r} = nw{r} & 0xffffu;")
            A(f"    const int skew{r} = (int)((nw{r} >> 16) & 0xffu);")
            A(f"    const bool root{r} = (nw{r} >> 24) & 1u; (void)root{r};")
            for j in range(p.npairs[r]):
                A(f"    const unsigned pw{r}_{j} = PAIR[((blk * R + {r}) * 32 + lane) * {npmax} + {j}];")
                A(f"    const double e{r}_{j} = (double)(int)(signed char)(pw{r}_{j} >> 8);")
                A(f"    const double *xp{r}_{j} = smem + ((sl * NROW + (int)(pw{r}_{j} & 0xffu)) * XLD + PAD - skew{r});")
            for qi in range(len(p.irregular[r])):
                A(f"    const unsigned iw{r}_{qi} = IRR[((blk * R + {r}) * 32 + lane) * {nimax} + {qi}];")
                A(f"    const int isrc{r}_{qi} = (int)(iw{r}_{qi} & 31u);")
                A(f"    const bool itake{r}_{qi} = (iw{r}_{qi} >> 8) & 1u;")
            for c in self.cols:
                # lanes without an emission never write: any threshold will do
                A(f"    const double {self.th(r, c)} = TH[(emit{r} == 0xffffu ? 0u : emit{r}) * {self.ntc} + {self.colpos[c]}];")
            A(f"    double S{r} = D_NINF;")
            for j in range(ncr):
                A(f"    unsigned cn{r}_{j} = 0u;")
            for k in range(3):
                if sv.avg[k]:
                    A(f"    double sm{k}_{r} = 0.0;")
            if sv.cnt[2]:
                A(f"    double d1_{r} = 0.0;")
            if sv.mx:
                A(f"    double mx{r} = D_NINF;")
            if sv.mn:
                A(f"    double mn{r} = D_INF;")
        # ---- time loop: fill (masked), steady (unmasked), drain (masked) ----
        A(f"    const int total = T + {p.max_skew};")
        # (the deepest lanes take their first step at s = max_skew: still masked)
        A(f"    const int fill_end = min({p.max_skew + 1}, total);")
        A(f"    const int steady_end = max(T, fill_end);")
        A("    int s = 0;")
        masked = ["        " + ln for ln in self.step(True)]
        plain = ["        " + ln for ln in self.step(False)]
        A("#pragma unroll 1")
        A("    for (; s < fill_end; s++) {")
        src.extend(masked)
        A("    }")
        A(f"#pragma unroll {self.unroll}")
        A("    for (; s < T; s++) {")
        src.extend(plain)
        A("    }")
        A("    s = steady_end;")
        A("#pragma unroll 1")
        A("    for (; s < total; s++) {")
        src.extend(masked)
        A("    }")
        # ---- epilogue ----
        A("    const bool mc = (a.sanitize & 2) != 0;")
        for r in range(R):
            for ln in self.epilogue(r):
                A("    " + ln)
        A("}")
        return "\n".join(src) + "\n"

    def smem_bytes(self, t: int) -> int:
        return self.spc * self.nrow * (t + 2 * self.pad) * 8


# ---------------------------------------------------------------------------
# run time
# ---------------------------------------------------------------------------

DEFAULT_OPTS = {"rows": 3, "unroll": 4, "minb": 1, "warps": 4, "regs": 128}


def options() -> dict:
    """Generator options; ``FRUITS_B200_CHAIN_OPTS="rows=4,unroll=1"`` overrides."""
    opts = dict(DEFAULT_OPTS)
    for item in filter(None, os.environ.get("FRUITS_B200_CHAIN_OPTS", "").split(",")):
        key, val = item.split("=")
        if key not in opts:
            raise ValueError(f"unknown chain kernel option {key!r}")
        opts[key] = int(val)
    return opts


def suitable(trie, semiring: int, weight_mode: int) -> bool:
    """Unweighted Arctic tries (the alternating-sign chains of the experiment
    scripts); everything else keeps its route."""
    return semiring == be.SEMIRING_ARCTIC and weight_mode == be.WEIGHT_NONE


def chain_like(trie) -> bool:
    """Long words with few branches (on average at least 8 nodes per leaf): the
    lane-per-node layout wires almost every parent for free.  Bushy tries are
    better served by the thread-per-series kernel when it can hold them."""
    like = trie.__dict__.get("_chain_like")
    if like is None:
        leaves = sum(1 for n in trie.nodes if not n.children)
        like = trie._chain_like = len(trie.nodes) >= 8 * max(leaves, 1)
    return like


@dataclass
class GeneratedChain:
    source: str
    em: ChainEmitter
    max_regs: int

    def digest(self) -> str:
        h = hashlib.sha256(f"chain v{CHAIN_VERSION} r{self.max_regs}\n".encode())
        h.update(self.source.encode())
        return h.hexdigest()[:24]


def generate(trie, semiring: int, weight_mode: int, sieves, dims: list, opts: dict = None):
    if not suitable(trie, semiring, weight_mode):
        raise NotImplementedError("the chain kernel covers unweighted Arctic plans")
    opts = options() if opts is None else opts
    if sieves.rank2:
        raise NotImplementedError("XPI / LPI / CUR / CPV accumulators: thread-per-series kernel")
    if len(dims) > 200 or any(abs(e) > 127 for n in trie.nodes for e in n.expo):
        raise NotImplementedError("letter outside the chain kernel's tables")
    if trie.max_depth > 250:
        raise NotImplementedError("word too long for the chain kernel")
    prog = ChainProgram(trie, sieves, opts["rows"])
    nb = len(prog.blocks)
    spc = max(1, opts["warps"] // nb)
    em = ChainEmitter(prog, dims, spc, opts["unroll"])
    if len(trie.emits) * em.ntc * 8 > 60 * 1024:
        raise NotImplementedError("threshold table exceeds the constant bank")
    nt = 32 * nb * spc
    if nt > 1024:
        raise NotImplementedError("too many blocks per series")
    warps = -(-(nt // 32) // 4) * 4
    max_regs = min(opts["regs"], 255, (65536 // (warps * 32 * opts["minb"])) // 8 * 8)
    return GeneratedChain(em.source(opts["minb"]), em, max_regs)


class JitChain:
    """The loaded chain kernel of one slice."""

    _loaded: dict = {}

    def __init__(self, gen: GeneratedChain) -> None:
        self.gen, self.em = gen, gen.em
        path = os.path.join(_jit.CACHE_DIR, gen.digest() + ".cubin")
        if os.path.exists(path):
            with open(path, "rb") as f:
                cubin = f.read()
        else:
            cubin = _jit._nvrtc(gen.source, "fb_jit_chain.cu", False, gen.max_regs)
            try:
                os.makedirs(_jit.CACHE_DIR, exist_ok=True)
                tmp = path + f".tmp{os.getpid()}"
                with open(tmp, "wb") as f:
                    f.write(cubin)
                os.replace(tmp, path)
            except OSError:
                pass
        handle = ctypes.c_void_p()
        be.check(be.lib().fb_jit_load(cubin, len(cubin), ctypes.byref(handle)))
        self.handle = handle
        em = gen.em
        self.geo = FbJitChainGeometry(em.nb, em.spc, em.nrow, em.pad)
        self.cols = list(em.cols)

    @classmethod
    def load(cls, gen: GeneratedChain, cached_only: bool = False) -> "JitChain":
        key = (gen.digest(), _device_index())
        obj = cls._loaded.get(key)
        if obj is None:
            if cached_only and not os.path.exists(os.path.join(_jit.CACHE_DIR, key[0] + ".cubin")):
                raise _jit.NotCompiled(key[0])
            obj = cls(gen)
            cls._loaded[key] = obj
        return obj

    def fits(self, t: int) -> bool:
        return self.em.smem_bytes(t) <= 200 * 1024

    def n_launches(self, n_series: int = 0, length: int = 0) -> int:
        return 1

    def launch(self, X, extra, extra_ld, thr_compact, out, col0, sanitize, multicast=None,
               cuts=None) -> None:
        batch = be.FbBatch()
        batch.X = X.data_ptr()
        batch.n, batch.d, batch.t = X.shape
        n_thr = 0 if thr_compact is None else thr_compact.numel()
        be.check(be.lib().fb_jit_chain_features(
            self.handle, ctypes.byref(self.geo), ctypes.byref(batch), be.ptr(thr_compact), n_thr,
            out.data_ptr() if multicast is None else int(multicast[0]),
            out.stride(0) if multicast is None else int(multicast[1]), int(col0),
            int(bool(sanitize)) | (2 if multicast is not None else 0), be.stream_ptr()))


def _device_index() -> int:
    import torch
    return torch.cuda.current_device() if torch.cuda.is_available() else -1
