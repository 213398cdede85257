"""Compile ISS words into the device plan of the CUDA kernel.

Host-side metadata only (no numerics).  The reference recomputes every word
from its first letter (fruits/iss/semiring.py:142-158) and deduplicates only
the *emission* of prefixes (fruits/iss/cache.py:17-37).  Here every distinct
prefix becomes one node of a trie and is computed once; nodes are laid out as
``blocks x rows x 32 lanes`` -- one warp owns one block for one series
(``csrc/lns.cuh``).  A block is closed under ancestors, so warps never have to
talk to each other; ancestors needed by several blocks are duplicated as
non-emitting slots.
"""
import ctypes
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _backend as be


@dataclass
class _Node:
    parent: int            # node id or -1
    expo: tuple            # exponent vector (trailing zeros stripped)
    alpha: float           # float32 alpha of this level (0.0 if unweighted)
    depth: int             # 1-based
    emit: int = -1         # emission index or -1
    children: list = field(default_factory=list)


class Trie:
    """Prefix trie over (exponent vector, alpha) letters with the emission
    order of the reference (iss.py:49-65: words in order, each contributing
    its last ``plan[i]`` prefixes, shortest first)."""

    def __init__(self, words: Sequence, plan: Optional[Sequence[int]], weighted: bool) -> None:
        self.nodes: list = []
        self.emits: list = []          # emission index -> node id
        index: dict = {}
        for wi, word in enumerate(words):
            mat = [tuple(int(x) for x in el) for el in word]
            alphas = word.alpha if weighted else np.zeros(len(mat), dtype=np.float32)
            p = len(mat)
            n_emit = 1 if plan is None else plan[wi]
            parent = -1
            for k, el in enumerate(mat):
                expo = list(el)
                while expo and expo[-1] == 0:
                    expo.pop()
                key = (parent, tuple(expo), float(np.float32(alphas[k])))
                canon = index.get(key)
                if canon is None:
                    canon = self._add(parent, key[1], key[2], k + 1)
                    index[key] = canon
                if p - k <= n_emit:
                    # the same prefix can be emitted twice ("[12]" and "[21]"
                    # are different strings but the same letter): the second
                    # emission gets a leaf of its own
                    tgt = canon if self.nodes[canon].emit < 0 else \
                        self._add(parent, key[1], key[2], k + 1)
                    self.nodes[tgt].emit = len(self.emits)
                    self.emits.append(tgt)
                parent = canon
        self.max_depth = max(n.depth for n in self.nodes)

    def _add(self, parent, expo, alpha, depth) -> int:
        nid = len(self.nodes)
        self.nodes.append(_Node(parent, expo, alpha, depth))
        if parent >= 0:
            self.nodes[parent].children.append(nid)
        return nid

    def used_dims(self) -> list:
        # (a trie is not modified once built; every transform call asks)
        used = self.__dict__.get("_used")
        if used is None:
            dims = set()
            for n in self.nodes:
                dims.update(d for d, e in enumerate(n.expo) if e != 0)
            used = self._used = sorted(dims)
        return list(used)

    def dfs(self) -> list:
        order = []
        roots = [i for i, n in enumerate(self.nodes) if n.parent < 0]
        stack = list(reversed(roots))
        while stack:
            v = stack.pop()
            order.append(v)
            stack.extend(reversed(self.nodes[v].children))
        return order

    def subset(self, emit_lo: int, emit_hi: int) -> "Trie":
        """Trie restricted to the emissions [emit_lo, emit_hi) and their
        ancestors, emissions renumbered from 0."""
        sub = object.__new__(Trie)
        sub.nodes, sub.emits = [], [None] * (emit_hi - emit_lo)
        keep: dict = {}

        def take(v):
            if v in keep:
                return keep[v]
            n = self.nodes[v]
            par = take(n.parent) if n.parent >= 0 else -1
            nid = len(sub.nodes)
            e = n.emit - emit_lo if emit_lo <= n.emit < emit_hi else -1
            sub.nodes.append(_Node(par, n.expo, n.alpha, n.depth, e))
            if par >= 0:
                sub.nodes[par].children.append(nid)
            if e >= 0:
                sub.emits[e] = nid
            keep[v] = nid
            return nid

        for e in range(emit_lo, emit_hi):
            take(self.emits[e])
        sub.max_depth = max(n.depth for n in sub.nodes)
        return sub


def _encode_letter(expo, dim_index, arctic: bool):
    """-> (lo, hi, weight) letter words of ``struct fb_slot``."""
    if arctic:
        pairs = [(dim_index[d], e) for d, e in enumerate(expo) if e != 0]
        if len(pairs) > 8:
            raise NotImplementedError("arctic letters with more than 8 distinct dimensions")
        val = 0
        for i, (u, e) in enumerate(pairs):
            if not -16 <= e <= 15:
                raise NotImplementedError("arctic exponents beyond [-16, 15]")
            val |= ((u & 7) | ((e & 31) << 3)) << (8 * i)
        return val & 0xFFFFFFFF, (val >> 32) & 0xFFFFFFFF, len(pairs)
    occ = []
    for d, e in enumerate(expo):
        if e != 0:
            occ += [(dim_index[d]) | (8 if e < 0 else 0)] * abs(e)
    if len(occ) > 15:
        raise NotImplementedError("letters with more than 15 occurrences")
    val = 0
    for i, o in enumerate(occ):
        val |= o << (4 * i)
    return val & 0xFFFFFFFF, (val >> 32) & 0xFFFFFFFF, len(occ)


class DevicePlan:
    """``struct fb_iss_plan`` plus the device buffers it points to."""

    def __init__(self, trie: Trie, semiring: int, weight_mode: int, rows_max: int,
                 dim_desc: Optional[list] = None) -> None:
        self.trie = trie
        self.semiring = semiring
        self.weight_mode = weight_mode
        arctic = semiring == be.SEMIRING_ARCTIC
        used = trie.used_dims()
        if len(used) > be.FB_MAX_USED_DIMS:
            raise NotImplementedError(
                f"words reference {len(used)} distinct dimensions; the kernel "
                f"supports {be.FB_MAX_USED_DIMS}")
        if trie.max_depth > be.FB_RING - 64:
            raise NotImplementedError(
                f"words longer than {be.FB_RING - 64} letters are not supported")
        self.used_dims = used
        dim_index = {d: u for u, d in enumerate(used)}
        one = len(used)   # index of the constant-one row
        alphas = sorted({n.alpha for n in trie.nodes}) if weight_mode != be.WEIGHT_NONE else [0.0]
        if len(alphas) > be.FB_MAX_ALPHAS:
            raise NotImplementedError(
                f"more than {be.FB_MAX_ALPHAS} distinct alpha values in one ISS")
        aidx = {a: i for i, a in enumerate(alphas)}

        # ---- partition the DFS order into ancestor-closed blocks ----
        cap = 32 * rows_max
        if trie.max_depth > cap:
            raise NotImplementedError("word longer than one kernel block")
        order = trie.dfs()
        n_blocks = max(1, -(-len(order) // cap))
        while True:
            target = -(-len(order) // n_blocks)
            blocks, cur, cur_set, owned = [], [], set(), 0
            for v in order:
                chain, a = [], v
                while a >= 0:
                    chain.append(a)
                    a = trie.nodes[a].parent
                need = [a for a in reversed(chain) if a not in cur_set]
                if cur and (len(cur) + len(need) > cap or owned >= target):
                    blocks.append(cur)
                    cur, cur_set, owned = [], set(), 0
                    need = list(reversed(chain))
                for a in need:
                    cur.append((a, a == v))   # (node, owned by this block)
                    cur_set.add(a)
                owned += 1
            blocks.append(cur)
            if len(blocks) <= n_blocks or all(len(b) <= cap for b in blocks):
                break
            n_blocks += 1

        n_rows = max(-(-len(b) // 32) for b in blocks)
        self.n_blocks, self.n_rows = len(blocks), n_rows
        slots = np.zeros((len(blocks), n_rows, 32), dtype=be.SLOT_DTYPE)
        slots["parent"] = -1
        slots["emit"] = -1
        slots["depth"] = 1
        slots["letter_lo"] = 0
        row_pub = np.zeros(len(blocks), dtype=np.uint32)
        row_weight = np.zeros((len(blocks), n_rows), dtype=np.uint8)
        pad_lo = 0
        for i in range(8):
            pad_lo |= one << (4 * i)
        for bi, b in enumerate(blocks):
            members = {v for v, _ in b}
            enc = {}
            for v, own in b:
                n = trie.nodes[v]
                lo, hi, w = _encode_letter(n.expo, dim_index, arctic)
                has_child = any(c in members for c in n.children)
                enc[v] = (lo, hi, w, has_child, own)
            # heaviest letters first (rows become homogeneous in the number of
            # multiplications), nodes with children first among equals
            layout = sorted(enc, key=lambda v: (-enc[v][2], not enc[v][3], v))
            pos = {v: i for i, v in enumerate(layout)}
            for i, v in enumerate(layout):
                if enc[v][3]:
                    row_pub[bi] |= np.uint32(1 << (i // 32))
            flat = slots[bi].reshape(-1)
            for i, v in enumerate(layout):
                n = trie.nodes[v]
                lo, hi, w, has_child, own = enc[v]
                if not arctic:
                    # pad unused occurrences with the constant-one row
                    for k in range(w, 8):
                        lo |= one << (4 * k)
                    for k in range(max(w, 8), 16):
                        hi |= one << (4 * (k - 8))
                s = flat[i]
                s["letter_lo"], s["letter_hi"] = lo, hi
                s["parent"] = pos[n.parent] if n.parent >= 0 else -1
                s["emit"] = n.emit if own else -1
                s["depth"] = n.depth
                s["aidx"] = aidx.get(n.alpha, 0)
                paidx = aidx.get(trie.nodes[n.parent].alpha, 0) if n.parent >= 0 else 0
                s["weight"] = w
                s["flags"] = 1 | (2 if has_child else 0) | (paidx << 4)
                flat[i] = s
                row_weight[bi, i // 32] = max(row_weight[bi, i // 32], w)
            if not arctic:
                for i in range(len(layout), n_rows * 32):
                    flat[i]["letter_lo"] = pad_lo
        self.n_slots_used = sum(len(b) for b in blocks)

        dev = be.require_cuda()
        self._slots = torch.from_numpy(slots.view(np.uint8).reshape(-1)).to(dev)
        self._row_pub = torch.from_numpy(row_pub).to(dev)
        self._row_weight = torch.from_numpy(row_weight.reshape(-1)).to(dev)

        plan = be.FbIssPlan()
        plan.semiring, plan.weight_mode = semiring, weight_mode
        plan.n_blocks, plan.n_rows = len(blocks), n_rows
        plan.n_emit = len(trie.emits)
        plan.n_used_dims = len(used)
        plan.n_alphas = len(alphas)
        plan.max_depth = trie.max_depth
        for i, a in enumerate(alphas):
            plan.alphas[i] = a
        for u, d in enumerate(used):
            desc = dim_desc[d] if dim_desc is not None else (d, 0, 0)
            plan.dims[u].raw_dim, plan.dims[u].inc, plan.dims[u].std = desc
        plan.slots = self._slots.data_ptr()
        plan.row_pub = self._row_pub.data_ptr()
        plan.row_weight = self._row_weight.data_ptr()
        self.c = plan

    @property
    def n_emit(self) -> int:
        return len(self.trie.emits)

    def byref(self):
        return ctypes.byref(self.c)
