"""Host side of the plan-specialised kernel (no GPU needed): partitioning of
the trie, source generation, NVRTC compilation and nvJitLink linking through
the C ABI."""
import numpy as np
import pytest

import fruits_b200 as fruits
import specs
from fruits_b200 import _backend as be
from fruits_b200 import _jit
from fruits_b200 import _jit_chain


def _program(name, si=0, **kw):
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    slc = fruit._slices[si]
    iss = slc._iss[0]
    feats, bhi, bmm = slc._fused_sieves()
    sieves = _jit.SieveSet.make(feats, bhi, bmm)
    trie = iss.trie()
    return trie, iss, sieves, _jit.Program(trie, iss.semiring._code, iss._weight_mode(), sieves, **kw)


@pytest.mark.parametrize("name,si", [("C1_readme", 0), ("C2_reduced", 0), ("C4_twi", 0),
                                     ("C5_sweep", 0)])
@pytest.mark.parametrize("budget,mult", [(70, 2), (150, 1)])
def test_partition_covers_every_emission_once(name, si, budget, mult):
    trie, _, _, prog = _program(name, si, reg_budget=budget, parts_multiple=mult)
    owned = [v for part in prog.parts for v in part.owned]
    emitted = [v for v, n in enumerate(trie.nodes) if n.emit >= 0]
    assert sorted(owned) == sorted(emitted)
    assert len(prog.parts) % mult == 0
    for part in prog.parts:
        have = set(part.snodes)
        assert set(part.owned) <= have
        for v in part.snodes:                     # closed under parents, parents first
            p = trie.nodes[v].parent
            assert p < 0 or (p in have and part.snodes.index(p) < part.snodes.index(v))
    assert 1.0 <= prog.overhead < 1.6


def test_sweep_partition_is_balanced():
    _, _, _, prog = _program("C5_sweep", reg_budget=70, parts_multiple=2)
    costs = [p.cost for p in prog.parts if p.owned]
    assert len(costs) == 50 and max(costs) <= 1.15 * (sum(costs) / len(costs))


def test_deep_arctic_chain_is_not_for_the_thread_per_series_kernel():
    trie, iss, sieves, _ = _program("C3_general", 1)
    with pytest.raises(NotImplementedError):
        _jit.generate(trie, iss.semiring._code, iss._weight_mode(), sieves,
                      [(d, 0) for d in trie.used_dims()], True, _jit.options())


@pytest.mark.parametrize("name,si,blocks,slots", [("C2_reduced", 1, 2, 192), ("C3_general", 1, 4, 384),
                                                  ("C4_twi", 1, 1, 96)])
def test_chain_layout_of_the_arctic_slices(name, si, blocks, slots):
    """Lane-per-node layout (fruits_b200/_jit_chain.py): every emission owned
    once, blocks closed under ancestors, parents before children, and the
    alternating-sign chains wire almost every parent to the previous row of
    the same lane (no instruction) or to the rotation shuffle."""
    trie, iss, sieves, _ = _program(name, si)
    prog = _jit_chain.ChainProgram(trie, sieves, 3)
    assert len(prog.blocks) == blocks and prog.n_slots == slots
    owned = [sl.node for b in prog.blocks for sl in b if sl.owned]
    assert sorted(owned) == sorted(v for v, n in enumerate(trie.nodes) if n.emit >= 0)
    irregular = 0
    for b in prog.blocks:
        assert len(b) <= 96
        pos = {}
        for i, sl in enumerate(b):
            par = trie.nodes[sl.node].parent
            assert (par < 0 and sl.parent == -1) or pos[par] == sl.parent
            pos[sl.node] = i
            irregular += sl.parent >= 0 and sl.parent != i - 1
    # (the second word of an alternating-sign pair branches off the first letter of the
    # first: that letter is duplicated in a free lane, so no parent needs an indexed shuffle)
    assert irregular == 0 and not any(prog.irregular)
    assert prog.max_skew == trie.max_depth - 1
    node_w, pair_w, irr_w = prog.tables()
    assert len(node_w) == len(pair_w) == len(irr_w) == blocks * 3 * 32
    emits = sorted(w & 0xffff for w in node_w if w & 0xffff != 0xffff)
    assert emits == list(range(len(trie.emits)))


def test_chain_kernel_compiles_without_gpu(tmp_path, monkeypatch):
    """C2 slice 1 (188-node arctic trie, seven sieves) -> CUDA source -> NVRTC."""
    monkeypatch.setattr(_jit, "CACHE_DIR", str(tmp_path))
    trie, iss, sieves, _ = _program("C2_reduced", 1)
    gen = _jit_chain.generate(trie, iss.semiring._code, iss._weight_mode(), sieves,
                              [(0, 0), (0, 1)])
    assert "fb_jit_slice" in gen.source and "__shfl_sync" in gen.source
    cubin = _jit._nvrtc(gen.source, "fb_jit_chain.cu", False, gen.max_regs)
    assert cubin[:4] == b"\x7fELF"
    # a masked copy of the step body for pipeline fill / drain and an unmasked one
    assert gen.source.count("for (; s <") == 3
    # Reals and weighted plans keep their routes
    trie0, iss0, sieves0, _ = _program("C2_reduced", 0)
    assert not _jit_chain.suitable(trie0, iss0.semiring._code, iss0._weight_mode())
    with pytest.raises(NotImplementedError):
        _jit_chain.generate(trie0, iss0.semiring._code, iss0._weight_mode(), sieves0, [(0, 0)])


def test_sibling_letters_share_products():
    """One multiplication per node: [112] reuses the product of [11]."""
    trie, iss, sieves, _ = _program("C5_sweep")
    gen = _jit.generate(trie, iss.semiring._code, iss._weight_mode(), sieves,
                        [(d, 0) for d in trie.used_dims()], True, _jit.options())
    em = gen.em
    muls = sum(sum(ln.count("__dmul_rn(") for ln in em.step(part)) for part in em.p.parts
               if part.owned)
    nodes = sum(len(part.snodes) for part in em.p.parts)
    # the reference multiplies 684 times per step (plus duplicated ancestors); sharing
    # leaves one product per node plus the few a part boundary separates from its sibling
    assert muls <= 1.2 * nodes and muls < 684, (muls, nodes)
    adds = sum(sum(1 for ln in em.step(part) if ln.startswith("S[")) for part in em.p.parts
               if part.owned)
    assert adds == nodes


def test_compile_and_link_without_gpu(tmp_path, monkeypatch):
    """README slice -> CUDA source -> NVRTC (one unit per part) -> nvJitLink."""
    monkeypatch.setattr(_jit, "CACHE_DIR", str(tmp_path))
    trie, iss, sieves, _ = _program("C1_readme")
    gen = _jit.generate(trie, iss.semiring._code, iss._weight_mode(), sieves,
                        [(d, 1) for d in trie.used_dims()], True, _jit.options())
    assert len(gen.parts) >= 1 and "fb_jit_slice" in gen.entry
    assert all("fb_part_" in src for src in gen.parts)
    cubin = _jit.build_cubin(gen)
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 10000
    assert (tmp_path / (gen.digest() + ".cubin")).exists()
    assert _jit.build_cubin(gen) == cubin         # second call: disk cache
    # same plan -> same digest; other preparateur descriptor -> other kernel
    gen2 = _jit.generate(trie, iss.semiring._code, iss._weight_mode(), sieves,
                         [(d, 0) for d in trie.used_dims()], True, _jit.options())
    assert gen2.digest() != gen.digest()


def test_compile_error_is_reported():
    import ctypes
    L = be.lib()
    cubin, size = ctypes.c_void_p(), ctypes.c_size_t()
    log = ctypes.create_string_buffer(4096)
    rc = L.fb_jit_compile(b"this is not CUDA", b"bad.cu", 0, 0, ctypes.byref(cubin),
                          ctypes.byref(size), log, len(log))
    assert rc != 0 and b"nvrtc" in L.fb_last_error().lower()
    assert b"error" in log.value.lower()


def test_route_rule(monkeypatch):
    monkeypatch.delenv("FRUITS_B200_JIT", raising=False)
    assert not _jit.enabled(100) and _jit.enabled(_jit.MIN_SERIES)
    monkeypatch.setenv("FRUITS_B200_JIT", "force")
    assert _jit.enabled(1)
    monkeypatch.setenv("FRUITS_B200_JIT", "0")
    assert not _jit.enabled(10 ** 6)
    assert np.isfinite(_jit.DEFAULT_OPTS["budget"])


@pytest.mark.parametrize("name,si,ppc,tt,ndims", [("C5_sweep", 0, 2, 16, 3), ("C4_twi", 0, 1, 16, 3),
                                                  ("C2_reduced", 0, 1, 8, 2), ("C1_readme", 1, 4, 16, 3)])
def test_staged_epilogue_writes_every_feature_column_once(name, si, ppc, tt, ndims):
    """The epilogue stages a warp's features in the group's tile buffers and
    writes them in chunks of consecutive columns: every feature column of every
    owned node exactly once, every chunk inside the staging row."""
    import re
    trie, iss, sieves, prog = _program(name, si, reg_budget=70, parts_multiple=ppc)
    dims = [(d, 0) for d in trie.used_dims()]
    em = _jit.Emitter(prog, dims, ppc, 8, True, tt)
    sw = em.staging_width()
    row = em.nrow * tt + 2
    assert 5 <= sw <= (row if ppc <= 2 else row // 2) and sw % 2 == 1
    nf = len(sieves.feats)
    written = []
    for part in prog.parts:
        if not part.owned:
            continue
        code = em.epilogue_code(part)
        slots = [int(m.group(1)) for ln in code for m in [re.match(r"stg\[lane \* \d+ \+ (\d+)\]", ln)] if m]
        assert max(slots) < sw - 1
        starts = [int(m.group(1)) for ln in code
                  for m in [re.search(r"a\.col0 \+ (\d+);", ln)] if m]
        counts = [int(m.group(1)) for ln in code for m in [re.search(r"e < (\d+); e \+= 32", ln)] if m]
        assert len(starts) == len(counts) and sum(counts) == len(slots) == len(part.owned) * nf
        for c0, cnt in zip(starts, counts):
            written += list(range(c0, c0 + cnt))
    assert sorted(written) == list(range(len(trie.emits) * nf))
    # without staging: one direct store per feature
    em0 = _jit.Emitter(prog, dims, ppc, 8, True, tt, stage=0)
    assert em0.staging_width() == 0
    direct = [ln for part in prog.parts if part.owned for ln in em0.epilogue_code(part)
              if ln.strip().startswith("put(o + ")]
    assert len(direct) == len(trie.emits) * nf


def test_rank2_sieves_compile_into_the_thread_per_series_kernel(tmp_path, monkeypatch):
    """XPI / LPI / CUR / CPV as accumulators of the
    generated kernel: source -> NVRTC -> nvJitLink without a GPU; the chain kernel
    and the generic kernel decline them."""
    import ctypes
    monkeypatch.setattr(_jit, "CACHE_DIR", str(tmp_path))
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended"}],
                        "sieves": [["NPI", {"q": [0.4, 1.0]}], ["XPI", {"q": [0.4, 1.0]}],
                                   ["LPI", {"inc": 2}], ["XPI", {"inc": 0, "q": [0.2, 0.9]}],
                                   ["CUR", {"q": [-1.0, 0.7]}], ["CPV", {}], ["PPV", {}], ["END", {}]]}]}
    fruit = specs.build_fruit(fruits, spec)
    slc = fruit._slices[0]
    feats, bhi, bmm = slc._fused_sieves()
    assert [k for k, _ in feats] == [be.FEAT_CNT, be.FEAT_XPI, be.FEAT_LPI, be.FEAT_XPI, be.FEAT_CUR,
                                     be.FEAT_CPV, be.FEAT_PPV, be.FEAT_END]
    sieves = _jit.SieveSet.make(feats, bhi, bmm)
    assert sieves.rank2 and sieves.xpi == (True, True, False) and sieves.lpi == (False, False, True)
    assert set(sieves.thr_cols()) >= {0, 1, 2, 3, 4, 5, 6, 7, 12, 13}
    trie = slc._iss[0].trie()
    gen = _jit.generate(trie, be.SEMIRING_REALS, be.WEIGHT_NONE, sieves,
                        [(d, 1) for d in trie.used_dims()], True, _jit.options())
    src = "\n".join(gen.parts)
    assert "XS1[" in src and "LL2[" in src and "SQ[" in src and "CPC[" in src
    assert _jit.build_cubin(gen)[:4] == b"\x7fELF"
    with pytest.raises(NotImplementedError):
        _jit_chain.generate(trie, be.SEMIRING_ARCTIC, be.WEIGHT_NONE, sieves, [(0, 0), (1, 0)])
    sp = be.FbSievePlan()
    sp.n_feats = len(feats)
    for f, (kind, arg) in enumerate(feats):
        sp.kind[f], sp.arg[f] = kind, arg
    assert be.lib().fb_slice_policy(ctypes.byref(sp), 0, 0) < 0


def test_bayesian_sums_compile_into_the_thread_per_series_kernel(tmp_path, monkeypatch):
    """Unweighted Bayesian (max, times) sums (fruits/iss/semiring.py:461-494) as a
    third semiring of the generated kernel: one multiplication / division per letter
    occurrence, then the running maximum; weighted ones are declined (scan kernel)."""
    monkeypatch.setattr(_jit, "CACHE_DIR", str(tmp_path))
    spec = {"slices": [{"iss": [{"words": ["[1][-2]", "[11][2]", "[2][22][1]"], "mode": "extended",
                                 "semiring": "bayesian"}],
                        "sieves": [["NPI", {"q": [0.4, 1.0]}], ["MAX", {}], ["END", {}]]}]}
    fruit = specs.build_fruit(fruits, spec)
    slc = fruit._slices[0]
    assert slc._is_fusable(2, None, 64)
    feats, bhi, bmm = slc._fused_sieves()
    sieves = _jit.SieveSet.make(feats, bhi, bmm)
    trie = slc._iss[0].trie()
    gen = _jit.generate(trie, be.SEMIRING_BAYESIAN, be.WEIGHT_NONE, sieves,
                        [(d, 0) for d in trie.used_dims()], True, _jit.options())
    src = "\n".join(gen.parts)
    assert "__ddiv_rn(" in src and "__dmul_rn(__dmul_rn(" in src and "fma(" not in src.split("epilogue")[0]
    assert _jit.build_cubin(gen)[:4] == b"\x7fELF"
    with pytest.raises(NotImplementedError):
        _jit.generate(trie, be.SEMIRING_BAYESIAN, be.WEIGHT_TOTAL, sieves,
                      [(d, 0) for d in trie.used_dims()], True, _jit.options())
    weighted = {"slices": [dict(spec["slices"][0], iss=[dict(spec["slices"][0]["iss"][0],
                                                             weighting=["Indices", {}])])]}
    assert not specs.build_fruit(fruits, weighted)._slices[0]._is_fusable(2, None, 64)


def test_chain_schedule_simulated_on_the_host_equals_the_oracle():
    """The tables of the chain kernel (position -> lane / row, parent wiring, skew)
    driven by a numpy simulation of its schedule -- every node works on
    t = step - (depth - 1) and reads its parent's value from before the step --
    reproduce the oracle's Arctic iterated sums bit for bit (C2 slice 1: letters
    +-x, so the FMA of the kernel is an exact addition here)."""
    from oracle import pipeline as orc
    desc = specs.SPECS["C2_reduced"]["slices"][1]["iss"][0]
    trie, iss, sieves, _ = _program("C2_reduced", 1)
    prog = _jit_chain.ChainProgram(trie, sieves, 3)
    rng = np.random.default_rng(3)
    X = rng.standard_normal((3, 2, 40)).cumsum(axis=2)
    n, _, T = X.shape
    want = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))        # [emit, n, T]
    got = np.full_like(want, np.nan)
    node_w, pair_w, irr_w = prog.tables()
    R = prog.rows
    for b, block in enumerate(prog.blocks):
        S = np.full((len(block), n), -np.inf)
        for s in range(T + prog.max_skew):
            old = S.copy()
            for i, sl in enumerate(block):
                r, lane = i % R, i // R
                w_ = node_w[(b * R + r) * 32 + lane]
                skew, root = (w_ >> 16) & 0xff, (w_ >> 24) & 1
                assert skew == trie.nodes[sl.node].depth - 1 and root == (sl.parent < 0)
                tl = s - skew
                if not 0 <= tl < T:
                    continue
                val = np.zeros(n) if root else old[sl.parent]
                for pw in pair_w[(b * R + r) * 32 + lane]:
                    e = ((pw >> 8) & 0xff) - (256 if (pw >> 8) & 0x80 else 0)
                    if e:
                        val = val + e * X[:, prog.used[pw & 0xff], tl]
                S[i] = np.maximum(S[i], val)
                emit = w_ & 0xffff
                if emit != 0xffff:
                    got[emit, :, tl] = S[i]
    assert np.array_equal(got, want)
