"""CPU restatement of the reference's remaining preparateurs.

TEST INFRASTRUCTURE ONLY (tests/, oracle/gen_golden.py): nothing under
``fruits_b200/`` imports this module.  Every function follows the cited lines
of the reference (``fruits/preparation/transform.py`` and ``filter.py``) in
plain numpy; ``fit`` consumes the global numpy RNG with the reference's calls
in the reference's order.  ``oracle/gen_golden.py preps2`` checks this file
against the real reference (bit for bit where the reference is plain numpy,
within 1e-12 where it is a numba ``fastmath`` loop) and freezes the
reference's outputs in ``tests/golden/preps2.npz``.

A preparateur is described as ``[name, kwargs]``; callables are named
(``"@sq"`` -> ``CALLABLES["sq"]``) so that the descriptions stay data.
"""
import numpy as np

from . import pipeline as orc

# functions a test may hand to FUN / SPE(function=) / RIN(width=)
CALLABLES = {
    "sq": lambda X: X * X,
    "cumsum_centered": lambda X: np.cumsum(X - X.mean(), axis=2),
    "cos": np.cos,
    "half": lambda length: length // 2,
}


def resolve(args):
    """kwargs with ``"@name"`` strings replaced by the named callables and
    lists under ``kernel`` turned into arrays."""
    out = {}
    for key, val in dict(args).items():
        if isinstance(val, str) and val.startswith("@"):
            val = CALLABLES[val[1:]]
        if key == "kernel" and val is not None:
            val = np.array(val, dtype=np.float64)
        out[key] = val
    return out


def _split_dims(n_in, n_out):
    quotient, remainder = divmod(n_in, n_out)
    return np.array([quotient + 1] * remainder + [quotient] * (n_out - remainder),
                    dtype=np.int32)


# ---------------------------------------------------------------------------
def fit_prep(desc, X):
    """State a fitted preparateur holds (dict); RNG draws as in the reference."""
    name, args = desc[0], resolve(desc[1])
    n, d, t = X.shape
    st = {}
    if name == "MAV":
        # transform.py:250-256
        w = args.get("width", 5)
        if isinstance(w, float):
            st["w"] = max(int(w * t), 1)
        elif w > 0:
            st["w"] = w
    elif name == "FFN":
        # transform.py:344-360
        h = 2 * d if args.get("d_hidden") is None else args["d_hidden"]
        st["w1"] = np.random.normal(loc=0, scale=1.0, size=(h, d))
        st["b"] = np.random.normal(loc=0, scale=1.0, size=(h,))
        st["w2"] = np.random.normal(loc=0, scale=1.0, size=(args.get("d_out", 1), h))
    elif name == "RIN":
        # transform.py:484-523
        if args.get("kernel") is not None:
            st["kernel"] = args["kernel"].copy()
            st["ndim"] = np.ones((d,), dtype=np.int32)
            st["dims"] = np.arange(d, dtype=np.int32)
            return st
        width = args.get("width", 1)
        width = width(t) if callable(width) else min(width, t - 1)
        out_dim = args.get("out_dim", -1)
        out_dim = out_dim if out_dim > 0 else d
        st["ndim"] = _split_dims(d, out_dim)
        st["dims"] = np.random.choice(d, size=d, replace=False).astype(np.int32)
        if args.get("force_sum_one", False):
            while True:
                kernel = np.random.uniform(-1., 1., size=(d, width))
                change = 1.0 - np.sum(kernel, axis=1)
                diff = 1.0 - np.abs(kernel)
                diffsum = np.sum(diff, axis=1)
                if np.sum(diffsum < 1e-5) > 0:
                    continue
                kernel += diff * (change / diffsum)[:, np.newaxis]
                break
        else:
            kernel = np.random.normal(size=(d, width))
            kernel -= np.mean(kernel, axis=1)[:, np.newaxis]
        st["kernel"] = kernel
    elif name == "RDW":
        # transform.py:590-599
        if args.get("dist", "dirichlet") == "dirichlet":
            alphas = np.max(np.mean(np.abs(X), axis=0), axis=1)
            alphas[alphas != 0] = alphas[alphas != 0] / np.max(alphas[alphas != 0])
            if np.sum(alphas == 0) >= 1:
                alphas += 1e-5
            st["weights"] = np.random.dirichlet(alphas)
        else:
            w = np.random.random(d)
            st["weights"] = w / np.sum(w)
    elif name == "JLD":
        # transform.py:687-722
        dim = args.get("dim", 0.99)
        if isinstance(dim, float):
            out_dim = int(24 * np.log(d) / (3 * dim**2 - 2 * dim**3)) + 1
        else:
            out_dim = dim
        if args.get("distribute", False):
            st["ndim"] = _split_dims(d, out_dim)
            st["dims"] = np.random.choice(d, size=d, replace=False).astype(np.int32)
            st["kernel"] = np.random.standard_normal(d)
        else:
            st["ndim"] = np.array(out_dim * [d], dtype=np.int32)
            st["dims"] = np.array(out_dim * list(range(d)), dtype=np.int32)
            st["kernel"] = np.random.standard_normal(d * out_dim)
        st["bias"] = (np.random.standard_normal(out_dim) if args.get("bias", False)
                      else np.zeros(out_dim))
    elif name == "QTC":
        st["quantile"] = np.quantile(X, args["q"])          # transform.py:987-988
    elif name == "DIL":
        # filter.py:33-53
        clusters = args.get("clusters")
        if clusters is not None:
            nclusters = int(clusters * t)
        else:
            upper = int(np.floor(t / 10.0))
            nclusters = 1 if upper <= 1 else np.random.randint(1, upper)
        if nclusters >= t:
            indices = np.arange(t)
        else:
            indices = np.sort(np.random.choice(t, size=nclusters, replace=False))
        lengths = []
        for i in range(nclusters):
            top = t - indices[i] if i == nclusters - 1 else indices[i + 1] - indices[i]
            lengths.append(np.random.randint(1, top + 1))
        st["indices"], st["lengths"] = indices, lengths
    elif name == "DOT":
        # filter.py:162-183
        n_given, first = args.get("n", 2), args.get("first")
        st["n"] = max(int(n_given * t), 1) if isinstance(n_given, float) else min(n_given, t)
        if isinstance(first, float):
            st["first"] = min(max(int(first * t), 1), t - 1)
        elif first is not None:
            st["first"] = min(first, t - 1)
        else:
            st["first"] = st["n"] - 1
    elif name == "PDD":
        # filter.py:238-251
        p = max(int(args.get("proportion", 0.5) * t), 1)
        points = max(int((1.0 - args.get("density", 0.1)) * t), 1)
        width = int(p / points)
        if points == t - width:
            points -= 1
        st["width"] = width
        st["indices"] = np.linspace(0, t - width, points, dtype="int")
    return st


# ---------------------------------------------------------------------------
def transform_prep(desc, st, X, cache=None):
    """Output of the fitted preparateur; ``cache`` (``pipeline.RawCache`` of
    the raw input) serves WIN and SPE(step_transform=...)."""
    name, args = desc[0], resolve(desc[1])
    n, d, t = X.shape
    if cache is None:
        cache = orc.RawCache(X)
    if name == "NRM":
        return orc.nrm(X, args.get("scale_dim", False))
    if name == "MAV":
        # transform.py:233-239
        w = st["w"]
        out = np.zeros_like(X)
        for k in range(w, t + 1):
            out[:, :, k - 1] = np.sum(X[:, :, k - w:k], axis=2) / w
        return out
    if name == "LAG":
        # transform.py:291-298
        out = np.zeros((n, 2 * d, 2 * t - 1))
        for i in range(d):
            out[:, 2 * i, 0::2] = X[:, i, :]
            out[:, 2 * i, 1::2] = X[:, i, 1:]
            out[:, 2 * i + 1, 0::2] = X[:, i, :]
            out[:, 2 * i + 1, 1::2] = X[:, i, :-1]
        return out
    if name == "FFN":
        # transform.py:362-376
        X_in = X - np.mean(X, axis=2)[:, :, np.newaxis] if args.get("center", True) else X
        temp = np.tensordot(st["w1"], X_in, axes=(1, 1)) + st["b"][:, np.newaxis, np.newaxis]
        out = np.tensordot(st["w2"], temp * (temp > 0), axes=(1, 0)).swapaxes(0, 1)
        return out * (out > 0) if args.get("relu_out", False) else out
    if name == "RIN":
        # transform.py:447-468, :525-544
        kernel, ndim, dims = st["kernel"], st["ndim"], st["dims"]
        w = kernel.shape[1]
        adaptive = args.get("adaptive_width", False)
        Xp = np.pad(X, ((0, 0), (0, 0), (w, 0))) if adaptive else X
        out = np.zeros((n, ndim.size, Xp.shape[2]))
        start = 0
        for new_dim in range(ndim.size):
            end = start + ndim[new_dim]
            for k in range(w, Xp.shape[2]):
                s = np.zeros(n)
                for j in range(start, end):
                    for l in range(k - w, k):
                        s += -Xp[:, dims[j], l] * kernel[j, l - k + w]
                    s += Xp[:, j, k]
                out[:, new_dim, k] = s
            start = end
        return out[:, :, w:] if adaptive else out
    if name == "RDW":
        with np.errstate(invalid="ignore"):
            return X ** st["weights"][np.newaxis, :, np.newaxis]      # transform.py:601-602
    if name == "JLD":
        # transform.py:651-670
        out = np.zeros((n, st["ndim"].size, t))
        start = 0
        for new_dim in range(st["ndim"].size):
            end = start + st["ndim"][new_dim]
            for j in range(start, end):
                out[:, new_dim, :] += X[:, st["dims"][j], :] * st["kernel"][j] + st["bias"][new_dim]
            start = end
        return out
    if name == "SPE":
        # transform.py:789-812
        freq, step = args["freq"], args.get("step_transform")
        max_length = args.get("max_length")
        if step is None:
            T = t if max_length is None else max_length
            range_ = np.arange(t) / (T**freq)
        else:
            range_ = cache.lsum(step)
            T = range_[:, -1:] if max_length is None else max_length
            range_ = range_ / (T**freq)
        fn = args.get("function")
        wave = np.sin(range_) if fn is None else fn(range_)
        wave = wave[np.newaxis, np.newaxis, :] if step is None else wave[:, np.newaxis, :]
        if args.get("operation", "multiplicative") == "multiplicative":
            return X * wave
        return X + wave
    if name == "RPE":
        # transform.py:859-875
        T = t if args.get("max_length") is None else args["max_length"]
        a = np.arange(t) / float(T)**args["freq"]
        out = np.zeros((n, 2, t))
        out[:, 0] = np.cos(a) * X[:, 0] - np.sin(a) * X[:, 1]
        out[:, 1] = np.sin(a) * X[:, 0] + np.cos(a) * X[:, 1]
        return out
    if name == "CTS":
        # transform.py:934-945
        s = args["s"]
        shift = max(1, int(s * t)) if 0 < s < 1 else int(s)
        Y = X.copy()
        if args.get("pseudo_shift", False):
            Y[:, :, :shift] = 0
        else:
            Y[:, :, :-shift] = Y[:, :, shift:]
            Y[:, :, -shift:] = Y[:, :, -1:]
        return Y
    if name == "QTC":
        # transform.py:990-1001
        q = st["quantile"]
        bound = q if args.get("bound") is None else args["bound"]
        if args.get("lower", False):
            return np.where(X < q, bound, X)
        return np.where(X > q, bound, X)
    if name == "FUN":
        return args["f"](X)                                           # transform.py:1038-1039
    if name == "DIL":
        # filter.py:55-61
        out = X.copy()
        for i, index in enumerate(st["indices"]):
            out[:, :, index:index + st["lengths"][i]] = 0
        return out
    if name == "WIN":
        # filter.py:97-113
        lo = cache.coquantile(args["start"], "L2")
        hi = cache.coquantile(args["end"], "L2")
        out = np.zeros_like(X)
        for i in range(n):
            out[i, :, lo[i] - 1:hi[i]] = X[i, :, lo[i] - 1:hi[i]]
        return out
    if name == "DOT":
        out = np.zeros(X.shape)                                       # filter.py:185-190
        out[:, :, st["first"]::st["n"]] = X[:, :, st["first"]::st["n"]]
        return out
    if name == "PDD":
        out = X.copy()                                                # filter.py:253-259
        for index in st["indices"]:
            out[:, :, index:index + st["width"]] = 0
        return out
    raise NotImplementedError(name)
