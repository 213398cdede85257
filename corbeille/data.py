"""Datasets of the harness (reference:
``experiments/corbeille/corbeille/data.py``): readers for the layouts of
timeseriesclassification.com -- univariate ``<name>/<name>_TRAIN.txt`` /
``_TEST.txt`` (first column = class label) and multivariate ``.arff`` with an
``.npy`` cache beside it -- returned as ``float64[n, n_dims, length]``, the
layout ``fruits.Fruit.fit`` takes; the synthetic ``multisine`` generator and the
resampling helpers (``implant_stuttering``, ``lengthen``, ``downsample``,
``upsample``).  All of it is host-side data handling ahead of the GPU path; the
random helpers draw from the global numpy RNG in the reference's order."""
import os
from typing import Callable, Generator, Optional, Sequence

import numpy as np

Dataset = tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]


def replace_nan(X: np.ndarray, value: Optional[float] = None) -> np.ndarray:
    """NaNs become ``value`` or, if none is given, the last observed value of
    the series (0 at the first time step) -- reference :125-147, here as one
    vectorised forward fill instead of a Python loop over the NaN positions."""
    if value is not None:
        return np.nan_to_num(X, nan=value)
    X = np.asarray(X, dtype=np.float64)
    nan = np.isnan(X)
    if not nan.any():
        return X.copy()
    t = X.shape[-1]
    # index of the last non-NaN position at or before every time step (-1: none)
    last = np.where(nan, -1, np.arange(t))
    last = np.maximum.accumulate(last, axis=-1)
    filled = np.take_along_axis(X, np.maximum(last, 0), axis=-1)
    return np.where(last < 0, 0.0, filled)


def _read_txt(path: str) -> tuple[np.ndarray, np.ndarray]:
    with open(path) as f:
        delimiter = "," if "," in f.readline() else None
    raw = np.loadtxt(path, delimiter=delimiter, ndmin=2)
    return raw[:, None, 1:].astype(np.float64), raw[:, 0].astype(np.int32)


def load(path: str, univariate: bool = True, cache: bool = True,
         keep_nan: bool = False) -> Dataset:
    """-> ``(X_train, y_train, X_test, y_test)`` of the dataset folder ``path``
    (reference :150-195; ``cache`` only concerns the .arff branch there)."""
    if not univariate:
        return _load_arff(path, cache, keep_nan)
    path = path.rstrip("/")
    name = os.path.basename(path)
    X_train, y_train = _read_txt(os.path.join(path, f"{name}_TRAIN.txt"))
    X_test, y_test = _read_txt(os.path.join(path, f"{name}_TEST.txt"))
    if not keep_nan:
        X_train, X_test = replace_nan(X_train), replace_nan(X_test)
    return X_train, y_train, X_test, y_test


def load_all(path: str, univariate: bool = True, cache: bool = True, keep_nan: bool = False,
             datasets: Optional[Sequence[str]] = None) -> Generator:
    """Yield ``(name, X_train, y_train, X_test, y_test)`` for every dataset
    folder in ``path``, sorted by name (reference :270-301)."""
    for folder in sorted(os.listdir(path)):
        full = os.path.join(path, folder)
        if os.path.isdir(full) and (datasets is None or folder in datasets):
            yield (folder,) + load(full, univariate=univariate, cache=cache, keep_nan=keep_nan)


def _read_arff(path: str):
    """``(series[n][n_dims][length] as nested tuples, labels[n])`` of one
    multivariate ``.arff`` file (relational attribute + class attribute)."""
    from scipy.io import arff
    with open(path, "r", encoding="utf8") as f:
        rows, _ = arff.loadarff(f)
    return [row[0].tolist() for row in rows], [row[1] for row in rows]


def _load_arff(path: str, cache: bool, keep_nan: bool) -> Dataset:
    """Multivariate branch of ``load`` (reference :197-267): ``<name>_TRAIN.arff``
    / ``_TEST.arff``; class labels are numbered in order of first appearance,
    train before test; with ``cache`` the four arrays are kept as
    ``<name>_XTRAIN.npy`` ... beside the files and read from there next time."""
    name = os.path.basename(os.path.normpath(path))
    stem = os.path.join(path, name)
    parts = ("_XTRAIN", "_yTRAIN", "_XTEST", "_yTEST")
    if cache and os.path.isfile(stem + "_XTRAIN.npy"):
        X_train, y_train, X_test, y_test = (np.load(stem + part + ".npy") for part in parts)
    else:
        train, train_labels = _read_arff(stem + "_TRAIN.arff")
        test, test_labels = _read_arff(stem + "_TEST.arff")
        X_train = np.array(train, dtype=np.float64).reshape(
            len(train), len(train[0]), len(train[0][0]))
        X_test = np.array(test, dtype=np.float64).reshape(
            len(test), len(test[0]), len(test[0][0]))
        number: dict = {}
        for label in train_labels + test_labels:
            number.setdefault(label, len(number))
        y_train = np.array([number[label] for label in train_labels], dtype=np.int32)
        y_test = np.array([number[label] for label in test_labels], dtype=np.int32)
        if cache:
            for part, array in zip(parts, (X_train, y_train, X_test, y_test)):
                np.save(stem + part, array)
    if not keep_nan:
        X_train, X_test = replace_nan(X_train), replace_nan(X_test)
    return X_train, y_train, X_test, y_test


def _class_sizes(total: int, n_classes: int) -> list:
    """``total`` samples over ``n_classes`` as evenly as the reference does it
    (:68-81: the remainder goes to classes ``remain % n_classes``, counting down)."""
    sizes = [total // n_classes] * n_classes
    for remain in range(total - sum(sizes), 0, -1):
        sizes[remain % n_classes] += 1
    return sizes


def multisine(train_size: int = 100, test_size: int = 1000, length: int = 100,
              n_classes: int = 2, used_sines: int = 3,
              coefficients: Optional[np.ndarray] = None,
              noise: Optional[Callable[[], float]] = None) -> Dataset:
    """Synthetic dataset: every class is a sum of ``used_sines`` sine waves
    ``amplitude * sin(frequency * x + phase)`` on ``[0, 2 pi]`` (random
    ``coefficients[n_classes, used_sines, 3]`` in ``[0, 2)`` unless given; a
    frequency may be a callable of ``x``), every sample its class model plus noise
    (N(0, 0.5) unless ``noise()`` is given) -- reference :25-123, same draws."""
    grid = np.linspace(0, 2 * np.pi, num=length)
    if coefficients is None:
        coefficients = 2 * np.random.rand(n_classes, used_sines, 3)

    def model(coeff):
        def value(x):
            total = 0.
            for amplitude, frequency, phase in coeff:
                f = frequency(x) if callable(frequency) else frequency
                total += amplitude * np.sin(f * x + phase)
            return total
        return np.vectorize(value)(grid)

    models = [model(coefficients[c]) for c in range(n_classes)]

    def draw(total):
        X, y = np.zeros((total, length)), np.zeros(total)
        row = 0
        for c, size in enumerate(_class_sizes(total, n_classes)):
            for _ in range(size):
                wobble = (np.random.normal(0, 0.5, length) if noise is None
                          else np.array([noise() for _ in range(length)]))
                X[row] = models[c] + wobble
                y[row] = c
                row += 1
        return X[:, np.newaxis, :], y

    X_train, y_train = draw(train_size)
    X_test, y_test = draw(test_size)
    return X_train, y_train, X_test, y_test


def implant_stuttering(X: np.ndarray, stutter_length: float = 0.1) -> np.ndarray:
    """Every series grows by ``int(stutter_length * length)`` steps: at random
    positions a value is repeated a random number of times, everything behind it
    moves back (reference :311-365, same draws: a run length, then a position)."""
    n, d, t = X.shape
    extra = int(stutter_length * t)
    out = np.zeros((n, d, t + extra))
    out[:, :, :t] = X
    for i in range(n):
        for j in range(d):
            added, floor = 0, 0          # steps added so far; end of the last stutter
            while added < extra:
                run = np.random.randint(1, extra - added + 1)
                at = np.random.randint(floor + 1, t + extra)
                if at >= t + added - 1:
                    # behind the last original value: repeat that one to the end
                    out[i, j, t + added - 1:] = X[i, j, -1]
                    break
                tail = t - (at + 1 - added)
                out[i, j, at + run + 1:at + run + 1 + tail] = out[i, j, at + 1:at + 1 + tail]
                out[i, j, at + 1:at + run + 1] = out[i, j, at]
                added += run
                floor = at + run
    return out


def lengthen(X: np.ndarray, length: float = 0.1) -> np.ndarray:
    """The last value of every series repeated ``int(length * T)`` more times
    (reference :368-386)."""
    extra = int(length * X.shape[2])
    return np.concatenate((X, np.repeat(X[:, :, -1:], extra, axis=2)), axis=2).astype(np.float64)


def downsample(X: np.ndarray, resolution: float = 0.5) -> np.ndarray:
    """Every ``int(1 / resolution)``-th value (reference :389-402)."""
    return X[:, :, ::int(1 / resolution)]


def upsample(X: np.ndarray) -> np.ndarray:
    """Midpoints between neighbours woven into the series: length ``2 T - 1``
    (reference :405-417).  For one dimension this is plain linear interpolation;
    with ``d > 1`` dimensions the reference's Fortran-order reshape walks values
    and midpoints of ALL dimensions in one sequence (position ``2 j + s + 2 d k``
    for dimension ``j``, value / midpoint ``s``, step ``k``) and deals that
    sequence out over the dimensions again -- reproduced as it is."""
    n, d, t = X.shape
    midpoints = (X + np.roll(X, -1, axis=2)) / 2          # (the wrapped last one is cut below)
    woven = np.stack((X, midpoints), axis=-1)             # [n, d, t, 2]
    sequence = woven.transpose(0, 2, 1, 3).reshape(n, 2 * t, d)
    return np.ascontiguousarray(sequence.transpose(0, 2, 1)[:, :, :-1])
