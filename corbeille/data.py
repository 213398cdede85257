"""Dataset reader of the harness (reference:
``experiments/corbeille/corbeille/data.py:125-301``): univariate datasets in
the ``<name>/<name>_TRAIN.txt`` / ``_TEST.txt`` layout of
timeseriesclassification.com (first column = class label), returned as
``float64[n, 1, length]`` -- the layout ``fruits.Fruit.fit`` takes.

The multivariate ``.arff`` reader of the reference is not mirrored (scipy's
arff parser, no GPU work involved); ``univariate=False`` raises."""
import os
from typing import Generator, Optional, Sequence

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
        raise NotImplementedError("only the univariate .txt layout is read")
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
