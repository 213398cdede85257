"""Timing / accuracy harness (reference:
``experiments/corbeille/corbeille/fruitifier.py:20-175``): fit a fruit on the
training series, extract features of both splits, classify with a ridge
classifier on standardised features, report (seconds, accuracy).  The timed
region is exactly the reference's -- ``fit`` + two ``transform`` calls with
host arrays in and out -- plus a device synchronisation, so the number
includes the host<->device copies."""
import os
import time
from collections.abc import Sequence
from typing import Callable, Optional, Union

import numpy as np

import fruits_b200 as fruits

from .data import load_all


def _default_classifier():
    from sklearn.linear_model import RidgeClassifierCV
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import FunctionTransformer, StandardScaler
    return Pipeline(steps=[
        ("scaler", StandardScaler()),
        ("nantonum", FunctionTransformer(np.nan_to_num)),
        ("ridge", RidgeClassifierCV(alphas=np.logspace(-3, 3, 10))),
    ])


def _sync() -> None:
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def fruitify(dataset, fruit: Union[fruits.Fruit, Callable[[np.ndarray, np.ndarray], fruits.Fruit]],
             classifier=None, mean_over_n_runs: int = 1) -> tuple[float, float]:
    """``dataset`` = ``(X_train, y_train, X_test, y_test)``; ``fruit`` a Fruit
    or a function ``(X_train, y_train) -> Fruit``; ``classifier`` anything with
    ``fit`` / ``score`` (default: StandardScaler -> nan_to_num ->
    RidgeClassifierCV).  Returns the mean feature-extraction time in seconds
    and the mean test accuracy over ``mean_over_n_runs`` repetitions."""
    X_train, y_train, X_test, y_test = dataset
    if classifier is None:
        classifier = _default_classifier()
    X_train = np.ascontiguousarray(np.nan_to_num(X_train), dtype=np.float64)
    X_test = np.ascontiguousarray(np.nan_to_num(X_test), dtype=np.float64)
    if callable(fruit) and not isinstance(fruit, fruits.Fruit):
        fruit = fruit(X_train, y_train)
    seconds, accuracies = [], []
    for _ in range(mean_over_n_runs):
        _sync()
        start = time.perf_counter()
        fruit.fit(X_train)
        train_features = fruit.transform(X_train)
        test_features = fruit.transform(X_test)
        _sync()
        seconds.append(time.perf_counter() - start)
        classifier.fit(train_features, y_train)
        accuracies.append(classifier.score(test_features, y_test))
    return float(np.mean(seconds)), float(np.mean(accuracies))


def fruitify_all(path: str, fruit, datasets: Optional[Sequence[str]] = None,
                 univariate: bool = True, classifier=None, output_csv: Optional[str] = None,
                 mean_over_n_runs: int = 1):
    """Run :func:`fruitify` on every dataset folder in ``path`` and keep a
    running CSV (``Dataset, Accuracy, Time``); an existing file is never
    overwritten (a timestamp is appended to the name instead)."""
    import pandas as pd
    stem = "results_fruits" if output_csv is None else output_csv.removesuffix(".csv")
    if os.path.exists(stem + ".csv"):
        stem += "_" + time.strftime("%Y-%m-%d-%H%M%S")
    rows = []
    for name, *data in load_all(path, univariate=univariate, datasets=datasets):
        seconds, accuracy = fruitify(tuple(data), fruit, classifier, mean_over_n_runs)
        rows.append((name, accuracy, seconds))
        pd.DataFrame(rows, columns=["Dataset", "Accuracy", "Time"]).to_csv(stem + ".csv",
                                                                           index=False)
    return pd.DataFrame(rows, columns=["Dataset", "Accuracy", "Time"])


def decide_which_fruit(choices, n_splits: int = 1, validation_size: float = 0.2,
                       classifier=None, mean_over_n_runs: int = 1) -> Callable:
    """-> ``choose(X_train, y_train) -> Fruit`` for :func:`fruitify`: every
    candidate in ``choices`` (a Fruit, or a pair ``(cheap fruit to judge by,
    fruit to return)``) is scored on ``n_splits`` stratified validation splits of
    the training data and a deep copy of the best one is returned; the first
    candidate if some class has a single sample (no stratified split exists).
    ``validation_size`` grows to one sample per class if it has to (reference
    :106-175; the splits draw from the global numpy RNG through sklearn)."""
    def second(choice):
        return (choice[1] if isinstance(choice, tuple) else choice).deepcopy()

    def choose(X: np.ndarray, y: np.ndarray):
        from sklearn.model_selection import train_test_split
        counts = np.unique(y, return_counts=True)[1]
        if np.sum(counts == 1) >= 1:
            return second(choices[0])
        share = max(validation_size, len(counts) / X.shape[0])
        means = []
        for choice in choices:
            judged = choice[0] if isinstance(choice, tuple) else choice
            accuracies = []
            for _ in range(n_splits):
                split = train_test_split(X, y, test_size=share, stratify=y)
                accuracies.append(fruitify((split[0], split[2], split[1], split[3]), judged,
                                           classifier=classifier,
                                           mean_over_n_runs=mean_over_n_runs)[1])
            means.append(np.mean(accuracies))
        return second(choices[int(np.argmax(means))])

    return choose
