"""Comparison helpers shared by the oracle and GPU parity tests."""
import numpy as np

from cases import unwrap


def assert_exact(got, ref, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    assert same.all(), (f"{what}: {(~same).sum()} of {same.size} values differ, "
                        f"first at {np.argwhere(~same)[:3].tolist()}")


def rowmax_rel_err(got, ref):
    """|got-ref| / max(|ref|, rowmax|ref|): the parity metric of SURVEY.md
    section 8(d) for floating point iterated sums."""
    got, ref = np.asarray(got), np.asarray(ref)
    if ref.size == 0:
        return np.zeros(ref.shape)
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    fin = np.where(np.isfinite(ref), ref, 0.0)
    scale = np.maximum(np.abs(fin), np.max(np.abs(fin), axis=-1, keepdims=True))
    with np.errstate(invalid="ignore"):
        err = np.where(same, 0.0, np.abs(got - ref))
    return err / (scale + 1e-300)


def assert_close(got, ref, rtol, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    rel = rowmax_rel_err(got, ref)
    assert np.all(rel <= rtol), f"{what}: max rel err {rel.max():.3e} > {rtol}"


def fitted_thresholds(fruit):
    """All fitted thresholds of a product Fruit in slice/node/sieve order."""
    rows = []
    for slc in fruit:
        for sieves in slc._sieves_extended:
            for sv in sieves:
                sv = unwrap(sv)                      # sieve wrappers: the wrapped sieve's
                q = getattr(sv, "_quantiles", None)
                if q is None:
                    q = getattr(sv, "_q", [])
                rows.append(np.asarray(q, dtype=np.float64).ravel())
    return np.concatenate(rows) if rows else np.zeros(0)


def oracle_thresholds(of):
    rows = []
    for slc in of.slices:
        for sieves in slc.sieves_extended:
            for sv in sieves:
                sv = unwrap(sv)
                q = sv.fitted_q if sv.name in ("PPV", "CPV") else sv.quantiles
                rows.append(np.asarray(q, dtype=np.float64).ravel())
    return np.concatenate(rows) if rows else np.zeros(0)


def parity_report(got, ref, rtol=1e-9):
    """Numbers SURVEY.md section 8(d) asks to report beside a tolerance check of
    weighted (floating point) features: how many elements violate the
    elementwise rtol, how many of those are integer counts that moved by exactly
    one (a value within rounding of its threshold), and the worst relative error
    of the rest."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.maximum(np.abs(ref), 1.0)
    diff = np.abs(got - ref)
    diff = np.where((got == ref) | (np.isnan(got) & np.isnan(ref)), 0.0, diff)
    bad = diff > rtol * scale
    integer = (ref == np.round(ref)) & (got == np.round(got))
    flips = bad & integer & (diff == 1.0)
    other = bad & ~flips
    return {"elements": int(ref.size), "rtol": rtol, "rtol_violations": int(bad.sum()),
            "count_flips": int(flips.sum()), "other_violations": int(other.sum()),
            "max_rel_err_non_flip": float(np.max(np.where(flips, 0.0, diff / scale), initial=0.0))}


REPORTS = {}


def record_report(name, report):
    """Keep a parity report for the session summary (tests/conftest.py writes
    them to gpurun_out/parity_report.json when that directory exists)."""
    REPORTS[name] = report
    print(f"[parity] {name}: {report}")
