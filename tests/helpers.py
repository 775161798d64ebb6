"""Comparison helpers shared by the parity tests (SURVEY.md §8d tolerances)."""
import numpy as np

RTOL, ATOL = 1e-5, 1e-6  # |a-b| <= RTOL*|b| + ATOL ; NaN positions and exact 0.0/1.0 cells must match exactly


def compare_layer(got, ref, name=""):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    nan_mismatch = int((np.isnan(got) != np.isnan(ref)).sum())
    both = ~np.isnan(got) & ~np.isnan(ref)
    diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    tol = RTOL * np.abs(ref.astype(np.float64)) + ATOL
    out_of_tol = int((both & (diff > tol)).sum())
    rel_only = int((both & (diff > RTOL * np.abs(ref.astype(np.float64)))).sum())
    branch = both & ((ref == 0.0) | (ref == 1.0))
    branch_mismatch = int((branch & (got != ref)).sum())
    bit_exact = int((got.view(np.uint32) == ref.view(np.uint32)).sum()) + int((np.isnan(got) & np.isnan(ref)).sum()) \
        - int(((got.view(np.uint32) == ref.view(np.uint32)) & np.isnan(got)).sum())
    return {"name": name, "cells": got.size, "nan_mismatch": nan_mismatch, "out_of_tol": out_of_tol,
            "rel_only_violations": rel_only, "branch_mismatch": branch_mismatch, "bit_exact": bit_exact,
            "max_abs": float(diff[both].max()) if both.any() else 0.0}


def assert_parity(got: dict, ref: dict, keys=("slope", "step", "roughness", "traversability"), allow_out_of_tol=0):
    reports = [compare_layer(got[k], ref[k], k) for k in keys]
    for r in reports:
        assert r["nan_mismatch"] == 0, r
        assert r["out_of_tol"] <= allow_out_of_tol, r
        assert r["branch_mismatch"] <= allow_out_of_tol, r
    return reports
