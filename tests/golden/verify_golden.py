"""Re-generate the fixtures from the UNMODIFIED reference into a scratch directory and compare them, array by array and bit
for bit, with the committed ones (build container only: needs /root/reference, like make_golden.py).

    python tests/golden/verify_golden.py [sampler module negatives feature replay bracket]     (default: all, ~2.5 min)

Exit status 0 when every regenerated array equals the committed one except the entries of NONDETERMINISTIC below, whose
values depend on the thread schedule of torch's CPU reductions or on ARPACK's random start vector — for those the bound
that the tests rely on is checked instead."""
from __future__ import annotations

import os
import runpy
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import common  # noqa: E402

# (file, key) -> max abs difference tolerated between two runs of the reference itself
NONDETERMINISTIC = {
    ("module_full.npz", "grad_fft_filter.weight"): 1e-5,  # backward of the complex filter: multi-threaded fp32 reductions
    ("module_full.npz", "grad_fft_agg.weight"): 1e-5,
    ("module_small.npz", "grad_fft_filter.weight"): 1e-5,
    ("module_small.npz", "grad_fft_agg.weight"): 1e-5,
    ("run_bracket.npz", "lappe_12"): None,  # ARPACK eigenvectors: sign / basis inside an eigenspace is arbitrary (the test
                                            # compares eigenvalues and invariant subspaces, tests/test_run_bracket_gpu.py)
}
# generators that read params_<tag>.npz written by an earlier one
NEEDS = {"replay": ["module"]}


def main(which):
    for w in list(which):
        for dep in NEEDS.get(w, []):
            if dep not in which:
                which.append(dep)
    real, tmp = common.HERE, tempfile.mkdtemp(prefix="golden_regen_")
    common.HERE = tmp  # golden_path() resolves against it at call time
    argv, sys.argv = sys.argv, ["make_golden.py"] + which
    t0 = time.time()
    try:
        runpy.run_path(os.path.join(real, "make_golden.py"), run_name="__main__")
    finally:
        sys.argv, common.HERE = argv, real
    print(f"regenerated {which} in {time.time() - t0:.1f} s -> {tmp}")
    bad = n_arrays = 0
    for fn in sorted(os.listdir(tmp)):
        a_path, b_path = os.path.join(tmp, fn), os.path.join(real, fn)
        if not os.path.exists(b_path):
            print("NOT COMMITTED", fn)
            bad += 1
            continue
        if not fn.endswith(".npz"):
            same = open(a_path, "rb").read() == open(b_path, "rb").read()
            print(("identical" if same else "differs (pickle framing)"), fn)
            continue
        a, b = np.load(a_path, allow_pickle=True), np.load(b_path, allow_pickle=True)
        if sorted(a.files) != sorted(b.files):
            print("KEYS DIFFER", fn, set(a.files) ^ set(b.files))
            bad += 1
            continue
        for k in a.files:
            x, y = a[k], b[k]
            n_arrays += 1
            if x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y, equal_nan=x.dtype.kind in "fc"):
                continue
            tol = NONDETERMINISTIC.get((fn, k), 0.0)
            diff = float(np.abs(x - y).max()) if x.shape == y.shape and x.dtype.kind in "fciu" else float("inf")
            ok = (fn, k) in NONDETERMINISTIC and (tol is None or diff <= tol)
            print(("nondeterministic in the reference" if ok else "MISMATCH"), fn, k, x.dtype, x.shape, diff)
            bad += 0 if ok else 1
    print(f"{n_arrays} arrays compared, {bad} mismatches")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1:] or ["sampler", "module", "replay", "bracket", "negatives", "feature"]) else 0)
