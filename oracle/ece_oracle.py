"""CPU oracle: ECE / MCE / reliability-histogram binning.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  numpy restatement of
``src/uncertainty.py:84-171`` (ECE/MCE, fp32 ``torch.linspace`` edges) and
``src/uncertainty.py:218-241`` (reliability diagram, float64 ``np.linspace``
edges).  Pinned by ``tests/golden/ece_*.npz`` (``oracle/make_golden.py``).

The binning itself is integer work and must be bit-exact; the per-bin mean
confidence depends on the reference's fp32 summation order and is compared
within 1e-6 relative.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def linspace_f32(num_bins: int) -> np.ndarray:
    """``torch.linspace(0, 1, num_bins + 1)`` in fp32 (uncertainty.py:109).

    ATen's linspace kernel computes ``start + step*i`` for the first half and
    ``end - step*(steps-1-i)`` for the second half with
    ``step = fp32((end-start)/(steps-1))``, the multiply-add being fused (one
    rounding) — this is why the edges equal neither ``float32(i/nb)`` nor
    ``i*float32(1/nb)``.  The fused op is restated
    exactly: fp32 x small-int products and their sum with 0/1 are exact in
    float64, so a single final rounding to fp32 reproduces the FMA.
    """
    steps = num_bins + 1
    if steps == 1:
        return np.array([0.0], dtype=np.float32)
    step = np.float64(np.float32(np.float32(1.0) / np.float32(steps - 1)))
    half = steps // 2
    out = np.empty(steps, dtype=np.float32)
    for i in range(steps):
        if i < half:
            out[i] = np.float32(0.0 + step * i)
        else:
            out[i] = np.float32(1.0 - step * (steps - 1 - i))
    return out


def linspace_f64(num_bins: int) -> np.ndarray:
    """``np.linspace(0.0, 1.0, num_bins + 1)`` (uncertainty.py:222)."""
    return np.linspace(0.0, 1.0, num_bins + 1)


def bin_masks(
    conf: np.ndarray, pred: np.ndarray, label: np.ndarray, edges: np.ndarray
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One masked pass per bin, as the reference does (uncertainty.py:113-126,
    231-241).  ``edges`` are compared in float64 (fp32 edges and fp32
    confidences promote exactly).  Returns ``(count i64, correct i64,
    conf_sum f64)`` per bin.  NaN / out-of-[0,1] confidences land in no bin."""
    c = np.asarray(conf, dtype=np.float64)
    e = np.asarray(edges, dtype=np.float64)
    nb = e.shape[0] - 1
    hit = np.asarray(pred) == np.asarray(label)
    count = np.zeros(nb, dtype=np.int64)
    correct = np.zeros(nb, dtype=np.int64)
    conf_sum = np.zeros(nb, dtype=np.float64)
    for i in range(nb):
        lower, upper = e[i], e[i + 1]
        if i == nb - 1:  # `upper == 1.0` (:114) / `idx == num_bins-1` (:232)
            in_bin = (c >= lower) & (c <= upper)
        else:
            in_bin = (c >= lower) & (c < upper)
        count[i] = int(in_bin.sum())
        correct[i] = int((in_bin & hit).sum())
        conf_sum[i] = float(c[in_bin].sum())
    return count, correct, conf_sum


def ece_from_bins(
    count: np.ndarray, correct: np.ndarray, conf_sum: np.ndarray, total: int
) -> Tuple[float, float]:
    """ECE and MCE from per-bin statistics, in the reference's arithmetic:
    fp32 bin means, Python-float weight, fp32 accumulator
    (uncertainty.py:110,119-131,153-171)."""
    ece = np.float32(0.0)
    mce = np.float32(0.0)
    for i in range(count.shape[0]):
        n = int(count[i])
        if n == 0:
            continue
        bin_conf = np.float32(conf_sum[i] / n)
        bin_acc = np.float32(correct[i] / n)
        err = np.float32(abs(np.float32(bin_acc - bin_conf)))
        ece = np.float32(ece + np.float32(np.float32(n / total) * err))
        mce = max(mce, err)
    return float(ece), float(mce)


def expected_calibration_error(conf, pred, label, num_bins: int = 15) -> float:
    cnt, cor, cs = bin_masks(conf, pred, label, linspace_f32(num_bins))
    return ece_from_bins(cnt, cor, cs, int(np.asarray(conf).shape[0]))[0]


def maximum_calibration_error(conf, pred, label, num_bins: int = 15) -> float:
    cnt, cor, cs = bin_masks(conf, pred, label, linspace_f32(num_bins))
    return ece_from_bins(cnt, cor, cs, int(np.asarray(conf).shape[0]))[1]


def reliability_bins(conf, pred, label, num_bins: int = 15):
    """``bin_counts, avg_confidences, accuracies`` float32 arrays of
    uncertainty.py:227-241 (float64 edges)."""
    cnt, cor, cs = bin_masks(conf, pred, label, linspace_f64(num_bins))
    nz = np.maximum(cnt, 1)
    avg = np.where(cnt > 0, cs / nz, 0.0).astype(np.float32)
    acc = np.where(cnt > 0, cor / nz, 0.0).astype(np.float32)
    return cnt.astype(np.float32), avg, acc


# ---------------------------------------------------------------------------
# C restatement (oracle/ece_oracle.c): same per-bin masked passes, used for
# large-N cross-checks and as the CPU baseline "port" in bench.py.
# ---------------------------------------------------------------------------
_LIB = None


def build_c(force: bool = False) -> str:
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libece_oracle.so")
    src = os.path.join(_HERE, "ece_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", so, src]
        )
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.ece_oracle_bin.restype = ctypes.c_int
        _LIB.ece_oracle_bin.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int,
        ]
    return _LIB


def bin_masks_c(conf, pred, label, edges, threads: int = 1):
    conf = np.ascontiguousarray(conf, dtype=np.float32)
    pred = np.ascontiguousarray(pred, dtype=np.int64)
    label = np.ascontiguousarray(label, dtype=np.int64)
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    nb = edges.shape[0] - 1
    count = np.zeros(nb, dtype=np.int64)
    correct = np.zeros(nb, dtype=np.int64)
    conf_sum = np.zeros(nb, dtype=np.float64)
    rc = _lib().ece_oracle_bin(
        conf.ctypes.data, pred.ctypes.data, label.ctypes.data, conf.shape[0],
        edges.ctypes.data, nb, count.ctypes.data, correct.ctypes.data,
        conf_sum.ctypes.data, threads,
    )
    if rc != 0:
        raise RuntimeError(f"ece_oracle_bin failed: {rc}")
    return count, correct, conf_sum
