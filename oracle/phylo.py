"""TEST INFRASTRUCTURE: ctypes loader for oracle/_build/libphylo_oracle.so
(oracle/phylo_oracle.cpp + oracle/beagle_cpu.cpp), the CPU restatement of the
reference's FatBeagle path.  Used as the checker only.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libphylo_oracle.so")
_P = ctypes.POINTER
_D = _P(ctypes.c_double)
_lib = None


def build():
    """Compiles the oracle's C++ restatement (gcc only; no GPU, no reference needed)."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        common = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int64, _P(ctypes.c_uint8), _D,
                  ctypes.c_int, ctypes.c_int, _P(ctypes.c_int32), _D, _D, ctypes.c_int, ctypes.c_int,
                  ctypes.c_int, ctypes.c_int]
        lib.sbno_log_likelihoods.restype = ctypes.c_int
        lib.sbno_log_likelihoods.argtypes = common + [_D, _D, _D, ctypes.c_int, _D]
        lib.sbno_gradients.restype = ctypes.c_int
        lib.sbno_gradients.argtypes = common + [_D, _D, _D, _D, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                _D, _D, _D, _D, _D, _D]
        lib.sbno_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_D)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def gamma_rates(shape, categories):
    """Discrete Gamma(shape, rate = shape), median discretisation (Yang 1994): the
    category rates, normalised to mean 1, and d rate / d shape -- from
    scipy.stats.gamma.ppf, independently of libsbn_b200's own incomplete-gamma code.
    The derivative is a 4th-order central difference of the normalised rates."""
    from scipy.stats import gamma as gamma_distribution

    def rates(a):
        q = (2.0 * np.arange(categories) + 1.0) / (2.0 * categories)
        g = gamma_distribution.ppf(q, a)
        return g / g.mean()

    h = 1e-3 * shape
    derivative = (-rates(shape + 2 * h) + 8 * rates(shape + h) - 8 * rates(shape - h) + rates(shape - 2 * h)) / (12 * h)
    return rates(shape), derivative


def _expand_site(site, params, substitution):
    """A "gamma+K" site model is handed to the C oracle as explicit category rates
    ("rates+K": K rates, K derivatives in place of the shape column)."""
    if not site.startswith("gamma"):
        return site, params
    categories = int(site.split("+")[1]) if "+" in site else 4
    params = np.asarray(params, dtype=np.float64)
    head = 10 if substitution == "GTR" else 0
    rows = []
    for row in params:
        r, d = gamma_rates(row[head], categories)
        rows.append(np.concatenate([row[:head], r, d, row[head + 1:]]))
    return f"rates+{categories}", np.array(rows)


def _common(substitution, site, patterns, weights, parent_ids, branch_lengths, params, rescaling,
            use_tip_states, rooted):
    site, params = _expand_site(site, params, substitution)
    patterns = _c(patterns, np.uint8)
    weights = _c(weights, np.float64)
    parent_ids = _c(parent_ids, np.int32)
    branch_lengths = _c(branch_lengths, np.float64)
    T = parent_ids.shape[0]
    if params is None or params.size == 0:
        params = np.zeros((T, 1))
    params = _c(params, np.float64)
    keep = (patterns, weights, parent_ids, branch_lengths, params)
    args = [substitution.encode(), site.encode(), patterns.shape[0], patterns.shape[1],
            patterns.ctypes.data_as(_P(ctypes.c_uint8)), _d(weights), T, branch_lengths.shape[1],
            parent_ids.ctypes.data_as(_P(ctypes.c_int32)), _d(branch_lengths), _d(params),
            params.shape[1], int(rescaling), int(use_tip_states), int(rooted)]
    return keep, args, T


def _check(code):
    if code != 0:
        raise RuntimeError(load().sbno_last_error().decode())


def log_likelihoods(substitution, site, patterns, weights, parent_ids, branch_lengths, params=None,
                    rescaling=False, use_tip_states=True, rooted=False, rates=None, node_heights=None,
                    node_bounds=None, threads=1):
    keep, args, T = _common(substitution, site, patterns, weights, parent_ids, branch_lengths, params,
                            rescaling, use_tip_states, rooted)
    rates, node_heights, node_bounds = (_c(x, np.float64) for x in (rates, node_heights, node_bounds))
    out = np.empty(T)
    _check(load().sbno_log_likelihoods(*args, _d(rates), _d(node_heights), _d(node_bounds), threads, _d(out)))
    return out


def gradients(substitution, site, patterns, weights, parent_ids, branch_lengths, params=None,
              rescaling=False, use_tip_states=True, rooted=False, rates=None, node_heights=None,
              node_bounds=None, height_ratios=None, rate_count=1, reference_quirks=False, threads=1):
    keep, args, T = _common(substitution, site, patterns, weights, parent_ids, branch_lengths, params,
                            rescaling, use_tip_states, rooted)
    rates, node_heights, node_bounds, height_ratios = (
        _c(x, np.float64) for x in (rates, node_heights, node_bounds, height_ratios))
    n = keep[0].shape[0]
    out = {
        "log_likelihood": np.zeros(T),
        "branch": np.zeros((T, 2 * n - 1)),
        "substitution_model": np.zeros((T, 8)),
        "site_model": np.zeros(T),
        "ratios_root_height": np.zeros((T, n - 1)),
        "clock_model": np.zeros((T, rate_count)),
    }
    _check(load().sbno_gradients(*args, _d(rates), _d(node_heights), _d(node_bounds), _d(height_ratios),
                                 rate_count, int(reference_quirks), threads, _d(out["log_likelihood"]),
                                 _d(out["branch"]), _d(out["substitution_model"]), _d(out["site_model"]),
                                 _d(out["ratios_root_height"]), _d(out["clock_model"])))
    return out
