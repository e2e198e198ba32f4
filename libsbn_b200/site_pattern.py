"""Host mirror of the reference's SitePattern (src/site_pattern.hpp): an alignment
compressed into unique site patterns and their multiplicities, computed on the
device through sbnb_compress_site_patterns (include/sbn_b200_patterns.h).

    pattern = SitePattern(sequences)          # list of equal-length strings, leaf-id order
    engine = Engine(spec, pattern.patterns, pattern.weights)
"""
import ctypes

import numpy as np

from . import _capi


class SitePattern:
    """patterns: uint8 [taxon][pattern] (0..3 = ACGT, 4 = gap), weights: float64
    [pattern]; patterns are in order of first appearance in the alignment."""

    def __init__(self, sequences, device=0):
        rows = [s.encode() if isinstance(s, str) else bytes(s) for s in sequences]
        if not rows:
            raise RuntimeError("Site pattern compression needs at least one sequence.")
        length = len(rows[0])
        if any(len(r) != length for r in rows):
            # Alignment::Length (alignment.hpp) asserts this in the reference
            raise RuntimeError("Sequences of the alignment are not all of the same length.")
        self.sequence_count, self.site_count = len(rows), length
        flat = b"".join(rows)
        patterns = np.zeros(max(len(rows) * length, 1), dtype=np.uint8)
        weights = np.zeros(max(length, 1), dtype=np.float64)
        count, device_ms = ctypes.c_int64(), ctypes.c_double()
        _capi.check(_capi.load().sbnb_compress_site_patterns(
            len(rows), length, flat, device, patterns.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
            _capi.as_double_ptr(weights), ctypes.byref(count), ctypes.byref(device_ms)))
        self.pattern_count = count.value
        self.patterns = patterns[:len(rows) * count.value].reshape(len(rows), count.value).copy()
        self.weights = weights[:count.value].copy()
        self.device_ms = device_ms.value
