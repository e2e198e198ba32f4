"""Host-side alignment handling: FASTA -> compressed site patterns.

Mirrors the reference's Alignment::ReadFasta (src/alignment.cpp:41-73) and
SitePattern (src/site_pattern.cpp:16-131): DNA symbol table with every
non-ACGT symbol treated as a gap (state 4), identical columns compressed into
weighted patterns.  Unlike the reference (whose pattern order is the iteration
order of a std::unordered_map and therefore stdlib-dependent) patterns are
kept in order of first appearance, which makes runs reproducible; the
log-likelihood is invariant to the order up to summation rounding.
"""
import numpy as np

_SYMBOLS = {c: i for i, c in enumerate("ACGT")}
_SYMBOLS.update({c.lower(): i for c, i in list(_SYMBOLS.items())})
_GAPS = "-NX?BDHKMRSUVWY"


def symbol_table():
    """SitePattern::GetSymbolTable (site_pattern.cpp:16-46)."""
    table = dict(_SYMBOLS)
    table.update({c: 4 for c in _GAPS})
    return table


def read_fasta(path):
    """Alignment::ReadFasta: {taxon name: sequence}; all sequences same length."""
    data, taxon, chunks = {}, None, []
    with open(path) as handle:
        for line in handle:
            line = line.rstrip("\n").rstrip("\r")
            if not line:
                continue
            if line[0] == ">":
                if taxon:
                    data[taxon] = "".join(chunks)
                taxon, chunks = line[1:], []
            else:
                chunks.append(line)
    if taxon:
        data[taxon] = "".join(chunks)
    if len({len(s) for s in data.values()}) > 1:
        raise RuntimeError("Sequences of the alignment are not all the same length.")
    return data


def encode(sequences, taxon_names):
    """Sequences (dict name -> str) in leaf-id order -> uint8 [taxon][site]."""
    table = symbol_table()
    lut = np.full(256, 255, dtype=np.uint8)
    for symbol, state in table.items():
        lut[ord(symbol)] = state
    rows = []
    for name in taxon_names:
        if name not in sequences:
            raise RuntimeError(f"Taxon {name} not found in alignment.")
        row = lut[np.frombuffer(sequences[name].encode("ascii"), dtype=np.uint8)]
        if (row == 255).any():
            bad = sequences[name][int(np.argmax(row == 255))]
            raise RuntimeError(f"Symbol '{bad}' not known.")
        rows.append(row)
    return np.stack(rows)


def compress(states):
    """SitePattern::Compress: uint8 [taxon][site] -> (patterns [taxon][P], weights [P])."""
    columns = np.ascontiguousarray(states.T)
    _, first, inverse, counts = np.unique(columns, axis=0, return_index=True, return_inverse=True,
                                          return_counts=True)
    order = np.argsort(first, kind="stable")  # first-appearance order
    patterns = np.ascontiguousarray(columns[first[order]].T)
    return patterns, counts[order].astype(np.float64)


def site_patterns_of_fasta(path, taxon_names):
    return compress(encode(read_fasta(path), taxon_names))
