"""Host-side alignment handling: FASTA parsing and the DNA symbol table.

Mirrors the reference's Alignment::ReadFasta (src/alignment.cpp:41-73) and
SitePattern::GetSymbolTable (src/site_pattern.cpp:16-46): every non-ACGT symbol is a
gap (state 4).  Compressing identical columns into weighted site patterns
(SitePattern::Compress) is done on the device: libsbn_b200.site_pattern.SitePattern.
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


def sequences_in_leaf_order(sequences, taxon_names):
    """Sequences (dict name -> str) as a list in leaf-id order."""
    for name in taxon_names:
        if name not in sequences:
            raise RuntimeError(f"Taxon {name} not found in alignment.")
    return [sequences[name] for name in taxon_names]


def site_patterns_of_fasta(path, taxon_names, device=0):
    """FASTA -> (patterns [taxon][P], weights [P]); the compression runs on the
    device (SitePattern::Compress, include/sbn_b200_patterns.h)."""
    from .site_pattern import SitePattern
    pattern = SitePattern(sequences_in_leaf_order(read_fasta(path), taxon_names), device)
    return pattern.patterns, pattern.weights
