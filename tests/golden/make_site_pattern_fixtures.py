"""Generates tests/golden/site_pattern_*.json.gz by running the UNMODIFIED
reference SitePattern::Compress (oracle/_ref/site_pattern_dump, built by
`make -C oracle ref` in the build container, where /root/reference is mounted):

    python tests/golden/make_site_pattern_fixtures.py

Each fixture holds the raw sequences (in the reference's taxon numbering) and the
patterns / weights the reference produced, in the reference's own (hash-dependent)
order.  Nothing here runs on the GPU box."""
import gzip
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DUMP = os.path.join(ROOT, "oracle", "_ref", "site_pattern_dump")
DATA = "/root/reference/data"

for name, fasta in [("hello", "hello.fasta"), ("ds1", "DS1.fasta"), ("flua", "fluA.fa"),
                    ("five_taxon", "five_taxon.fasta"), ("seven_taxon", "7-taxon-slice-of-ds1.fasta")]:
    out = subprocess.run([DUMP, os.path.join(DATA, fasta)], capture_output=True, text=True, check=True).stdout
    record = json.loads(out)
    record["source"] = f"data/{fasta}"
    path = os.path.join(HERE, f"site_pattern_{name}.json.gz")
    with gzip.open(path, "wt") as handle:
        json.dump(record, handle)
    print(name, len(record["sequences"]), "sequences x", len(record["sequences"][0]), "sites ->",
          len(record["weights"]), "patterns", os.path.getsize(path), "bytes")
