"""One op program of the DS1 DAG, a few times over (an ncu target):
    ncu --set full --import-source on -k regex:GpInterpret -s 3 -c 1 -o gpurun_out/gp python tools/gp_profile.py populate_plvs
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libsbn_b200.gp_engine import GPEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "populate_plvs"
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "gp_ds1_dag.npz")))
engine = GPEngine(fx["tip_states"], fx["pattern_weights"], int(fx["site_count"]), int(fx["plv_count"]),
                  int(fx["gpcsp_count"]), rescaling_threshold=float(fx["rescaling_threshold"]),
                  sbn_prior=fx["sbn_prior"], unconditional_node_probabilities=fx["unconditional_node_probabilities"],
                  inverted_sbn_prior=fx["inverted_sbn_prior"])
engine.set_branch_lengths(fx["initial_branch_lengths"])
for _ in range(3):
    engine.process_operations(fx["program_populate_plvs"])
for _ in range(3):
    engine.process_operations(fx["program_" + name])
print(engine.last_kernel_ms)
