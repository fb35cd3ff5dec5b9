"""Times one build of the library (BCM3B200_LIB) on a config-5-shaped PopPK batch (two compartments, 100 000 patients x 16 chains)
at the block sizes given on the command line; the per-chain results are compared with those of the first run that wrote
gpurun_out/poppk_occ_ref.npy (run the default build first)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_TWO
from bcm3_b200.poppk import PopPKEvaluator

tag = os.path.basename(os.environ.get("BCM3B200_LIB", "default"))
blocks = [int(b) for b in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"])]
prob = syn.make_poppk_problem(PK_TWO, P=100_000, T=10, t_end=72.0, seed=1)
vals = syn.make_chain_values(prob, 16)
ref_path = os.path.join("gpurun_out", "poppk_occ_ref.npy")
for b in blocks:
    ev = PopPKEvaluator(prob, block_size=b)
    times = []
    for it in range(5):
        logp, status = ev.evaluate(vals)
        times.append(ev.get_stat("last_kernel_us") / 1e3)
    ev.close()
    if not os.path.exists(ref_path):
        np.save(ref_path, logp)
    ref = np.load(ref_path)
    print(f"{tag:28s} block={b:3d}: kernel best {min(times[1:]):8.3f} ms median {np.median(times[1:]):8.3f} ms   max rel diff vs first run {np.max(np.abs(logp - ref) / np.abs(ref)):.2e}", flush=True)
