"""1 GPU: short-range parity of the C5 set (uniform, ~37 particles per chaining cell) against the fp64 brute-force sum,
per dense-cell threshold / kernel choice (run with P3M_TUNE_* set)."""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from particlesimulation_b200 import capi
args = types.SimpleNamespace(warmup=3, steps=5)
prm, ic, grid = bench.c5_setup(capi, int(os.environ.get("WORLD", 1)), int(os.environ.get("N", 1 << 23)), margin=bench.uniform_margin_cells(args))
ctx = capi.Context(prm)
ctx.generate_particles(ic); ctx.green_init(); ctx.force()
ids = bench.parity_sample_ids(ic.n)
pos, acc, sr = ctx.sample(ids)
ref = ctx.direct_sum(pos, capi.SUM_SHORT_RANGE)
err = np.linalg.norm(sr - ref, axis=1)
print(os.environ.get("TAG", ""), "grid", grid, "sr_rel_l2 %.3e" % bench.rel_l2(sr, ref), "max row err %.3e" % err.max(), "|sr| rms %.3e" % np.sqrt((ref ** 2).sum(1).mean()),
      "rows with err > 1e-3 |sr|rms:", int((err > 1e-3 * np.sqrt((ref ** 2).sum(1).mean())).sum()))
ctx.close()
