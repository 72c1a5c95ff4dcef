"""2 GPUs (torchrun): where does the multi-GPU short-range acceleration of the C5 set differ from the brute-force sum?"""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from particlesimulation_b200 import capi, dist as pdist
rank, world, local = pdist.init_process_group("nccl")
args = types.SimpleNamespace(warmup=3, steps=5)
prm, ic, grid = bench.c5_setup(capi, world, 1 << 23, device=local, margin=bench.uniform_margin_cells(args))
ctx = pdist.create_context(prm, capi)
ctx.generate_particles(ic); ctx.green_init(); ctx.force()
ids = bench.parity_sample_ids(ic.n)
pos_l, acc_l, sr_l = ctx.sample(ids)
held = (np.abs(pos_l).sum(1) > 0).astype(np.float64)[:, None].repeat(3, 1)
g = lambda a: bench.gather_rows(dist, torch, a, len(ids), 3)
pos, sr, nheld = g(pos_l), g(sr_l), g(held)
ref = g(ctx.direct_sum(pos, capi.SUM_SHORT_RANGE))
shell = g(ctx.direct_sum(pos, capi.SUM_CUTOFF_SHELL))[:, 0]
if rank == 0:
    print("rank 0: cutoff-shell rows", np.nonzero(shell > 0)[0], "worst row", int(np.argmax(np.linalg.norm(sr - ref, axis=1))), "rel_l2 without them %.3e" % bench.rel_l2(sr[shell == 0], ref[shell == 0]))
info = ctx.rank_info(); b = ctx.binning()
if rank == 0:
    err = np.linalg.norm(sr - ref, axis=1); mag = np.sqrt((ref ** 2).sum(1).mean())
    print("rel_l2 %.3e" % bench.rel_l2(sr, ref), "rms |sr| %.3e" % mag, "info", info, "binning", b, "held counts", np.unique(nheld[:, 0], return_counts=True))
    hc = (float(prm.box[2]) / float(prm.H)) / b["mz"]
    for i in np.argsort(err)[::-1][:12]:
        print("  id", ids[i], "pos", pos[i], "layer %.3f" % (pos[i, 2] / hc), "err %.3e" % err[i], "sr", sr[i], "ref", ref[i])
# ---- which source accounts for the worst row's difference?
err_all = np.linalg.norm(sr - ref, axis=1)
w = int(np.argmax(err_all))
tgt = pos[w]
idsl, lp, _, _ = ctx.get_local(capi.UNITS_CODE, want=("pos",))
re = float(np.float32(np.float32(prm.cutoff_radius) / np.float32(prm.H)))
d = tgt[None, :] - lp.astype(np.float64)
r2 = (d * d).sum(1)
sel = np.nonzero((r2 < re * re * 1.01) & (r2 > 0))[0]
tab = ctx.sr_table()
mcode = bench.mass_code(ctx.params, np.float32(ic.total_mass) / np.float32(ic.n))
xi = r2[sel] / (re * re / 499.0)
t = np.minimum(xi.astype(np.int64), 498)
F = np.where(r2[sel] < re * re, tab[t] + (xi - t) * (tab[t + 1] - tab[t]), 0.0)
contrib = (mcode * F)[:, None] * d[sel]
diff = sr[w] - ref[w]
print(f"rank {rank}: {len(sel)} local sources near the target, HOSTSUM {contrib.sum(0)} sr {sr[w]} ref {ref[w]} diff {diff}")
inr = r2[sel] < re * re
print(f"rank {rank}: in range {int(inr.sum())}; sources with r/re in (0.99, 1.0): {int(((r2[sel] > 0.9801 * re * re) & inr).sum())}; re {re!r} re2 {re * re!r}")
for sgn in (1, -1):
    k = np.argmin(np.linalg.norm(contrib - sgn * diff[None, :], axis=1))
    print(f"rank {rank}: best single-pair match of {'+' if sgn > 0 else '-'}diff: local index {sel[k]} id {idsl[sel[k]]} pos {lp[sel[k]]} r/re {np.sqrt(r2[sel[k]]) / re:.5f} contrib {contrib[k]} residual {np.linalg.norm(contrib[k] - sgn * diff):.2e}")
# local array neighbourhood of the target itself
me = np.nonzero(idsl == ids[w])[0]
if len(me):
    i = int(me[0]); print(f"rank {rank}: target local index {i} (mod 32 = {i % 32}), n_local {len(idsl)}")
ctx.close()
dist.barrier(); dist.destroy_process_group()
