"""2-GPU debug: where do multi-GPU short-range accelerations differ from the single-GPU ones?"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from common import plummer_case, to_p3m, uniform_case
from particlesimulation_b200 import capi, dist as pdist

pdist.init_process_group("nccl")
rank = dist.get_rank()
def allsum(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda(); dist.all_reduce(t); return t.cpu().numpy()
for name, (p, pos, vel, mass) in {"plummer": plummer_case(20000), "uniform": uniform_case(20000)}.items():
    prm = to_p3m(p, p3m=True, zero_degenerate=True); prm.device = int(os.environ["LOCAL_RANK"])
    ctx = pdist.create_context(prm, capi)
    ctx.set_particles(pos, vel, mass); ctx.green_init(); ctx.force()
    pm, sr = ctx.acc_parts(); pm, sr = allsum(pm), allsum(sr)
    info = ctx.rank_info(); b = ctx.binning()
    gp = allsum(ctx.get_particles(capi.UNITS_CODE, want=("pos",))[0])
    ctx.close()
    if rank == 0:
        s = capi.Context(prm); s.set_particles(pos, vel, mass); s.green_init(); s.force()
        pm1, sr1 = s.acc_parts(); s.close()
        e_sr = np.abs(sr - sr1).max(1); e_pm = np.abs(pm - pm1).max(1)
        hc = (float(prm.box[2]) / b["mz"]) / float(prm.H)
        layer = np.floor(gp[:, 2] / hc).astype(int)
        print(name, "binning", b, "info", info)
        print("  sr rel", np.linalg.norm(sr - sr1) / np.linalg.norm(sr1), "pm rel", np.linalg.norm(pm - pm1) / np.linalg.norm(pm1))
        bad = np.argsort(e_sr)[::-1][:10]
        print("  worst sr errs", e_sr[bad], "layers", layer[bad], "|sr|", np.abs(sr1[bad]).max(1))
        for L in range(b["mz"]):
            m = layer == L
            if m.any():
                print(f"   layer {L}: n={m.sum()} max sr err {e_sr[m].max():.3e} (|sr| max {np.abs(sr1[m]).max():.3e}) max pm err {e_pm[m].max():.3e}")
dist.barrier(); dist.destroy_process_group()
