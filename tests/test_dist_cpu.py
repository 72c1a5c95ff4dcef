"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group (no GPU needed).
Covers the unique-id broadcast, the slab cuts (identical on every rank, covering every layer) and the
ownership rule that partitions a particle set exactly once."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(here)r)
import torch, torch.distributed as dist
from particlesimulation_b200 import capi, dist as pdist, ics
import bench
rank, world, _ = pdist.init_process_group("gloo")
payload = bytes(range(128)) if rank == 0 else None
got = pdist.broadcast_bytes(payload, 128, 0)
assert got == bytes(range(128)), "unique id broadcast"
prm = bench.multi_params(capi, world)
cuts, layers = capi.slab_cuts(prm, world)
t = torch.tensor(cuts.tolist()); gathered = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(gathered, t)
assert all(torch.equal(g, t) for g in gathered), "cuts differ between ranks"
assert cuts[0] == 0 and cuts[-1] == layers and np.all(np.diff(cuts) > 0)
pos, vel, mass = bench.multi_particles(world, 4096)
hc = (float(prm.box[2]) / layers) / float(prm.H)
owner = pdist.owner_of(pos[:, 2] / np.float32(prm.H), cuts, hc)
mine = int((owner == rank).sum())
tot = torch.tensor([mine]); dist.all_reduce(tot)
assert int(tot) == len(mass), "every particle must have exactly one owner"
assert abs(mine - len(mass) / world) < 0.05 * len(mass), "clusters are one per slab"
# work-balanced cuts: every rank derives them from the full set it was handed, without communication
bc, _ = capi.balanced_cuts(prm, world, pos)
tb = torch.tensor(bc.tolist()); gb = [torch.zeros_like(tb) for _ in range(world)]
dist.all_gather(gb, tb)
assert all(torch.equal(g, tb) for g in gb), "balanced cuts differ between ranks"
owner_b = pdist.owner_of(pos[:, 2] / np.float32(prm.H), bc, hc)
tot_b = torch.tensor([int((owner_b == rank).sum())]); dist.all_reduce(tot_b)
assert int(tot_b) == len(mass), "balanced cuts must partition the set exactly once"
dist.barrier()
if rank == 0: print("DIST_CPU_OK")
dist.destroy_process_group()
'''


def test_gloo_world2_host_logic(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, here=HERE))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "DIST_CPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
