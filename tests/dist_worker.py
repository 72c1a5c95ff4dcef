"""Worker of tests/test_multi_gpu.py (run under torchrun, one rank per GPU, NCCL): the z-slab path on
`world` GPUs must reproduce the single-GPU result on the same particles."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from common import plummer_case, to_p3m, uniform_case  # noqa: E402
from particlesimulation_b200 import capi  # noqa: E402
from particlesimulation_b200 import dist as pdist  # noqa: E402
from refapi import rel_l2  # noqa: E402


def note(msg):
    if dist.get_rank() == 0:
        print("DIST_PROGRESS " + msg, flush=True)


def allsum(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def run_case(name, p, pos, vel, mass, p3m, steps):
    rank, world = dist.get_rank(), dist.get_world_size()
    note(name)
    prm = to_p3m(p, p3m=p3m, zero_degenerate=True)  # 0/0 modes of the optimal G set to 0
    prm.device = int(os.environ.get("LOCAL_RANK", 0))
    out = {}
    ctx = pdist.create_context(prm, capi)
    ctx.set_particles(pos, vel, mass)
    n_local0 = ctx.n
    ctx.green_init()
    ctx.force()
    info = ctx.rank_info()
    acc = allsum(ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2])
    slab = bool(ctx.rank_info().get("slab", 0))
    # slab-decomposed mesh: every rank returns its own planes (zeros elsewhere); replicated: rank 0 has all
    rho = allsum(ctx.density(f64=True)) if slab else ctx.density(f64=True)
    phi = allsum(ctx.potential(f64=True)) if slab else ctx.potential(f64=True)
    ctx.kick(0.5)
    done = ctx.step(steps)
    gp, gv, _ = ctx.get_particles(capi.UNITS_CODE)
    gp, gv = allsum(gp), allsum(gv)
    counts = allsum(np.array([ctx.n], np.int64))
    diag = ctx.diagnostics()
    ctx.close()
    if rank == 0:
        single = capi.Context(prm)
        single.set_particles(pos, vel, mass)
        single.green_init()
        single.force()
        acc1 = single.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        rho1 = single.density(f64=True)
        phi1 = single.potential(f64=True)
        single.kick(0.5)
        single.step(steps)
        sp, sv, _ = single.get_particles(capi.UNITS_CODE)
        diag1 = single.diagnostics()
        single.close()
        out = dict(case=name, n=int(p.n), world=world, n_local0=int(n_local0), ghosts=info["ghosts"],
                   total_after=int(counts[0]), steps_done=int(done),
                   acc=rel_l2(acc, acc1), rho=rel_l2(rho, rho1), phi=rel_l2(phi, phi1), slab=int(slab), pos=rel_l2(gp, sp), vel=rel_l2(gv, sv),
                   diag=float(np.abs(diag - diag1).max() / (np.abs(diag1).max() + 1e-300)))
    return out


def run_generated(name, p, ic, p3m):
    """Device-side initial conditions: every rank generates only its own z-slab (p3m_generate_particles); the
    union must be the sampled set, and the force must equal the single-GPU force on the uploaded set."""
    rank, world = dist.get_rank(), dist.get_world_size()
    note(name)
    prm = to_p3m(p, p3m=p3m, zero_degenerate=True)
    prm.device = int(os.environ.get("LOCAL_RANK", 0))
    ctx = pdist.create_context(prm, capi)
    ctx.generate_particles(ic)
    note(name + " generated")
    n_local0 = ctx.n
    counts = allsum(np.array([ctx.n], np.int64))
    gp = allsum(ctx.get_particles(capi.UNITS_ORIGINAL, want=("pos",))[0])
    ctx.green_init()
    ctx.force()
    acc = allsum(ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2])
    ctx.close()
    out = {}
    if rank == 0:
        pos, vel, mass = capi.sample_particles(ic)
        single = capi.Context(prm)
        single.set_particles(pos, vel, mass)
        single.green_init()
        single.force()
        acc1 = single.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        single.close()
        out = dict(case=name, generated=1, n=int(ic.n), total=int(counts[0]), n_local0=int(n_local0),
                   pos=rel_l2(gp, pos), acc=rel_l2(acc, acc1))
    return out


def main():
    pdist.init_process_group("nccl")
    results = []
    p, _, _, _ = plummer_case(20000)
    results.append(run_generated("generated_plummer_p3m", p, capi.ic_plummer(20000, seed=42), True))
    p, _, _, _ = uniform_case(20000, gfunc=0)
    results.append(run_generated("generated_uniform_pm", p, capi.ic_uniform(20000, [5.0] * 3, [55.0] * 3, seed=1), False))
    # one Plummer sphere centred ON the slab boundary (world = 2): heavy ghost traffic, migration
    p, pos, vel, mass = plummer_case(20000)
    results.append(run_case("plummer_p3m", p, pos, vel, mass, True, 5))
    # uniform set, PM only (tile layers), fast particles so that many migrate
    p, pos, vel, mass = uniform_case(20000, gfunc=0)
    rng = np.random.default_rng(5)
    vel = (0.6 * rng.standard_normal(vel.shape)).astype(np.float32)
    results.append(run_case("uniform_pm", p, pos, vel, mass, False, 5))
    # uniform P3M: ghosts on both sides for interior ranks
    p, pos, vel, mass = uniform_case(20000)
    results.append(run_case("uniform_p3m", p, pos, vel, mass, True, 3))
    if dist.get_rank() == 0:
        print("DIST_RESULTS " + json.dumps(results))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        # a rank that fails must not linger in interpreter shutdown (destroying an NCCL communicator waits for the
        # peers, which are themselves waiting in a collective): report and leave at once, torchrun ends the job
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(1)
