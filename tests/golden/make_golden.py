"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libp3m_ref.so, built by
`make -C oracle ref` from /root/reference).  Run in the build container:

    python tests/golden/make_golden.py

The reference's own tests hold no vectors for the P3M path, so these outputs of
the reference itself are the pin: inputs (fixed numpy seeds), every intermediate of one force
evaluation, and the diagnostics of a 20-step run.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refapi  # noqa: E402
from common import disk_case, plummer_case  # noqa: E402

CASES = {
    # name: (builder, n, kwargs, p3m)
    "plummer_p3m_tsc_s1": (plummer_case, 768, dict(grid=(16, 16, 16)), True),
    "plummer_p3m_cic_s2_4pt": (plummer_case, 512, dict(grid=(16, 16, 16), is_=refapi.CIC, fds=refapi.FOUR_POINT,
                                                       gfunc=refapi.S2_OPTIMAL, cloud=refapi.S2), True),
    "disk_pm_tsc_laplacian_ext": (disk_case, 768, dict(grid=(32, 32, 16), gfunc=refapi.DISCRETE_LAPLACIAN), False),
    "plummer_pm_ngp_poorman": (plummer_case, 512, dict(grid=(16, 16, 16), is_=refapi.NGP, gfunc=refapi.POOR_MAN), False),
    "plummer_p3m_analytic": (plummer_case, 512, dict(grid=(16, 16, 16), use_table=False, y_sort=False), True),
}


def params_dict(p):
    d = {}
    for name, _ in p._fields_:
        v = getattr(p, name)
        d[name] = np.array(list(v)) if hasattr(v, "__len__") else v
    return d


def main():
    ref = refapi.Ref()
    for name, (mk, n, kw, p3m) in CASES.items():
        p, pos, vel, mass = mk(n, **kw)
        out = dict(pos=pos, vel=vel, mass=mass, p3m=np.int32(p3m))
        out.update({"param_" + k: v for k, v in params_dict(p).items()})
        r = ref.pm_force(p, pos, vel, mass, want_green=True)
        assert np.abs(r["green"][..., 1]).max() == 0
        out.update(pos_code=r["pos_code"], mass_code=r["mass_code"], green=r["green"][..., 0],
                   density=r["density"], potential=r["potential"], field=r["field"], acc_pm=r["acc"])
        if p3m:
            r3 = ref.p3m_force(p, pos, vel, mass)
            out.update(sr_force=r3["sr_force"], acc=r3["acc"], cell=r3["cell"], order=r3["order"],
                       chain_dims=r3["dims"], ftable=r3["ftable"])
        with tempfile.TemporaryDirectory() as d:
            diag, po, vo, ao = ref.run(p, pos, vel, mass, 20, p3m, d)
        out.update(run_diag=diag, run_pos=po, run_vel=vo, run_acc=ao)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
