#!/bin/bash
# 1-GPU box: what the driver runs at round end -- GPU suite, smoke, the N = 1 bench line, the reference arm -- plus C3 / C5
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_suite.log 2>&1; tail -3 gpurun_out/r2_final_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2_final_n1.log 2> gpurun_out/r2_final_n1.err; echo "bench rc=$? ${SECONDS}s"
SECONDS=0; timeout 600 python bench.py --impl reference > gpurun_out/r2_final_ref.log 2>&1; echo "reference arm rc=$? ${SECONDS}s"; tail -1 gpurun_out/r2_final_ref.log | cut -c1-300
timeout 300 python bench.py --config mesh --steps 5 --warmup 3 > gpurun_out/r2_final_mesh.log 2>&1
timeout 300 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_final_c5.log 2>&1
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r2_final_n1.log") if l.startswith("{")][-1]
print("N=1 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print(" phases", {k: round(v, 4) for k, v in d["ms_per_step_by_phase"].items() if v > 0})
print(" parity", {k: v for k, v in d["parity"].items() if k not in ("what", "cutoff_shell_note")})
print(" companion", d["same_n_companion"]); print(" cpu", {k: v for k, v in (d["cpu_baseline"] or {}).items() if k != "sample"})
print(" c5_n1", d["extra"]["c5_n1"]["ms_per_step"], {k: round(v, 3) for k, v in d["extra"]["c5_n1"]["ms_per_step_by_phase"].items() if v > 0})
for f in ("mesh", "c5"):
    x = [json.loads(l) for l in open(f"gpurun_out/r2_final_{f}.log") if l.startswith("{")][-1]
    print(f, round(x["ms_per_step"], 3), {k: round(v, 3) for k, v in x["ms_per_step_by_phase"].items() if v > 0}, {k: round(v["frac"], 3) for k, v in x["roofline_kernels"].items()})
PY
