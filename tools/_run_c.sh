timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err; echo rc=$?
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo rc=$?
timeout 400 python bench.py > gpurun_out/r2_final_n1_default.json 2> gpurun_out/r2_final_n1_default.err; echo rc=$?
python - <<'PY'
import json
for f in ("r2_final_n1", "r2_final_n1_default", "r2_final_ref"):
    d = json.loads(open(f"gpurun_out/{f}.json").read())
    print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), d.get("passes", {}).get("timed_ms"), d.get("roofline", {}).get("frac"), (d.get("roofline_integrate_hbm") or {}).get("frac"), (d.get("refexact_leg") or {}).get("value"), d.get("variants"))
PY
