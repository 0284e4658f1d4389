#!/bin/bash
# GPU box: large-volume integrate time vs resident CTAs per SM (register cap)
cd "$(dirname "$0")/.."
for n in 6 8 10 12; do
  VH_EXTRA_NVCC_FLAGS="-DVH_INTEGRATE_MIN_CTAS=$n" python -m voxelhashing_demo_b200._build -f > /dev/null
  echo "min CTAs/SM = $n"
  python - <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
s = torch.cuda.Stream()
r = bench.integrate_hbm_roofline(s)
print("  us", round(r["us"], 1), "frac", round(r["frac"], 3))
PY
done
python -m voxelhashing_demo_b200._build -f > /dev/null
