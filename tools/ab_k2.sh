# A/B of two library builds on the same box: exaconstit_b200/lib_base (baseline) vs exaconstit_b200/lib (current)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-ab}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for cfg in "base:lib_base" "new:lib" "base2:lib_base" "new2:lib"; do
  tag=${cfg%%:*}; dir=${cfg#*:}
  EXAB200_LIBDIR=$GRAFT_REPO_ROOT/exaconstit_b200/$dir timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$tag.json 2> gpurun_out/${TAG}_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$tag.json"))
    print("$tag", round(d["value"], 4), "K1", round(d["model_setup"]["avg_ms"], 2), "K2", round(d["roofline"]["avg_launch_ms"], 4), "us/it", round(d["step_anatomy"]["us_per_pcg_iteration_outside_k1"], 1), d["parity_fingerprint"]["max_rel_err"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag FAILED", e)
PY
done
