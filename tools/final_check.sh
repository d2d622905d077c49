set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; tail -2 gpurun_out/r2z_pytest.log
timeout 300 python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_config5.json 2> gpurun_out/r2z_config5.err; tail -c 300 gpurun_out/r2z_config5.json
timeout 300 python bench.py --deterministic --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_determ.json 2> gpurun_out/r2z_determ.err; tail -c 300 gpurun_out/r2z_determ.json
