# two-rank validation on a 2-GPU box: the multi-GPU parity tests and a short bench (run under gpurun --gpus 2)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_multi_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_multi_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
tail -c 1500 gpurun_out/${TAG}_bench2.json
