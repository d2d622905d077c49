# short N-rank bench on an N-GPU box (run under gpurun --gpus N): bash tools/check_ngpu.sh N TAG
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}; TAG=${2:-r2h}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
tail -c 1200 gpurun_out/${TAG}_bench$N.json; tail -5 gpurun_out/${TAG}_bench$N.err
