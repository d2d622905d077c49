# Round-end evidence on one B200 (run under gpurun): the driver's bench command, the ncu launch list of a shortened
# bench run and one `ncu --set full` capture of the dominant kernel (PA gradient apply at 128^3).
# usage: bash tools/capture_profiles.sh TAG
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r2f}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt
timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
tail -c 600 gpurun_out/${TAG}_bench1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 2 --krylov-iter 150 --no-cpu-baseline --no-same-config > gpurun_out/${TAG}_launches.out 2> gpurun_out/${TAG}_launches.err
gzip -f gpurun_out/${TAG}_launches.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_grad_mult_pa_c -s 3 -c 1 -f -o /tmp/${TAG}_k2 \
  python tools/bench_kernels.py 128 > gpurun_out/${TAG}_k2.log 2>&1
ncu -i /tmp/${TAG}_k2.ncu-rep --page raw --csv > gpurun_out/${TAG}_k2.raw.csv
ncu -i /tmp/${TAG}_k2.ncu-rep --page details > gpurun_out/${TAG}_k2.details.txt
