# linked-slab protocol A/B on N GPUs (tuning aid): bash tools/probe8.sh N
N=${1:-8}
run() { echo "## $1 | $3"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 tools/slab_probe.py $3 --T 45 --reps 2 --modes 0:2 2>&1 | grep slabs; }
S="--nxg 1024 --ny 1024 --nz 1024"
W="--nxg $((512*N)) --ny 512 --nz 512"
run "FDTD_B200_HALO_PULL=1" 29701 "$S"
run "FDTD_B200_HALO_PULL=0" 29702 "$S"
run "FDTD_B200_HALO_PULL=1" 29703 "$W"
run "FDTD_B200_HALO_PULL=0" 29704 "$W"
