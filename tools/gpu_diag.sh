mkdir -p gpurun_out
for hh in 12 14; do for s in 4 8; do
  echo "fifo_high=$hh slack=$s"; MN_FIFO_HIGH=$hh MN_SYNC_SLACK=$s timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 4 2>&1 | tail -1 | cut -c1-110
done; done
