# ncu --set full of the publication kernels with a host copy registered: k_emit_early of round 2 and the final k_emit
mkdir -p gpurun_out
python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 1 --host-mirror 2>&1 | tail -1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_emit_early -s 2 -c 1 -f -o gpurun_out/decor_r2_emit_early \
    python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 1 --host-mirror > gpurun_out/decor_ncu_r2_emit_early.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_emit<' -s 0 -c 1 -f -o gpurun_out/decor_r2_emit_final \
    python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 1 --host-mirror > gpurun_out/decor_ncu_r2_emit_final.log 2>&1
tail -2 gpurun_out/decor_ncu_r2_emit_early.log gpurun_out/decor_ncu_r2_emit_final.log
ls -la gpurun_out/decor_r2_emit_*.ncu-rep
