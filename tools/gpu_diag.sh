mkdir -p gpurun_out
echo "single partner"; timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -1 | cut -c1-120
cp manette_b200/libmanette_b200.so /tmp/lib_keep.so
MN_BUILD_DEFS="-DMN_PARTNERS=2" python -m manette_b200.build > /dev/null 2>&1
echo "two partners"; MN_DIAG=2 timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -2 | cut -c1-160
cp /tmp/lib_keep.so manette_b200/libmanette_b200.so
