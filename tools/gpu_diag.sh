mkdir -p gpurun_out
cp manette_b200/libmanette_b200.so /tmp/lib_keep.so
for defs in "" "-DMN_PICAPP_IN" "-DMN_PICW_IN" "-DMN_PICAPP_IN -DMN_PICW_IN" "-DMN_PICAPP_IN -DMN_PICADV_IN" "-DMN_PICAPP_IN -DMN_PICADV_IN -DMN_PICW_IN"; do
  MN_BUILD_DEFS="$defs" python -m manette_b200.build > /dev/null 2>&1 || echo "build failed: $defs"
  echo "defs: [$defs]"; timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -1 | cut -c1-110
done
cp /tmp/lib_keep.so manette_b200/libmanette_b200.so
