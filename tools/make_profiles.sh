#!/bin/bash
# gpurun_out/ (scratch) -> profiles/ (tracked): summaries of the reports tools/gpu_final_profiles.sh <tag> brought back.
# usage: tools/make_profiles.sh <tag> [name-prefix under profiles/, default r2]
tag=${1:-r2}; pre=${2:-r2}; out=profiles
python tools/ncu_counters.py gpurun_out/decor_${tag}_round0.ncu-rep $out/${pre}_k_round_counters.json --next-calls 16384 \
  --note "one k_round launch = round 0 of a macro step of a decorrelated 16,384-env Ms Pacman pool (every env on the work list)"
python tools/ncu_counters.py gpurun_out/decor_${tag}_k3.ncu-rep $out/${pre}_k3_counters.json --next-calls 16384 \
  --note "one k_push_frames launch over the 16,384 gray envs of round 0"
for r in round0 round5; do
  python tools/ncu_summary.py gpurun_out/decor_${tag}_$r.ncu-rep > $out/${pre}_k_round_${r}_summary.txt
  ncu -i gpurun_out/decor_${tag}_$r.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); H,U,R=rows[0],rows[1],rows[2]
items=[]
for i,h in enumerate(H):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
        try: items.append((float(R[i]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except ValueError: pass
print('warp-state cycles per issued instruction (ncu smsp__average_warps_issue_stalled_*_per_issue_active), k_round, $r')
for v,h in sorted(items,reverse=True): print('%8.3f  %s'%(v,h))
for k in ('sm__icc_request_hit_rate.pct','gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active'):
    if k in H: print('%-80s %s %s'%(k,R[H.index(k)],U[H.index(k)]))
" > $out/${pre}_k_round_${r}_stalls.txt
done
python tools/ncu_summary.py gpurun_out/decor_${tag}_k3.ncu-rep > $out/${pre}_k3_push_frames_summary.txt
ncu -i gpurun_out/decor_${tag}_round0.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src_final.csv 2>/dev/null
python tools/ncu_funcs.py /tmp/src_final.csv > $out/${pre}_k_round_round0_by_function.txt 2>&1
cp gpurun_out/${tag}_launches.csv $out/${pre}_launches.csv
cp gpurun_out/${tag}_per_game_n16384.txt $out/${pre}_per_game_n16384.txt
for w in pong_paac_n32 breakout_figar10_n256 seaquest_figar10_rgb_n4096 ms_pacman_figar10_n16384 mixed12_figar10_n16384; do
  tail -1 gpurun_out/${tag}_bench_$w.json > $out/${pre}_bench_$w.json
done
ls -la $out | grep ${pre}_
