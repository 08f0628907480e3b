mkdir -p gpurun_out
cat > /tmp/seq.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, manette_b200 as mb
from manette_b200.learner import PAACLearner
from util import rom_bytes
tab = mb.tab_repetitions(10, 11)
os.environ.pop("MN_DIAG", None)
pool = mb.DevicePool([("pong", rom_bytes("pong"), 32)], tab_rep=tab)
pool.reset_all()
L = PAACLearner(pool, arch="NIPS", seed=5)
L.train_rollout()
L.close(); pool.close()
os.environ["MN_DIAG"] = sys.argv[1]
game, n, hist = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
try:
    p2 = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=tab, history=hist)
    print("diag", sys.argv[1:], "second pool ok")
except Exception as e:
    print("diag", sys.argv[1:], "FAIL", str(e)[:90])
PY
for d in 0 32 36 40 48; do CUDA_LAUNCH_BLOCKING=1 timeout 120 python /tmp/seq.py $d breakout 8 5 2>&1 | tail -1; done
for cfg in "breakout 8 0" "breakout 32 0" "pong 8 0" "pong 32 5" "breakout 64 5"; do CUDA_LAUNCH_BLOCKING=1 timeout 120 python /tmp/seq.py 0 $cfg 2>&1 | tail -1; done
