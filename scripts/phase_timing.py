"""Debug: per-phase cycle stamps of the fast forward kernel (build with -DMMG_PHASE_TIMING into scripts/dbg)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodalgame_b200 import capi, engine as eng
from oracle import game_oracle as go
from tests import parity_util as pu
lib = capi.Library(os.path.join(ROOT, "scripts", "dbg", sys.argv[1] if len(sys.argv) > 1 else "libmmg_dbg.so"))
HEAD = dict(img_h_dim=256, baseline_hid_dim=500, sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100,
            entropy_sen=0.01, entropy_rec=0.01, top_k_train=6)
cfg = go.GameConfig(batch_size=64, img_feat_dim=2048, n_classes=30, max_exchange=10, fixed_exchange=True, use_binary=True, **HEAD)
e = eng.GameEngine(pu.config_from(cfg), device="cuda", lib=lib)
e.load_params(go.init_params(cfg, seed=0))
x, desc, target = go.synthetic_batch(cfg, seed=0)
for _ in range(3):
    e.forward(x, desc, target, train=True)
torch.cuda.synchronize()
T, B, M = 10, 64, 32
st = e.ws("g_sen_probs", (T * B * M,)).view(torch.int32).cpu().numpy().astype(np.int64) & 0xffffffff
st = st[:T * 64].reshape(T, 8, 8)
names = ["P1 hidden", "P2 sender msg", "P3 gru", "P4 heads", "P5 scores", "P6 softmax/mix", "P8 rec msg"]
for t in (1, 5):
    s = st[t]
    print("step", t)
    for p in range(7):
        d = (s[p + 1] - s[p]) & 0xffffffff
        print("  %-16s per-warp cycles: min %5d max %5d  warps %s" % (names[p], d.min(), d.max(), " ".join("%4d" % v for v in d)))
    print("  whole step (warp 0): %d cycles" % ((st[t + 1][0][0] - s[0][0]) & 0xffffffff))
