"""Debug: phase stamps of the fast backward kernel (receiver CTA 0, sender CTA n_rec)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodalgame_b200 import capi, engine as eng
from oracle import game_oracle as go
from tests import parity_util as pu
lib = capi.Library(os.path.join(ROOT, "scripts", "dbg", "libmmg_dbg.so"))
HEAD = dict(img_h_dim=256, baseline_hid_dim=500, sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100,
            entropy_sen=0.01, entropy_rec=0.01, top_k_train=6)
cfg = go.GameConfig(batch_size=64, img_feat_dim=2048, n_classes=30, max_exchange=10, fixed_exchange=True, use_binary=True, **HEAD)
e = eng.GameEngine(pu.config_from(cfg), device="cuda", lib=lib)
e.load_params(go.init_params(cfg, seed=0))
x, desc, target = go.synthetic_batch(cfg, seed=0)
for _ in range(3):
    e.forward(x, desc, target, train=True); e.loss(); e.backward()
torch.cuda.synchronize()
e.forward(x, desc, target, train=True); e.loss()
torch.cuda.synchronize()
e.backward()
torch.cuda.synchronize()
st = e.ws("g_bs", (640,)).view(torch.int32).cpu().numpy().astype(np.int64) & 0xffffffff
rec, sen = st[:6], st[32:36]
names = ["A0 issue loads", "wait TMA + loads", "A1 d_hw / class head", "A2 injections", "B chain (T steps)"]
for i, n in enumerate(names):
    print("receiver %-22s %6d cycles" % (n, (rec[i + 1] - rec[i]) & 0xffffffff))
print("receiver total %d" % ((rec[5] - rec[0]) & 0xffffffff))
for i, n in enumerate(["load a_s / d logits", "T-step loop", "d code partial"]):
    print("sender   %-22s %6d cycles" % (n, (sen[i + 1] - sen[i]) & 0xffffffff))
print("sender total %d" % ((sen[3] - sen[0]) & 0xffffffff))
