"""Extra data point (not a bench.py line): one stage-2 VPTR-FAR training iteration at the cfg2 shape on ONE GPU --
KTH-shape 10 -> 20 (T = 29 input frames), 64x64x1, 16 clips, 12 encoder layers, causal temporal attention, dropout 0.1,
train_FAR.single_iter (reference train_FAR.py:49-101): Enc (no_grad) -> Transformer -> Dec -> MSE + GDL -> backward -> clip -> AdamW."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vptr_b200.model import GDL, MSELoss, VPTRDec, VPTREnc, VPTRFormerFAR, init_weights
import contextlib, io

dev = torch.device("cuda", 0)
N, Tp, Tf = int(os.environ.get("CLIPS", "16")), 10, 20
torch.manual_seed(2021)
enc = VPTREnc(1, feat_dim=528, n_downsampling=3).to(dev).eval()
dec = VPTRDec(1, feat_dim=528, n_downsampling=3, out_layer="Tanh").to(dev).eval()
with contextlib.redirect_stdout(io.StringIO()):
    init_weights(enc); init_weights(dec)
T = VPTRFormerFAR(Tp, Tf, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=12, dropout=0.1, window_size=4, rpe=True).to(dev)
opt = torch.optim.AdamW(T.parameters(), lr=1e-4)
mse, gdl = MSELoss(), GDL(alpha=1)
g = torch.Generator().manual_seed(2021)
past = (torch.rand(N, Tp, 1, 64, 64, generator=g) * 2 - 1).to(dev)
fut = (torch.rand(N, Tf, 1, 64, 64, generator=g) * 2 - 1).to(dev)


def step():
    x = torch.cat([past, fut[:, :-1]], 1)                       # frames 0..T-2 predict 1..T-1
    with torch.no_grad():
        feats = enc(x)
    T.train(); T.zero_grad(set_to_none=True)
    pred = dec(T(feats))
    target = torch.cat([past[:, 1:], fut], 1)
    loss = mse(pred, target) + gdl(target, pred)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(T.parameters(), 1.0)
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("FAR cfg2 shape, %d clips x (T=29), 12 layers, 1 GPU: %.1f ms/step, %.0f predicted frames/s (29 per clip), %.0f (20 future per clip), loss %.4f, peak mem %.1f GiB"
      % (N, ms, N * 29 / ms * 1e3, N * 20 / ms * 1e3, float(loss), torch.cuda.max_memory_allocated() / 2 ** 30))
