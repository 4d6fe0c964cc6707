"""GPU: the whole training iteration captured as ONE CUDA graph (vptr_b200.trainer.GraphedStep) -- the C-ABI's claim that every entry
point is capturable (borrowed pointers, explicit stream, no allocation, no synchronisation), as a test.  Replays must (1) reproduce
the eager iterations when dropout is off, optimizer step count included (it lives on the device), and (2) draw NEW dropout masks on
every replay although the seeds are frozen kernel arguments (device-side epoch, vptr_rng_advance)."""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def _build(kind, dropout):
    from vptr_b200.model import VPTRDec, VPTREnc, VPTRFormerFAR, VPTRFormerNAR
    from vptr_b200.trainer import Stage2Trainer
    dev = torch.device("cuda")
    torch.manual_seed(21)
    enc = VPTREnc(1, feat_dim=528, n_downsampling=3).to(dev).eval()
    dec = VPTRDec(1, feat_dim=528, n_downsampling=3, out_layer="Sigmoid").to(dev).eval()
    if kind == "nar":
        T = VPTRFormerNAR(3, 3, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=dropout, window_size=4, rpe=True).to(dev)
    else:
        T = VPTRFormerFAR(3, 3, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=2, dropout=dropout, window_size=4, rpe=True).to(dev)
    return Stage2Trainer(kind, enc, dec, T, lr=1e-4, use_bpnce=True), T


@pytest.mark.parametrize("kind", ["nar", "far"])
def test_graph_replay_matches_eager_steps(kind):
    from vptr_b200 import _lib, ops
    from vptr_b200.trainer import GraphedStep
    g = torch.Generator().manual_seed(2)
    past, fut = torch.rand(2, 3, 1, 64, 64, generator=g).cuda(), torch.rand(2, 3, 1, 64, 64, generator=g).cuda()
    tr_e, T_e = _build(kind, 0.0)
    losses_e = [float(tr_e.step(past, fut)) for _ in range(6)]
    tr_g, T_g = _build(kind, 0.0)
    gs = GraphedStep(tr_g, past, fut, warmup=3)                       # 3 eager warm-up iterations, then the capture (executes nothing)
    losses_g = [float(gs.step(past, fut)) for _ in range(3)]          # iterations 4, 5 and 6
    _lib.call("vptr_rng_advance", 0, ops._s())                        # back to the eager default epoch for the other tests
    for a, b in zip(losses_g, losses_e[3:]):
        assert abs(a - b) <= 2e-5 * abs(b), (losses_e, losses_g)
    worst = 0.0
    for (k, a), (_, b) in zip(T_e.named_parameters(), T_g.named_parameters()):
        worst = max(worst, float((a - b).abs().max()))
    assert worst <= 6.5e-4, worst                                     # 6 AdamW steps of lr 1e-4: sign-level noise of cancelling gradients at most
    st = tr_g.tail.opt.state_dict()["state"]
    assert all(float(v["step"]) == 6.0 for v in st.values())          # host-side step count follows the replays (checkpoints)


def test_graph_replays_draw_new_dropout_masks():
    from vptr_b200 import _lib, ops
    from vptr_b200.trainer import GraphedStep
    g = torch.Generator().manual_seed(3)
    past, fut = torch.rand(2, 3, 1, 64, 64, generator=g).cuda(), torch.rand(2, 3, 1, 64, 64, generator=g).cuda()
    tr, T = _build("nar", 0.3)
    for p in T.parameters():
        p.requires_grad_(True)
    tr.tail.opt.param_groups[0]["lr"] = 0.0                           # frozen weights: only the masks can change the loss
    tr.tail.opt.param_groups[0]["weight_decay"] = 0.0
    gs = GraphedStep(tr, past, fut, warmup=3)
    losses = [float(gs.step(past, fut)) for _ in range(4)]
    _lib.call("vptr_rng_advance", 0, ops._s())
    assert len(set(round(l, 7) for l in losses)) == 4, losses         # four replays, four different dropout realisations
    assert max(losses) - min(losses) < 0.2 * abs(losses[0])           # ... of the same function
