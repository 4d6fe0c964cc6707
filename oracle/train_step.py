"""TEST / BASELINE INFRASTRUCTURE ONLY -- CPU restatement of one stage-2 NAR/FAR training iteration
(reference train_NAR.py:49-107 `single_iter` + `cal_lossT` :33-47; train_FAR.py:48-101), built on oracle/vptr_oracle.py
with torch autograd on the CPU.  Used by bench.py's cpu_baseline / --impl reference legs (the unmodified reference tree
is not present on the GPU box) and by tests.  Never imported by the product package."""
import torch
import torch.nn.functional as F

import vptr_oracle as O


def bi_patch_nce(gt_f, pred_f, temperature):
    """criterion.py:206-259 (BiPatchNCE): positives on the diagonal, stop-gradient on negatives."""
    gt = gt_f.permute(0, 1, 3, 4, 2).flatten(0, 1).flatten(1, 2)
    pr = pred_f.permute(0, 1, 3, 4, 2).flatten(0, 1).flatten(1, 2)
    L = gt.shape[1]
    eye = torch.eye(L)[None]

    def scores(a, b):
        return (torch.matmul(a, b.transpose(1, 2)) * eye + torch.matmul(a, b.detach().transpose(1, 2)) * (1 - eye)) / temperature

    target = torch.arange(L).repeat(gt.shape[0])
    return 0.5 * (F.cross_entropy(scores(gt, pr).flatten(0, 1), target) + F.cross_entropy(scores(pr, gt).flatten(0, 1), target))


def nar_step(sd_enc, sd_dec, sd_T, params, optimizer, past, future, out_layer="Sigmoid", padding_type="reflect", lam_pc=0.1,
             max_grad_norm=1.0, nhead=8, ws=4):
    """One iteration of train_NAR.single_iter with dropout = 0 (the oracle has no RNG; the reference's dropout masks are
    extra CPU work the baseline number therefore does not include).  params: dict name -> leaf tensors of the Transformer
    (a subset view of sd_T's entries).  Returns the loss value."""
    with torch.no_grad():
        past_f = O.resnet_encoder(sd_enc, past, 3, padding_type)
        fut_f = O.resnet_encoder(sd_enc, future, 3, padding_type)
    optimizer.zero_grad(set_to_none=True)
    sd = dict(sd_T)
    sd.update(params)
    pred_f = O.vptr_former_nar(sd, past_f, nhead=nhead, ws=ws, rpe=True, training=True, bn_updates={})
    pred = O.resnet_decoder(sd_dec, pred_f, 3, out_layer)
    proj = lambda t: (F.relu(t.permute(0, 1, 3, 4, 2) @ sd["NCE_projector.0.weight"].t() + sd["NCE_projector.0.bias"])
                      @ sd["NCE_projector.2.weight"].t() + sd["NCE_projector.2.bias"]).permute(0, 1, 4, 2, 3)
    pf, gf = proj(pred_f), proj(fut_f)
    loss = O.mse_loss(pred, future) + O.gdl_loss(future, pred) + lam_pc * bi_patch_nce(F.normalize(gf, p=2.0, dim=2), F.normalize(pf, p=2.0, dim=2), 1.0)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(params.values()), max_norm=max_grad_norm, norm_type=2)
    optimizer.step()
    return float(loss.detach())
