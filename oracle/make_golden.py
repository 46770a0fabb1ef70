"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run once, here (the GPU box has no /root/reference):  python oracle/make_golden.py
Imports /root/reference/AttnGAN2/code/GlobalAttention.py and miscc/losses.py as they are;
the only shim is a stand-in for the missing ``easydict`` package.
Inputs come from oracle/synth.py (numpy RandomState: frozen stream), so fixtures hold
only the reference's OUTPUTS plus input checksums that detect generator drift.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/AttnGAN2/code"
sys.path.insert(0, ROOT)

from oracle.synth import synth_attention_inputs, synth_words_loss_inputs  # noqa: E402


def import_reference():
    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    mod = types.ModuleType("easydict")
    mod.EasyDict = EasyDict
    sys.modules.setdefault("easydict", mod)
    sys.path.insert(0, REF)
    import GlobalAttention as ref_ga          # noqa
    from miscc import losses as ref_losses    # noqa
    from miscc.config import cfg              # noqa
    cfg.CUDA = False
    return ref_ga, ref_losses, cfg


def import_reference_models():
    """model_bert.py of the reference (G_NET, NEXT_STAGE_G, ADAIN_NORM ...) with stand-ins for the two
    uninstalled packages it names at import time; bird_style.yml dimensions."""
    import_reference()
    ppb = types.ModuleType("pytorch_pretrained_bert")
    ppb.BertModel = type("BertModel", (), {})
    sys.modules.setdefault("pytorch_pretrained_bert", ppb)
    from miscc.config import cfg
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM = 32, 64, 100, 2      # cfg/bird_style.yml
    cfg.TREE.BRANCH_NUM, cfg.TEXT.EMBEDDING_DIM, cfg.GAN.CONDITION_DIM = 3, 256, 100
    import model_bert
    return model_bert, cfg


STAGE_CASES = {
    # name: (B, idf, cdf, L, ih, iw, seed)      attention + AdaIN + cat of NEXT_STAGE_G.forward (model_bert.py:458-461)
    "stage_b3_16x8": (3, 32, 256, 18, 16, 8, 51),
    "stage_b2_16x16": (2, 32, 256, 12, 16, 16, 52),
}


def synth_stage_inputs(B, idf, cdf, L, ih, iw, seed, dtype, w_dim):
    d = synth_attention_inputs(B, idf, cdf, L, ih, iw, seed=seed, dtype=dtype, with_g_attn=True)
    rs = np.random.RandomState(seed + 1000)
    d["w_code"] = torch.from_numpy(rs.standard_normal((B, w_dim))).to(dtype)
    d["style_weight"] = torch.from_numpy(rs.standard_normal((2 * idf, w_dim)) / np.sqrt(w_dim)).to(dtype)
    d["style_bias"] = torch.from_numpy(0.1 * rs.standard_normal(2 * idf)).to(dtype)
    d["g_buf"] = torch.from_numpy(rs.standard_normal((B, 2 * idf, ih, iw))).to(dtype)
    return d


def run_stage(model_bert, cfg, ref_ga, spec, dtype):
    B, idf, cdf, L, ih, iw, seed = spec
    d = synth_stage_inputs(B, idf, cdf, L, ih, iw, seed, dtype, cfg.GAN.W_DIM)
    att = ref_ga.GlobalAttentionGeneral(idf, cdf).to(dtype)
    adain = model_bert.ADAIN_NORM(idf).to(dtype)
    with torch.no_grad():
        att.conv_context.weight.copy_(d["weight"])
        adain.style.weight.copy_(d["style_weight"])
        adain.style.bias.copy_(d["style_bias"])
    h = d["x"].clone().requires_grad_(True)
    w = d["w_code"].clone().requires_grad_(True)
    ctx = d["context"].clone().requires_grad_(True)
    att.applyMask(d["mask"])
    c_code, attn = att(h, ctx)                       # model_bert.py:458-459
    hn = adain(h, w)                                 # :460
    h_c = torch.cat((hn, c_code), 1)                 # :461
    ((h_c * d["g_buf"]).sum() + (attn * d["g_attn"]).sum()).backward()
    return dict(h_c_code=h_c.detach().numpy(), attn=attn.detach().numpy(), dX=h.grad.numpy(), dCtx=ctx.grad.numpy(),
                dW=att.conv_context.weight.grad.numpy(), d_w_code=w.grad.numpy(),
                d_style_weight=adain.style.weight.grad.numpy(), d_style_bias=adain.style.bias.grad.numpy(),
                in_sum=np.concatenate([checksum(d["x"]), checksum(d["w_code"]), checksum(d["g_buf"])]))


ATTN_CASES = {
    # name: (B, idf, cdf, L, ih, iw, seed, masked, with_g_attn)
    "attn_b3_q64_rag": (3, 32, 256, 18, 8, 8, 11, True, True),       # Q % B != 0
    "attn_b4_q64_div": (4, 32, 256, 18, 8, 8, 12, True, False),      # Q % B == 0
    "attn_b2_idf48_nomask": (2, 48, 256, 12, 6, 6, 13, False, True),
    "attn_b5_idf128_l25": (5, 128, 256, 25, 4, 4, 14, True, False),
    "attn_b7_q289_odd": (7, 32, 64, 20, 17, 17, 15, True, True),     # Q odd (no 128-bit path)
    "attn_cfg1_b10_64x64": (10, 32, 256, 18, 64, 64, 1234, True, False),   # BASELINE configs[0]
}
SUBSAMPLE = {"attn_cfg1_b10_64x64": 97}
# fp64 runs of the reference (tie-breaker) are kept only for the small cases
F64_SKIP = {"attn_b7_q289_odd", "wl_b4_full", "wl_b12_dupclass"}

WL_CASES = {
    # name: (B, nef, L, ih, iw, seed, gammas, use_class_ids, use_labels)
    "wl_b4_full": (4, 256, 18, 17, 17, 21, (4.0, 5.0, 10.0), True, True),
    "wl_b6_small": (6, 32, 7, 5, 5, 22, (5.0, 5.0, 10.0), True, True),
    "wl_b5_noclass": (5, 64, 12, 6, 6, 23, (4.0, 5.0, 10.0), False, True),
    "wl_b3_nolabels": (3, 256, 18, 17, 17, 24, (4.0, 5.0, 10.0), False, False),
    "wl_b12_dupclass": (12, 64, 18, 17, 17, 25, (4.0, 5.0, 10.0), True, True),
}


SL_CASES = {
    # name: (B, nef, seed, gamma3, use_class_ids)       sent_loss (losses.py:20-59)
    "sl_b6": (6, 32, 41, 10.0, True),
    "sl_b12_dupclass": (12, 256, 42, 10.0, True),
    "sl_b5_noclass": (5, 64, 43, 5.0, False),
}


def synth_sent_inputs(B, nef, seed, dtype):
    rs = np.random.RandomState(seed)
    cnn = torch.from_numpy(rs.standard_normal((B, nef))).to(dtype)
    rnn = torch.from_numpy(np.tanh(rs.standard_normal((B, nef)))).to(dtype)
    cls = rs.randint(1, max(2, B // 2) + 1, size=B)
    return cnn, rnn, cls


def run_sent_loss(ref_losses, cfg, spec, dtype):
    B, nef, seed, g3, use_cls = spec
    cnn, rnn, cls = synth_sent_inputs(B, nef, seed, dtype)
    cfg.TRAIN.SMOOTH.GAMMA3 = g3
    a, b = cnn.clone().requires_grad_(True), rnn.clone().requires_grad_(True)
    l0, l1 = ref_losses.sent_loss(a, b, torch.arange(B), cls if use_cls else None, B)
    (l0 + 2.0 * l1).backward()              # different weights: the two gradients are told apart
    return dict(loss0=l0.detach().numpy(), loss1=l1.detach().numpy(), d_cnn=a.grad.numpy(), d_rnn=b.grad.numpy(),
                class_ids=cls, in_sum=np.concatenate([checksum(cnn), checksum(rnn)]))


def checksum(t):
    t = t.double()
    return np.array([t.sum().item(), t.abs().sum().item()])


def run_attention(ref_ga, name, spec, dtype):
    B, idf, cdf, L, ih, iw, seed, masked, with_ga = spec
    d = synth_attention_inputs(B, idf, cdf, L, ih, iw, seed=seed, dtype=dtype, with_g_attn=with_ga)
    mod = ref_ga.GlobalAttentionGeneral(idf, cdf).to(dtype)
    with torch.no_grad():
        mod.conv_context.weight.copy_(d["weight"])
    x = d["x"].clone().requires_grad_(True)
    ctx = d["context"].clone().requires_grad_(True)
    mod.applyMask(d["mask"] if masked else None)
    c, attn = mod(x, ctx)
    loss = (c * d["g_c"]).sum()
    if with_ga:
        loss = loss + (attn * d["g_attn"]).sum()
    loss.backward()
    out = dict(c_code=c.detach(), attn=attn.detach(), dX=x.grad, dW=mod.conv_context.weight.grad, dCtx=ctx.grad)
    step = SUBSAMPLE.get(name, 1)
    rec = {}
    for k, v in out.items():
        flat = v.reshape(-1)
        rec[k] = flat[::step].numpy() if step > 1 and flat.numel() > 100000 else v.numpy()
        rec[k + "_sum"] = checksum(v)
    rec["in_sum"] = np.concatenate([checksum(d["x"]), checksum(d["context"]), checksum(d["weight"]), checksum(d["g_c"])])
    return rec


def run_words_loss(ref_losses, cfg, name, spec, dtype):
    B, nef, L, ih, iw, seed, gammas, use_cls, use_lab = spec
    d = synth_words_loss_inputs(B, nef, L, ih, iw, seed=seed, dtype=dtype, n_classes=max(2, B // 2))
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3 = gammas
    img = d["img_features"].clone().requires_grad_(True)
    words = d["words_emb"].clone().requires_grad_(True)
    captured = []

    real_ce = torch.nn.CrossEntropyLoss

    class CapturingCE(real_ce):
        def forward(self, inp, tgt):
            captured.append(inp.detach().clone())
            return super().forward(inp, tgt)

    ref_losses.nn.CrossEntropyLoss = CapturingCE
    try:
        loss0, loss1, att_maps = ref_losses.words_loss(
            img, words, d["labels"] if use_lab else None, d["cap_lens"],
            d["class_ids"] if use_cls else None, B)
    finally:
        ref_losses.nn.CrossEntropyLoss = real_ce
    rec = {"in_sum": np.concatenate([checksum(d["img_features"]), checksum(d["words_emb"])]),
           "cap_lens": d["cap_lens"].numpy(), "class_ids": d["class_ids"]}
    for i, a in enumerate(att_maps):
        rec[f"att_{i}"] = a.detach().numpy()
    if use_lab:
        (loss0 + loss1).backward()
        rec.update(loss0=loss0.detach().numpy(), loss1=loss1.detach().numpy(),
                   sim_masked=captured[0].numpy(), d_img=img.grad.numpy(), d_words=words.grad.numpy())
    return rec


def run_func_attention(ref_ga, dtype):
    rs = np.random.RandomState(31)
    q = torch.from_numpy(np.tanh(rs.standard_normal((3, 64, 9)))).to(dtype)
    c = torch.from_numpy(rs.standard_normal((3, 64, 5, 7))).to(dtype)
    wc, attn = ref_ga.func_attention(q, c, 4.0)
    return dict(query=q.numpy(), context=c.numpy(), wc=wc.numpy(), attn=attn.numpy())


def main():
    ref_ga, ref_losses, cfg = import_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)          # fixed summation order in the CPU kernels
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        for name, spec in ATTN_CASES.items():
            if tag == "f64" and name in F64_SKIP:
                continue
            rec = run_attention(ref_ga, name, spec, dtype)
            np.savez_compressed(os.path.join(outdir, f"{name}_{tag}.npz"), **rec)
        for name, spec in WL_CASES.items():
            if tag == "f64" and name in F64_SKIP:
                continue
            rec = run_words_loss(ref_losses, cfg, name, spec, dtype)
            np.savez_compressed(os.path.join(outdir, f"{name}_{tag}.npz"), **rec)
        np.savez_compressed(os.path.join(outdir, f"func_attention_{tag}.npz"), **run_func_attention(ref_ga, dtype))
        for name, spec in SL_CASES.items():
            np.savez_compressed(os.path.join(outdir, f"{name}_{tag}.npz"), **run_sent_loss(ref_losses, cfg, spec, dtype))
        model_bert, _ = import_reference_models()
        for name, spec in STAGE_CASES.items():
            np.savez_compressed(os.path.join(outdir, f"{name}_{tag}.npz"), **run_stage(model_bert, cfg, ref_ga, spec, dtype))
    total = sum(os.path.getsize(os.path.join(outdir, f)) for f in os.listdir(outdir))
    print(f"wrote {len(os.listdir(outdir))} fixtures, {total / 1e6:.2f} MB, torch {torch.__version__}")


if __name__ == "__main__":
    main()
