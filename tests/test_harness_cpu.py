"""The end-to-end measurement harness (harness/gan_step.py) runs one full G+D step on CPU with the
eager (reference op sequence) hot path - shapes, losses finite, parameters move.  The fused path
of the same step is exercised on the B200 box by tools/e2e_train.py."""
import torch

from harness.gan_step import EagerAttention, Trainer, eager_words_loss
import oracle


def test_eager_ops_match_oracle():
    d = oracle.synth_attention_inputs(3, 32, 256, 18, 8, 8, seed=5)
    m = EagerAttention(32, 256)
    with torch.no_grad():
        m.conv_context.weight.copy_(d["weight"])
    m.applyMask(d["mask"])
    c, a = m(d["x"], d["context"])
    cr, ar, _ = oracle.attn_forward(d["x"], d["context"], d["weight"], d["mask"])
    assert oracle.normalised_max_err(c, cr) < 1e-5 and oracle.normalised_max_err(a, ar) < 1e-5
    w = oracle.synth_words_loss_inputs(4, 256, 18, 17, 17, seed=6)
    cls = torch.as_tensor(w["class_ids"])
    l0, l1, _ = eager_words_loss(w["img_features"], w["words_emb"], w["labels"], w["cap_lens"], cls, 4)
    r0, r1, _ = oracle.words_loss(w["img_features"], w["words_emb"], w["labels"], w["cap_lens"], w["class_ids"], 4, 4.0, 5.0, 10.0)
    assert abs(l0.item() - r0.item()) < 1e-4 and abs(l1.item() - r1.item()) < 1e-4


def test_one_training_step_on_cpu():
    torch.set_num_threads(4)
    tr = Trainer(2, torch.device("cpu"), "eager", 1, seed=3)
    before = [p.detach().clone() for p in tr.G.parameters()]
    errD, errG = tr.step()
    assert torch.isfinite(errD) and torch.isfinite(errG)
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.G.parameters()))
