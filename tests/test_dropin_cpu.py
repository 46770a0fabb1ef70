"""Drop-in test against the REAL reference generator (SURVEY.md §4 "drop-in test", §8 a-5): with
``sba_gan_b200.install()`` in effect, the unmodified ``/root/reference/AttnGAN2/code/model_bert.py``
builds ``G_NET`` with this repo's ``GlobalAttentionGeneral`` as ``ATT_NET`` (model_bert.py:12, 445-459),
a state dict produced by the all-reference ``G_NET`` loads strictly (keys ``h_net2.att.conv_context.weight``
/ ``h_net3.att.conv_context.weight``, model_bert.py:562-594), the reference's ``weights_init``
(miscc/utils.py:286-296) still initialises ``conv_context``, and ``miscc.losses`` / the trainers get the
fused ``func_attention`` / ``words_loss``.  Runs where the reference tree exists (this container);
the reference is imported in a subprocess so that its module names never leak into the test session."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/AttnGAN2/code"

SCRIPT = textwrap.dedent('''
    import sys, types
    import torch
    sys.path.insert(0, %(root)r)
    from oracle.make_golden import import_reference          # easydict stand-in + sys.path of the reference
    import_reference()
    # pytorch_pretrained_bert is not installed: model_bert only needs the name at import time
    ppb = types.ModuleType("pytorch_pretrained_bert")
    ppb.BertModel = type("BertModel", (), {})
    sys.modules["pytorch_pretrained_bert"] = ppb
    sk = types.ModuleType("skimage"); skt = types.ModuleType("skimage.transform"); sk.transform = skt
    sys.modules["skimage"] = sk; sys.modules["skimage.transform"] = skt
    from miscc.config import cfg
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM = 32, 64, 100, 2      # cfg/bird_style.yml
    cfg.TREE.BRANCH_NUM, cfg.TEXT.EMBEDDING_DIM, cfg.GAN.CONDITION_DIM = 3, 256, 100
    cfg.GAN.W_DIM = getattr(cfg.GAN, "W_DIM", 256)

    import GlobalAttention as ref_ga
    import model_bert
    assert model_bert.ATT_NET is ref_ga.GlobalAttentionGeneral
    torch.manual_seed(0)
    ref_net = model_bert.G_NET()
    ref_sd = ref_net.state_dict()
    assert "h_net2.att.conv_context.weight" in ref_sd and "h_net3.att.conv_context.weight" in ref_sd
    assert tuple(ref_sd["h_net2.att.conv_context.weight"].shape) == (32, 256, 1, 1)

    import sba_gan_b200
    sba_gan_b200.install()                  # patches the already imported model_bert / miscc.losses
    assert model_bert.ATT_NET is sba_gan_b200.GlobalAttentionGeneral
    net = model_bert.G_NET()
    for stage in (net.h_net2, net.h_net3):
        assert type(stage.att) is sba_gan_b200.GlobalAttentionGeneral, type(stage.att)
    assert list(net.state_dict()) == list(ref_sd), "state_dict layout differs from the reference generator"
    missing, unexpected = net.load_state_dict(ref_sd, strict=True)
    assert not missing and not unexpected
    assert torch.equal(net.h_net3.att.conv_context.weight, ref_net.h_net3.att.conv_context.weight)

    # the reference initialiser still reaches conv_context (class name contains "Conv")
    from miscc.utils import weights_init
    net.apply(weights_init)
    w = net.h_net2.att.conv_context.weight.detach().reshape(32, 256)
    assert torch.allclose(w @ w.t(), torch.eye(32), atol=1e-4), "orthogonal init did not reach conv_context"

    # the loss module and (when imported) the trainers call the fused operators
    import miscc.losses as ref_losses
    assert ref_losses.func_attention is sba_gan_b200.func_attention
    assert ref_losses.words_loss is sba_gan_b200.words_loss
    # applyMask / forward signature as NEXT_STAGE_G.forward uses them (model_bert.py:458-459); the forward itself
    # needs a GPU: on a CPU tensor it must refuse instead of falling back
    net.h_net2.att.applyMask(torch.zeros(2, 18, dtype=torch.bool))
    try:
        net.h_net2.att(torch.zeros(2, 32, 8, 8), torch.zeros(2, 256, 18))
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("CPU forward did not refuse")
    print("DROPIN-OK", len(ref_sd))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_generator_builds_with_the_replacement_attention():
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DROPIN-OK" in r.stdout
