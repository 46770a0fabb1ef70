"""CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU) and exports
every symbol include/sba_attn.h declares; the Python surface mirrors the reference's
(SURVEY.md §8b); the product path refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sba_attn.h")).read()
    return re.findall(r"SBA_API\s+[\w\s\*]+?\b(sba_\w+)\s*\(", src)


def test_header_symbols_are_exported_and_bound():
    from sba_gan_b200 import _abi
    names = _declared_symbols()
    assert len(names) >= 9
    assert set(names) == set(_abi.SYMBOLS), "binding table and header disagree"
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libsba_attn.so does not export {n}"
    assert _abi.load().sba_abi_version() == _abi.ABI_VERSION


def test_no_oracle_import_in_product_path():
    """The oracle is test infrastructure: nothing under sba_gan_b200/ (nor the measurement harness or the
    tools) may import it; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do."""
    pkg = os.path.join(ROOT, "sba_gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} references the oracle"
    for sub in ("harness", "tools"):
        for f in os.listdir(os.path.join(ROOT, sub)):
            if f.endswith(".py"):
                src = open(os.path.join(ROOT, sub, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{sub}/{f} imports the oracle"


def test_module_surface_matches_reference():
    from sba_gan_b200 import GlobalAttention as ga
    assert {"conv1x1", "func_attention", "GlobalAttentionGeneral"} <= set(dir(ga))
    m = ga.GlobalAttentionGeneral(32, 256)
    sd = m.state_dict()
    assert list(sd) == ["conv_context.weight"] and tuple(sd["conv_context.weight"].shape) == (32, 256, 1, 1)
    assert isinstance(m.conv_context, torch.nn.Conv2d) and m.conv_context.bias is None
    assert type(m.conv_context).__name__.find("Conv") != -1        # weights_init hook, miscc/utils.py:287
    assert m.mask is None
    mask = torch.zeros(2, 5, dtype=torch.bool)
    m.applyMask(mask)
    assert m.mask is mask
    conv = ga.conv1x1(256, 32)
    assert conv.kernel_size == (1, 1) and conv.bias is None
    # a reference-built state dict (same key/shape) loads
    m.load_state_dict({"conv_context.weight": torch.randn(32, 256, 1, 1)})


def test_cpu_tensors_fail_loudly():
    import sba_gan_b200 as pkg
    m = pkg.GlobalAttentionGeneral(32, 256)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 32, 4, 4), torch.zeros(1, 256, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.words_loss(torch.zeros(2, 256, 17, 17), torch.zeros(2, 256, 18), torch.arange(2), torch.tensor([5, 4]),
                       None, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.func_attention(torch.zeros(2, 256, 5), torch.zeros(2, 256, 17, 17), 4.0)


def test_install_registers_reference_import_names():
    import sys
    import sba_gan_b200 as pkg
    saved = sys.modules.get("GlobalAttention")
    try:
        pkg.install()
        import GlobalAttention
        assert GlobalAttention.GlobalAttentionGeneral is pkg.GlobalAttentionGeneral
        assert GlobalAttention.func_attention is pkg.func_attention
    finally:
        if saved is None:
            sys.modules.pop("GlobalAttention", None)
        else:
            sys.modules["GlobalAttention"] = saved
