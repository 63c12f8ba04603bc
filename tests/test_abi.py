"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ducosy.h declares; the Python modules keep the reference's parameter tree; no compute is launched."""
import os
import re

import pytest
import torch

from oracle import ducosy_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ducosy.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ducosy_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ducosy_gan_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ducosy.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.ducosy_version() == 100


def test_no_gpu_means_loud_error_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ducosy_gan_b200 import _lib
    from ducosy_gan_b200.modules.model import Generator
    assert _lib.load().ducosy_check_device() != 0
    assert b"no CUDA device" in _lib.load().ducosy_last_error() or b"sm_" in _lib.load().ducosy_last_error()
    with torch.no_grad(), pytest.raises(RuntimeError):
        Generator(1, 1)(torch.zeros(1, 1, 128, 128))


@pytest.mark.parametrize("cin,nb,cbam", [(1, 9, True), (3, 9, True), (2, 3, False)])
def test_generator_state_dict_layout_matches_reference(cin, nb, cbam):
    from ducosy_gan_b200.modules.model import Generator, weights_init_normal
    G = Generator(cin, nb, cbam)
    shapes = orc.generator_param_shapes(cin, nb, cbam)      # pinned to the reference by tests/golden
    sd = G.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] and sd[k].dtype == torch.float32 for k in shapes)
    # checkpoints with the DataParallel 'module.' prefix are stripped by the callers (generate.py:38-43)
    G.load_state_dict(orc.make_state_dict(shapes, 3), strict=True)
    n_conv = sum(1 for m in G.modules() if "Conv" in type(m).__name__)
    assert n_conv == (6 + nb * (5 if cbam else 2))
    before = G.model[1].bias.clone()
    G.apply(weights_init_normal)
    assert torch.equal(G.model[1].bias, before)              # biases untouched by weights_init_normal
    assert abs(G.model[1].weight.std().item() - 0.02) < 0.01
    assert sum(p.numel() for p in Generator(3).parameters()) == 11446515


def test_discriminator_state_dict_layout_matches_reference():
    from ducosy_gan_b200.modules.model import Discriminator
    D = Discriminator(1)
    shapes = orc.discriminator_param_shapes(1)
    assert list(D.state_dict().keys()) == list(shapes.keys())
    assert sum(p.numel() for p in D.parameters()) == 2762689


def test_discriminator_cpu_input_is_loud():
    from ducosy_gan_b200.modules.model import Discriminator
    with pytest.raises(RuntimeError):
        Discriminator(1)(torch.zeros(1, 1, 256, 256))


def test_generator_cpu_input_is_loud():
    from ducosy_gan_b200.modules.model import Generator
    with pytest.raises(RuntimeError):
        Generator(1, 1)(torch.zeros(1, 1, 128, 128))
    with torch.no_grad(), pytest.raises(RuntimeError):
        Generator(1, 1)(torch.zeros(1, 1, 128, 128))


def test_workspace_and_shape_validation_without_gpu():
    import ctypes as C
    from ducosy_gan_b200 import _lib
    lib = _lib.load()
    cfg = _lib.GenConfig(1, 9, 1, _lib.F16)
    assert lib.ducosy_generator_num_params(C.byref(cfg)) == 75
    assert lib.ducosy_generator_num_launches(C.byref(cfg)) == 97      # 16 + 9 blocks x (6 + pool, spatial conv, channel MLP)
    cfg_x2 = _lib.GenConfig(1, 9, 1, _lib.F16X2)                          # split-operand arm: packed weights and workspace double
    assert lib.ducosy_generator_packed_bytes(C.byref(cfg_x2)) > 44_000_000
    assert lib.ducosy_generator_workspace_bytes(C.byref(cfg_x2), 1, 512, 512) > 380_000_000
    assert lib.ducosy_generator_packed_bytes(C.byref(cfg)) > 22_000_000
    assert lib.ducosy_generator_workspace_bytes(C.byref(cfg), 1, 512, 512) > 200_000_000
    assert lib.ducosy_generator_workspace_bytes(C.byref(cfg), 1, 500, 500) == 0      # unsupported shape
    assert lib.ducosy_generator_workspace_bytes(C.byref(cfg), 1, 128, 384) == 0
