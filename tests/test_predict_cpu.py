"""Host-side pieces of the predictor's fast path that need no GPU: the space-to-depth restatements of the ResNet stem
(deephisto_b200/examples/predict_full_patched.py FusedResNetForward) against conv1 + folded BatchNorm + ReLU of the torchvision
model the reference builds (models/patch_cls_simple/model.py:5-11), in float32 on the CPU."""

import pytest
import torch
import torch.nn.functional as F


@pytest.fixture(scope="module")
def model():
    from deephisto_b200.examples import predict_full_patched as pfp

    torch.manual_seed(0)
    m = pfp.get_model(5).eval()
    for mod in m.modules():                                   # non-trivial BatchNorm statistics, as a trained model has
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    return m


@pytest.mark.parametrize("stem,ps", [("s2d4", 224), ("s2d4", 64), ("s2d2", 224), ("s2d2", 62)])
def test_space_to_depth_stem_equals_conv1(model, stem, ps):
    from deephisto_b200.examples import predict_full_patched as pfp

    x = torch.rand(2, ps, ps, 3, generator=torch.Generator().manual_seed(ps))
    want = F.relu(pfp.fold_batchnorm(model).conv1(x.permute(0, 3, 1, 2)))                     # [2, 64, ps/2, ps/2]
    f = pfp.FusedResNetForward(model, dtype=torch.float32, stem=stem)
    s2d = f.space_to_depth(x)
    assert tuple(s2d.shape) == f.s2d_shape(2, ps) and s2d.is_contiguous(memory_format=torch.channels_last)
    if stem == "s2d4":
        y = F.relu(F.conv2d(s2d, f.stem_w, f.stem_b, padding=1))                              # depth-to-space: channel (P*2 + Q)*64 + o
        got = y.reshape(2, 2, 2, 64, ps // 4, ps // 4).permute(0, 3, 4, 1, 5, 2).reshape(2, 64, ps // 2, ps // 2)
        assert f.gather_layout == "S2D48"
    else:
        got = F.relu(F.conv2d(s2d, f.stem_w, f.stem_b))
        assert f.gather_layout == "S2D16"
        assert float(s2d[:, :, :2].abs().max()) == 0 and float(s2d[:, :, -1].abs().max()) == 0 and float(s2d[:, 6:8].abs().max()) == 0
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    # the rest of the restated forward: same blocks, same weights (BatchNorm folded)
    assert len(f.blocks) == 8 and f.fc_w.shape == (5, 512)
    assert next(model.parameters()).dtype == torch.float32


def test_fused_forward_rejects_other_models(model):
    from torchvision import models

    from deephisto_b200.examples import predict_full_patched as pfp

    with pytest.raises(TypeError):
        pfp.FusedResNetForward(models.resnet50(weights=None))
    with pytest.raises(TypeError):
        pfp.FusedResNetForward(torch.nn.Linear(3, 3))
    with pytest.raises(ValueError):
        pfp.FusedResNetForward(model, stem="s2d8")
