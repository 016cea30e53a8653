"""The fused ResNet kernel (csrc/azb_resnet.cu) against a plain PyTorch fp32
evaluation of the same module.  Convolution operands are bf16 (fp32
accumulation), so the tolerance is the bf16 one, stated here: 3e-2 absolute on
probabilities; the folded-weight algebra itself is checked in float64 on the
CPU to 1e-9."""
import numpy as np
import pytest
import torch

from azb200 import nnet as aznet


def _model(seed=0, depth=4):
    torch.manual_seed(seed)
    m = aznet.ResNet((4, 6, 7), 7, 3, **dict(aznet.DEFAULT_NET_ARGS, depth=depth)).eval()
    with torch.no_grad():                      # non-trivial BN statistics / affine
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.3); mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.2)
    return m


def _obs(n, seed=1):
    rs = np.random.RandomState(seed)
    o = np.zeros((n, 4, 6, 7), np.float32)
    cells = rs.randint(0, 3, size=(n, 6, 7))
    o[:, 0] = cells == 1; o[:, 1] = cells == 2
    o[:, 2] = rs.randint(0, 2, size=(n, 1, 1)); o[:, 3] = rs.randint(0, 43, size=(n, 1, 1)) / 42.0
    return torch.from_numpy(o)


def test_folded_weights_reproduce_the_module_in_float64():
    """BN folding + affine head folding, emulated with the folded tensors in float64."""
    from azb200.fused_nn import fold
    m = _model()
    f = fold(m, 296)
    x = _obs(16).double()
    md = m.double()
    with torch.no_grad():
        lp, lv = md(x)
        want = torch.cat([lp.exp(), lv.exp()], 1)
        ch, depth = 32, f["depth"]
        w = f["wconv"].double()              # bf16-rounded operands: compare against the same rounding
        def conv(inp, layer, cin, kper):
            k = w[layer, :, :9 * kper].view(ch, 3, 3, kper)[..., :cin].permute(0, 3, 1, 2)
            return torch.nn.functional.conv2d(inp, k, padding=1)
        t = torch.relu(conv(x, 0, 4, 16) + f["cbias"][0].double().view(1, -1, 1, 1))
        for i in range(depth):
            a = torch.relu(t * f["bn_scale"][i].double().view(1, -1, 1, 1) + f["bn_shift"][i].double().view(1, -1, 1, 1))
            b = torch.relu(conv(a, 1 + 2 * i, ch, ch) + f["cbias"][1 + 2 * i].double().view(1, -1, 1, 1))
            t = t + conv(b, 2 + 2 * i, ch, ch)
        feat = t.permute(0, 2, 3, 1).reshape(len(x), 42, ch)
        logits = torch.einsum("npc,jpc->nj", feat, f["whead"].double()) + f["bhead"].double()
        got = torch.cat([torch.softmax(logits[:, :7], 1), torch.softmax(logits[:, 7:], 1)], 1)
    m.float()
    # only the bf16 rounding of the conv weights separates the two
    assert torch.allclose(got, want, atol=2e-2)
    # and with unrounded weights the algebra is exact
    f64 = fold(_model(), 296)
    assert f64["whead"].shape == (10, 42, 32) and f64["wconv"].shape == (9, 32, 296)


def test_tc_weight_layouts_match_the_mma_layouts():
    """fold_tc (UMMA operand chunks with the horizontal taps side by side, frame-row head matrix)
    is a pure re-layout of fold."""
    from azb200.fused_nn import fold, fold_tc
    m = _model()
    a, b = fold(m, 296), fold_tc(m)
    w, t = a["wconv"].float(), b["wconv"].float()
    assert t.shape == (9, 12, 96, 8)
    for l in range(1, 9):
        ref = w[l, :, :288].view(32, 3, 3, 4, 8)                 # [cout][dy][dx][cin/8][8]
        assert torch.equal(ref.permute(1, 3, 2, 0, 4).reshape(12, 96, 8), t[l])
    ref0 = w[0, :, :144].view(32, 3, 3, 16)[..., :8]             # [cout][dy][dx][cin]
    assert torch.equal(ref0.permute(1, 2, 0, 3).reshape(3, 96, 8), t[0, :3]) and bool((t[0, 3:] == 0).all())
    wh = a["whead16"].float()[:10, :42 * 32].view(10, 6, 7, 32)
    th = b["whead16"].float()[:, :56 * 32].view(10, 7, 8, 32)
    assert torch.equal(wh, th[:, :6, :7]) and bool((th[:, 6] == 0).all()) and bool((th[:, :, 7] == 0).all())


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["tc", "mma"])
@pytest.mark.parametrize("batch", [8, 100, 8192])
def test_fused_kernel_matches_pytorch_fp32(batch, kernel):
    from azb200.fused_nn import FusedResNetEvaluator
    dev = torch.device("cuda")
    m = _model().to(dev)
    obs = _obs(batch).to(dev)
    pol = torch.zeros(batch, 7, device=dev); val = torch.zeros(batch, 3, device=dev)
    ev = FusedResNetEvaluator(m, obs, pol, val, kernel=kernel)
    ev()
    torch.cuda.synchronize()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        lp, lv = m(obs)
    torch.backends.cudnn.allow_tf32 = old
    wp, wv = lp.exp(), lv.exp()
    assert torch.isfinite(pol).all() and torch.isfinite(val).all()
    assert torch.allclose(pol.sum(1), torch.ones(batch, device=dev), atol=1e-5)
    assert torch.allclose(val.sum(1), torch.ones(batch, device=dev), atol=1e-5)
    ep, evl = (pol - wp).abs().max().item(), (val - wv).abs().max().item()
    assert ep < 3e-2 and evl < 3e-2, (ep, evl)
    # typical error is far below the bound
    assert (pol - wp).abs().mean().item() < 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [0, 1, 4])
def test_tc_kernel_layer_by_layer(depth):
    """Every epilogue of the tcgen05 kernel (stem, conv1, conv2 of each block) against the fp32
    activations of the PyTorch module; tolerance = bf16 operand rounding, 3e-2 absolute."""
    import torch.nn.functional as F
    from azb200.fused_nn import FusedResNetEvaluator
    dev = torch.device("cuda")
    m = _model(depth=depth).to(dev)
    batch = 40
    obs = _obs(batch).to(dev)
    pol = torch.zeros(batch, 7, device=dev); val = torch.zeros(batch, 3, device=dev)
    ev = FusedResNetEvaluator(m, obs, pol, val, kernel="tc")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    want = []
    with torch.no_grad():
        x = F.relu(m.bn1(m.conv1(obs)))
        blocks = list(m.resnet)
        a = F.relu(blocks[0].bn1(x)) if blocks else x
        want.append(a)
        for i, blk in enumerate(blocks):
            b = F.relu(blk.bn2(blk.conv1(a)))
            want.append(b)
            x = x + blk.conv2(b)
            a = F.relu(blocks[i + 1].bn1(x)) if i + 1 < len(blocks) else x
            want.append(a)
    torch.backends.cudnn.allow_tf32 = old
    for l, w in enumerate(want):
        got = ev.debug_layer(l)
        torch.cuda.synchronize()
        err = (got - w.permute(0, 2, 3, 1)).abs().max().item()
        assert err < 3e-2, (l, err)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,keep", [(8192, 0.85), (1000, 0.5), (64, 0.1), (40, 0.0)])
def test_tc_compact_rows_equal_dense(batch, keep):
    """azb_nn_forward_tc_rows: the listed rows get bit-identical answers to the dense evaluation, the others are left
    untouched; the row count is read from device memory (wave shaping on the device)."""
    from azb200.fused_nn import FusedResNetEvaluator
    dev = torch.device("cuda")
    m = _model().to(dev)
    obs = _obs(batch).to(dev)
    pol = torch.zeros(batch, 7, device=dev); val = torch.zeros(batch, 3, device=dev)
    FusedResNetEvaluator(m, obs, pol, val, kernel="tc")()
    rs = np.random.RandomState(batch)
    sel = np.flatnonzero(rs.random_sample(batch) < keep).astype(np.int32)
    rs.shuffle(sel)
    rows = torch.zeros(batch, dtype=torch.int32, device=dev)
    rows[:len(sel)] = torch.from_numpy(sel).to(dev)
    count = torch.tensor([len(sel)], dtype=torch.int32, device=dev)
    pol2 = torch.full((batch, 7), -1.0, device=dev); val2 = torch.full((batch, 3), -1.0, device=dev)
    FusedResNetEvaluator(m, obs, pol2, val2, kernel="tc", rows=rows, count=count)()
    torch.cuda.synchronize()
    mask = torch.zeros(batch, dtype=torch.bool, device=dev)
    mask[torch.from_numpy(sel).long().to(dev)] = True
    assert torch.equal(pol2[mask], pol[mask]) and torch.equal(val2[mask], val[mask])
    assert bool((pol2[~mask] == -1).all()) and bool((val2[~mask] == -1).all())
