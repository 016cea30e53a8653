"""The generic tcgen05 leaf evaluator (csrc/azb_resnet_g.cu, azb200/nn_tc.py) against the reference network.

Oracle for this floating-point kernel: the PyTorch module of alphazero/NNetArchitecture.py:69-120 (mirror in
azb200.nnet, state_dict-compatible, checked equal to the reference's own module in test_host_logic.py) evaluated in
strict fp32 (no TF32), i.e. NNetWrapper.process (NNetWrapper.py:225-232).

Stated tolerances on probabilities (absolute):
  bf16x2  1e-5   the north star's bound -- the default precision of every product path (32 / 64 channels)
  fp16x2  1e-5   the same, default for the 128-channel network
  fp16    max(2 x the error of cuDNN's TF32 evaluation of the same boards, 1e-4): TF32-class (11-bit significand)
  bf16    3e-2   performance mode
CPU half: the folded operand layouts, decoded again, reproduce the module in float64."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from azb200 import nnet as aznet
from azb200 import nn_tc

GEOMS = {
    "connect4": dict(obs=(4, 6, 7), A=7, args=aznet.DEFAULT_NET_ARGS),
    "brandubh": dict(obs=(5, 7, 7), A=588, args=aznet.BRANDUBH_TRAIN_NET_ARGS),
    "brandubh32": dict(obs=(5, 7, 7), A=588, args=aznet.DEFAULT_NET_ARGS),
    "connect4_64": dict(obs=(4, 6, 7), A=7, args=aznet.BRANDUBH_TRAIN_NET_ARGS),
    # envs/connect4/train.py:44-49 (128 channels x 8 blocks): k_trunk_wide; the same trunk on a 7x7 board with 588 actions
    "connect4_train": dict(obs=(4, 6, 7), A=7, args=aznet.CONNECT4_TRAIN_NET_ARGS),
    "brandubh128": dict(obs=(5, 7, 7), A=588, args=aznet.CONNECT4_TRAIN_NET_ARGS),
}


def _model(geom, seed=0, depth=None, sharpen=1.0):
    g = GEOMS[geom]
    torch.manual_seed(seed)
    args = dict(g["args"])
    if depth is not None:
        args["depth"] = depth
    m = aznet.ResNet(g["obs"], g["A"], 3, **args).eval()
    with torch.no_grad():                      # non-trivial BN statistics / affine, optionally peaked outputs
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.3); mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.2)
        for fc in (m.pi_fc, m.v_fc):
            [l for l in fc if isinstance(l, torch.nn.Linear)][-1].weight.mul_(sharpen)
    return m


def _obs(geom, n, seed=1):
    c, h, w = GEOMS[geom]["obs"]
    rs = np.random.RandomState(seed)
    o = np.zeros((n, c, h, w), np.float32)
    cells = rs.randint(0, c, size=(n, h, w))
    for k in range(c - 2):
        o[:, k] = cells == k + 1
    o[:, c - 2] = rs.randint(0, 2, size=(n, 1, 1))
    o[:, c - 1] = (rs.randint(0, 43, size=(n, 1, 1)) / 42.0).astype(np.float32)
    return torch.from_numpy(o)


def _want(m, obs):
    """strict-fp32 evaluation on the device the module lives on"""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            lp, lv = m(obs)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return lp.exp(), lv.exp()


@pytest.mark.parametrize("geom", ["connect4", "brandubh", "connect4_train"])
@pytest.mark.parametrize("precision", ["bf16x2", "fp16x2", "fp16"])
def test_folded_operand_layouts_reproduce_the_module(geom, precision):
    """Decode wtrunk / whead exactly as the kernel addresses them and evaluate in float64."""
    m = _model(geom)
    f = nn_tc.fold_g(m, precision)
    lay = nn_tc.layout(f["channels"], precision)
    assert lay["parts"] == f["parts"] and lay["slab_bytes"] == f["wtrunk"].shape[1] * 2
    ch, depth, cin, H, W, A = (f[k] for k in ("channels", "depth", "in_channels", "board_h", "board_w", "action_size"))
    parts, dys, c8, nacc = lay["parts"], lay["dys"], ch // 8, 3 * ch
    wt = f["wtrunk"].double()

    def wide_stem_w():                                  # k_trunk_wide: slab j = dx, K chunks (dy=-1, dy=0, zero, dy=+1)
        sp = 4 * ch * 8
        w = torch.stack([sum(wt[j, p * sp:(p + 1) * sp] for p in range(parts)).view(4, ch, 8) for j in range(3)])
        assert bool((w[:, 2] == 0).all())
        return w[:, [0, 1, 3]][..., :cin].permute(2, 3, 1, 0)                             # [dx][dy][cout][cin] -> [cout][cin][dy][dx]

    def wide_layer_w(l):                                # slab (tap, kq): [4 K chunks][cout][8]
        sp = 4 * ch * 8
        s0 = 3 + (l - 1) * 36
        w = sum(wt[s0:s0 + 36, p * sp:(p + 1) * sp] for p in range(parts)).view(3, 3, c8 // 4, 4, ch, 8)
        return w.permute(4, 2, 3, 5, 0, 1).reshape(ch, ch, 3, 3)                          # [cout][cin][dy][dx]

    def stem_w():
        if dys == 0:
            return wide_stem_w()
        sp = 4 * nacc * 8
        w = sum(wt[0, p * sp:(p + 1) * sp] for p in range(parts)).view(4, 3, ch, 8)       # [dy][dx][cout][cin]
        assert bool((w[3] == 0).all())
        return w[:3, :, :, :cin].permute(2, 3, 0, 1)                                      # [cout][cin][dy][dx]

    def layer_w(l):
        if dys == 0:
            return wide_layer_w(l)
        sp = dys * c8 * nacc * 8
        per = 3 // dys
        rows = []
        for j in range(per):
            s = 1 + (l - 1) * per + j
            rows.append(sum(wt[s, p * sp:(p + 1) * sp] for p in range(parts)).view(dys, c8, 3, ch, 8))
        w = torch.cat(rows, 0)                                                            # [dy][cin/8][dx][cout][8]
        return w.permute(3, 1, 4, 0, 2).reshape(ch, ch, 3, 3)

    x = _obs(geom, 4 if ch == 128 else 8).double()
    b = lambda v: v.double().view(1, -1, 1, 1)
    t = torch.relu(F.conv2d(x, stem_w(), padding=1) + b(f["cbias"][0]))
    for i in range(depth):
        a = torch.relu(t * b(f["bn_scale"][i]) + b(f["bn_shift"][i]))
        bb = torch.relu(F.conv2d(a, layer_w(1 + 2 * i), padding=1) + b(f["cbias"][1 + 2 * i]))
        t = t + F.conv2d(bb, layer_w(2 + 2 * i), padding=1)
    nt, ntiles, kc = f["head_nt"], f["head_ntiles"], f["head_kc"]
    assert (nt, ntiles) == nn_tc.head_tiles(A + 3) and kc % lay["head_kgran"] == 0 and nt * ntiles >= A + 3
    wh = f["whead"].double().sum(0)                                                       # [n tile][kc][row][8]
    wh = wh.permute(0, 2, 1, 3).reshape(ntiles * nt, kc * 8)[:, :H * W * ch]              # [out][pos*ch + c]
    feat = t.permute(0, 2, 3, 1).reshape(len(x), -1)
    logits = feat @ wh.T + f["bhead"].double()
    got = torch.cat([torch.softmax(logits[:, :A], 1), torch.softmax(logits[:, A:A + 3], 1)], 1)
    with torch.no_grad():
        lp, lv = m.double()(x)
    want = torch.cat([lp.exp(), lv.exp()], 1)
    # only the rounding of the weights to 16 (bf16x2) / 11 (fp16) significant bits separates the two
    tol = 2e-4 if precision == "fp16" else 2e-6
    assert (got - want).abs().max().item() < tol
    assert bool((wh[A + 3:] == 0).all())


@pytest.mark.parametrize("geom,depth", [("connect4", 4), ("brandubh", 4), ("connect4", 0), ("connect4_64", 1)])
def test_pair_layout_is_the_single_cta_layout_split_by_rows(geom, depth):
    """AZB_NNG_PAIR: slab s of CTA rank r holds the N rows [r * 3 ch / 2, (r + 1) * 3 ch / 2) of every K chunk of slab s."""
    m = _model(geom, depth=depth)
    one, two = nn_tc.fold_g(m, "bf16x2"), nn_tc.fold_g(m, "bf16x2", pair=True)
    ch, parts = one["channels"], one["parts"]
    nacc, nb, c8 = 3 * ch, 3 * ch // 2, ch // 8
    dys = 3 if ch == 32 else 1
    a, b = one["wtrunk"], two["wtrunk"]
    assert a.shape == b.shape and a.dtype == b.dtype
    sp = 4 * nacc * 8                                                    # stem part
    for p in range(parts):
        full = a[0, p * sp:(p + 1) * sp].view(4, nacc, 8)
        for r in range(2):
            half = b[0].view(2, -1)[r, p * (sp // 2):(p + 1) * (sp // 2)].view(4, nb, 8)
            assert torch.equal(half, full[:, r * nb:(r + 1) * nb])
    for s in range(1, a.shape[0]):
        full = a[s].view(parts * dys * c8, nacc, 8)
        for r in range(2):
            assert torch.equal(b[s].view(2, parts * dys * c8, nb, 8)[r], full[:, r * nb:(r + 1) * nb])
    for k in ("whead", "bhead", "cbias", "bn_scale", "bn_shift"):
        assert torch.equal(one[k], two[k])


def _tol(precision, m, obs, want):
    if precision in ("bf16x2", "fp16x2"):
        return 1e-5
    if precision == "bf16":
        return 3e-2
    # fp16 = TF32-class: measured against what the reference's own default (cuDNN TF32) does on the same boards
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    with torch.no_grad():
        lp, lv = m(obs)
    torch.backends.cudnn.allow_tf32 = old
    err = max((lp.exp() - want[0]).abs().max().item(), (lv.exp() - want[1]).abs().max().item())
    return max(2 * err, 1e-4)


def _run(geom, batch, precision, sharpen=1.0, depth=None, pair=None, persist=None):
    dev = torch.device("cuda")
    m = _model(geom, depth=depth, sharpen=sharpen).to(dev)
    obs = _obs(geom, batch).to(dev)
    A = GEOMS[geom]["A"]
    pol = torch.zeros(batch, A, device=dev); val = torch.zeros(batch, 3, device=dev)
    ev = nn_tc.TensorCoreEvaluator(m, obs, pol, val, precision=precision, pair=pair, persist=persist)
    ev()
    torch.cuda.synchronize()
    return m, obs, pol, val, ev


@pytest.mark.gpu
@pytest.mark.parametrize("geom", ["connect4", "brandubh", "brandubh32", "connect4_64", "connect4_train", "brandubh128"])
@pytest.mark.parametrize("precision", ["bf16x2", "fp16x2", "fp16", "bf16"])
@pytest.mark.parametrize("batch", [6, 1000])
def test_evaluator_matches_the_fp32_module(geom, precision, batch):
    m, obs, pol, val, _ = _run(geom, batch, precision)
    want = _want(m, obs)
    tol = _tol(precision, m, obs, want)
    assert torch.isfinite(pol).all() and torch.isfinite(val).all()
    assert torch.allclose(pol.sum(1), torch.ones(batch, device=pol.device), atol=1e-5)
    assert torch.allclose(val.sum(1), torch.ones(batch, device=pol.device), atol=1e-5)
    ep, evl = (pol - want[0]).abs().max().item(), (val - want[1]).abs().max().item()
    assert ep < tol and evl < tol, (geom, precision, ep, evl, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("geom,batch", [("connect4", 8192), ("brandubh", 4096), ("connect4_train", 8192)])
def test_default_precision_at_baseline_sizes_and_peaked_outputs(geom, batch):
    """BASELINE batch sizes, heads sharpened x8 so that the probabilities are peaked like a trained network's:
    the default precision stays within 1e-5 of the fp32 module."""
    m, obs, pol, val, ev = _run(geom, batch, None, sharpen=8.0)
    assert ev.precision == ("fp16x2" if geom == "connect4_train" else "bf16x2")
    want = _want(m, obs)
    assert want[0].max().item() > 2.0 / GEOMS[geom]["A"]            # visibly non-uniform
    ep, evl = (pol - want[0]).abs().max().item(), (val - want[1]).abs().max().item()
    assert ep < 1e-5 and evl < 1e-5, (ep, evl)


@pytest.mark.gpu
def test_wide_network_precisions_against_the_exact_network():
    """Why the 128-channel network defaults to fp16x2: 17 layers of K = 1152 with x8-sharpened heads, every split
    precision against the module in float64 (the exact network) next to the fp32 module's own distance to it."""
    dev = torch.device("cuda")
    batch, errs = 2048, {}
    for prec in ("fp16x2", "bf16x2"):
        m, obs, pol, val, _ = _run("connect4_train", batch, prec, sharpen=8.0)
        want = _want(m, obs)
        with torch.no_grad():
            lp, lv = m.double()(obs.double())
        exact = (lp.exp(), lv.exp())
        errs[prec] = max((pol.double() - exact[0]).abs().max().item(), (val.double() - exact[1]).abs().max().item())
        errs["fp32 module"] = max((want[0].double() - exact[0]).abs().max().item(), (want[1].double() - exact[1]).abs().max().item())
    print("connect4_train x8-sharpened, max |dp| against float64:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["fp16x2"] < 1e-5, errs                 # measured 7.7e-6 (what is left is the accumulation itself)
    assert errs["bf16x2"] < 3e-5, errs                 # measured 1.07e-5: at the edge, hence not this network's default


@pytest.mark.gpu
@pytest.mark.parametrize("geom,depth", [("connect4", 0), ("connect4", 1), ("connect4", 4), ("brandubh", 0), ("brandubh", 1),
                                        ("brandubh", 4), ("connect4_train", 0), ("connect4_train", 1), ("connect4_train", 8),
                                        ("brandubh128", 2)])
def test_layer_by_layer(geom, depth):
    """Every epilogue (stem, conv1, conv2 of each block) against the fp32 activations of the module; reports all layers."""
    dev = torch.device("cuda")
    m, obs, pol, val, ev = _run(geom, 10, "bf16x2", depth=depth)
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
    errs = []
    for l, w in enumerate(want):
        got = ev.debug_layer(l)
        torch.cuda.synchronize()
        errs.append((got - w.permute(0, 2, 3, 1)).abs().max().item() / max(w.abs().max().item(), 1.0))
    assert max(errs) < 2e-5, errs


@pytest.mark.gpu
@pytest.mark.parametrize("geom,batch,keep", [("connect4", 8192, 0.85), ("connect4", 1000, 0.5), ("brandubh", 300, 0.7),
                                             ("connect4", 40, 0.0), ("connect4_train", 1000, 0.6)])
def test_compact_rows_equal_dense(geom, batch, keep):
    """rows / count: the listed rows get bit-identical answers to the dense evaluation, the others are left untouched;
    the row count is read from device memory."""
    dev = torch.device("cuda")
    m, obs, pol, val, _ = _run(geom, batch, "bf16x2")
    A = GEOMS[geom]["A"]
    rs = np.random.RandomState(batch)
    sel = np.flatnonzero(rs.random_sample(batch) < keep).astype(np.int32)
    rs.shuffle(sel)
    rows = torch.zeros(batch, dtype=torch.int32, device=dev)
    rows[:len(sel)] = torch.from_numpy(sel).to(dev)
    count = torch.tensor([len(sel)], dtype=torch.int32, device=dev)
    pol2 = torch.full((batch, A), -1.0, device=dev); val2 = torch.full((batch, 3), -1.0, device=dev)
    nn_tc.TensorCoreEvaluator(m, obs, pol2, val2, precision="bf16x2", rows=rows, count=count)()
    torch.cuda.synchronize()
    mask = torch.zeros(batch, dtype=torch.bool, device=dev)
    mask[torch.from_numpy(sel).long().to(dev)] = True
    assert torch.equal(pol2[mask], pol[mask]) and torch.equal(val2[mask], val[mask])
    assert bool((pol2[~mask] == -1).all()) and bool((val2[~mask] == -1).all())


@pytest.mark.gpu
@pytest.mark.parametrize("geom", ["connect4", "brandubh", "brandubh32", "connect4_64"])
@pytest.mark.parametrize("precision", ["bf16x2", "fp16x2", "fp16"])
@pytest.mark.parametrize("batch", [1, 2, 3, 6, 29, 1000, 4133])
def test_cta_pairs_equal_single_ctas_bit_for_bit(geom, precision, batch):
    """AZB_NNG_PAIR (two CTAs in lockstep, one M = 256 MMA stream, half of the weight rows staged per CTA) changes where
    the operands come from, not the arithmetic: policy and value are bit-identical to the single-CTA kernel -- for batch
    sizes that leave the peer CTA's tile half empty or empty, too."""
    m, obs, pol, val, ev = _run(geom, batch, precision, pair=True)
    assert ev.pair
    m2, obs2, pol2, val2, ev2 = _run(geom, batch, precision, pair=False)
    assert not ev2.pair
    assert torch.equal(pol, pol2) and torch.equal(val, val2)


@pytest.mark.parametrize("channels", [32, 64])
@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("persist", [0, 1])
def test_tile_plan_deals_every_tile_exactly_once(channels, pair, persist):
    """azb_nng_tile_plan (host only): the units' shares are consecutive, disjoint and cover all tiles (pair mode: pairs of
    tiles); a unit's rounds hold exactly its share (`big` rounds of `tiles` tiles, the others of tiles - 1, every round
    giving each of the kernel's MMA-issuing threads a tile) or, where that is impossible, rounds x tiles with fewer than
    `rounds` empty tiles; tiles per round fit the CTA's tensor memory; persistent plans never exceed one wave."""
    import ctypes as C
    from azb200 import _capi
    lib = _capi.load()
    cap, issuers = (7, 3) if channels == 32 else (2, 2)
    for sms in (148, 132, 3):
        resident = sms // 2 if pair else sms
        if resident < 1:
            continue
        for boards in [0, 1, 2, 3, 4, 5, 27, 28, 29, 147, 148, 296, 297, 1000, 2048, 3915, 4133, 5255, 5831, 6300, 6960, 8192,
                       16384, 65536]:
            out = (C.c_int32 * 6)()
            assert lib.azb_nng_tile_plan(channels, boards, sms, pair, persist, 0, out) == 0
            units = out[0]
            tiles = (boards + 1) // 2
            items = (tiles + 1) // 2 if pair else tiles
            assert (units > 0) == (boards > 0)
            if persist:
                assert units == min(items, resident)
            nxt = 0
            for u in range(units):
                assert lib.azb_nng_tile_plan(channels, boards, sms, pair, persist, u, out) == 0
                _, n, first, rounds, per, big = out
                assert n >= 1 and first == nxt
                assert 1 <= per <= cap and 1 <= big <= rounds
                if big < rounds:
                    assert big * per + (rounds - big) * (per - 1) == n and per - 1 >= issuers
                else:
                    assert rounds * per >= n and rounds * per - n < rounds
                    assert rounds * per == n or per - 1 < issuers
                if not persist:
                    assert rounds == 1
                nxt += n
            assert nxt == items
            assert lib.azb_nng_tile_plan(channels, boards, sms, pair, persist, units, out) == 0 and list(out)[1:] == [0, 0, 0, 0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("geom", ["connect4", "brandubh", "brandubh32", "connect4_64"])
@pytest.mark.parametrize("precision", ["bf16x2", "fp16"])
@pytest.mark.parametrize("pair", [True, False])
@pytest.mark.parametrize("batch", [3, 1000, 4133, 8192])
def test_persistent_units_equal_one_round_units_bit_for_bit(geom, precision, pair, batch):
    """AZB_NNG_PERSIST (one wave of CTAs running their share of the tiles in several rounds, the next round's observations
    staged by the last layer's epilogue) against AZB_NNG_ONE_ROUND (whole waves of CTAs): the same arithmetic per board,
    so policy and value are bit-identical -- at batch sizes of one round, of several rounds, and with empty tiles."""
    m, obs, pol, val, ev = _run(geom, batch, precision, pair=pair, persist=True)
    m2, obs2, pol2, val2, ev2 = _run(geom, batch, precision, pair=pair, persist=False)
    assert torch.equal(pol, pol2) and torch.equal(val, val2)
    want = _want(m, obs)
    tol = _tol(precision, m, obs, want)
    assert float((pol - want[0]).abs().max()) <= tol and float((val - want[1]).abs().max()) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("geom", ["connect4", "brandubh"])
def test_persistent_units_compact_rows(geom):
    """compact evaluation (rows / device-side count) through several rounds of persistent units"""
    dev = torch.device("cuda")
    batch = 8192
    m = _model(geom).to(dev)
    obs = _obs(geom, batch).to(dev)
    A = GEOMS[geom]["A"]
    rng = np.random.RandomState(5)
    sel = np.sort(rng.choice(batch, size=6960, replace=False)).astype(np.int32)
    rows = torch.zeros(batch, dtype=torch.int32, device=dev)
    rows[:len(sel)] = torch.from_numpy(sel).to(dev)
    count = torch.tensor([len(sel)], dtype=torch.int32, device=dev)
    outs = []
    for persist in (True, False):
        pol = torch.full((batch, A), -1.0, device=dev); val = torch.full((batch, 3), -1.0, device=dev)
        nn_tc.TensorCoreEvaluator(m, obs, pol, val, precision="bf16x2", rows=rows, count=count, persist=persist)()
        torch.cuda.synchronize()
        outs.append((pol, val))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    mask = torch.zeros(batch, dtype=torch.bool, device=dev)
    mask[torch.from_numpy(sel).long().to(dev)] = True
    assert bool((outs[0][0][~mask] == -1).all()) and bool((outs[0][0][mask] >= 0).all())


@pytest.mark.gpu
@pytest.mark.parametrize("geom", ["connect4", "brandubh"])
@pytest.mark.parametrize("depth", [0, 1])
@pytest.mark.parametrize("pair", [True, False])
def test_persistent_units_shallow_networks(geom, depth, pair):
    """depth 0 (the stem is the last layer and hands the tile on itself) and depth 1 through several rounds: equal to the
    one-round plan bit for bit and within 1e-5 of the module; layer activations of a board of the LAST round are the
    module's (the debug kernel dumps the layer for every round)."""
    batch = 8192
    m, obs, pol, val, ev = _run(geom, batch, "bf16x2", depth=depth, pair=pair, persist=True)
    m2, obs2, pol2, val2, ev2 = _run(geom, batch, "bf16x2", depth=depth, pair=pair, persist=False)
    assert torch.equal(pol, pol2) and torch.equal(val, val2)
    want = _want(m, obs)
    assert float((pol - want[0]).abs().max()) <= 1e-5 and float((val - want[1]).abs().max()) <= 1e-5
    got = ev.debug_layer(0)                                  # [batch, H, W, channels]: what the stem hands on
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        x = F.relu(m.bn1(m.conv1(obs)))
        blocks = list(m.resnet)
        a = F.relu(blocks[0].bn1(x)) if blocks else x
    torch.backends.cudnn.allow_tf32 = old
    a = a.permute(0, 2, 3, 1)
    scale = float(a.abs().max())
    assert float((got - a).abs().max()) <= 2e-5 * max(scale, 1.0)      # as test_layer_by_layer
