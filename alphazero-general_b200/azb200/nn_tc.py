"""TensorCoreEvaluator -- the reference ResNet (alphazero/NNetArchitecture.py:69-120, evaluated by
NNetWrapper.process, alphazero/NNetWrapper.py:225-232) on the tcgen05 tensor cores for every shipped
geometry (boards up to 7x7, 32 / 64 / 128 trunk channels, any action size): csrc/azb_resnet_g.cu behind
azb_nng_forward (include/azb200_nn.h).

precision (operands of the convolutions and of the head GEMM; accumulation is fp32 throughout):
  "bf16x2"  default -- every operand as hi + lo bf16 (16 significant bits; TF32, the reference's cuDNN
            default, has 11), three MMAs per K step: probabilities within 1e-5 of the fp32 module
  "fp16x2"  the same three products on hi + lo fp16 pairs (22 significant bits for |v| >= 0.03, saturation at 65504):
            default of the 128-channel network (default_precision)
  "fp16"    one pass, 11 significant bits (TF32's), activations saturate at 65504
  "bf16"    one pass, 8 significant bits
Host-side folding (BN into the convolutions, the affine heads into one matrix) is fused_nn._folded,
evaluated in float64; this module only lays the folded tensors out as UMMA operands."""
import ctypes as C

import torch

from . import _capi
from .fused_nn import _folded

PRECISIONS = {"bf16": 0, "fp16": 1, "bf16x2": 2, "fp16x2": 3}
DEFAULT_PRECISION = "bf16x2"
SPLIT = ("bf16x2", "fp16x2")


def default_precision(model=None):
    """The precision product paths use when none is given: the one that keeps probabilities within 1e-5 of the fp32
    module with margin -- bf16x2, and fp16x2 for the 128-channel network (measured: bf16x2 1.07e-5 against the exact
    network with x8-sharpened heads, tests/test_nn_tc.py)."""
    if model is not None and model.conv1.out_channels == 128:
        return "fp16x2"
    return DEFAULT_PRECISION


class _NNGNet(C.Structure):
    _fields_ = [("channels", C.c_int32), ("depth", C.c_int32), ("in_channels", C.c_int32), ("board_h", C.c_int32),
                ("board_w", C.c_int32), ("action_size", C.c_int32), ("precision", C.c_int32), ("max_boards", C.c_int32),
                ("head_nt", C.c_int32), ("head_ntiles", C.c_int32), ("head_kc", C.c_int32), ("flags", C.c_int32),
                ("wtrunk", C.c_void_p), ("cbias", C.c_void_p), ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p),
                ("whead", C.c_void_p), ("bhead", C.c_void_p), ("gact", C.c_void_p), ("logits", C.c_void_p),
                ("h_params", C.c_void_p)]


def supported(model):
    """Geometry azb_nng_forward covers."""
    ch = model.conv1.out_channels
    return (ch in (32, 64, 128) and 1 <= model.board_x <= 7 and 1 <= model.board_y <= 7 and model.channels <= 8
            and len(model.resnet) <= (8 if ch == 128 else 6))


def head_tiles(nout):
    """(N tile, number of N tiles) of the head GEMM for `nout` outputs."""
    ntiles = -(-nout // 208)
    nt = 16 * -(-nout // (16 * ntiles))
    return nt, ntiles


def _split(w64, precision):
    """float64 tensor -> list of operand parts in the kernel's element type."""
    if precision in ("fp16", "fp16x2"):
        hi = w64.clamp(-65504.0, 65504.0).to(torch.float16)
        return [hi] if precision == "fp16" else [hi, (w64 - hi.double()).to(torch.float16)]
    hi = w64.to(torch.bfloat16)
    if precision == "bf16":
        return [hi]
    return [hi, (w64 - hi.double()).to(torch.bfloat16)]


def layout(channels, precision):
    lib = _capi.load()
    out = (C.c_int32 * 8)()
    rc = lib.azb_nng_layout(channels, PRECISIONS[precision], out)
    if rc != 0:
        raise NotImplementedError(f"tcgen05 evaluator: {channels} channels / {precision} not supported")
    return dict(parts=out[0], dys=out[1], slab_bytes=out[2], boards_per_cta=out[3], head_kgran=out[4], max_depth=out[5],
                layer_slabs=out[6], stem_slabs=out[7])


NNG_PAIR = 1      # azb_nng_net.flags: CTA pairs (tcgen05 cta_group::2), include/azb200_nn.h
NNG_PERSIST, NNG_ONE_ROUND = 2, 4     # one wave of persistent CTAs / whole waves of one-round CTAs (neither: kernel default)


def pair_default(channels):
    """CTA pairs are the default for the 32 / 64-channel trunk kernel (AZB_NN_PAIR=0 switches them off)."""
    import os
    return channels in (32, 64) and os.environ.get("AZB_NN_PAIR", "1") != "0"


@torch.no_grad()
def fold_g(model, precision=DEFAULT_PRECISION, lay=None, pair=False):
    """-> dict of CPU tensors in the layouts of azb_nng_net (include/azb200_nn.h).  pair: the trunk weights in the
    AZB_NNG_PAIR layout (every slab as two halves, CTA rank r holding the N rows [r * 3 ch / 2, (r + 1) * 3 ch / 2) of
    every K chunk)."""
    f = _folded(model)
    ch, depth, cin, H, W = f["channels"], f["depth"], f["in_channels"], f["board_h"], f["board_w"]
    if lay is None:       # the library's constants, restated (tests compare them with azb_nng_layout)
        parts = 2 if precision in SPLIT else 1
        if ch == 128:
            lay = dict(parts=parts, dys=0, slab_bytes=parts * 4 * ch * 16, head_kgran=4, layer_slabs=36, stem_slabs=3)
        else:
            dys = 3 if ch == 32 else 1
            lay = dict(parts=parts, dys=dys, slab_bytes=parts * dys * (ch // 8) * 3 * ch * 16, head_kgran=4,
                       layer_slabs=3 // dys, stem_slabs=1)
    parts, dys, c8, nacc = lay["parts"], lay["dys"], ch // 8, 3 * ch
    slab_elems = lay["slab_bytes"] // 2
    L = 1 + 2 * depth
    slabs_per_layer = lay["layer_slabs"]
    nslabs = lay["stem_slabs"] + (L - 1) * slabs_per_layer
    edt = torch.float16 if precision in ("fp16", "fp16x2") else torch.bfloat16
    wtrunk = torch.zeros(nslabs, slab_elems, dtype=edt)
    if dys == 0:
        # k_trunk_wide: every tap is its own MMA.  Stem slab j (dx = j - 1): [part][4 K chunks: dy = -1, 0, zero, +1][cout][8 cin];
        # trunk slab (tap = 3 (dy+1) + (dx+1), kq): [part][4 K chunks = cin/8 in 4 kq .. 4 kq + 3][cout][8 cin]
        slab_part = 4 * ch * 8
        w0 = torch.zeros(3, 4, ch, 8, dtype=torch.float64)                                 # [dx][chunk][cout][cin]
        ws = f["convs"][0].view(ch, 3, 3, cin).permute(2, 1, 0, 3)                         # [dx][dy][cout][cin]
        w0[:, 0, :, :cin], w0[:, 1, :, :cin], w0[:, 3, :, :cin] = ws[:, 0], ws[:, 1], ws[:, 2]
        for j in range(3):
            for p, t in enumerate(_split(w0[j], precision)):
                wtrunk[j, p * slab_part:(p + 1) * slab_part] = t.reshape(-1)
        for l in range(1, L):
            w = f["convs"][l].view(ch, 9, c8 // 4, 4, 8).permute(1, 2, 3, 0, 4)             # [tap][kq][chunk][cout][8]
            w = w.reshape(slabs_per_layer, slab_part)
            s0 = 3 + (l - 1) * slabs_per_layer
            for p, t in enumerate(_split(w, precision)):
                wtrunk[s0:s0 + slabs_per_layer, p * slab_part:(p + 1) * slab_part] = t
    else:
        # stem slab: [part][4 K chunks: dy = -1, 0, +1, zero][dx*ch + cout][8 cin]
        w0 = torch.zeros(4, nacc, 8, dtype=torch.float64)
        w0[:3, :, :cin] = f["convs"][0].view(ch, 3, 3, cin).permute(1, 2, 0, 3).reshape(3, nacc, cin)
        stem_part = 4 * nacc * 8
        for p, t in enumerate(_split(w0, precision)):
            wtrunk[0, p * stem_part:(p + 1) * stem_part] = t.reshape(-1)
        # trunk slabs: [part][dy in slab][cin/8][dx*ch + cout][8 cin]
        slab_part = dys * c8 * nacc * 8
        for l in range(1, L):
            w = f["convs"][l].view(ch, 3, 3, c8, 8).permute(1, 3, 2, 0, 4).reshape(3, c8, nacc, 8)    # [dy][cin/8][dx*ch+cout][8]
            for j in range(slabs_per_layer):
                s = 1 + (l - 1) * slabs_per_layer + j
                for p, t in enumerate(_split(w[j * dys:(j + 1) * dys], precision)):
                    wtrunk[s, p * slab_part:(p + 1) * slab_part] = t.reshape(-1)
    if pair:
        assert dys != 0, "CTA pairs: 32 / 64 channels"
        nb = nacc // 2
        wp = torch.zeros(nslabs, 2, slab_elems // 2, dtype=edt)
        sp = 4 * nacc * 8
        for p in range(parts):                                             # stem: [part][4 K chunks][rows][8]
            src = wtrunk[0, p * sp:(p + 1) * sp].view(4, nacc, 8)
            for r in range(2):
                wp[0, r, p * (sp // 2):(p + 1) * (sp // 2)] = src[:, r * nb:(r + 1) * nb].reshape(-1)
        if nslabs > 1:
            body = wtrunk[1:].view(nslabs - 1, parts * dys * c8, nacc, 8)   # [slab][part x dy x K chunk][rows][8]
            for r in range(2):
                wp[1:, r] = body[:, :, r * nb:(r + 1) * nb].reshape(nslabs - 1, slab_elems // 2)
        wtrunk = wp.view(nslabs, slab_elems)
    # heads: [part][n tile][K chunk = pos*c8 + ch/8][row in tile][8]
    whead, bias = f["whead"], f["bhead"]                                   # [nout, pos, ch], [nout]
    nout = whead.shape[0]
    nt, ntiles = head_tiles(nout)
    kgran = lay["head_kgran"]
    kc = -(-(H * W * c8) // kgran) * kgran
    wh = torch.zeros(ntiles * nt, kc, 8, dtype=torch.float64)
    wh[:nout, :H * W * c8] = whead.reshape(nout, H * W * c8, 8)
    wh = wh.view(ntiles, nt, kc, 8).permute(0, 2, 1, 3).contiguous()       # [n tile][kc][row][8]
    whead_t = torch.stack(_split(wh, precision))                           # [part][n tile][kc][row][8]
    bhead = torch.zeros(ntiles * nt, dtype=torch.float64)
    bhead[:nout] = bias
    return dict(wtrunk=wtrunk, cbias=f["cbias"].float(), bn_scale=f["bn_scale"].float(), bn_shift=f["bn_shift"].float(),
                whead=whead_t, bhead=bhead.float(), head_nt=nt, head_ntiles=ntiles, head_kc=kc, parts=parts,
                channels=ch, depth=depth, in_channels=cin, board_h=H, board_w=W, action_size=f["action_size"])


class TensorCoreEvaluator:
    """Same call surface as azb200.nnet.LeafEvaluator: evaluator(stream) enqueues one evaluation of ``obs`` into
    ``policy`` / ``value`` (engine-owned device rows).  rows / count: compact evaluation of rows[0 .. count) (device
    int32; count a tensor or a callable returning the device address of the counter)."""

    def __init__(self, model, obs, policy, value, precision=None, rows=None, count=None, max_batch=None, pair=None,
                 persist=None):
        precision = precision or default_precision(model)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        if not supported(model):
            raise NotImplementedError("tcgen05 evaluator: boards up to 7x7, 32 / 64 / 128 channels, <= 8 planes, depth <= 6 (8 at 128)")
        self.precision, self.kernel = precision, "tcg"
        self.lib = _capi.load()
        dev = obs.device
        ch = model.conv1.out_channels
        self.layout = layout(ch, precision)
        self.pair = pair_default(ch) if pair is None else bool(pair) and ch in (32, 64)
        f = fold_g(model, precision, self.layout, pair=self.pair)
        self.t = {k: v.to(dev).contiguous() for k, v in f.items() if torch.is_tensor(v)}
        # host copy of the per-channel parameters: a kernel argument of the 32 / 64-channel trunk (azb_nng_net.h_params)
        self.h_params = torch.cat([f["cbias"].reshape(-1), f["bn_scale"].reshape(-1), f["bn_shift"].reshape(-1)]).float().contiguous()
        assert obs.is_contiguous() and policy.is_contiguous() and value.is_contiguous()
        assert obs.dtype == policy.dtype == value.dtype == torch.float32
        self.obs, self.policy, self.value = obs, policy, value
        self.batch = obs.shape[0]
        self.max_batch = int(max_batch or self.batch)
        mt = -(-self.max_batch // 128)
        nout_pad = f["head_nt"] * f["head_ntiles"]
        # scratch: trunk output in the head GEMM's operand layout (zero: unused rows of the last M tile must be finite)
        self.gact = torch.zeros(f["parts"] * mt * f["head_kc"] * 2048, dtype=torch.uint8, device=dev)
        fused_softmax = f["head_ntiles"] == 1 and f["head_nt"] == 16
        self.logits = None if fused_softmax else torch.zeros(self.max_batch, nout_pad, device=dev)
        self.net = _NNGNet(f["channels"], f["depth"], f["in_channels"], f["board_h"], f["board_w"], f["action_size"],
                           PRECISIONS[precision], self.max_batch, f["head_nt"], f["head_ntiles"], f["head_kc"],
                           (NNG_PAIR if self.pair else 0) | (0 if persist is None else NNG_PERSIST if persist else NNG_ONE_ROUND),
                           *(self.t[k].data_ptr() for k in ("wtrunk", "cbias", "bn_scale", "bn_shift", "whead", "bhead")),
                           self.gact.data_ptr(), self.logits.data_ptr() if self.logits is not None else None,
                           self.h_params.data_ptr())
        self.kernels_per_call = 2 if fused_softmax else 3       # trunk + head (+ softmax over the logits)
        self.rows, self.count = rows, count
        if rows is not None:
            assert count is not None and rows.dtype == torch.int32
        self.stream = torch.cuda.Stream(device=dev)

    def __call__(self, stream=None):
        stream = stream or torch.cuda.current_stream()
        if self.rows is not None:
            cnt = self.count() if callable(self.count) else self.count.data_ptr()
            rc = self.lib.azb_nng_forward(C.byref(self.net), self.obs.data_ptr(), self.policy.data_ptr(), self.value.data_ptr(),
                                          self.max_batch, self.rows.data_ptr(), cnt, C.c_void_p(stream.cuda_stream))
        else:
            rc = self.lib.azb_nng_forward(C.byref(self.net), self.obs.data_ptr(), self.policy.data_ptr(), self.value.data_ptr(),
                                          self.batch, None, None, C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"azb_nng_forward failed with status {rc}")

    def debug_layer(self, layer):
        """Evaluate and return the activation the epilogue of `layer` hands on, as [batch, H, W, channels]."""
        dump = torch.zeros(self.batch, 8, 8, self.net.channels, device=self.obs.device)
        rc = self.lib.azb_nng_forward_debug(C.byref(self.net), self.obs.data_ptr(), self.policy.data_ptr(),
                                            self.value.data_ptr(), self.batch,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream), dump.data_ptr(), layer)
        if rc != 0:
            raise RuntimeError(f"azb_nng_forward_debug failed with status {rc}")
        return dump[:, :self.net.board_h, :self.net.board_w]


def make_evaluator(model, obs, policy, value, precision=None, kernel=None, rows=None, count=None, max_batch=None,
                   use_graph=True, channels_last=False):
    """The leaf evaluator for `model` at `precision`:
      "bf16x2" (default) / "fp16" / "bf16"  hand-written tcgen05 kernels (TensorCoreEvaluator); kernel = "tc-r1" / "mma"
                                          selects the round-1 bf16-only kernels of fused_nn (6x7 boards, 32 channels)
      "fp32" / "tf32" / "cudnn-bf16"         PyTorch / cuDNN (azb200.nnet.LeafEvaluator; strict fp32 = the parity oracle)
    A geometry the hand-written kernels do not cover falls back to cuDNN TF32 -- the reference's own arithmetic -- never
    to a narrower type."""
    precision = precision or default_precision(model)
    if kernel in ("tc-r1", "mma"):
        from .fused_nn import FusedResNetEvaluator
        if precision != "bf16":
            raise ValueError("the round-1 kernels compute in bf16 only")
        return FusedResNetEvaluator(model, obs, policy, value, kernel="tc" if kernel == "tc-r1" else "mma", rows=rows,
                                    count=count, max_batch=max_batch)
    if precision in PRECISIONS and supported(model):
        return TensorCoreEvaluator(model, obs, policy, value, precision=precision, rows=rows, count=count, max_batch=max_batch)
    if rows is not None:
        raise NotImplementedError("compact evaluation needs the tcgen05 evaluator")
    from .nnet import LeafEvaluator
    lib_prec = {"cudnn-bf16": "bf16"}.get(precision, precision if precision in ("fp32", "tf32") else "tf32")
    return LeafEvaluator(model, obs, policy, value, precision=lib_prec, use_graph=use_graph, channels_last=channels_last)
