"""FusedResNetEvaluator -- the reference ResNet for Connect4-sized boards as one
CUDA kernel launch (csrc/azb_resnet.cu, C ABI in include/azb200_nn.h).

Host-side weight preparation (exact algebra, evaluated in float64):
  stem     conv -> BN -> ReLU          => conv' = conv * s, bias' = beta - mean * s,  s = gamma / sqrt(var + eps)
  block    BN1 -> ReLU -> conv1 -> BN2 -> ReLU -> conv2 (+x)
           BN1 stays elementwise (scale, shift); BN2 is folded into conv1
  heads    conv1x1 -> BN -> flatten -> Linear..Linear with Identity activations
           (NNetArchitecture.py:86-102) is affine in the trunk output: its matrix
           is obtained by pushing the canonical basis through the modules.
Convolution operands are rounded to bf16 (fp32 accumulation); everything else
is fp32.  This is the bf16 performance mode of the leaf evaluation; parity runs
use azb200.nnet.LeafEvaluator(precision="fp32")."""
import ctypes as C

import torch

from . import _capi


class _NNWeights(C.Structure):
    _fields_ = [("channels", C.c_int32), ("depth", C.c_int32), ("in_channels", C.c_int32), ("board_h", C.c_int32),
                ("board_w", C.c_int32), ("action_size", C.c_int32), ("wconv", C.c_void_p), ("cbias", C.c_void_p),
                ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p), ("whead", C.c_void_p), ("bhead", C.c_void_p)]


def supported(model):
    ch = model.conv1.out_channels
    return (ch == 32 and (model.board_x, model.board_y) == (6, 7) and model.action_size == 7
            and model.channels <= 16)


def _bn_affine(bn):
    s = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
    return s, bn.bias.double() - bn.running_mean.double() * s


@torch.no_grad()
def _folded(model):
    """Folded network in float64, layout-free: conv weights [L][cout][tap][cin], biases, BN1
    scale / shift per block, the affine head map [A+3][pos][ch] and its bias."""
    # a private copy: nn.Module.cpu() / .double() replace the parameters' storage in place, which would leave every
    # pointer captured from the caller's model (optimizer state aside: a CUDA-graphed training step) dangling
    import copy
    m = copy.deepcopy(model).eval().cpu()
    ch, depth, cin = m.conv1.out_channels, len(m.resnet), m.channels
    L = 1 + 2 * depth
    convs = []
    cbias = torch.zeros(L, ch, dtype=torch.float64)
    s, t = _bn_affine(m.bn1)
    w = m.conv1.weight.double() * s[:, None, None, None]                 # [co, ci, ky, kx]
    convs.append(w.permute(0, 2, 3, 1).reshape(ch, 9, cin))
    cbias[0] = t
    bn_scale = torch.zeros(max(depth, 1), ch, dtype=torch.float64)
    bn_shift = torch.zeros(max(depth, 1), ch, dtype=torch.float64)
    for i, blk in enumerate(m.resnet):
        bn_scale[i], bn_shift[i] = _bn_affine(blk.bn1)
        s2, t2 = _bn_affine(blk.bn2)
        w1 = blk.conv1.weight.double() * s2[:, None, None, None]
        convs.append(w1.permute(0, 2, 3, 1).reshape(ch, 9, ch))
        cbias[1 + 2 * i] = t2
        convs.append(blk.conv2.weight.double().permute(0, 2, 3, 1).reshape(ch, 9, ch))
    # heads as one affine map of the trunk output [ch, H, W]
    H, W = m.board_x, m.board_y
    md = m.double()

    def heads(x):
        v = md.v_fc(torch.flatten(md.v_bn(md.v_conv(x)), 1))
        p = md.pi_fc(torch.flatten(md.pi_bn(md.pi_conv(x)), 1))
        return torch.cat([p, v], dim=1)                                   # [n, A + 3]
    zero = torch.zeros(1, ch, H, W, dtype=torch.float64)
    bias = heads(zero)[0]
    basis = torch.eye(ch * H * W, dtype=torch.float64).view(ch * H * W, ch, H, W)
    mat = (heads(basis) - bias[None]).T.contiguous()                       # [A+3, ch*H*W], feature = c*HW + pos
    whead = mat.view(-1, ch, H * W).permute(0, 2, 1).contiguous()          # [A+3, pos, ch]
    return dict(convs=convs, cbias=cbias, bn_scale=bn_scale, bn_shift=bn_shift, whead=whead, bhead=bias,
                channels=ch, depth=depth, in_channels=cin, board_h=H, board_w=W, action_size=m.action_size)


def fold(model, row_stride, head_stride=None):
    """-> dict of CPU tensors in the layouts azb200_nn.h documents for azb_nn_forward (mma.sync kernel)."""
    f = _folded(model)
    ch, depth, cin, H, W = f["channels"], f["depth"], f["in_channels"], f["board_h"], f["board_w"]
    L = 1 + 2 * depth
    wconv = torch.zeros(L, ch, row_stride, dtype=torch.float64)
    wconv[0, :, :9 * 16].view(ch, 9, 16)[:, :, :cin] = f["convs"][0]
    for l in range(1, L):
        wconv[l, :, :9 * ch] = f["convs"][l].reshape(ch, 9 * ch)
    whead, bias = f["whead"], f["bhead"]
    nout = whead.shape[0]
    hs = head_stride or (H * W * ch + 8)
    whead16 = torch.zeros(16, hs, dtype=torch.float64)                     # padded to two n-tiles, k = pos*ch + c
    whead16[:nout, :H * W * ch] = whead.reshape(nout, -1)
    bhead16 = torch.zeros(16, dtype=torch.float64)
    bhead16[:nout] = bias
    return dict(wconv=wconv.to(torch.bfloat16), cbias=f["cbias"].float(), bn_scale=f["bn_scale"].float(),
                bn_shift=f["bn_shift"].float(), whead=whead.float(), bhead=bias.float(),
                whead16=whead16.to(torch.bfloat16), bhead16=bhead16.float(),
                channels=ch, depth=depth, in_channels=cin, board_h=H, board_w=W, action_size=f["action_size"])


def fold_tc(model, layer_bytes=18432, head_stride=1800, frame_rows=56):
    """-> dict of CPU tensors in the layouts of azb_nn_forward_tc (tcgen05 kernel, azb_resnet_tc.cu):
    conv weights as the K-major no-swizzle UMMA B operand with the three horizontal taps side by
    side in N: [layer][16-byte K chunk][dx*32 + cout][8 cin], K chunk = (dy+1)*4 + cin/8 (trunk) or
    dy+1 with cin < 8 (stem); the head matrix [A+3][k] over frame rows, k = (y*8 + x)*channels + ch
    (the frame's padding rows carry zero weights)."""
    f = _folded(model)
    ch, depth, cin, H, W = f["channels"], f["depth"], f["in_channels"], f["board_h"], f["board_w"]
    assert cin <= 8 and ch == 32 and frame_rows == (H + 1) * (W + 1)
    L = 1 + 2 * depth
    kch = layer_bytes // (3 * ch * 16)
    wconv = torch.zeros(L, kch, 3 * ch, 8, dtype=torch.float64)
    w0 = f["convs"][0].view(ch, 3, 3, cin)                                  # [cout][dy][dx][cin]
    wconv[0, :3, :, :cin] = w0.permute(1, 2, 0, 3).reshape(3, 3 * ch, cin)  # [dy][dx*32+cout][cin]
    for l in range(1, L):
        w = f["convs"][l].view(ch, 3, 3, ch // 8, 8)                        # [cout][dy][dx][cin/8][8]
        wconv[l] = w.permute(1, 3, 2, 0, 4).reshape(kch, 3 * ch, 8)         # [dy][cin/8][dx][cout][8]
    whead, bias = f["whead"], f["bhead"]
    nout = whead.shape[0]
    fr = torch.zeros(nout, H + 1, W + 1, ch, dtype=torch.float64)
    fr[:, :H, :W] = whead.view(nout, H, W, ch)
    wh = torch.zeros(nout, head_stride, dtype=torch.float64)
    wh[:, :frame_rows * ch] = fr.reshape(nout, -1)
    bhead16 = torch.zeros(16, dtype=torch.float64)
    bhead16[:nout] = bias
    return dict(wconv=wconv.to(torch.bfloat16), cbias=f["cbias"].float(), bn_scale=f["bn_scale"].float(),
                bn_shift=f["bn_shift"].float(), whead16=wh.to(torch.bfloat16), bhead16=bhead16.float(),
                channels=ch, depth=depth, in_channels=cin, board_h=H, board_w=W, action_size=f["action_size"])


def supported_tc(model):
    """Geometry the tcgen05 kernel covers (azb_resnet_tc.cu): the mma.sync one, with <= 8 input planes."""
    return supported(model) and model.channels <= 8


def boards_per_cta(model, kernel=None):
    """Boards one CTA of the chosen kernel evaluates (launch-wave arithmetic for batch splits)."""
    lib = _capi.load()
    kernel = kernel or ("tc" if supported_tc(model) else "mma")
    return lib.azb_nn_tc_boards_per_cta() if kernel == "tc" else lib.azb_nn_boards_per_cta()


class FusedResNetEvaluator:
    """Same call surface as azb200.nnet.LeafEvaluator: evaluator(stream) enqueues one
    evaluation of ``obs`` into ``policy`` / ``value`` (engine-owned device rows).

    kernel = "tc"  : tcgen05 / TMEM kernel (azb_nn_forward_tc), the default where it applies
             "mma" : the mma.sync kernel (azb_nn_forward)"""

    precision = "bf16"

    def __init__(self, model, obs, policy, value, kernel=None, rows=None, count=None, max_batch=None):
        if not supported(model):
            raise NotImplementedError("fused evaluator: 6x7 boards, 32 channels, 7 actions only")
        kernel = kernel or ("tc" if supported_tc(model) else "mma")
        if kernel == "tc" and not supported_tc(model):
            raise NotImplementedError("tcgen05 evaluator: at most 8 observation planes")
        self.kernel = kernel
        self.lib = _capi.load()
        dev = obs.device
        model_dev = next(model.parameters()).device
        if kernel == "tc":
            f = fold_tc(model, self.lib.azb_nn_tc_layer_bytes(), self.lib.azb_nn_tc_head_row_stride(),
                        self.lib.azb_nn_tc_frame_rows_per_board())
            self._fwd = self.lib.azb_nn_forward_tc
        else:
            f = fold(model, self.lib.azb_nn_weight_row_stride(), self.lib.azb_nn_head_row_stride())
            self._fwd = self.lib.azb_nn_forward
        model.to(model_dev)
        self.t = {k: v.to(dev).contiguous() for k, v in f.items() if torch.is_tensor(v)}
        self.w = _NNWeights(f["channels"], f["depth"], f["in_channels"], f["board_h"], f["board_w"], f["action_size"],
                            *(self.t[k].data_ptr() for k in ("wconv", "cbias", "bn_scale", "bn_shift", "whead16", "bhead16")))
        assert obs.is_contiguous() and policy.is_contiguous() and value.is_contiguous()
        self.obs, self.policy, self.value = obs, policy, value
        self.batch = obs.shape[0]
        self.stream = torch.cuda.Stream(device=dev)
        # compact mode (tcgen05 kernel): evaluate only rows[0 .. count) -- device int32 tensors, e.g. the engine's
        # list of non-terminal leaves (SelfPlayEngine.nn_rows / nn_count); the other rows keep their old answers
        self.rows, self.count = rows, count
        if rows is not None:
            # count: int32 device tensor, or a callable returning the device address of the count (the engine alternates
            # between two counters, SelfPlayEngine.nn_count_ptr)
            assert self.kernel == "tc" and count is not None and rows.dtype == torch.int32
            self.max_batch = int(max_batch or self.batch)

    def __call__(self, stream=None):
        stream = stream or torch.cuda.current_stream()
        if self.rows is not None:
            rc = self.lib.azb_nn_forward_tc_rows(C.byref(self.w), self.obs.data_ptr(), self.policy.data_ptr(),
                                                 self.value.data_ptr(), self.rows.data_ptr(),
                                                 self.count() if callable(self.count) else self.count.data_ptr(),
                                                 self.max_batch, C.c_void_p(stream.cuda_stream))
        else:
            rc = self._fwd(C.byref(self.w), self.obs.data_ptr(), self.policy.data_ptr(), self.value.data_ptr(),
                           self.batch, C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"azb_nn_forward ({self.kernel}) failed with status {rc}")

    def debug_layer(self, layer):
        """tcgen05 kernel only: evaluate and return the fp32 activation the epilogue of `layer`
        produced, as [batch, 6, 7, 32] (padding rows dropped) -- for layer-by-layer checks."""
        assert self.kernel == "tc"
        nb, fb = self.lib.azb_nn_tc_boards_per_cta(), self.lib.azb_nn_tc_frame_rows_per_board()
        ctas = -(-self.batch // nb)
        dump = torch.zeros(ctas * nb + nb, fb // 8, 8, 32, device=self.obs.device)
        rc = self.lib.azb_nn_forward_tc_debug(C.byref(self.w), self.obs.data_ptr(), self.policy.data_ptr(),
                                              self.value.data_ptr(), self.batch,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream), dump.data_ptr(), layer)
        if rc != 0:
            raise RuntimeError(f"azb_nn_forward_tc_debug failed with status {rc}")
        return dump[:self.batch, :6, :7]
