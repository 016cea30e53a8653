"""Leaf evaluation: the reference's ResNet run through PyTorch/cuDNN.

The north star keeps the leaf evaluation on the reference's own network
(alphazero/NNetArchitecture.py:69-120) and library kernels; what this module
adds is the plumbing that lets it overlap with the tree kernels:
  * ``ResNet``       -- same architecture and state_dict layout as the
                        reference module, so its checkpoints load unchanged;
  * ``NNetWrapper``  -- the ``process(batch) -> (pi, v)`` surface of
                        alphazero/NNetWrapper.py:225-232 (probabilities, not logs);
  * ``LeafEvaluator``-- a CUDA-graph capture of ``process`` that reads the
                        engine's observation rows and writes the engine's
                        policy / value rows in place, on a caller-chosen stream.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# hyper-parameters the architecture reads from args (Coach.py:103-116)
DEFAULT_NET_ARGS = dict(num_channels=32, depth=4, value_head_channels=16, policy_head_channels=16,
                        value_dense_layers=[512, 64], policy_dense_layers=[512, 256])
# alphazero/envs/connect4/train.py:44-49 and envs/hnefatafl/train_brandubh.py:50-55
CONNECT4_TRAIN_NET_ARGS = dict(num_channels=128, depth=8, value_head_channels=32, policy_head_channels=32,
                               value_dense_layers=[1024, 256], policy_dense_layers=[1024])
BRANDUBH_TRAIN_NET_ARGS = dict(num_channels=64, depth=4, value_head_channels=16, policy_head_channels=16,
                               value_dense_layers=[1024, 128], policy_dense_layers=[1024])


def _conv3x3(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, bias=False)


def _conv1x1(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0, bias=False)


def _head(sizes):
    # Linear layers separated by Identity activations (indices 0, 2, 4, ... hold the Linears)
    layers = []
    for i in range(len(sizes) - 1):
        layers += [nn.Linear(sizes[i], sizes[i + 1]), nn.Identity()]
    return nn.Sequential(*layers)


class _PreActBlock(nn.Module):
    """Pre-activation residual block (NNetArchitecture.py:36-66, no downsample)."""

    def __init__(self, ch):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(ch)
        self.conv1 = _conv3x3(ch, ch)
        self.bn2 = nn.BatchNorm2d(ch)
        self.conv2 = _conv3x3(ch, ch)

    def forward(self, x):
        out = self.conv1(F.relu(self.bn1(x)))
        out = self.conv2(F.relu(self.bn2(out)))
        return out + x


class ResNet(nn.Module):
    def __init__(self, observation_size, action_size, value_size=3, num_channels=32, depth=4,
                 value_head_channels=16, policy_head_channels=16, value_dense_layers=(512, 64),
                 policy_dense_layers=(512, 256)):
        super().__init__()
        self.channels, self.board_x, self.board_y = observation_size
        self.action_size = action_size
        cells = self.board_x * self.board_y
        self.conv1 = _conv3x3(self.channels, num_channels)
        self.bn1 = nn.BatchNorm2d(num_channels)
        self.resnet = nn.Sequential(*[_PreActBlock(num_channels) for _ in range(depth)])
        self.v_conv = _conv1x1(num_channels, value_head_channels)
        self.v_bn = nn.BatchNorm2d(value_head_channels)
        self.v_fc = _head([cells * value_head_channels, *value_dense_layers, value_size])
        self.pi_conv = _conv1x1(num_channels, policy_head_channels)
        self.pi_bn = nn.BatchNorm2d(policy_head_channels)
        self.pi_fc = _head([cells * policy_head_channels, *policy_dense_layers, action_size])

    @classmethod
    def for_game(cls, game_cls, args):
        keys = ("num_channels", "depth", "value_head_channels", "policy_head_channels",
                "value_dense_layers", "policy_dense_layers")
        kw = {k: args[k] for k in keys if k in args}
        return cls(tuple(game_cls.observation_size()), game_cls.action_size(),
                   game_cls.num_players() + int(game_cls.has_draw()), **kw)

    def forward(self, s):
        s = s.view(-1, self.channels, self.board_x, self.board_y)
        s = F.relu(self.bn1(self.conv1(s)))
        s = self.resnet(s)
        v = self.v_fc(torch.flatten(self.v_bn(self.v_conv(s)), 1))
        pi = self.pi_fc(torch.flatten(self.pi_bn(self.pi_conv(s)), 1))
        return F.log_softmax(pi, dim=1), F.log_softmax(v, dim=1)


class NNetWrapper:
    """Inference half of alphazero/NNetWrapper.py (process / predict); training
    stays with the reference wrapper."""

    def __init__(self, game_cls=None, args=None, nnet=None, cuda=None, fused=False, precision=None):
        self.game_cls, self.args = game_cls, args
        self.nnet = nnet if nnet is not None else ResNet.for_game(game_cls, args)
        self.cuda = torch.cuda.is_available() if cuda is None else cuda
        if self.cuda:
            self.nnet.cuda()
        # fused: evaluate with the hand-written tcgen05 kernels (csrc/azb_resnet_g.cu) when they cover the network, at
        # `precision` (azb200.nn_tc: default "bf16x2", within 1e-5 of this module in fp32; "fp16" / "bf16" are opt-in)
        self.fused, self.precision = fused, precision
        self._fused_eval = {}

    def fused_supported(self):
        from . import nn_tc
        return bool(self.fused) and self.cuda and nn_tc.supported(self.nnet)

    def _process_fused(self, batch):
        n = batch.shape[0]
        # one evaluator (device staging rows + folded weights) per caller tensor: agents served on different streams
        # must not share staging buffers
        key = (n, batch.data_ptr() if not batch.is_cuda else 0)
        ev = self._fused_eval.get(key)
        if ev is None:
            from .nn_tc import TensorCoreEvaluator
            dev = next(self.nnet.parameters()).device
            obs = torch.empty((n,) + tuple(batch.shape[1:]), device=dev)
            ev = TensorCoreEvaluator(self.nnet, obs, torch.empty(n, self.nnet.action_size, device=dev),
                                     torch.empty(n, 3, device=dev), precision=self.precision)
            self._fused_eval[key] = ev
        upload(ev.obs, batch)
        ev()
        return ev.policy, ev.value

    def process(self, batch):
        if self.fused_supported():
            return self._process_fused(batch)
        batch = batch.type(torch.FloatTensor) if not batch.is_cuda else batch.float()
        if self.cuda and not batch.is_cuda:
            batch = batch.cuda()
        self.nnet.eval()
        with torch.no_grad():
            pi, v = self.nnet(batch)
            return torch.exp(pi), torch.exp(v)


_UPLOAD_KERNEL = __import__("os").environ.get("AZB_UPLOAD", "sm") != "dma"


def upload(dst, src):
    """dst (device) <- src (host), asynchronously on the current stream.  A pinned, contiguous source is read by a
    copy kernel through its device mapping (azb_upload_pinned: a few-MB cudaMemcpyAsync pays a DMA start-up that the
    SM path does not); anything else goes through Tensor.copy_."""
    if (_UPLOAD_KERNEL and not src.is_cuda and src.is_pinned() and src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype
            and src.numel() == dst.numel() and src.data_ptr() % 16 == 0 and dst.data_ptr() % 16 == 0):
        import ctypes as C
        from . import _capi
        lib = _capi.load()
        rc = lib.azb_upload_pinned(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size(),
                C.c_void_p(torch.cuda.current_stream(dst.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"azb_upload_pinned failed with status {rc}")
        return dst
    return dst.copy_(src, non_blocking=True)


_DOWNLOAD_KERNEL = __import__("os").environ.get("AZB_DOWNLOAD", "dma") == "sm"


def download(dst, src):
    """dst (pinned host) <- src (device), asynchronously on the current stream: Tensor.copy_ (the DMA engine) by default.
    AZB_DOWNLOAD=sm sends it through the copy kernel instead (posted PCIe writes through the host buffer's device
    mapping) -- measured SLOWER in the host-tensor protocol (e2e 16.8 M instead of 22.2 M sims/s): unlike the upload, the
    download's CTAs sit on SMs that the next trunk launch needs, for as long as PCIe takes."""
    if (_DOWNLOAD_KERNEL and src.is_cuda and not dst.is_cuda and dst.is_pinned() and src.is_contiguous() and dst.is_contiguous()
            and src.dtype == dst.dtype and src.numel() == dst.numel() and src.data_ptr() % 16 == 0 and dst.data_ptr() % 16 == 0):
        import ctypes as C
        from . import _capi
        lib = _capi.load()
        rc = lib.azb_upload_pinned(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size(),
                                   C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"azb_upload_pinned (device -> pinned host) failed with status {rc}")
        return dst
    return dst.copy_(src, non_blocking=True)


_CAPTURE_LOCK = __import__("threading").Lock()


def capture_graph(fn, stream):
    """Capture fn()'s launches on `stream` as a CUDA graph from a worker thread.  Unlike the torch.cuda.graph context
    manager this does not synchronise the device (not permitted while another thread captures) -- fn must not allocate;
    thread-local capture mode, so other threads keep launching work meanwhile."""
    g = torch.cuda.CUDAGraph()
    with _CAPTURE_LOCK, torch.cuda.stream(stream):
        g.capture_begin(capture_error_mode="thread_local")
        try:
            fn()
        finally:
            g.capture_end()
    return g


class HostBatchServer:
    """The body of Coach.processSelfPlayBatches (Coach.py:337-342) for stream-ordered agents
    (azb200.selfplay.SelfPlayAgent(stream_ordered=True)): upload the agent's host observation batch, evaluate, download
    policy / value into the agent's host tensors -- captured once per agent as a CUDA graph on the server stream and
    ordered against the agent's stream by CUDA events (no host synchronisation)."""

    def __init__(self, wrapper, stream=None, stream_per_agent=True):
        self.wrapper = wrapper
        dev = next(wrapper.nnet.parameters()).device
        self.stream = stream or torch.cuda.Stream(device=dev)
        # one server stream per agent: the upload of one agent's batch overlaps the evaluation of another's
        self.stream_per_agent, self._streams, self._dev = stream_per_agent, {}, dev
        self._graphs = {}

    def serve(self, agent, batch_ready):
        """One served batch of `agent` (whose id came out of the ready queue); `batch_ready` is its host event."""
        S = self.stream
        if self.stream_per_agent:
            S = self._streams.get(agent.id)
            if S is None:
                S = self._streams[agent.id] = torch.cuda.Stream(device=self._dev)
        S.wait_event(agent.batch_event)
        g = self._graphs.get(agent.id)
        with torch.cuda.stream(S):
            if g is None and agent.id in self._graphs:           # second batch of this agent: capture
                g = capture_graph(lambda: self._body(agent), S)
                self._graphs[agent.id] = g
            if g is not None:
                g.replay()
            else:                                                # first batch: eager (lazy one-time setup)
                self._body(agent)
                self._graphs[agent.id] = None
            ev = torch.cuda.Event()
            ev.record(S)
        agent.answer_event = ev
        batch_ready.set()

    def _body(self, agent):
        policy, value = self.wrapper.process(agent.batch_tensor)
        download(agent.policy_tensor, policy)
        download(agent.value_tensor, value)


class LeafEvaluator:
    """process() captured once as a CUDA graph over fixed device buffers.

    ``obs`` / ``policy`` / ``value`` are (row slices of) the engine's NN I/O
    tensors; replay() enqueues the whole forward as one graph launch on
    ``stream``, so the tree kernels of another cohort can run beside it.
    precision: "fp32" (strict, no TF32 -- the parity setting), "tf32" (what the
    reference gets from PyTorch's cuDNN default) or "bf16" (autocast)."""

    def __init__(self, nnet, obs, policy, value, precision="tf32", use_graph=True, channels_last=False):
        self.nnet = nnet.eval()
        self.obs, self.policy, self.value = obs, policy, value
        self.precision = precision
        self.channels_last = channels_last
        if channels_last:
            self.nnet.to(memory_format=torch.channels_last)
        self.graph = None
        self.stream = torch.cuda.Stream(device=obs.device)
        if use_graph:
            self._capture()

    def _forward(self):
        tf32 = self.precision != "fp32"
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32 and self.precision == "bf16"
        try:
            with torch.no_grad():
                x = self.obs
                if self.channels_last:
                    x = x.contiguous(memory_format=torch.channels_last)
                if self.precision == "bf16":
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        lp, lv = self.nnet(x)
                    lp, lv = lp.float(), lv.float()
                else:
                    lp, lv = self.nnet(x)
                torch.exp(lp, out=self.policy)
                torch.exp(lv, out=self.value)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    def _capture(self):
        s = self.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._forward()
        s.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=s):
            self._forward()
        s.synchronize()

    def __call__(self, stream=None):
        """Enqueue one evaluation on ``stream`` (default: the current stream)."""
        stream = stream or torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._forward()
