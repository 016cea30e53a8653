"""GPU-resident training-sample window: the input side of Coach.train (alphazero/Coach.py:436-520).

The reference writes every iteration's examples to three files (Coach.saveIterationSamples, Coach.py:364-386), then
for every training phase re-loads the files of the history window, wraps them in TensorDataset / ConcatDataset and
iterates a ``DataLoader(shuffle=True, num_workers=workers, pin_memory=True)`` whose batches NNetWrapper.train uploads
one by one (NNetWrapper.py:143-150).  Here the window stays on the device:

``SampleWindow``   holds the (obs, pi, z) tensors of the last iterations on the GPU (straight from
                   SelfPlayEngine.drain_samples_into, or loaded from the reference's three-file format, which it
                   also writes), and computes the reference's window size and train-step count;
``WindowLoader``   the iterable NNetWrapper.train consumes: every epoch is one random permutation of the window cut
                   into ``train_batch_size`` batches -- the same permutation torch's RandomSampler draws for
                   ``DataLoader(ConcatDataset(...), shuffle=True)`` from the same global RNG state, so a training run
                   sees the same batches in the same order as the reference's (tests/test_samples.py) -- gathered on
                   the device with one index_select per tensor: no worker processes, no pinned staging, no uploads.
"""
import os
import pickle

import torch


def history_window(iteration, args):
    """Coach.train (Coach.py:509-516): number of past iterations trained on at `iteration`."""
    g = lambda k, d: (args[k] if k in args else d)
    lo, hi, inc = g("minTrainHistoryWindow", 4), g("maxTrainHistoryWindow", 20), g("trainHistoryIncrementIters", 2)
    return min(max(lo, (iteration + lo) // inc), hi)


def iter_file(iteration):
    """alphazero/utils.py:15-16 without the .pkl suffix"""
    return f"iteration-{iteration:04d}"


class WindowLoader:
    """Drop-in for the DataLoader of Coach.train's train_data (Coach.py:466-469): ``for boards, pis, vs in loader``."""

    def __init__(self, tensors, batch_size, drop_last=False):
        self.data, self.policy, self.value = tensors
        self.n = int(self.data.shape[0])
        self.batch_size = int(batch_size)
        self.drop_last = drop_last

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else -(-self.n // self.batch_size)

    def __iter__(self):
        # the global-RNG draws of one DataLoader epoch, in order: the iterator's worker base seed
        # (_BaseDataLoaderIter.__init__), then RandomSampler.__iter__ (generator=None): a fresh generator seeded from
        # the global RNG and one randperm(n)
        torch.empty((), dtype=torch.int64).random_()
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        gen = torch.Generator()
        gen.manual_seed(seed)
        perm = torch.randperm(self.n, generator=gen).to(self.data.device)
        for b in range(len(self)):
            idx = perm[b * self.batch_size:(b + 1) * self.batch_size]
            yield self.data.index_select(0, idx), self.policy.index_select(0, idx), self.value.index_select(0, idx)


class SampleWindow:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.iters = {}          # iteration -> (obs, pi, z) device tensors

    # ---- filling ------------------------------------------------------------------------------
    def add_iteration(self, iteration, obs, pi, z):
        f = lambda t: torch.as_tensor(t).to(self.device, dtype=torch.float32).contiguous()
        obs, pi, z = f(obs), f(pi), f(z)
        assert obs.shape[0] == pi.shape[0] == z.shape[0]
        self.iters[int(iteration)] = (obs, pi, z)

    def add_from_engine(self, iteration, engine):
        """Everything currently in the engine's sample ring, device to device (no host copy)."""
        n = engine.sample_count()
        obs = torch.empty((n,) + tuple(engine.obs_shape), device=self.device)
        pi = torch.empty(n, engine.A, device=self.device)
        z = torch.empty(n, 3, device=self.device)
        if n:
            engine.drain_samples_into(obs, pi, z)
        if int(iteration) in self.iters:                        # several drains of one self-play phase
            o0, p0, z0 = self.iters[int(iteration)]
            obs, pi, z = torch.cat([o0, obs]), torch.cat([p0, pi]), torch.cat([z0, z])
        self.iters[int(iteration)] = (obs, pi, z)
        return n

    def load_iteration(self, iteration, data_dir, run_name):
        """The reference's three files (Coach.py:443-447); False if they do not exist (Coach.py:448-450)."""
        base = os.path.join(data_dir, run_name, iter_file(iteration))
        try:
            t = [torch.load(base + s, weights_only=False) for s in ("-data.pkl", "-policy.pkl", "-value.pkl")]
        except FileNotFoundError:
            return False
        self.add_iteration(iteration, *t)
        return True

    def save_iteration(self, iteration, data_dir, run_name):
        """Coach.saveIterationSamples (Coach.py:364-386)."""
        folder = os.path.join(data_dir, run_name)
        os.makedirs(folder, exist_ok=True)
        base = os.path.join(folder, iter_file(iteration))
        for t, s in zip(self.iters[int(iteration)], ("-data.pkl", "-policy.pkl", "-value.pkl")):
            torch.save(t.cpu(), base + s, pickle_protocol=pickle.HIGHEST_PROTOCOL)
        return base

    def evict_before(self, iteration):
        for i in [i for i in self.iters if i < iteration]:
            del self.iters[i]

    # ---- Coach.train ------------------------------------------------------------------------------
    def window(self, iteration, args):
        """Iterations Coach.train uses at `iteration` (Coach.py:509-518), those present here, ascending."""
        size = history_window(iteration, args)
        return [i for i in range(max(1, iteration - size), iteration + 1) if i in self.iters]

    def tensors(self, iterations):
        """ConcatDataset order: iterations ascending, samples in emission order."""
        parts = [self.iters[i] for i in iterations]
        return tuple(torch.cat([p[k] for p in parts]) for k in range(3))

    def train_steps(self, iterations, args, train_on_all=False):
        """Coach.train's step count (Coach.py:452-478): autoTrainSteps / averageTrainSteps / train_steps_per_iteration."""
        g = lambda k, d: (args[k] if k in args else d)
        sizes = [int(self.iters[i][0].shape[0]) for i in iterations]
        total, bs = sum(sizes), int(g("train_batch_size", 1024))
        if train_on_all:
            return total // bs
        if not g("autoTrainSteps", True):
            return int(g("train_steps_per_iteration", 64))
        n = (sum(sizes) // len(sizes)) if g("averageTrainSteps", False) else sizes[-1]
        return n // bs

    def loader(self, iteration, args):
        its = self.window(iteration, args)
        return WindowLoader(self.tensors(its), int(args["train_batch_size"] if "train_batch_size" in args else 1024)), its

    def train(self, train_net, iteration, args):
        """train_data of Coach.train for the window at `iteration`: -> (loss_pi, loss_v) of NNetWrapper.train."""
        loader, its = self.loader(iteration, args)
        return train_net.train(loader, self.train_steps(its, args))


def loss_pi(targets, outputs):
    """NNetWrapper.loss_pi (NNetWrapper.py:234-235); outputs are log-probabilities"""
    return -torch.sum(targets * outputs) / targets.size()[0]


def loss_v(targets, outputs, value_loss_weight=1.0):
    """NNetWrapper.loss_v (NNetWrapper.py:237-238)"""
    return -value_loss_weight * torch.sum(targets * outputs) / targets.size()[0]
