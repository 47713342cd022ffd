r"""
:mod:`signal.correlate` -- 1-D / 2-D FFT correlation of arrays (timelapse kernel)
===================================================================================

``correlate2d`` of the reference (``signal/correlate.py:162-291``) pads two
``F x T`` arrays to ``(2F-1) x (2T-1)``, multiplies ``fft2(f)`` by
``conj(fft2(g))`` and returns ``fftshift(real(ifft2(.)))``; the timelapse script
evaluates it for every day pair of CCF spectrograms
(``scripts/timelapse.py:218-275``) and keeps the arg-max.

Here the 2-D linear correlation runs on the 1-D kernels of the main path by
embedding: write each array into rows of width ``W >= 2T-1`` (zeros to the
right) and flatten.  The 1-D linear correlation of the flattened arrays at lag
``dF*W + dT`` (``|dT| < T``) is exactly ``sum f[u+dF, v+dT] g[u, v]`` -- a
shifted column either stays inside its row or lands on padding, never on a
neighbouring row's data.  So every image is transformed ONCE (``N`` transforms
+ ``P`` products for ``P`` image pairs, the reference recomputes both
spectra per pair), with ``n = F*W`` samples; for the named shape 1024 x 128,
``W = 256`` gives ``n = 262144 = 64^3``.
"""
import numpy as np
import torch

from ..plan import CorrelationPlan

__all__ = ['correlate1d', 'correlate2d', 'Correlate2dPlan', 'timelapse_pairs']


def _row_width(T):
    """Smallest multiple of 16 that is >= 2T-1 (keeps rows 64-byte aligned)."""
    return max(16, -(-(2 * T - 1) // 16) * 16)


def timelapse_pairs(n_days, k=1):
    """Day pairs ``(time1, time2)`` the timelapse script evaluates: everything on and below the k-th
    super-diagonal (``signal/tri.py:19-55`` ``np.tri(N, N, k)``, applied with k = 1 at
    ``scripts/timelapse.py:348-352``): N(N+1)/2 + N-1 pairs for k = 1."""
    return [(i, j) for i in range(n_days) for j in range(min(n_days, i + k + 1))]


class Correlate2dPlan:
    """All 2-D correlations of ``n_images`` arrays ``F x T`` over a pair list.

    :meth:`correlate_argmax` is the timelapse kernel proper (``scripts/timelapse.py:255-268`` keeps
    ``cc2.argmax()`` -> value, delta-freq, delta-lag per day pair): the maximum over the
    ``(2F-1) x (2T-1)`` window is taken inside the last inverse pass and no ``cc2`` is stored.
    :meth:`correlate` still returns the full ``cc2`` (parity, plots).

    With ``day_shard=(rank, world)`` the images are split over the ranks of ``torch.distributed``'s
    default group: every rank transforms ``ceil(n_images / world)`` of them and ONE all-gather of the
    half spectra (NCCL) hands every rank all of them; `pairs` is then this rank's share of the pair list
    (SURVEY 8e, config 5)."""

    def __init__(self, F, T, n_images, pairs, dtype=torch.float32, normalize=True, device=None, day_shard=None,
                 max_work_bytes=0):
        self.F, self.T, self.W = int(F), int(T), _row_width(T)
        self.n = self.F * self.W
        self.pairs = [(int(i), int(j)) for i, j in pairs]
        self.n_images = int(n_images)
        self.rank, self.world = (int(day_shard[0]), int(day_shard[1])) if day_shard else (0, 1)
        self.per = -(-self.n_images // self.world)                       # images transformed per rank
        lo = min(self.n_images, self.rank * self.per)
        self.day_slice = slice(lo, min(self.n_images, lo + self.per))
        self.plan = CorrelationPlan(self.n, self.per * self.world, self.pairs, dtype=dtype, normalize=normalize,
                                    device=device, n_blocks=1, max_work_bytes=max_work_bytes)
        self.fwd = self.plan if self.world == 1 else \
            CorrelationPlan(self.n, self.per, [], dtype=dtype, normalize=normalize, device=device, n_blocks=1)
        self.dtype, self.device = self.plan.dtype, self.plan.device
        self._x = torch.zeros((self.per if self.world > 1 else self.n_images, self.F, self.W), dtype=self.dtype,
                              device=self.device)                        # row-embedded images; the padding stays zero
        self._index = None

    def embed(self, images):
        """(k, F, T) -> (k, F*W): every image row is followed by W - T zeros, so that the 1-D lag
        ``dF*W + dT`` of the flattened arrays is the 2-D shift (dF, dT)."""
        k = images.shape[0]
        self._x[:k, :, :self.T].copy_(images)
        return self._x.reshape(self._x.shape[0], self.n)

    def forward(self, images):
        """Transform this rank's images (all of them without sharding) and, when sharded, all-gather
        the spectra and the norms."""
        x = self.embed(images)
        self.fwd.forward(x)
        self.all_gather()

    def all_gather(self):
        """The one collective of the sharded plan: every rank's half spectra and norms to every rank (NCCL)."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(torch.view_as_real(self.plan.spec), torch.view_as_real(self.fwd.spec))
            dist.all_gather_into_tensor(self.plan.sumsq, self.fwd.sumsq)

    def correlate(self):
        """cc2 of every pair, shape (n_pairs, 2F-1, 2T-1)."""
        if self._index is None:
            a = torch.arange(-(self.F - 1), self.F, device=self.device).view(-1, 1)
            b = torch.arange(-(self.T - 1), self.T, device=self.device).view(1, -1)
            # 1-D output index of shift (dF, dT): lag dF*W + dT sits at (n-1) + lag
            self._index = ((self.n - 1) + a * self.W + b).reshape(-1)
        ccf = self.plan.correlate()[:, 0]
        return ccf.index_select(1, self._index).reshape(len(self.pairs), 2 * self.F - 1, 2 * self.T - 1)

    def correlate_argmax(self, images=None):
        """(value, dF, dT) of the maximum of every pair's cc2 (first maximal index in C order, like
        ``np.unravel_index(cc2.argmax(), cc2.shape)``), fused into the inverse transform."""
        if images is not None:
            self.forward(images)
        idx, val = self.plan.correlate_argmax(mask=(self.W, self.T, self.F))
        t = idx[:, 0] - (self.n - 1) + (self.T - 1) + self.F * self.W      # lag + (T-1) + F*W > 0
        return val[:, 0], t // self.W - self.F, t % self.W - (self.T - 1)

    def argmax(self, cc2):
        """(value, dF, dT) of the maximum of each cc2 (first maximal index,
        like ``np.argmax`` + ``np.unravel_index`` in ``scripts/timelapse.py:261-268``)."""
        idx, val = self.plan.argmax(cc2.reshape(cc2.shape[0], -1))
        nb = 2 * self.T - 1
        return val, idx // nb - (self.F - 1), idx % nb - (self.T - 1)

    def __call__(self, images):
        self.forward(images)
        return self.correlate()


def _as_cuda(x, dtype=None):
    t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda() if not t.is_cuda else t


def correlate2d(in1, in2, normalize: bool = True, dtype=None):
    """
    Two-dimensional cross-correlation of two arrays over their last two axes
    (array core of ``signal/correlate.py:162-291``; labels and attributes are
    the caller's).  Returns ``(..., 2F-1, 2T-1)``; index ``(a, b)`` is the shift
    ``(a-(F-1), b-(T-1))``.  numpy in -> numpy out, CUDA tensor in -> tensor out.
    """
    was_numpy = not torch.is_tensor(in1)
    a, b = _as_cuda(in1), _as_cuda(in2)
    if a.shape != b.shape:
        raise ValueError('in1 and in2 should have the same shape!')
    if a.dtype not in (torch.float32, torch.float64):
        a, b = a.double(), b.double()
    if dtype is not None:
        a, b = a.to(dtype), b.to(dtype)
    lead = a.shape[:-2]
    F, T = a.shape[-2:]
    a2, b2 = a.reshape(-1, F, T), b.reshape(-1, F, T)
    m = a2.shape[0]
    plan = Correlate2dPlan(F, T, 2 * m, [(k, m + k) for k in range(m)], dtype=a.dtype, normalize=normalize)
    cc = plan(torch.cat((a2, b2), dim=0)).reshape(*lead, 2 * F - 1, 2 * T - 1)
    return cc.cpu().numpy() if was_numpy else cc


def correlate1d(in1, in2, normalize: bool = True, dtype=None):
    """
    One-dimensional cross-correlation along the last axis (array core of
    ``signal/correlate.py:32-160``): ``(..., 2n-1)``, zero lag at index ``n-1``.
    """
    was_numpy = not torch.is_tensor(in1)
    a, b = _as_cuda(in1), _as_cuda(in2)
    if a.shape != b.shape:
        raise ValueError('in1 and in2 should have the same shape!')
    if a.dtype not in (torch.float32, torch.float64):
        a, b = a.double(), b.double()
    if dtype is not None:
        a, b = a.to(dtype), b.to(dtype)
    lead, n = a.shape[:-1], a.shape[-1]
    a2, b2 = a.reshape(-1, n), b.reshape(-1, n)
    m = a2.shape[0]
    plan = CorrelationPlan(n, 2 * m, [(k, m + k) for k in range(m)], dtype=a.dtype, normalize=normalize, n_blocks=1)
    cc = plan(torch.cat((a2, b2), dim=0).contiguous())[:, 0].reshape(*lead, 2 * n - 1)
    return cc.cpu().numpy() if was_numpy else cc
