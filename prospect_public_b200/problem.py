"""Host-side (FP64) preparation of the Gaussian target and per-point proposals, and their
device images.  Everything here happens once per job, not per step.

Reference behaviour reproduced (paths relative to the reference root):
  * -log L = 1/2 r^T inv(C) r with r = [varying..., fixed] - means contracted in the ORIGINAL
    parameter order (prospect/kernels/analytical_kernel.py:59-66) -- only
    S = sym(inv(C)) matters and its blocks are taken by slot, not by parameter index;
  * proposal covariance = covmat * step_size**2 with covmat per profile point
    (prospect/mcmc.py:186-187, tasks/initialise_optimiser_task.py:126-130);
  * prior box on the varying parameters only (prospect/kernels/base_kernel.py:89-102).

What is new: instead of an SVD of the proposal covariance and an inverse of C per step
 the proposal factor F (F F^T = covmat) is rotated by the eigenvectors of
F^T S_aa F so that the change of -log L along a proposal is an O(n) expression in the
whitened draw z (see include/pb200.h, struct pb200_problem).
"""
from dataclasses import dataclass

import numpy as np

PADS = (32, 64, 128, 256)


def pad_of(n):
    for p in PADS:
        if n <= p:
            return p
    raise ValueError(f'at most {PADS[-1]} varying parameters are supported, got {n}')


def sym_precision(covmat):
    a = np.linalg.inv(np.asarray(covmat, dtype=np.float64))
    return 0.5 * (a + a.T)


def proposal_factor(covmat):
    """Any F with F F^T = covmat.  Cholesky when positive definite; otherwise the clipped
    symmetric eigen-factor (numpy's multivariate_normal tolerates PSD input through its SVD,
    prospect/mcmc.py:187)."""
    cov = np.asarray(covmat, dtype=np.float64)
    cov = 0.5 * (cov + cov.T)
    try:
        return np.linalg.cholesky(cov)
    except np.linalg.LinAlgError:
        w, v = np.linalg.eigh(cov)
        return v * np.sqrt(np.clip(w, 0.0, None))


@dataclass
class HostProblem:
    d: int
    n: int
    npad: int
    n_points: int
    saa: np.ndarray      # [npad, npad] f64
    sab: np.ndarray      # [npad]
    sbb: float
    mu: np.ndarray       # [npad] means of the varying parameters
    lo: np.ndarray       # [npad]
    hi: np.ndarray       # [npad]
    f: np.ndarray        # [P]
    c0: np.ndarray       # [P]
    lt_t: np.ndarray     # [P, npad, npad] f32, lt_t[p, j, i] = Lt[i, j]
    lam: np.ndarray      # [P, npad] f32
    lt_norm: np.ndarray  # [P, npad] f32, Euclidean row norms of the FP32 Lt, rounded up
    g_t: np.ndarray      # [P, npad, npad] f64, g_t[p, j, i] = G[i, j]
    h: np.ndarray        # [P, npad]
    lt_hi: object        # [P, npad, npad] f32, lt_hi[p, i, j] = TF32 part of Lt[i, j] (npad >= 128 only, else None)
    lt_lo: object        # [P, npad, npad] f32, the exact remainder Lt - lt_hi
    varying: list        # indices of the varying parameters in x1..xd
    fixed_index: object  # index of the profiled parameter or None
    fixed_values: np.ndarray  # [P] (empty for free runs)

    def loglkl(self, x, point=0):
        """FP64 host evaluation with the same formula the device uses (for set-up work such as
        initial_loglkl; the batched path is CudaGaussianKernel / loglkl_batch)."""
        q = np.zeros(self.npad)
        q[:self.n] = np.asarray(x, dtype=np.float64)[:self.n] - self.mu[:self.n]
        return float(q @ (0.5 * (self.saa @ q) + self.f[point] * self.sab) + self.c0[point])


def build_problem(means, covmat, lo, hi, fixed_index, fixed_values, proposal_covmats, ignore_prior=False):
    """means [d], covmat [d, d] in x1..xd order; lo/hi [d] (None entries = open);
    fixed_index None -> free run (one 'point'); fixed_values [P]; proposal_covmats [P, n, n]."""
    means = np.asarray(means, dtype=np.float64)
    d = len(means)
    varying = [i for i in range(d) if i != fixed_index]
    n = len(varying)
    npad = pad_of(n)
    s = sym_precision(covmat)
    saa = np.zeros((npad, npad))
    sab = np.zeros(npad)
    saa[:n, :n] = s[:n, :n]
    if fixed_index is None:
        fixed_values = np.zeros(0)
        f = np.zeros(1)
        sbb = 0.0
    else:
        fixed_values = np.asarray(fixed_values, dtype=np.float64).reshape(-1)
        sab[:n] = s[:n, n]
        sbb = float(s[n, n])
        f = fixed_values - means[fixed_index]
    n_points = len(f)
    c0 = 0.5 * sbb * f * f
    mu = np.zeros(npad)
    mu[:n] = means[varying]
    lo_p = np.full(npad, -np.inf)
    hi_p = np.full(npad, np.inf)
    if not ignore_prior:
        for k, i in enumerate(varying):
            if lo is not None and lo[i] is not None:
                lo_p[k] = lo[i]
            if hi is not None and hi[i] is not None:
                hi_p[k] = hi[i]
    if proposal_covmats is None:         # likelihood-only image (pb200_loglkl_batch): one zero proposal slot
        slots = 1
    else:
        proposal_covmats = np.asarray(proposal_covmats, dtype=np.float64)
        if proposal_covmats.ndim == 2:
            proposal_covmats = proposal_covmats[None]
        if proposal_covmats.shape != (n_points, n, n):
            raise ValueError(f'proposal covmats must have shape {(n_points, n, n)}, got {proposal_covmats.shape}')
        slots = n_points
    lt_t = np.zeros((slots, npad, npad), dtype=np.float32)
    lam = np.zeros((slots, npad), dtype=np.float32)
    lt_norm = np.zeros((slots, npad), dtype=np.float32)
    g_t = np.zeros((slots, npad, npad))
    h = np.zeros((slots, npad))
    s_aa = saa[:n, :n]
    for p in range(n_points if proposal_covmats is not None else 0):
        fac = proposal_factor(proposal_covmats[p])
        m = fac.T @ s_aa @ fac
        w, q = np.linalg.eigh(0.5 * (m + m.T))
        lt = fac @ q
        g = lt.T @ s_aa
        lt_t[p, :n, :n] = lt.T
        lam[p, :n] = w
        # bound used by the deferred prior-box test: norms of the rows of the FP32 factor, nudged upwards
        rows = np.linalg.norm(lt_t[p, :n, :n].astype(np.float64), axis=0) * (1.0 + 1e-6)
        lt_norm[p, :n] = np.nextafter(rows.astype(np.float32), np.float32(np.inf))
        g_t[p, :n, :n] = g.T
        h[p, :n] = lt.T @ sab[:n] * f[p]
    lt_hi = lt_lo = None
    if npad >= 128 and proposal_covmats is not None:       # operands of the tcgen05 (3xTF32) position GEMM of the wide path: K-major, hi + lo == Lt exactly
        lt_rows = np.ascontiguousarray(lt_t.transpose(0, 2, 1))              # [P, i, j] = Lt[i][j]
        lt_hi = (lt_rows.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        lt_lo = lt_rows - lt_hi
    return HostProblem(d, n, npad, n_points, saa, sab, sbb, mu, lo_p, hi_p, f, c0, lt_t, lam, lt_norm, g_t, h, lt_hi,
                       lt_lo, varying, fixed_index, fixed_values)


_TORCH_DTYPES = {'<f8': 'float64', '<f4': 'float32', '<i4': 'int32', '<u4': 'int32', '<i8': 'int64'}


def upload_packed(arrays, device):
    """One pinned staging buffer + ONE host-to-device copy for a dict of numpy arrays; returns {name: tensor}
    views (256-byte aligned) into a single device allocation.  Keeps set-up latency off the end-to-end path."""
    import torch
    offsets, total = {}, 0
    for name, a in arrays.items():
        total = (total + 255) // 256 * 256
        offsets[name] = total
        total += a.nbytes
    host = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    for name, a in arrays.items():
        hv[offsets[name]:offsets[name] + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    dev = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
    dev.copy_(host, non_blocking=True)
    out = {}
    for name, a in arrays.items():
        dt = getattr(torch, _TORCH_DTYPES[a.dtype.str])
        out[name] = dev[offsets[name]:offsets[name] + a.nbytes].view(dt).view(a.shape)
    out['_storage'] = (dev, host)      # keep both alive until the copy has been consumed
    return out


class DeviceProblem:
    """Device image of a HostProblem + the ctypes struct handed to libpb200."""

    def __init__(self, host: HostProblem, device):
        import torch
        from . import capi
        self.host = host
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise capi.Pb200Error('prospect_public_b200 runs on CUDA devices only (no CPU fallback)')
        names = ('saa', 'sab', 'mu', 'lo', 'hi', 'f', 'c0', 'lt_t', 'lam', 'lt_norm', 'g_t', 'h')
        if host.lt_hi is not None:
            names += ('lt_hi', 'lt_lo')
        self.t = upload_packed({name: getattr(host, name) for name in names}, self.device)
        self._storage = self.t.pop('_storage')
        self.c = capi.Problem(host.d, host.n, host.npad, host.n_points,
                              *(capi.ptr(self.t.get(k)) for k in ('saa', 'sab', 'mu', 'lo', 'hi', 'f', 'c0', 'lt_t',
                                                                  'lam', 'lt_norm', 'g_t', 'h', 'lt_hi', 'lt_lo')))

    @property
    def n(self):
        return self.host.n

    @property
    def npad(self):
        return self.host.npad
