"""Device engine: torch tensors hold every buffer, ctypes calls the sm_100a
kernels.  One engine per process (one process per GPU)."""
import ctypes

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


class Engine(object):
    _instance = None

    @classmethod
    def get(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self, device=None):
        torch = _torch()
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("miniframe_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device)
        self.handle = ctypes.c_void_p()
        st = self.lib.mf_create(ctypes.byref(self.handle), self.device.index)
        if st != _lib.MF_OK:
            raise _lib.MiniframeError("mf_create(device=%d) failed: %s" % (
                self.device.index, self.lib.mf_status_string(st).decode()))
        self._ws = None
        self.max_workspace_bytes = None     # None -> 70 % of the free device memory
        self.launches = 0                   # kernels this engine launched (bench bookkeeping)

    # -- helpers -----------------------------------------------------------
    def stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def tensor(self, x, dtype=None):
        torch = _torch()
        dtype = dtype or torch.float64
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device)

    @staticmethod
    def ptr(t):
        return ctypes.c_void_p(0 if t is None else t.data_ptr())

    def ld(self, n):
        return int(self.lib.mf_ld(n))

    def matrix_stride(self, n):
        return int(self.lib.mf_matrix_stride(n))

    def workspace(self, batch, n):
        """A cached byte buffer holding as many of `batch` matrices as memory allows."""
        torch = _torch()
        per = self.matrix_stride(n) * 8
        if (self._ws is not None and self._ws.numel() >= batch * per and
                (self.max_workspace_bytes is None or batch * per <= self.max_workspace_bytes)):
            return self._ws, batch * per          # the cached buffer already holds the whole batch
        cap = self.max_workspace_bytes
        if cap is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            have = 0 if self._ws is None else self._ws.numel()
            cap = int(0.7 * (free + have))
        count = max(1, min(batch, cap // per))
        need = count * per
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    # -- kernels -----------------------------------------------------------
    def build_cov(self, prob, t, diag_add, hyper, extra, full=True):
        """Returns the (B, rows, ld) workspace tensor (rows = stride / ld >= n + 1); [:, :n, :n] is K."""
        torch = _torch()
        B = hyper.shape[0]
        n = prob.n_series * prob.n_epochs
        ld, stride = self.ld(n), self.matrix_stride(n)
        K = torch.zeros((B, stride // ld, ld), dtype=torch.float64, device=self.device)
        st = self.lib.mf_build_cov(self.handle, ctypes.byref(prob), B, self.ptr(t), self.ptr(diag_add),
                                   self.ptr(hyper), self.ptr(extra), 1 if full else 0,
                                   self.ptr(K), ld, stride, self.stream())
        _lib.check(self.lib, self.handle, st, "mf_build_cov")
        self.launches += 1
        return K

    def potrf_loglike(self, K, n, resid):
        torch = _torch()
        B = K.shape[0]
        ld, stride = K.shape[2], K.shape[1] * K.shape[2]
        ll = torch.empty(B, dtype=torch.float64, device=self.device)
        info = torch.empty(B, dtype=torch.int32, device=self.device)
        rs = 0 if (resid is None or resid.shape[0] == 1) else resid.shape[1]
        st = self.lib.mf_potrf_loglike(self.handle, B, n, self.ptr(K), ld, stride, self.ptr(resid), rs,
                                       self.ptr(ll), self.ptr(info), self.stream())
        _lib.check(self.lib, self.handle, st, "mf_potrf_loglike")
        self.launches += 1
        return ll, info

    def solve_logdet(self, L, n, resid, info=None, shared=False):
        """z = L^-1 resid (and the log-likelihood terms) per matrix; ``shared``: ONE factor,
        L[0], for every row of ``resid`` (a multi-right-hand-side forward substitution)."""
        torch = _torch()
        B = resid.shape[0] if shared else L.shape[0]
        ld, stride = L.shape[2], (0 if shared else L.shape[1] * L.shape[2])
        ll = torch.empty(B, dtype=torch.float64, device=self.device)
        z = torch.empty((B, n), dtype=torch.float64, device=self.device)
        rs = 0 if resid.shape[0] == 1 else resid.shape[1]
        st = self.lib.mf_solve_logdet(self.handle, B, n, self.ptr(L), ld, stride, self.ptr(resid), rs,
                                      self.ptr(z), self.ptr(ll), self.ptr(info), self.stream())
        _lib.check(self.lib, self.handle, st, "mf_solve_logdet")
        self.launches += 1
        return z, ll

    def loglike(self, prob, t, diag_add, hyper, extra, resid, out_ll=None, out_info=None):
        """The whole hot path for a batch held on the device.  Returns (ll, info)."""
        torch = _torch()
        B = hyper.shape[0]
        n = prob.n_series * prob.n_epochs
        ws, ws_bytes = self.workspace(B, n)
        ll = out_ll if out_ll is not None else torch.empty(B, dtype=torch.float64, device=self.device)
        info = out_info if out_info is not None else torch.empty(B, dtype=torch.int32, device=self.device)
        rs = 0 if (resid is None or resid.shape[0] == 1) else resid.shape[1]
        st = self.lib.mf_loglike(self.handle, ctypes.byref(prob), B, self.ptr(t), self.ptr(diag_add),
                                 self.ptr(hyper), self.ptr(extra), self.ptr(resid), rs,
                                 self.ptr(ws), ws_bytes, self.ptr(ll), self.ptr(info), self.stream())
        _lib.check(self.lib, self.handle, st, "mf_loglike")
        per = self.matrix_stride(n) * 8
        self.launches += 2 * (-(-B // max(1, ws_bytes // per)))
        return ll, info

    def probe_fp64(self, mode, iters=20000):
        ms = ctypes.c_float()
        flops = ctypes.c_double()
        st = self.lib.mf_probe_fp64(self.handle, mode, iters, ctypes.byref(ms), ctypes.byref(flops),
                                    self.stream())
        _lib.check(self.lib, self.handle, st, "mf_probe_fp64")
        return flops.value / (ms.value * 1e-3) / 1e12     # TFLOP/s
