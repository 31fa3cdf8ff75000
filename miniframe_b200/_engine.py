"""Device engine: torch tensors hold every buffer, ctypes calls the sm_100a
kernels.  One engine per process (one process per GPU); inside it one library
handle per CUDA stream (a handle's scratch - the hand-out counter and group
barriers of the factorisation - must not be shared by concurrent launches)."""
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
        self._handles = {}                  # cuda_stream -> mf_handle_t
        self._options = {}
        self._ws = {}                       # cuda_stream -> workspace buffer
        self.max_workspace_bytes = None     # None -> 70 % of the free device memory
        self.launches = 0                   # kernels this engine launched (bench bookkeeping)
        self.handle                         # fail now if the device is not a B200

    # -- helpers -----------------------------------------------------------
    def _stream_id(self):
        return int(_torch().cuda.current_stream(self.device).cuda_stream)

    @property
    def handle(self):
        """The library handle of torch's current stream (created on first use)."""
        sid = self._stream_id()
        h = self._handles.get(sid)
        if h is None:
            h = ctypes.c_void_p()
            st = self.lib.mf_create(ctypes.byref(h), self.device.index)
            if st != _lib.MF_OK:
                raise _lib.MiniframeError("mf_create(device=%d) failed: %s" % (
                    self.device.index, self.lib.mf_status_string(st).decode()))
            for name, value in self._options.items():
                self.lib.mf_set_option(h, name.encode(), int(value))
            self._handles[sid] = h
        return h

    def set_option(self, name, value):
        """mf_set_option on every handle of this engine (see include/miniframe_b200.h)."""
        self._options[name] = int(value)
        for h in self._handles.values():
            st = self.lib.mf_set_option(h, name.encode(), int(value))
            _lib.check(self.lib, h, st, "mf_set_option(%s)" % name)

    def stream(self):
        return ctypes.c_void_p(self._stream_id())

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
        sid = self._stream_id()
        ws = self._ws.get(sid)
        if (ws is not None and ws.numel() >= batch * per and
                (self.max_workspace_bytes is None or batch * per <= self.max_workspace_bytes)):
            return ws, batch * per                # the cached buffer already holds the whole batch
        cap = self.max_workspace_bytes
        if cap is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            have = 0 if ws is None else ws.numel()
            cap = int(0.7 * (free + have))
        count = max(1, min(batch, cap // per))
        need = count * per
        if ws is None or ws.numel() < 0.9 * need:
            # grow (a cached buffer within 10 % of what memory allows is kept: free memory
            # fluctuates): give the old buffer back first, then size the new one by what is
            # really free (somebody else may still hold a reference to the old one)
            self._ws.pop(sid, None)
            ws = None
            if self.max_workspace_bytes is None:
                torch.cuda.empty_cache()
                free, _ = torch.cuda.mem_get_info(self.device)
                count = max(1, min(batch, int(0.7 * free) // per))
                need = count * per
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[sid] = ws
        else:
            need = min(need, ws.numel() // per * per)
        return ws, need

    def drop_workspaces(self):
        self._ws.clear()

    # -- kernels -----------------------------------------------------------
    def build_cov(self, prob, t, diag_add, hyper, extra, full=True, interleaved=False):
        """Returns the (B, rows, ld) workspace tensor (rows = stride / ld >= n + 1); [:, :n, :n] is K
        (``interleaved``: lower triangle in the epoch-interleaved order r = n_series * i + P)."""
        torch = _torch()
        B = hyper.shape[0]
        n = prob.n_series * prob.n_epochs
        ld, stride = self.ld(n), self.matrix_stride(n)
        K = torch.zeros((B, stride // ld, ld), dtype=torch.float64, device=self.device)
        st = self.lib.mf_build_cov(self.handle, ctypes.byref(prob), B, self.ptr(t), self.ptr(diag_add),
                                   self.ptr(hyper), self.ptr(extra),
                                   _lib.COV_LOWER_INTERLEAVED if interleaved else (_lib.COV_FULL if full else _lib.COV_LOWER),
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

    def current_workspace(self):
        """The workspace buffer of torch's current stream (after loglike: the factors, first chunk first)."""
        return self._ws[self._stream_id()]

    def loglike(self, prob, t, diag_add, hyper, extra, resid, out_ll=None, out_info=None, max_chunk=None):
        """The whole hot path for a batch held on the device.  Returns (ll, info)."""
        torch = _torch()
        B = hyper.shape[0]
        n = prob.n_series * prob.n_epochs
        ws, ws_bytes = self.workspace(B if max_chunk is None else min(B, max_chunk), n)
        hyper = hyper.contiguous()
        if extra is not None:
            extra = extra.contiguous()
        ll = out_ll if out_ll is not None else torch.empty(B, dtype=torch.float64, device=self.device)
        info = out_info if out_info is not None else torch.empty(B, dtype=torch.int32, device=self.device)
        rs = 0 if (resid is None or resid.shape[0] == 1) else resid.shape[1]
        st = self.lib.mf_loglike(self.handle, ctypes.byref(prob), B, self.ptr(t), self.ptr(diag_add),
                                 self.ptr(hyper), self.ptr(extra), self.ptr(resid), rs,
                                 self.ptr(ws), ws_bytes, self.ptr(ll), self.ptr(info), self.stream())
        _lib.check(self.lib, self.handle, st, "mf_loglike")
        per = self.matrix_stride(n) * 8
        interleaved = prob.framework != _lib.FRAMEWORK_SMALLGP and prob.nugget == 0.0 and resid is not None
        self.launches += (3 if interleaved else 2) * (-(-B // max(1, ws_bytes // per)))
        return ll, info

    def probe_fp64(self, mode, iters=20000):
        ms = ctypes.c_float()
        flops = ctypes.c_double()
        st = self.lib.mf_probe_fp64(self.handle, mode, iters, ctypes.byref(ms), ctypes.byref(flops),
                                    self.stream())
        _lib.check(self.lib, self.handle, st, "mf_probe_fp64")
        return flops.value / (ms.value * 1e-3) / 1e12     # TFLOP/s
