"""Device engine: owns a ``liblmc`` model handle and the PyTorch device buffers.

PyTorch is plumbing only (device memory, streams, pinned host copies); every kernel on the
path is ours and is launched through the C ABI in ``include/lmc.h``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .model import PackedModel


def _torch():
    import torch
    return torch


_POOL = None


def _copy_into(dst: np.ndarray, src: np.ndarray):
    """dst[...] = src with dtype conversion; large arrays are split over a few threads (numpy releases
    the GIL inside copyto)."""
    global _POOL
    if src.nbytes < (4 << 20) or src.shape[0] < 4:
        np.copyto(dst, src, casting="unsafe")
        return
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=4)
    bounds = np.linspace(0, src.shape[0], 5).astype(int)
    list(_POOL.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]], casting="unsafe"),
                   zip(bounds[:-1], bounds[1:])))


def _on_device(fn):
    """run the method with the engine's device current (launches and allocations of the C ABI go to the current
    device; a sampler built with ``device="cuda:1"`` must work while another device is current)"""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        torch = _torch()
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


class LmcEngine:
    """One model resident on one CUDA device."""

    def __init__(self, packed: PackedModel, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("smol_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = capi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.packed = packed
        self.N = int(packed.desc.num_sites)
        self.F = int(packed.desc.num_features)
        self.row_stride = int(self.lib.lmc_row_stride(self.N))
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            capi.check(self.lib.lmc_model_create(C.byref(packed.desc), C.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.lmc_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _upload_stream(self):
        torch = _torch()
        if getattr(self, "_up_stream", None) is None:
            self._up_stream = torch.cuda.Stream(device=self.device)
        return self._up_stream

    def _h2d_int32(self, pinned):
        """page-locked int32 ``[W, N]`` -> device, on the upload stream: the copy overlaps whatever the launching
        stream is still running (the tail of the previous run); the launching stream waits for it."""
        torch = _torch()
        main = torch.cuda.current_stream(self.device)
        up = self._upload_stream()
        with torch.cuda.stream(up):
            src = pinned.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(up)
        src.record_stream(main)
        main.wait_event(ev)
        return src, ev

    @_on_device
    def upload_occupancy(self, occ_host):
        """occupancies ``[W, N]`` (host array, page-locked int32 tensor or CUDA tensor) -> int8
        ``[W, row_stride]`` device tensor.  The input is never modified."""
        torch = _torch()
        if isinstance(occ_host, torch.Tensor) and occ_host.is_cuda:
            # device-resident entry: no host round trip
            if occ_host.device != self.device:
                occ_host = occ_host.to(self.device)
            W = occ_host.shape[0]
            src = occ_host if (occ_host.dtype == torch.int32 and occ_host.is_contiguous()) \
                else occ_host.to(torch.int32).contiguous()
            dst = torch.empty((W, self.row_stride), dtype=torch.int8, device=self.device)
            capi.check(self.lib.lmc_cast_i32_to_i8(src.data_ptr(), dst.data_ptr(), W, self.N, self._stream()))
            return dst
        if isinstance(occ_host, torch.Tensor) and occ_host.device.type == "cpu" and occ_host.is_pinned() \
                and occ_host.dtype == torch.int32 and occ_host.is_contiguous():
            # the caller's buffer is already page-locked int32: copy straight from it (no staging pass)
            W = occ_host.shape[0]
            src, self._pin_evt = self._h2d_int32(occ_host)
            dst = torch.empty((W, self.row_stride), dtype=torch.int8, device=self.device)
            capi.check(self.lib.lmc_cast_i32_to_i8(src.data_ptr(), dst.data_ptr(), W, self.N, self._stream()))
            return dst
        if isinstance(occ_host, torch.Tensor):
            occ_host = occ_host.cpu().numpy()
        occ_host = np.asarray(occ_host)
        if occ_host.dtype.kind not in "iu":
            occ_host = occ_host.astype(np.int32)      # raises for non-numeric input like the reference
        W = occ_host.shape[0]
        key = tuple(occ_host.shape)
        if getattr(self, "_pin_key", None) != key:   # cached page-locked H2D staging buffer
            self._pin_buf = torch.empty(key, dtype=torch.int32, pin_memory=True)
            self._pin_key = key
        if getattr(self, "_pin_evt", None) is not None:
            self._pin_evt.synchronize()          # the previous async copy has consumed the buffer
        _copy_into(self._pin_buf.numpy(), occ_host)   # one pass: copy + int32 conversion
        src, self._pin_evt = self._h2d_int32(self._pin_buf)
        dst = torch.empty((W, self.row_stride), dtype=torch.int8, device=self.device)
        capi.check(self.lib.lmc_cast_i32_to_i8(src.data_ptr(), dst.data_ptr(), W, self.N,
                                               self._stream()))
        return dst

    @_on_device
    def occupancy_to_int32(self, occ_dev, rows: int, stride: int):
        """int8 device rows -> int32 ``[rows, N]`` device tensor."""
        torch = _torch()
        out = torch.empty((rows, self.N), dtype=torch.int32, device=self.device)
        capi.check(self.lib.lmc_cast_i8_to_i32(occ_dev.data_ptr(), out.data_ptr(), rows, self.N,
                                               stride, self._stream()))
        return out

    @_on_device
    def full_features(self, occ_dev, field=None):
        """Full evaluation of every walker's occupancy.  With a factorising Ewald matrix the Ewald term goes
        through the walkers' potential (one tiled product instead of a pair sum per walker): ``field`` (``[W, N]``
        float64, allocated and cached here when not given) is filled on the way and IS the potential cache of
        these occupancies afterwards (``self.last_field``)."""
        torch = _torch()
        W = occ_dev.shape[0]
        feat = torch.empty((W, self.F), dtype=torch.float64, device=self.device)
        enth = torch.empty((W,), dtype=torch.float64, device=self.device)
        self.last_field = None
        if self.model_info()[0]:
            if field is None:
                field = getattr(self, "_field_scratch", None)
                if field is None or field.shape[0] != W:
                    field = self._field_scratch = torch.empty((W, self.N), dtype=torch.float64, device=self.device)
            capi.check(self.lib.lmc_full_features_field(self.handle, occ_dev.data_ptr(), W, feat.data_ptr(),
                                                        enth.data_ptr(), field.data_ptr(), self._stream()))
            self.last_field = field
            return feat, enth
        capi.check(self.lib.lmc_full_features(self.handle, occ_dev.data_ptr(), W, feat.data_ptr(),
                                              enth.data_ptr(), self._stream()))
        return feat, enth

    @_on_device
    def delta_features(self, occ_dev, sites, codes):
        """sites/codes: int32 device tensors ``[W, k]``."""
        torch = _torch()
        W, k = sites.shape
        out = torch.empty((W, self.F), dtype=torch.float64, device=self.device)
        capi.check(self.lib.lmc_delta_features(self.handle, occ_dev.data_ptr(), W, sites.data_ptr(),
                                               codes.data_ptr(), k, out.data_ptr(), self._stream()))
        return out

    def model_info(self):
        """(Ewald matrix factorises, speculative tables built, table blob bytes, speculative records per site,
        bytes per walker of the environment-word workspace ``LmcRunConfig.spec_env_dev`` or 0, bits per site of the
        compact environment words or 0)"""
        if getattr(self, "_info", None) is None:
            info = (C.c_int32 * 6)()
            capi.check(self.lib.lmc_model_info(self.handle, info, 6))
            self._info = tuple(int(x) for x in info)
        return self._info

    @_on_device
    def ewald_field(self, occ_dev, out=None):
        """Ewald potential cache ``[W, N]`` (float64) of the walkers' occupancies, see ``lmc.h``."""
        torch = _torch()
        W = occ_dev.shape[0]
        if out is None:
            out = torch.empty((W, self.N), dtype=torch.float64, device=self.device)
        capi.check(self.lib.lmc_ewald_field(self.handle, occ_dev.data_ptr(), W, out.data_ptr(), self._stream()))
        return out

    @_on_device
    def distance_tables(self, processor):
        """device copies of a DistanceProcessor's target vector and orbit groups (cached per engine)"""
        torch = _torch()
        if getattr(self, "_dist_tabs", None) is None or self._dist_tabs[0] is not processor:
            target, tol, goff, gidx, gdiam = processor.distance_tables()
            up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)   # noqa: E731
            self._dist_tabs = (processor, dict(target=up(target, np.float64), tol=float(tol), goff=up(goff, np.int32),
                                               gidx=up(gidx if len(gidx) else np.zeros(1), np.int32),
                                               gdiam=up(gdiam if len(gdiam) else np.zeros(1), np.float64),
                                               ngrp=len(gdiam)))
        return self._dist_tabs[1]

    @_on_device
    def distance_init(self, processor, feat, enth=None):
        """extensive features ``feat [W, F]`` -> distance vectors in place; returns the vectors per supercell."""
        torch = _torch()
        d = self.distance_tables(processor)
        W = feat.shape[0]
        vec = torch.empty_like(feat)
        capi.check(self.lib.lmc_distance_init(self.handle, W, feat.data_ptr(), vec.data_ptr(),
                                              enth.data_ptr() if enth is not None else None, d["target"].data_ptr(),
                                              d["tol"], d["ngrp"], d["goff"].data_ptr(), d["gidx"].data_ptr(),
                                              d["gdiam"].data_ptr(), self._stream()))
        return vec

    @_on_device
    def run(self, cfg: capi.LmcRunConfig):
        capi.check(self.lib.lmc_run(self.handle, C.byref(cfg), self._stream()))

    def launch_count(self) -> int:
        return int(self.lib.lmc_launch_count())

    def c64_launch_count(self) -> int:
        """launches of the speculative kernel over compact environment words (process-wide)"""
        return int(self.lib.lmc_c64_launch_count())

    def env_launch_count(self) -> int:
        """launches of the speculative kernel's environment-word variants (process-wide)"""
        return int(self.lib.lmc_env_launch_count())
