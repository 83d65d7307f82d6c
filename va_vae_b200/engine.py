"""ctypes binding of libvavae_b200.so (C ABI: include/vavae_b200.h).

There is NO CPU fallback: if the shared library is missing, or no sm_100 device is present,
every entry point raises.  PyTorch is used only for device memory, streams and
torch.distributed plumbing; all arithmetic happens in the library's own CUDA kernels.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# VAVAE_B200_LIB: developer knob for same-box A/B runs of two builds of the library (tools/ab_gemms.py)
LIB_PATH = os.environ.get("VAVAE_B200_LIB") or os.path.join(_HERE, "libvavae_b200.so")

PREC_BF16, PREC_FP32 = 0, 1
ODE_EULER, ODE_HEUN = 0, 1
ABI_VERSION = 3
SAMPLE_SKIP_DISCARDED = 1
EPI_QKV, EPI_SWIGLU, EPI_GATED_RESIDUAL = 0, 1, 2

EXPORTS = [
    "vavae_abi_version", "vavae_last_error", "vavae_create", "vavae_destroy", "vavae_load_weight", "vavae_finalize",
    "vavae_forward", "vavae_forward_with_cfg", "vavae_sample_ode", "vavae_sample_ode_host", "vavae_launch_count",
    "vavae_device_bytes", "vavae_test_gemm_bf16", "vavae_test_attention_bf16", "vavae_time_kernel",
    "vavae_profile_begin", "vavae_profile_end", "vavae_set_gemm_two_cta", "vavae_sample_dopri5",
    "vavae_device_status", "vavae_forward_rows", "vavae_test_gemm_epilogue", "vavae_sde_update",
    "vavae_decoder_create", "vavae_decoder_destroy", "vavae_decoder_load_weight", "vavae_decoder_finalize",
    "vavae_decoder_decode", "vavae_decoder_launch_count", "vavae_decoder_device_bytes", "vavae_gemm_plan",
]


class VavaeConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "abi_version", "input_size", "patch_size", "in_channels", "hidden_size", "depth", "num_heads", "mlp_hidden",
        "num_classes", "label_rows", "learn_sigma", "use_qknorm", "use_swiglu", "use_rope", "use_rmsnorm", "wo_shift",
        "precision", "max_chunk_rows")]


_lib = None


class VavaeEpilogueTest(ctypes.Structure):
    """vavae_epilogue_test_t (include/vavae_b200.h)."""
    _fields_ = [("kind", ctypes.c_int32), ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
                ("T", ctypes.c_int32), ("A", ctypes.c_void_p), ("W", ctypes.c_void_p), ("bias", ctypes.c_void_p),
                ("in_part", ctypes.c_void_p), ("in_nparts", ctypes.c_int32), ("in_dim", ctypes.c_int32),
                ("sp", ctypes.c_void_p), ("sp_ld", ctypes.c_int32), ("H", ctypes.c_int32), ("hd", ctypes.c_int32),
                ("cos", ctypes.c_void_p), ("sin", ctypes.c_void_p), ("qn_w", ctypes.c_void_p), ("kn_w", ctypes.c_void_p),
                ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("vt", ctypes.c_void_p), ("out_bf16", ctypes.c_void_p),
                ("ldo", ctypes.c_int32), ("X", ctypes.c_void_p), ("mod", ctypes.c_void_p), ("mod_ld", ctypes.c_int32),
                ("gate_off", ctypes.c_int32), ("nx_w", ctypes.c_void_p), ("nx_scale_off", ctypes.c_int32),
                ("nx_A", ctypes.c_void_p), ("nx_part", ctypes.c_void_p), ("nparts_out", ctypes.c_int32),
                ("in_fine", ctypes.c_int32), ("fine_out", ctypes.c_int32), ("part_ld", ctypes.c_int32)]


class VavaeSdeUpdate(ctypes.Structure):
    """vavae_sde_update_t (include/vavae_b200.h)."""
    _fields_ = [("mode", ctypes.c_int32), ("x", ctypes.c_void_p), ("v", ctypes.c_void_p), ("w", ctypes.c_void_p),
                ("k1", ctypes.c_void_p), ("xh", ctypes.c_void_p), ("out", ctypes.c_void_p), ("out2", ctypes.c_void_p),
                ("n", ctypes.c_int64)] + [(n, ctypes.c_float) for n in (
                    "t", "dt", "diffusion", "var", "sqrt_2diffusion", "sqrt_dt", "c0", "c1")]


class VavaeDecoderConfig(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("z_channels", ctypes.c_int32), ("ch", ctypes.c_int32),
                ("num_levels", ctypes.c_int32), ("ch_mult", ctypes.c_int32 * 8), ("num_res_blocks", ctypes.c_int32),
                ("num_attn_res", ctypes.c_int32), ("attn_res", ctypes.c_int32 * 4), ("resolution", ctypes.c_int32),
                ("out_ch", ctypes.c_int32), ("precision", ctypes.c_int32), ("max_batch", ctypes.c_int32)]


def load_library() -> ctypes.CDLL:
    """Loads the in-tree shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m va_vae_b200.build` (nvcc, sm_100a). "
            "This engine has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    c_void, c_int, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.vavae_abi_version.restype = c_int
    lib.vavae_last_error.restype = ctypes.c_char_p
    lib.vavae_create.argtypes = [ctypes.POINTER(VavaeConfig), c_int, ctypes.POINTER(c_void)]
    lib.vavae_destroy.argtypes = [c_void]
    lib.vavae_load_weight.argtypes = [c_void, ctypes.c_char_p, c_void, ctypes.POINTER(ctypes.c_int64), c_int, c_int]
    lib.vavae_finalize.argtypes = [c_void, c_void]
    lib.vavae_forward.argtypes = [c_void, c_void, c_void, c_void, c_void, c_int, c_void]
    lib.vavae_forward_with_cfg.argtypes = [c_void, c_void, c_void, c_void, c_float, c_int, c_float, c_int, c_void, c_int,
                                           c_void]
    lib.vavae_sample_ode.argtypes = [c_void, c_void, c_void, ctypes.POINTER(c_float), c_int, c_int, c_int, c_float,
                                     c_int, c_float, c_int, c_void, c_void, c_int, c_int, c_void]
    lib.vavae_sample_ode_host.argtypes = [c_void, c_void, c_void, ctypes.POINTER(c_float), c_int, c_int, c_int, c_float,
                                          c_int, c_float, c_int, c_void, c_int, c_int, c_void]
    lib.vavae_device_status.argtypes = [c_void, c_void]
    lib.vavae_forward_rows.argtypes = [c_void]
    lib.vavae_forward_rows.restype = ctypes.c_uint64
    lib.vavae_test_gemm_epilogue.argtypes = [ctypes.POINTER(VavaeEpilogueTest), c_void]
    lib.vavae_sde_update.argtypes = [ctypes.POINTER(VavaeSdeUpdate), c_void]
    lib.vavae_sample_dopri5.argtypes = [c_void, c_void, c_void, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, c_int, c_float, c_int, c_float, c_int, c_void, c_int,
                                        ctypes.POINTER(ctypes.c_int64), c_void]
    lib.vavae_launch_count.argtypes = [c_void]
    lib.vavae_launch_count.restype = ctypes.c_uint64
    lib.vavae_device_bytes.argtypes = [c_void]
    lib.vavae_device_bytes.restype = ctypes.c_uint64
    lib.vavae_test_gemm_bf16.argtypes = [c_void, c_void, c_void, c_void, c_int, c_int, c_int, c_void]
    lib.vavae_test_attention_bf16.argtypes = [c_void, c_void, c_void, c_void, c_int, c_int, c_int, c_int, c_void]
    lib.vavae_time_kernel.argtypes = [c_void, ctypes.c_char_p, c_int, c_int, ctypes.POINTER(c_float),
                                      ctypes.POINTER(ctypes.c_double), c_void]
    lib.vavae_set_gemm_two_cta.argtypes = [c_int]
    lib.vavae_profile_begin.argtypes = [c_void, ctypes.c_char_p]
    lib.vavae_profile_end.argtypes = [c_void, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_double)]
    lib.vavae_decoder_create.argtypes = [ctypes.POINTER(VavaeDecoderConfig), c_int, ctypes.POINTER(c_void)]
    lib.vavae_decoder_destroy.argtypes = [c_void]
    lib.vavae_decoder_load_weight.argtypes = [c_void, ctypes.c_char_p, c_void, ctypes.POINTER(ctypes.c_int64), c_int,
                                              c_int]
    lib.vavae_decoder_finalize.argtypes = [c_void]
    lib.vavae_decoder_decode.argtypes = [c_void, c_void, c_int, c_void, c_void, c_void]
    lib.vavae_decoder_launch_count.argtypes = [c_void]
    lib.vavae_decoder_launch_count.restype = ctypes.c_uint64
    lib.vavae_decoder_device_bytes.argtypes = [c_void]
    lib.vavae_decoder_device_bytes.restype = ctypes.c_uint64
    if "VAVAE_GEMM_2CTA" in os.environ:      # 0 = never, 1 = main-loop-bound GEMMs only, 2 = wherever possible (default)
        lib.vavae_set_gemm_two_cta(int(os.environ["VAVAE_GEMM_2CTA"]))
    if lib.vavae_abi_version() != ABI_VERSION:
        raise RuntimeError("libvavae_b200.so ABI version mismatch; rebuild with python -m va_vae_b200.build --force")
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise RuntimeError("vavae_b200: " + load_library().vavae_last_error().decode())


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """One engine per (process, GPU).  Mirrors the lifetime of a LightningDiT module on a device."""

    def __init__(self, config: dict, state_dict: Dict[str, torch.Tensor], device, precision: str = "bf16",
                 max_chunk_rows: int = 0):
        lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("vavae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("vavae_b200 engines live on CUDA devices only")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.precision = precision
        cfg = VavaeConfig(abi_version=ABI_VERSION, precision={"bf16": PREC_BF16, "fp32": PREC_FP32}[precision],
                          max_chunk_rows=int(max_chunk_rows), **{k: int(v) for k, v in config.items()})
        self.config = dict(config)
        self._h = ctypes.c_void_p()
        _check(lib.vavae_create(ctypes.byref(cfg), dev_index, ctypes.byref(self._h)))
        self._lib = lib
        try:
            with torch.cuda.device(self.device):
                for name, tensor in state_dict.items():
                    if not torch.is_floating_point(tensor):
                        continue
                    t = tensor.detach()
                    on_dev = t.is_cuda and t.device == self.device and t.dtype == torch.float32 and t.is_contiguous()
                    t = t if on_dev else t.to(dtype=torch.float32).contiguous()
                    shape = (ctypes.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
                    _check(lib.vavae_load_weight(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), shape,
                                                 max(t.dim(), 1), 1 if t.is_cuda else 0))
                _check(lib.vavae_finalize(self._h, ctypes.c_void_p(_stream(self.device))))
        except Exception:
            lib.vavae_destroy(self._h)
            self._h = None
            raise
        self.C, self.img = int(config["in_channels"]), int(config["input_size"])

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vavae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _labels(self, y: torch.Tensor, batch: int) -> torch.Tensor:
        """Labels as a contiguous int64 device tensor of shape (batch,).  A host tensor is range-checked here with
        nn.Embedding's error (models/lightningdit.py:168); a device tensor is checked by the kernels themselves
        (clamped read + device flag, raised by check_device_status / the *_host entry points) so that no
        synchronisation is added to the call."""
        if y.dim() != 1 or y.shape[0] != batch:
            raise ValueError(f"y must have shape ({batch},), got {tuple(y.shape)}")
        if torch.is_floating_point(y) or y.dtype == torch.bool:
            raise RuntimeError("Expected tensor for argument #1 'indices' to have one of the following scalar types: "
                               "Long, Int")
        if not y.is_cuda and y.numel():
            lo, hi = int(y.min()), int(y.max())
            if lo < 0 or hi >= int(self.config["label_rows"]):
                raise IndexError("index out of range in self")
        return y.to(device=self.device, dtype=torch.int64).contiguous()

    def check_device_status(self):
        """Synchronises the current stream and raises IndexError if a kernel saw a label outside the embedding table
        since the last check (nn.Embedding raises at once; a device-resident y cannot be checked without a sync)."""
        with torch.cuda.device(self.device):
            rc = self._lib.vavae_device_status(self._h, ctypes.c_void_p(_stream(self.device)))
        if rc != 0:
            msg = self._lib.vavae_last_error().decode()
            raise IndexError(msg) if "index out of range" in msg else RuntimeError("vavae_b200: " + msg)

    # -- LightningDiT.forward (models/lightningdit.py:391-418)
    def forward(self, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        x, t = _f32(x, self.device), _f32(t, self.device)
        self._check_x(x)
        if t.shape != (x.shape[0],):
            raise ValueError(f"t must have shape ({x.shape[0]},), got {tuple(t.shape)}")
        y = self._labels(y, x.shape[0])
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _check(self._lib.vavae_forward(self._h, x.data_ptr(), t.data_ptr(), y.data_ptr(), out.data_ptr(),
                                           x.shape[0], ctypes.c_void_p(_stream(self.device))))
        return out

    # -- LightningDiT.forward_with_cfg (models/lightningdit.py:420-442)
    def forward_with_cfg(self, x, t, y, cfg_scale, cfg_interval=None, cfg_interval_start=None, cfg_channels=3):
        x, t = _f32(x, self.device), _f32(t, self.device)
        self._check_x(x)
        if x.shape[0] % 2:
            raise ValueError("forward_with_cfg expects 2n rows (conditional rows then null-label rows)")
        if t.shape != (x.shape[0],):
            raise ValueError(f"t must have shape ({x.shape[0]},), got {tuple(t.shape)}")
        y = self._labels(y, x.shape[0])
        interval = 1 if cfg_interval is True else 0
        start = float(cfg_interval_start) if (interval and cfg_interval_start is not None) else 0.0
        if interval and cfg_interval_start is None:
            raise TypeError("'<' not supported between instances of 'Tensor' and 'NoneType'")  # reference behaviour
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _check(self._lib.vavae_forward_with_cfg(self._h, x.data_ptr(), t.data_ptr(), y.data_ptr(), float(cfg_scale),
                                                    interval, start, int(cfg_channels), out.data_ptr(), x.shape[0],
                                                    ctypes.c_void_p(_stream(self.device))))
        return out

    # -- ode.sample over a fixed grid (transport/integrators.py:107-126)
    def sample_ode(self, z, y, t_grid, method="euler", use_cfg=True, cfg_scale=1.0, cfg_interval=None,
                   cfg_interval_start=None, cfg_channels=3, return_trace=False, skip_discarded=False):
        """skip_discarded: with the CFG-interval gate, do not run the unconditional forward of evaluations at
        t < cfg_interval_start (its output is discarded, models/lightningdit.py:436-439).  Rows [0, n) of the result
        are bit-identical to the literal mode; rows [n, 2n) are not the reference's (include/vavae_b200.h)."""
        z = _f32(z, self.device)
        self._check_x(z)
        if use_cfg and z.shape[0] % 2:
            raise ValueError("CFG sampling expects 2n rows (conditional rows then null-label rows)")
        y = self._labels(y, z.shape[0])
        grid = torch.as_tensor(t_grid, dtype=torch.float32).cpu().contiguous()
        meth = {"euler": ODE_EULER, "heun": ODE_HEUN, "heun2": ODE_HEUN}[method]
        n_evals = (len(grid) - 1) * (2 if meth == ODE_HEUN else 1)
        out = torch.empty_like(z)
        trace = torch.empty((n_evals,) + tuple(z.shape), device=self.device, dtype=torch.float32) if return_trace else None
        interval = 1 if cfg_interval is True else 0
        start = float(cfg_interval_start) if (interval and cfg_interval_start is not None) else 0.0
        gptr = ctypes.cast(grid.data_ptr(), ctypes.POINTER(ctypes.c_float))
        with torch.cuda.device(self.device):
            _check(self._lib.vavae_sample_ode(
                self._h, z.data_ptr(), y.data_ptr(), gptr, len(grid), meth, 1 if use_cfg else 0, float(cfg_scale),
                interval, start, int(cfg_channels), out.data_ptr(), trace.data_ptr() if return_trace else None,
                z.shape[0], SAMPLE_SKIP_DISCARDED if skip_discarded else 0, ctypes.c_void_p(_stream(self.device))))
        return (out, trace) if return_trace else out

    # -- adaptive dopri5 (torchdiffeq.odeint(method="dopri5") as reached from transport/integrators.py:118-125)
    def sample_dopri5(self, z, y, t0=0.0, t1=1.0, atol=1e-6, rtol=1e-3, use_cfg=True, cfg_scale=1.0, cfg_interval=None,
                      cfg_interval_start=None, cfg_channels=3, return_stats=False):
        z = _f32(z, self.device)
        self._check_x(z)
        if use_cfg and z.shape[0] % 2:
            raise ValueError("CFG sampling expects 2n rows (conditional rows then null-label rows)")
        y = self._labels(y, z.shape[0])
        out = torch.empty_like(z)
        interval = 1 if cfg_interval is True else 0
        start = float(cfg_interval_start) if (interval and cfg_interval_start is not None) else 0.0
        stats = (ctypes.c_int64 * 3)()
        with torch.cuda.device(self.device):
            _check(self._lib.vavae_sample_dopri5(
                self._h, z.data_ptr(), y.data_ptr(), float(t0), float(t1), float(atol), float(rtol), 1 if use_cfg else 0,
                float(cfg_scale), interval, start, int(cfg_channels), out.data_ptr(), z.shape[0], stats,
                ctypes.c_void_p(_stream(self.device))))
        if return_stats:
            return out, dict(accepted=stats[0], rejected=stats[1], evals=stats[2])
        return out

    def sample_ode_host(self, z_host, y_host, t_grid, out_host, method="euler", use_cfg=True, cfg_scale=1.0,
                        cfg_interval=None, cfg_interval_start=None, cfg_channels=3, skip_discarded=False):
        """End-to-end entry: HOST (ideally pinned) buffers in and out; H2D/D2H inside the call."""
        assert z_host.dtype == torch.float32 and not z_host.is_cuda and z_host.is_contiguous()
        assert y_host.dtype == torch.int64 and not y_host.is_cuda and y_host.is_contiguous()
        assert out_host.dtype == torch.float32 and not out_host.is_cuda and out_host.shape == z_host.shape
        self._check_x(z_host)
        if y_host.shape != (z_host.shape[0],):
            raise ValueError(f"y must have shape ({z_host.shape[0]},), got {tuple(y_host.shape)}")
        grid = torch.as_tensor(t_grid, dtype=torch.float32).cpu().contiguous()
        meth = {"euler": ODE_EULER, "heun": ODE_HEUN, "heun2": ODE_HEUN}[method]
        interval = 1 if cfg_interval is True else 0
        start = float(cfg_interval_start) if (interval and cfg_interval_start is not None) else 0.0
        gptr = ctypes.cast(grid.data_ptr(), ctypes.POINTER(ctypes.c_float))
        with torch.cuda.device(self.device):
            _check(self._lib.vavae_sample_ode_host(
                self._h, z_host.data_ptr(), y_host.data_ptr(), gptr, len(grid), meth, 1 if use_cfg else 0,
                float(cfg_scale), interval, start, int(cfg_channels), out_host.data_ptr(), z_host.shape[0],
                SAMPLE_SKIP_DISCARDED if skip_discarded else 0, ctypes.c_void_p(_stream(self.device))))
        return out_host

    def time_kernel(self, kernel: str, rows: int, iters: int = 10):
        """(ms per launch, algorithmic work) of one production kernel on layer-0 weights; CUDA events on the
        current stream."""
        ms, work = ctypes.c_float(), ctypes.c_double()
        with torch.cuda.device(self.device):
            _check(self._lib.vavae_time_kernel(self._h, kernel.encode(), rows, iters, ctypes.byref(ms),
                                               ctypes.byref(work), ctypes.c_void_p(_stream(self.device))))
        return ms.value, work.value

    def profile_begin(self, kernel):
        _check(self._lib.vavae_profile_begin(self._h, kernel.encode() if kernel else None))

    def profile_end(self):
        """(summed ms, launches, summed algorithmic work) of the kernel class selected by profile_begin."""
        ms, n, work = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
        _check(self._lib.vavae_profile_end(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(work)))
        return ms.value, n.value, work.value

    @property
    def launch_count(self) -> int:
        return int(self._lib.vavae_launch_count(self._h))

    @property
    def forward_rows(self) -> int:
        """Model rows pushed through the layer stack since creation (forwards actually executed)."""
        return int(self._lib.vavae_forward_rows(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.vavae_device_bytes(self._h))

    def _check_x(self, x):
        if x.dim() != 4 or x.shape[1] != self.C or x.shape[2] != self.img or x.shape[3] != self.img:
            raise ValueError(f"expected x of shape (N, {self.C}, {self.img}, {self.img}), got {tuple(x.shape)}")


def set_gemm_two_cta(on: bool):
    """Process-wide switch between the single-CTA and the CTA-pair (cta_group::2) GEMM kernels."""
    load_library().vavae_set_gemm_two_cta(int(on) if not isinstance(on, bool) else (2 if on else 0))


def test_gemm_bf16(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    lib = load_library()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        _check(lib.vavae_test_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None,
                                        out.data_ptr(), M, N, K, ctypes.c_void_p(_stream(A.device))))
    return out


def attention_hdv(hd: int) -> int:
    """Rows of the transposed-V operand: head_dim + the ones row, rounded up to 16."""
    return (hd + 1 + 15) // 16 * 16


def test_attention_bf16(q, k, vt, hd: int) -> torch.Tensor:
    """q,k: (B,H,T,hdp) bf16; vt: (B,H,hdv,T) bf16 (row hd = ones) -> (B*T, H*hd) bf16."""
    lib = load_library()
    B, H, T, _ = q.shape
    out = torch.empty(B * T, H * hd, device=q.device, dtype=torch.bfloat16)
    with torch.cuda.device(q.device):
        _check(lib.vavae_test_attention_bf16(q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr(), B, H, T, hd,
                                             ctypes.c_void_p(_stream(q.device))))
    return out


def test_gemm_epilogue(**fields) -> int:
    """One launch of the production tcgen05 GEMM with a fused epilogue on caller-provided CUDA tensors
    (vavae_test_gemm_epilogue).  Tensor-valued fields are passed as torch tensors; returns nparts_out."""
    lib = load_library()
    t = VavaeEpilogueTest()
    dev = None
    keep = []
    for k, v in fields.items():
        if isinstance(v, torch.Tensor):
            assert v.is_cuda and v.is_contiguous(), k
            dev = v.device
            keep.append(v)
            setattr(t, k, v.data_ptr())
        else:
            setattr(t, k, int(v))
    with torch.cuda.device(dev):
        _check(lib.vavae_test_gemm_epilogue(ctypes.byref(t), ctypes.c_void_p(_stream(dev))))
    test_gemm_epilogue.fine_out = int(t.fine_out)   # kind of partial-sum slots a producer launch wrote (1 = 32-column)
    return int(t.nparts_out)


def gemm_plan(kind: int, M: int, N: int, K: int, T: int, H: int = 0, hd: int = 0, fused: bool = True) -> dict:
    """Host-only (vavae_gemm_plan): column-tile widths, epilogue variant and partial-sum layout the library picks for one
    layer GEMM of M token rows.  No GPU needed."""
    lib = load_library()
    out = (ctypes.c_int32 * 6)()
    _check(lib.vavae_gemm_plan(int(kind), int(M), int(N), int(K), int(T), int(H), int(hd), int(bool(fused)), out))
    bn, bn_last, tma, warps, slots, coarse = (int(v) for v in out)
    return dict(bn=bn, bn_last=bn_last, tma=bool(tma), warps=warps, slots=slots, coarse=bool(coarse))


def sde_update(mode: int, x: torch.Tensor, out: torch.Tensor, v=None, w=None, k1=None, xh=None, out2=None, **scalars):
    """One fused pass of the SDE sampler's elementwise update (vavae_sde_update; modes in include/vavae_b200.h)."""
    lib = load_library()
    u = VavaeSdeUpdate(mode=int(mode), n=x.numel())
    for name, ten in (("x", x), ("v", v), ("w", w), ("k1", k1), ("xh", xh), ("out", out), ("out2", out2)):
        if ten is not None:
            if not (ten.is_cuda and ten.is_contiguous() and ten.dtype == torch.float32 and ten.numel() == x.numel()):
                raise ValueError(f"sde_update: {name} must be a contiguous fp32 CUDA tensor of x's size")
            setattr(u, name, ten.data_ptr())
    for k, val in scalars.items():
        setattr(u, k, float(val))
    with torch.cuda.device(x.device):
        _check(lib.vavae_sde_update(ctypes.byref(u), ctypes.c_void_p(_stream(x.device))))
    return out
