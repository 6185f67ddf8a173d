"""Fused position-attention: host side above the C ABI.

``position_attention`` is the one functional entry point every ``posatt*`` module of
``position_induced_transformer_b200.pit`` calls.  It replaces the reference's
``dist2att`` + ``convolution`` pair (pit.py:46-57, 133-144, 190-200, 247-258) and, for the
self stage, the ``torch.cat`` of pit.py:44.  Nothing of size N x M is materialised.

``torch.compile``: the entry points are marked ``torch.compiler.disable`` -- the scripts' ``torch.compile(model)``
(train_darcy.py:112) compiles the surrounding MLPs and runs these ops as opaque eager calls (a graph break each);
the kernels never depend on a tracing compiler.

Autograd: gradients flow to ``values`` and to the per-head ``scale`` (and from there to
``lmda`` through ordinary torch ops) and -- when a mesh requires grad, which no call site of the
reference does -- to the mesh coordinates (pit_posatt_backward_coords).  No gradient
flows through the quantile.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _cabi

_VARIANTS = ("euclid", "periodic1d", "periodic2d")


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _check_tensor(name: str, t: torch.Tensor, device: torch.device) -> None:
    _require(t.is_cuda, f"{name} must be a CUDA tensor (position-attention has no CPU path), got {t.device}")
    _require(t.device == device, f"{name} is on {t.device}, expected {device}")
    _require(t.dtype == torch.float32, f"{name} must be float32, got {t.dtype}")


class _Stage:
    """Validated description of one call: the C problem struct plus python-side shape facts."""

    def __init__(self, mesh_out, mesh_in, values, n_head: int, variant: str):
        _require(variant in _VARIANTS, f"unknown variant {variant!r}")
        _require(values.dim() == 3, f"values must be (batch, L_in, dim), got {tuple(values.shape)}")
        device = values.device
        for name, t in (("values", values), ("mesh_out", mesh_out), ("mesh_in", mesh_in)):
            _check_tensor(name, t, device)
        _require(mesh_out.dim() == mesh_in.dim() and mesh_in.dim() in (2, 3),
                 f"meshes must both be (L, sd) or (batch, L, sd), got {tuple(mesh_out.shape)} / {tuple(mesh_in.shape)}")
        self.batched = mesh_in.dim() == 3
        self.B, self.M, self.D = values.shape
        self.N = mesh_out.shape[-2]
        self.sd = mesh_in.shape[-1]
        self.H = int(n_head)
        _require(mesh_out.shape[-1] == self.sd, "mesh_out / mesh_in disagree on space_dim")
        _require(self.sd in (1, 2), f"space_dim must be 1 or 2, got {self.sd}")
        _require(mesh_in.shape[-2] == self.M, f"mesh_in has {mesh_in.shape[-2]} points but values has {self.M}")
        if self.batched:
            _require(mesh_in.shape[0] == self.B and mesh_out.shape[0] == self.B, "per-sample meshes must match the batch")
        self.variant = variant
        self.device = device
        self.rows = (self.B if self.batched else 1) * self.N
        self.problem = _cabi.Problem(_cabi.VARIANT_CODE[variant], self.sd, int(self.batched), self.B, self.H,
                                     self.N, self.M, self.D)

    def stat_shape(self) -> Tuple[int, ...]:
        return (self.B, self.N) if self.batched else (self.N,)

    def rowsum_shape(self) -> Tuple[int, ...]:
        return (self.B, self.H, self.N) if self.batched else (self.H, self.N)


def wrap_period(mesh_in: torch.Tensor, variant: str) -> Optional[torch.Tensor]:
    """Wrap length of the periodic variants as a 1-element device tensor (no host sync).

    periodic1d: |x1 - x0| * M (pit.py:191-192); periodic2d: (max x - min x) / (res - 1) * res with
    res = int(sqrt(M)) (pit.py:248-250).  Same torch ops, same rounding.
    """
    if variant == "periodic1d":
        return (torch.abs(mesh_in[1, 0] - mesh_in[0, 0]) * mesh_in.shape[0]).reshape(1)
    if variant == "periodic2d":
        res = int(mesh_in.shape[0] ** 0.5)
        return ((torch.max(mesh_in[:, 0]) - torch.min(mesh_in[:, 0])) / (res - 1) * res).reshape(1)
    return None


_DENSE_PRECISIONS = {"fp32": 0, "tf32": 1, "bf16": 2}


def set_dense_precision(mode: str) -> None:
    """Operand precision of the global (locality >= 1) stages that run on the dense tcgen05 kernel; accumulation stays fp32.

    "fp32" (default)  3xTF32 products, parity with the fp32 reference to ~1e-6;
    "tf32"            single TF32 products -- what torch runs the reference's einsum as under its own 'high' setting (pit.py:2);
                      stated bound 1e-3;
    "bf16"            operands rounded to BF16; stated bound 5e-3.
    The fused processor kernel (small latent grids) and every masked stage are not affected: they stay fp32-exact."""
    _require(mode in _DENSE_PRECISIONS, f"dense precision must be one of {sorted(_DENSE_PRECISIONS)}, got {mode!r}")
    _cabi.check(_cabi.lib.pit_set_dense_precision(_DENSE_PRECISIONS[mode]), "pit_set_dense_precision")


def get_dense_precision() -> str:
    return {v: k for k, v in _DENSE_PRECISIONS.items()}[int(_cabi.lib.pit_get_dense_precision())]


class KernelTimer:
    """Optional per-call CUDA-event timing of the C-ABI launches (used by bench.py for the roofline figures).

    Events are recorded on the stream the kernels are enqueued on; nothing synchronises until ``summary()``.
    """

    def __init__(self):
        self.records = []          # (key, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for key, a, b in self.records:
            tot, n = agg.get(key, (0.0, 0))
            agg[key] = (tot + a.elapsed_time(b), n + 1)
        return {k: {"ms_total": tot, "calls": n, "ms_avg": tot / n} for k, (tot, n) in agg.items()}


_TIMER: Optional[KernelTimer] = None


def set_kernel_timer(timer: Optional[KernelTimer]) -> None:
    global _TIMER
    _TIMER = timer


class _timed:
    def __init__(self, tag, st, concat):
        self.key = (tag, st.variant, int(st.batched), st.B, st.H, st.N, st.M, st.D, st.sd, int(concat))
        self.dev = st.device

    def __enter__(self):
        if _TIMER is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record(torch.cuda.current_stream(self.dev))

    def __exit__(self, *exc):
        if _TIMER is not None and exc[0] is None:
            end = torch.cuda.Event(enable_timing=True)
            end.record(torch.cuda.current_stream(self.dev))
            _TIMER.records.append((self.key, self.start, end))
        return False


class _MeshEntry:
    """What the cache keeps for one pair of shared meshes: contiguous copies, wrap period, row statistics and -- built on
    first use by a decoder -- the tile plan of the fused decoder tail."""

    __slots__ = ("mesh_out", "mesh_in", "period", "stats", "tail_plan", "column_plan", "keep_alive")

    def __init__(self, mesh_out, mesh_in, period, stats, keep_alive):
        self.mesh_out, self.mesh_in, self.period, self.stats = mesh_out, mesh_in, period, stats
        self.tail_plan = None        # decoder side: tiles of mesh_out rows (built on first use by the fused decoder tail)
        self.column_plan = None      # encoder side: tiles of mesh_in columns (built on first use by a wide masked stage)
        self.keep_alive = keep_alive


class _MeshCache:
    """Per-mesh-pair constants of SHARED meshes, reused across steps.

    For a pair of shared meshes the contiguous copies, the wrap period, the order statistics v_min / v_lo / v_hi and the
    decoder's tile plan depend on the meshes and the locality only -- not on lmda -- and the scripts pass the same mesh
    tensors at every step (train_darcy.py:88-96, 128; note that they are column-major views of a transposed numpy array,
    so even `.contiguous()` is a copy), while the reference re-sorts every row each time (pit.py:136).

    An entry is keyed by the storage address, shape, strides and autograd version counter of both incoming tensors and
    keeps a strong reference to them, so the address cannot be recycled while the entry lives and any in-place write
    through torch (which bumps the version) misses.  CAVEAT: writes that do NOT bump the version counter -- `mesh.data.copy_`,
    a CUDA-graph replay that fills a static mesh buffer, an external kernel, NCCL -- are invisible here; code that rewrites
    a shared mesh that way must call `mesh_cache.clear()` (or set `mesh_cache.enabled = False`).  Tensors without a version
    counter (created under `torch.inference_mode()`) and per-sample meshes (B, L, sd), which change with every batch, are
    never cached: their statistics are recomputed by every call.  Least recently used entries are evicted first.
    """

    def __init__(self, capacity: int = 16):
        self.capacity = capacity
        self.entries = {}
        self.hits = self.misses = 0
        self.enabled = True

    @staticmethod
    def _sig(t: torch.Tensor):
        try:
            version = t._version
        except RuntimeError:          # inference tensors do not track a version counter
            return None
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), version, t.device.index)

    def key(self, mesh_out, mesh_in, variant, locality):
        a, b = self._sig(mesh_out), self._sig(mesh_in)
        return None if a is None or b is None else (a, b, variant, float(locality))

    def get(self, key):
        hit = self.entries.get(key)
        if hit is None:
            self.misses += 1
            return None
        self.hits += 1
        self.entries[key] = self.entries.pop(key)      # most recently used last
        return hit

    def put(self, key, entry):
        while len(self.entries) >= self.capacity:
            self.entries.pop(next(iter(self.entries)))
        self.entries[key] = entry

    def clear(self):
        self.entries.clear()


mesh_cache = _MeshCache()
rowstat_cache = mesh_cache          # historical name


_ZEROS = {}


def _zeros(shape, device) -> torch.Tensor:
    """Shared read-only zero tensors (row minima of global self stages): one fill per shape, not one per call."""
    key = (tuple(shape), device)
    z = _ZEROS.get(key)
    if z is None:
        if torch.cuda.is_current_stream_capturing():
            return torch.zeros(shape, dtype=torch.float32, device=device)
        z = _ZEROS[key] = torch.zeros(shape, dtype=torch.float32, device=device)
    return z


def row_statistics(st: _Stage, mesh_out, mesh_in, period, locality: float):
    """(v_min, v_lo, v_hi, w, masked): order statistics replacing torch.quantile's row sort (uncached)."""
    _require(0.0 <= locality <= 1.0, f"quantile() q values must be in the range [0, 1], got locality={locality}")
    masked = locality < 1.0
    k_lo, k_hi, w = _cabi.quantile_ranks(locality, st.M) if masked else (0, 0, 0.0)
    if not masked and mesh_out.data_ptr() == mesh_in.data_ptr() and mesh_out.shape == mesh_in.shape:
        # global self stage: the row minimum is d2(i, i) = 0 exactly, no statistics needed
        zeros = _zeros(st.stat_shape(), st.device)
        return zeros, zeros, zeros, w, masked
    stats = torch.empty((3,) + st.stat_shape(), dtype=torch.float32, device=st.device)
    v_min = stats[0]
    with _timed("rowstat", st, False):
        if st.batched and st.M <= 1024 and k_hi <= _NBR_MAX_RANK and st.N >= 4 * st.M and st.D % 4 == 0 and st.D <= 256 and st.H <= 2:
            # per-sample meshes: the sweep also leaves every row's neighbour list (columns with d2 <= v_hi), so that the masked
            # kernels of the stage never sweep the N x M pairs again (csrc/sample_tile.cuh)
            rows = st.rows
            lists = (torch.empty((rows, 32), dtype=torch.int16, device=st.device), torch.empty((rows, 32), dtype=torch.float32, device=st.device),
                     torch.empty(rows, dtype=torch.int32, device=st.device))
            _cabi.check(_cabi.lib.pit_rowstat_lists(C.byref(st.problem), mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), k_lo, k_hi,
                                                    stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), lists[0].data_ptr(),
                                                    lists[1].data_ptr(), lists[2].data_ptr(), _stream(st.device)), "pit_rowstat_lists")
            v_min.pit_neighbour_lists = lists          # rides along with the statistics (see _rowstat_struct)
        else:
            _cabi.check(_cabi.lib.pit_rowstat(C.byref(st.problem), mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period),
                                              k_lo, k_hi, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(),
                                              _stream(st.device)), "pit_rowstat")
    return v_min, stats[1], stats[2], w, masked


def meshes_need_grad(mesh_out, mesh_in) -> bool:
    return torch.is_grad_enabled() and (mesh_out.requires_grad or mesh_in.requires_grad)


def _meshes_are_constants(mesh_out, mesh_in) -> None:
    # dist2att is differentiable w.r.t. the coordinates in the reference (pit.py:47); only `position_attention` produces that
    # gradient (pit_posatt_backward_coords) -- the fused decoder tail and processor refuse loudly instead of returning a silent
    # zero (pit.decoder / pit.processor route around them when a mesh requires grad)
    if meshes_need_grad(mesh_out, mesh_in):
        raise RuntimeError("fused decoder tail / processor: gradients w.r.t. mesh coordinates are only produced by position_attention")


def prepare_meshes(mesh_out, mesh_in, values, n_head: int, variant: str, locality: float, coords_grad: bool = False):
    """(mesh_out, mesh_in, stage, period, stats, entry): contiguous meshes, the validated stage, the wrap period and the
    row statistics; `entry` is the cache entry for shared meshes (None when nothing is cached)."""
    if not coords_grad:
        _meshes_are_constants(mesh_out, mesh_in)
    # a mesh that is being trained changes between steps -- possibly through a kernel that does not bump its version counter
    # (fused_optimizer.FusedAllReduceAdam writes the flat parameter buffer directly) -- so its statistics are never cached
    learnable = mesh_out.requires_grad or mesh_in.requires_grad
    mesh_out, mesh_in = mesh_out.detach(), mesh_in.detach()
    capturing = mesh_in.is_cuda and torch.cuda.is_current_stream_capturing()
    cacheable = mesh_cache.enabled and mesh_in.is_cuda and mesh_in.dim() == 2 and not learnable
    key = mesh_cache.key(mesh_out, mesh_in, variant, locality) if cacheable else None
    hit = mesh_cache.get(key) if key is not None else None
    if hit is not None:
        return hit.mesh_out, hit.mesh_in, _Stage(hit.mesh_out, hit.mesh_in, values, n_head, variant), hit.period, hit.stats, hit
    mo, mi = mesh_out.contiguous(), mesh_in.contiguous()
    st = _Stage(mo, mi, values, n_head, variant)
    with torch.cuda.device(st.device):
        period = wrap_period(mi, variant)
        stats = row_statistics(st, mo, mi, period, float(locality))
    entry = None
    # entries created while a CUDA graph is being captured would point into the graph's private pool: do not keep them
    if key is not None and not capturing:
        entry = _MeshEntry(mo, mi, period, stats, (mesh_out, mesh_in))
        mesh_cache.put(key, entry)
    return mo, mi, st, period, stats, entry


class TailPlan:
    """Tile plan of a decoder stage (include/pit_posatt.h, pit_tail_plan_t): rows sorted by candidate set, 32 per tile."""

    def __init__(self, rec, tile_off, tile_cnt, cand, d2):
        self.rec, self.tile_off, self.tile_cnt, self.cand, self.d2 = rec, tile_off, tile_cnt, cand, d2
        self.n_tiles = tile_cnt.numel()
        self.n_cand = cand.numel()          # candidate entries including the padding of every list to a multiple of 8
        self.struct = _cabi.TailPlan(rec.data_ptr(), tile_off.data_ptr(), tile_cnt.data_ptr(), cand.data_ptr(), d2.data_ptr(), self.n_tiles)


_TAIL_PLAN = True
_TAIL_PLAN_MAX_MEAN = 16      # candidates per tile (mean) above which the plan is not used


def use_tail_plan(enabled: bool) -> None:
    """Switch the cached tile plan of the fused decoder tail on or off (off: the kernels scan the latent mesh per launch)."""
    global _TAIL_PLAN
    _TAIL_PLAN = bool(enabled)


def build_tail_plan(st: _Stage, mesh_out, mesh_in, period, stats, side: int = _cabi.PLAN_ROWS) -> TailPlan:
    """Runs the two construction stages of the C ABI; one 4-byte device->host read in between sizes the candidate array.
    side = PLAN_ROWS tiles the mesh_out rows (decoder stages), PLAN_COLUMNS the mesh_in columns (local encoder stages)."""
    v_min, v_lo, v_hi, w, masked = stats
    rs = _rowstat_struct(v_min, v_lo, v_hi, w, masked)
    dev = st.device
    with torch.cuda.device(dev):
        n_tiles = ((st.N if side == _cabi.PLAN_ROWS else st.M) + 31) // 32
        ws_bytes = int(_cabi.lib.pit_tail_plan_workspace_bytes(C.byref(st.problem), side))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        tile_off = torch.empty(n_tiles + 1, dtype=torch.int32, device=dev)
        tile_cnt = torch.empty(n_tiles, dtype=torch.int32, device=dev)
        _cabi.check(_cabi.lib.pit_tail_plan_rows(C.byref(st.problem), side, mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), C.byref(rs),
                                                 tile_off.data_ptr(), tile_cnt.data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev)),
                    "pit_tail_plan_rows")
        total = int(tile_off[-1].item())
        rec = torch.empty((n_tiles * 32, 4), dtype=torch.float32, device=dev)
        cand = torch.empty(max(total, 1), dtype=torch.int16, device=dev)
        d2 = torch.empty((max(total, 1), 32), dtype=torch.float32, device=dev)
        _cabi.check(_cabi.lib.pit_tail_plan_fill(C.byref(st.problem), side, mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), C.byref(rs),
                                                 tile_off.data_ptr(), rec.data_ptr(), cand.data_ptr(), d2.data_ptr(), ws.data_ptr(),
                                                 ws_bytes, _stream(dev)), "pit_tail_plan_fill")
    return TailPlan(rec, tile_off, tile_cnt, cand, d2)


def tail_plan_for(entry: Optional[_MeshEntry], st: _Stage, hidden: int) -> Optional[TailPlan]:
    """The cached plan of a shared-mesh decoder stage, built on first use (never while a CUDA graph is being captured)."""
    if not _TAIL_PLAN or entry is None or st.batched or st.M > 1024 or hidden not in (32, 64):
        return None
    if entry.tail_plan is None and not torch.cuda.is_current_stream_capturing():
        plan = build_tail_plan(st, entry.mesh_out, entry.mesh_in, entry.period, entry.stats)
        # small or incoherent meshes do not pack (a 43 x 43 grid has more distinct candidate sets than 32-row tiles): with more
        # than one 16-candidate block per tile on average the per-launch scan of the latent mesh is the better kernel
        entry.tail_plan = plan if plan.n_cand <= _TAIL_PLAN_MAX_MEAN * plan.n_tiles else False
    return entry.tail_plan or None


_NBR_MAX_RANK = 23      # a row's candidates (rank_hi + 1, plus ties at the cut) must fit a 32-entry neighbour list


def _rowstat_struct(v_min, v_lo, v_hi, w: float, masked: bool, rank_hi: int = 0, lists=None) -> _cabi.RowStat:
    lists = lists if lists is not None else getattr(v_min, "pit_neighbour_lists", None)
    nbr = (None, None, None) if lists is None else tuple(t.data_ptr() for t in lists)
    return _cabi.RowStat(v_min.data_ptr(), v_lo.data_ptr() if masked else None, v_hi.data_ptr() if masked else None,
                         w, int(masked), int(rank_hi), *nbr)


def column_plan_for(entry: Optional[_MeshEntry], st: _Stage, masked: bool) -> Optional[TailPlan]:
    """The cached encoder-side plan of a shared-mesh stage with few rows and a huge column set (the conditions under which
    the C ABI takes the wide path: masked, H <= 2, B*D <= 32, M >= 4096 and M >= 8 N), built on first use."""
    if (not _TAIL_PLAN or entry is None or st.batched or not masked or st.H > 2 or st.B * st.D > 32 or st.N > 1024
            or st.M < 4096 or st.M < 8 * st.N or st.B * st.M * st.D >= 2 ** 31):
        return None
    if entry.column_plan is None and not torch.cuda.is_current_stream_capturing():
        entry.column_plan = build_tail_plan(st, entry.mesh_out, entry.mesh_in, entry.period, entry.stats, _cabi.PLAN_COLUMNS)
    return entry.column_plan


class _PositionAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, scale, mesh_out, mesh_in, n_head, locality, variant, self_concat):
        values = values.contiguous()
        _require(values.is_cuda, f"values must be a CUDA tensor (position-attention has no CPU path), got {values.device}")
        scale_shape = scale.shape
        scale = scale.reshape(-1).contiguous()
        mesh_out, mesh_in, st, period, (v_min, v_lo, v_hi, w, masked), entry = prepare_meshes(
            mesh_out, mesh_in, values, n_head, variant, float(locality), coords_grad=True)
        plan = None if self_concat else column_plan_for(entry, st, masked)
        _check_tensor("scale", scale, st.device)
        _require(scale.numel() == st.H, f"scale must have n_head={st.H} entries, got {scale.numel()}")
        if self_concat:
            _require(st.N == st.M, "self stage needs mesh_out and mesh_in of equal length")
        with torch.cuda.device(st.device):
            width = (1 + st.H) * st.D if self_concat else st.H * st.D
            col_off = st.D if self_concat else 0
            out = torch.empty((st.B, st.N, width), dtype=torch.float32, device=st.device)
            rowsum = torch.empty(st.rowsum_shape(), dtype=torch.float32, device=st.device)
            ws_bytes = int(_cabi.lib.pit_workspace_bytes(C.byref(st.problem)))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=st.device)
            rank_hi = _cabi.quantile_ranks(float(locality), st.M)[1] if masked else 0
            lists = getattr(v_min, "pit_neighbour_lists", None)
            rs = _rowstat_struct(v_min, v_lo, v_hi, w, masked, rank_hi, lists)
            with _timed("fwd", st, self_concat):
                _cabi.check(_cabi.lib.pit_posatt_forward(
                    C.byref(st.problem), mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), values.data_ptr(),
                    scale.data_ptr(), C.byref(rs), out.data_ptr(), width, col_off, int(self_concat), rowsum.data_ptr(),
                    ws.data_ptr(), ws_bytes, C.byref(plan.struct) if plan is not None else None, _stream(st.device)), "pit_posatt_forward")
        ctx.save_for_backward(values, scale, mesh_out, mesh_in, period if period is not None else scale.new_empty(0),
                              v_min, v_lo, v_hi, rowsum)
        ctx.meta = (n_head, variant, self_concat, w, masked, scale_shape, rank_hi)
        ctx.lists = lists        # neighbour lists of per-sample meshes (int16 / float32 / int32 tensors), reused by the backward
        ctx.plan = plan          # keeps the plan's tensors alive until the backward has run
        return out

    @staticmethod
    def backward(ctx, d_out):
        values, scale, mesh_out, mesh_in, period, v_min, v_lo, v_hi, rowsum = ctx.saved_tensors
        n_head, variant, self_concat, w, masked, scale_shape, rank_hi = ctx.meta
        period = period if period.numel() else None
        st = _Stage(mesh_out, mesh_in, values, n_head, variant)
        d_out = d_out.contiguous()
        need_values, need_scale = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_values = d_scale = None
        with torch.cuda.device(st.device):
            if need_values:
                d_values = torch.empty_like(values)
            if need_scale:
                d_scale = torch.empty(st.H, dtype=torch.float32, device=st.device)
            width = d_out.shape[-1]
            col_off = st.D if self_concat else 0
            ws_bytes = int(_cabi.lib.pit_workspace_bytes(C.byref(st.problem)))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=st.device)
            rs = _rowstat_struct(v_min, v_lo, v_hi, w, masked, rank_hi, ctx.lists)
            if need_values or need_scale:
                with _timed("bwd", st, self_concat):
                    _cabi.check(_cabi.lib.pit_posatt_backward(
                        C.byref(st.problem), mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), values.data_ptr(),
                        scale.data_ptr(), C.byref(rs), rowsum.data_ptr(), d_out.data_ptr(), width, col_off,
                        int(self_concat), _ptr(d_values), _ptr(d_scale), ws.data_ptr(), ws_bytes,
                        C.byref(ctx.plan.struct) if ctx.plan is not None else None, _stream(st.device)), "pit_posatt_backward")
            d_mesh_out = d_mesh_in = None
            if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
                # dX (pit_posatt_backward_coords): what autograd gives the reference when a mesh requires grad
                d_mesh_out, d_mesh_in = torch.empty_like(mesh_out), torch.empty_like(mesh_in)
                d_period = torch.empty(1, dtype=torch.float32, device=st.device) if period is not None else None
                with _timed("bwd_coords", st, self_concat):
                    _cabi.check(_cabi.lib.pit_posatt_backward_coords(
                        C.byref(st.problem), mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), values.data_ptr(),
                        scale.data_ptr(), C.byref(rs), rowsum.data_ptr(), d_out.data_ptr(), width, col_off,
                        d_mesh_out.data_ptr(), d_mesh_in.data_ptr(), _ptr(d_period), _stream(st.device)), "pit_posatt_backward_coords")
                if d_period is not None:        # the wrap length is a function of mesh_in (pit.py:191-192, 248-250): chain through it
                    with torch.enable_grad():
                        leaf = mesh_in.detach().requires_grad_(True)
                        (through_period,) = torch.autograd.grad(wrap_period(leaf, variant), leaf, d_period)
                    d_mesh_in = d_mesh_in + through_period
                d_mesh_out = d_mesh_out if ctx.needs_input_grad[2] else None
                d_mesh_in = d_mesh_in if ctx.needs_input_grad[3] else None
        if need_scale:
            d_scale = d_scale.reshape(scale_shape)
        return d_values, d_scale, d_mesh_out, d_mesh_in, None, None, None, None


@torch.compiler.disable
def position_attention(mesh_out: torch.Tensor, mesh_in: torch.Tensor, values: torch.Tensor, scale: torch.Tensor,
                       locality: float, variant: str = "euclid", self_concat: bool = False) -> torch.Tensor:
    """out[b, n, h*D + d] = sum_j softmax_j(-s_h d2(n, j) | quantile mask)[j] * values[b, j, d].

    mesh_out (N, sd) / mesh_in (M, sd) shared by the batch, or (B, N, sd) / (B, M, sd) per sample;
    values (B, M, D); scale: H positive per-head scales (any shape with H elements).
    Returns (B, N, H*D), or (B, N, (1+H)*D) = cat(values, out) when ``self_concat``.
    """
    return _PositionAttention.apply(values, scale, mesh_out, mesh_in, scale.numel(), float(locality), variant,
                                    bool(self_concat))


# ----------------------------------------------------------------------------------------------
# per-head scale map s = tan(c (1 + sin lmda)) (pit.py:48): one launch forward, one backward
# ----------------------------------------------------------------------------------------------
class _HeadScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lmda):
        _require(lmda.is_cuda and lmda.dtype == torch.float32, "head_scale: lmda must be a float32 CUDA tensor")
        flat = lmda.detach().reshape(-1).contiguous()
        scale = torch.empty_like(flat)
        with torch.cuda.device(lmda.device):
            _cabi.check(_cabi.lib.pit_head_scale_forward(flat.data_ptr(), scale.data_ptr(), flat.numel(), _stream(lmda.device)),
                        "pit_head_scale_forward")
        ctx.save_for_backward(flat, scale)
        ctx.lmda_shape = lmda.shape
        return scale.view(lmda.shape)

    @staticmethod
    def backward(ctx, d_scale):
        flat, scale = ctx.saved_tensors
        d_scale = d_scale.reshape(-1).contiguous()
        d_lmda = torch.empty_like(flat)
        with torch.cuda.device(flat.device):
            _cabi.check(_cabi.lib.pit_head_scale_backward(flat.data_ptr(), scale.data_ptr(), d_scale.data_ptr(), d_lmda.data_ptr(),
                                                          flat.numel(), _stream(flat.device)), "pit_head_scale_backward")
        return d_lmda.view(ctx.lmda_shape)


@torch.compiler.disable
def head_scale_cuda(lmda: torch.Tensor) -> torch.Tensor:
    """tan(c * (1 + sin(lmda))) with c = fp32(0.25 pi (1 - 1e-7)), rounded operation by operation like the reference's
    chain of torch ops (checked bit for bit against them in tests/test_posatt_gpu.py)."""
    return _HeadScale.apply(lmda)


# ----------------------------------------------------------------------------------------------
# bias + GELU epilogue of the MLP Linears (pit.py:21-26, 111, 121): one pass forward, one backward (with the bias gradient)
# ----------------------------------------------------------------------------------------------
class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, bias, apply_gelu):
        z = z.contiguous()
        bias = bias.contiguous()
        cols = z.shape[-1]
        rows = z.numel() // cols
        out = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _cabi.check(_cabi.lib.pit_bias_act_forward(z.data_ptr(), bias.data_ptr(), out.data_ptr(), rows, cols, int(apply_gelu),
                                                       _stream(z.device)), "pit_bias_act_forward")
        ctx.save_for_backward(z, bias)
        ctx.apply_gelu = bool(apply_gelu)
        return out

    @staticmethod
    def backward(ctx, d_out):
        z, bias = ctx.saved_tensors
        d_out = d_out.contiguous()
        cols = z.shape[-1]
        rows = z.numel() // cols
        d_z = torch.empty_like(z)
        d_bias = torch.empty_like(bias)
        with torch.cuda.device(z.device):
            _cabi.check(_cabi.lib.pit_bias_act_backward(z.data_ptr(), bias.data_ptr(), d_out.data_ptr(), d_z.data_ptr(),
                                                        d_bias.data_ptr(), rows, cols, int(ctx.apply_gelu), _stream(z.device)),
                        "pit_bias_act_backward")
        return d_z, d_bias, None


@torch.compiler.disable
def bias_act_supported(z: torch.Tensor, bias: torch.Tensor) -> bool:
    """True when the fused epilogue covers this activation: float32 CUDA tensors, width a multiple of 4 dividing 1024."""
    return (z.is_cuda and z.dtype == torch.float32 and bias.dtype == torch.float32 and bias.device == z.device and z.numel() > 0
            and bool(_cabi.lib.pit_bias_act_supported(z.numel() // z.shape[-1], z.shape[-1])))


@torch.compiler.disable
def bias_act(z: torch.Tensor, bias: torch.Tensor, apply_gelu: bool) -> torch.Tensor:
    """act(z + bias) with act = exact GELU or identity; backward returns d_z and d_bias = column sums in the same pass."""
    return _BiasAct.apply(z, bias, bool(apply_gelu))


# ----------------------------------------------------------------------------------------------
# fused narrow-input MLP: the encoder lift en_layer (+ the caller's GELU), pit.py:110-111, one launch per direction
# ----------------------------------------------------------------------------------------------
class _MlpFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act_out):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        w1, b1, w2, b2 = w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous()
        rows, k, d = x2.shape[0], x2.shape[1], w1.shape[0]
        z = torch.empty((2, rows, d), dtype=torch.float32, device=x.device)
        out = torch.empty((rows, d), dtype=torch.float32, device=x.device)
        lin3 = _linear_3xtf32()
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib.pit_mlp_fused_forward(x2.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), rows, k, d,
                                                        int(act_out), lin3, z[0].data_ptr(), z[1].data_ptr(), out.data_ptr(),
                                                        _stream(x.device)), "pit_mlp_fused_forward")
        ctx.save_for_backward(x2, w1, w2, z)
        ctx.meta = (shape, bool(act_out), lin3)
        return out.view(*shape[:-1], d)

    @staticmethod
    def backward(ctx, d_out):
        x2, w1, w2, z = ctx.saved_tensors
        shape, act_out, lin3 = ctx.meta
        rows, k, d = x2.shape[0], x2.shape[1], w1.shape[0]
        d_out = d_out.reshape(rows, d).contiguous()
        d_x = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        grads = torch.empty(d * d + 2 * d + d * k, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            _cabi.check(_cabi.lib.pit_mlp_fused_backward(x2.data_ptr(), w1.data_ptr(), w2.data_ptr(), z[0].data_ptr(), z[1].data_ptr(),
                                                         d_out.data_ptr(), rows, k, d, int(act_out), lin3, _ptr(d_x), grads.data_ptr(),
                                                         _stream(x2.device)), "pit_mlp_fused_backward")
        d_w2, d_b2, d_b1, d_w1 = grads[:d * d].view(d, d), grads[d * d:d * d + d], grads[d * d + d:d * d + 2 * d], grads[d * d + 2 * d:].view(d, k)
        return (None if d_x is None else d_x.view(shape)), d_w1, d_b1, d_w2, d_b2, None


@torch.compiler.disable
def mlp_fused_supported(x: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> bool:
    """True when the one-launch MLP covers this case: float32 CUDA input of width <= 32, hidden = output width 32 or 64."""
    return (x.is_cuda and x.dtype == torch.float32 and x.numel() > 0 and w1.dtype == torch.float32 and w1.device == x.device
            and w1.shape[1] == x.shape[-1] and tuple(w2.shape) == (w1.shape[0], w1.shape[0])
            and bool(_cabi.lib.pit_mlp_fused_supported(x.numel() // x.shape[-1], x.shape[-1], w1.shape[0], w2.shape[0])))


@torch.compiler.disable
def mlp_fused(x, w1, b1, w2, b2, act_out: bool) -> torch.Tensor:
    """act(W2 gelu(W1 x + b1) + b2) over the last dimension of x, act = exact GELU or identity."""
    return _MlpFused.apply(x, w1, b1, w2, b2, bool(act_out))


# ----------------------------------------------------------------------------------------------
# fused processor (pit.py:114-122): every block of a shared-mesh model in one launch per direction
# ----------------------------------------------------------------------------------------------
def _processor_problem(mesh: torch.Tensor, x: torch.Tensor, n_head: int, variant: str) -> _cabi.Problem:
    return _cabi.Problem(_cabi.VARIANT_CODE[variant], mesh.shape[-1], 0, x.shape[0], int(n_head), mesh.shape[0], mesh.shape[0], x.shape[-1])


def _linear_3xtf32() -> int:
    """torch's matmul precision decides how the Linear products run: 'highest' -> 3xTF32 (fp32 parity), else single TF32
    products -- what cuBLAS runs nn.Linear as under 'high' (pit.py:2)."""
    return int(torch.get_float32_matmul_precision() == "highest")


class _Processor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, mesh, period, scales, n_head, variant, *weights):
        x0 = x0.contiguous()
        n_blocks = len(weights) // 4
        weights = tuple(w.contiguous() for w in weights)
        scales = scales.contiguous()
        prob = _processor_problem(mesh, x0, n_head, variant)
        blocks = (_cabi.ProcessorBlock * n_blocks)(*[_cabi.ProcessorBlock(*(w.data_ptr() for w in weights[4 * k:4 * k + 4]))
                                                     for k in range(n_blocks)])
        lin3 = _linear_3xtf32()
        saved = torch.empty(int(_cabi.lib.pit_processor_saved_floats(C.byref(prob), n_blocks)), dtype=torch.float32, device=x0.device)
        out = torch.empty_like(x0)
        key = _Stage.__new__(_Stage)
        key.variant, key.batched, key.B, key.H, key.N, key.M, key.D, key.sd, key.device = (
            variant, False, x0.shape[0], int(n_head), mesh.shape[0], mesh.shape[0], x0.shape[-1], mesh.shape[-1], x0.device)
        with torch.cuda.device(x0.device), _timed("processor_fwd", key, n_blocks):
            _cabi.check(_cabi.lib.pit_processor_forward(C.byref(prob), n_blocks, mesh.data_ptr(), _ptr(period), x0.data_ptr(),
                                                        scales.data_ptr(), blocks, lin3, saved.data_ptr(), out.data_ptr(),
                                                        _stream(x0.device)), "pit_processor_forward")
        ctx.save_for_backward(x0, mesh, scales, saved, *weights)
        ctx.period, ctx.prob, ctx.lin3, ctx.key, ctx.n_blocks = period, prob, lin3, key, n_blocks
        return out

    @staticmethod
    def backward(ctx, d_out):
        x0, mesh, scales, saved, *weights = ctx.saved_tensors
        n_blocks, prob = ctx.n_blocks, ctx.prob
        d_out = d_out.contiguous()
        blocks = (_cabi.ProcessorBlock * n_blocks)(*[_cabi.ProcessorBlock(*(w.data_ptr() for w in weights[4 * k:4 * k + 4]))
                                                     for k in range(n_blocks)])
        grads = torch.empty(int(_cabi.lib.pit_processor_grad_floats(C.byref(prob), n_blocks)), dtype=torch.float32, device=x0.device)
        scratch = torch.empty(int(_cabi.lib.pit_processor_scratch_floats(C.byref(prob))), dtype=torch.float32, device=x0.device)
        d_x0 = torch.empty_like(x0)
        with torch.cuda.device(x0.device), _timed("processor_bwd", ctx.key, n_blocks):
            _cabi.check(_cabi.lib.pit_processor_backward(C.byref(prob), n_blocks, mesh.data_ptr(), _ptr(ctx.period), x0.data_ptr(),
                                                         scales.data_ptr(), blocks, ctx.lin3, saved.data_ptr(), d_out.data_ptr(),
                                                         d_x0.data_ptr(), grads.data_ptr(), scratch.data_ptr(), _stream(x0.device)),
                        "pit_processor_backward")
        outs, off = [], 0
        for w in weights:
            outs.append(grads[off:off + w.numel()].view(w.shape))
            off += w.numel()
        d_scales = grads[off:off + scales.numel()].view(scales.shape)
        return (d_x0, None, None, d_scales, None, None, *outs)


@torch.compiler.disable
def processor_supported(mesh: torch.Tensor, x: torch.Tensor, n_head: int, n_blocks: int, variant: str) -> bool:
    """True when the fused processor covers this case: a shared latent mesh of 32..256 points (a multiple of 32), float32 CUDA
    features of width 32 or 64, at most two heads and eight blocks."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and mesh.dim() == 2 and mesh.is_cuda and mesh.dtype == torch.float32
            and mesh.shape[0] == x.shape[1] and mesh.shape[-1] in (1, 2) and variant in _VARIANTS and x.numel() > 0):
        return False
    if torch.is_grad_enabled() and mesh.requires_grad:
        return False
    return bool(_cabi.lib.pit_processor_supported(C.byref(_processor_problem(mesh, x, n_head, variant)), int(n_blocks)))


@torch.compiler.disable
def processor_blocks(x: torch.Tensor, mesh: torch.Tensor, scales: torch.Tensor, n_head: int, variant: str, weights) -> torch.Tensor:
    """for k: x = gelu(mlp_k(cat(x, A_k x)))  with A_k the global position-attention of block k on the shared mesh `mesh`
    (pit.py:114-122).  scales [n_blocks, H] = head scales of the blocks; weights = (mlp1.weight, mlp1.bias, mlp2.weight, mlp2.bias)
    per block, flattened.  Gradients flow to x, scales and every weight."""
    mesh, _, _, period, _, _ = prepare_meshes(mesh, mesh, x, n_head, variant, 1.0)   # cached contiguous copy and wrap length
    return _Processor.apply(x, mesh, period, scales, int(n_head), variant, *weights)


# ----------------------------------------------------------------------------------------------
# relative Lp loss (utils.py:60-98): partial sums + finalize forward, one elementwise pass backward
# ----------------------------------------------------------------------------------------------
class _RelLp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, truth, p):
        pred, truth = pred.contiguous(), truth.contiguous()      # (B, L, O)
        B, L, O = pred.shape
        norms = torch.empty((B, O, 2), dtype=torch.float32, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            _cabi.check(_cabi.lib.pit_rel_lp_forward(truth.data_ptr(), pred.data_ptr(), B, L, O, int(p), norms.data_ptr(),
                                                     loss.data_ptr(), _stream(pred.device)), "pit_rel_lp_forward")
        ctx.save_for_backward(pred, truth, norms)
        ctx.p = int(p)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        pred, truth, norms = ctx.saved_tensors
        B, L, O = pred.shape
        d_loss = d_loss.contiguous()
        d_pred = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            _cabi.check(_cabi.lib.pit_rel_lp_backward(truth.data_ptr(), pred.data_ptr(), norms.data_ptr(), d_loss.data_ptr(), B, L, O,
                                                      ctx.p, d_pred.data_ptr(), _stream(pred.device)), "pit_rel_lp_backward")
        return d_pred, None, None


@torch.compiler.disable
def rel_lp_supported(truth: torch.Tensor, pred: torch.Tensor, p) -> bool:
    """True when the fused loss covers this case: float32 CUDA tensors (B, L, O) with O <= 4, p in {1, 2}, no gradient to `truth`."""
    return (pred.is_cuda and truth.is_cuda and pred.dtype == torch.float32 and truth.dtype == torch.float32 and pred.dim() == 3
            and pred.shape == truth.shape and not truth.requires_grad and p in (1, 2) and pred.numel() > 0
            and bool(_cabi.lib.pit_rel_lp_supported(pred.shape[0], pred.shape[1], pred.shape[2], int(p))))


@torch.compiler.disable
def rel_lp_loss(truth: torch.Tensor, pred: torch.Tensor, p) -> torch.Tensor:
    """sum_b mean_o ||truth - pred||_p / ||truth||_p over (B, L, O) tensors; gradient flows to `pred`."""
    return _RelLp.apply(pred, truth, int(p))


# ----------------------------------------------------------------------------------------------
# fused decoder tail: cross position-attention + two-layer MLP (pit.decoder, pit.py:124-127)
# ----------------------------------------------------------------------------------------------
class _DecoderTail(torch.autograd.Function):
    """out = W2 gelu(b1 + sum_h A_h Y_h) + b2 with Y = first Linear applied on the latent mesh (see decoder_tail)."""

    @staticmethod
    def forward(ctx, y, scale, b1, w2, b2, mesh_out, mesh_in, locality, variant):
        y = y.contiguous()                                   # (B, M, H, C)
        B, M, H, Cw = y.shape
        scale_shape = scale.shape
        scale = scale.reshape(-1).contiguous()
        b1, w2, b2 = b1.contiguous(), w2.contiguous(), b2.contiguous()
        out_dim = w2.shape[0]
        # the stage descriptor treats the hidden width as the value width
        mesh_out, mesh_in, st, period, (v_min, v_lo, v_hi, w, masked), entry = prepare_meshes(
            mesh_out, mesh_in, y.reshape(B, M, H * Cw)[:, :, :Cw], H, variant, float(locality))
        st.problem.dim = Cw
        plan = tail_plan_for(entry, st, Cw)
        with torch.cuda.device(st.device):
            out = torch.empty((B, st.N, out_dim), dtype=torch.float32, device=st.device)
            # opaque to the caller: row sums (and, with a tile plan, sum_j P^ d2 per row in tile order), consumed by the backward
            rowsum = torch.empty((2 * H, (st.N + 31) // 32 * 32), dtype=torch.float32, device=st.device)
            rs = _rowstat_struct(v_min, v_lo, v_hi, w, masked)
            with _timed("tail_fwd", st, False):
                _cabi.check(_cabi.lib.pit_decoder_tail_forward(
                    C.byref(st.problem), mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), y.data_ptr(), scale.data_ptr(),
                    C.byref(rs), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), out_dim, out.data_ptr(), rowsum.data_ptr(),
                    C.byref(plan.struct) if plan is not None else None, _stream(st.device)), "pit_decoder_tail_forward")
        ctx.save_for_backward(y, scale, b1, w2, b2, mesh_out, mesh_in, period if period is not None else scale.new_empty(0),
                              v_min, v_lo, v_hi, rowsum)
        ctx.meta = (variant, w, masked, scale_shape, out_dim)
        ctx.plan = plan          # keeps the plan's tensors alive until the backward has run
        return out

    @staticmethod
    def backward(ctx, d_out):
        y, scale, b1, w2, b2, mesh_out, mesh_in, period, v_min, v_lo, v_hi, rowsum = ctx.saved_tensors
        variant, w, masked, scale_shape, out_dim = ctx.meta
        plan = ctx.plan
        period = period if period.numel() else None
        B, M, H, Cw = y.shape
        st = _Stage(mesh_out, mesh_in, y.reshape(B, M, H * Cw)[:, :, :Cw], H, variant)
        st.problem.dim = Cw
        d_out = d_out.contiguous()
        with torch.cuda.device(st.device):
            # one allocation, gradients back to back (each piece padded to 16 bytes): the library clears them with ONE memset
            # instead of five when it finds them contiguous
            sizes = [y.numel(), b1.numel(), w2.numel(), b2.numel(), H]
            offs = [0]
            for n in sizes:
                offs.append(offs[-1] + (n + 3) // 4 * 4)
            flat = torch.empty(offs[-1], dtype=torch.float32, device=st.device)
            d_y, d_b1, d_w2, d_b2, d_scale = (flat[o:o + n] for o, n in zip(offs, sizes))
            d_y, d_b1, d_w2, d_b2 = d_y.view_as(y), d_b1.view_as(b1), d_w2.view_as(w2), d_b2.view_as(b2)
            rs = _rowstat_struct(v_min, v_lo, v_hi, w, masked)
            with _timed("tail_bwd", st, False):
                _cabi.check(_cabi.lib.pit_decoder_tail_backward(
                    C.byref(st.problem), mesh_out.data_ptr(), mesh_in.data_ptr(), _ptr(period), y.data_ptr(), scale.data_ptr(),
                    C.byref(rs), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), out_dim, rowsum.data_ptr(), d_out.data_ptr(),
                    d_y.data_ptr(), d_scale.data_ptr(), d_b1.data_ptr(), d_w2.data_ptr(), d_b2.data_ptr(),
                    C.byref(plan.struct) if plan is not None else None, _stream(st.device)), "pit_decoder_tail_backward")
        return d_y, d_scale.reshape(scale_shape), d_b1, d_w2, d_b2, None, None, None, None


@torch.compiler.disable
def decoder_tail_supported(mesh_in: torch.Tensor, values: torch.Tensor, n_head: int, hidden: int, out_dim: int) -> bool:
    """True when the fused decoder-tail kernels cover this configuration (shared mesh, M <= 1024, H <= 2, ...)."""
    if mesh_in.dim() != 2 or not values.is_cuda or values.dtype != torch.float32:
        return False
    prob = _cabi.Problem(0, mesh_in.shape[-1], 0, values.shape[0], n_head, 1, mesh_in.shape[0], hidden)
    return bool(_cabi.lib.pit_decoder_tail_supported(C.byref(prob), out_dim))


@torch.compiler.disable
def decoder_tail(mesh_out, mesh_in, values, scale, locality, w1, b1, w2, b2, variant: str = "euclid") -> torch.Tensor:
    """Fused `de(up(mesh_out, mesh_in, values))`: (B, M, D) latent features -> (B, N, out_dim).

    w1 (C, H*D), b1 (C), w2 (O, C), b2 (O) are the parameters of the two Linears of ``kaiming_mlp``.  The first
    Linear is applied on the latent mesh -- Y[b,j,h,:] = W1[:, hD:(h+1)D] U[b,j,:] -- which is exact because
    position-attention is linear in its values; gradients of w1 and of the features flow through this einsum.
    """
    _meshes_are_constants(mesh_out, mesh_in)
    B, M, D = values.shape
    H = scale.numel()
    Cw = w1.shape[0]
    y = torch.einsum("bjd,chd->bjhc", values, w1.reshape(Cw, H, D))
    return _DecoderTail.apply(y, scale, b1, w2, b2, mesh_out, mesh_in, float(locality), variant)
