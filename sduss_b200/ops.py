"""Thin Python wrappers over the C ABI: they pass data_ptr()s and the current CUDA stream and
raise on a non-zero status. PyTorch is plumbing here (memory + streams); all arithmetic
happens inside libsduss_b200.so."""
import ctypes
import os

import torch

from . import _lib
import numpy as np

from ._lib import AttnSource, EpilogueDesc, check, lib

EPI_BIAS, EPI_GELU_TANH, EPI_GATE_RESID, EPI_QK_RMSNORM, EPI_GEGLU, EPI_ROWVEC = range(6)


# Launch accounting (bench.py reads these): every wrapper below is exactly one kernel launch of
# libsduss_b200.so. When `profile` is a dict, per-launch CUDA events are recorded by kernel name.
launch_count = 0
profile = None


def _count(name, tag=None):
    global launch_count
    launch_count += 1
    if profile is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        profile.setdefault(name, []).append(ev)
        if tag is not None:
            profile.setdefault("tags", []).append((name, tag, ev))
        ev[0].record()
        return ev[1]
    return None


def graphs_enabled():
    import os
    return os.environ.get("SDUSS_B200_NO_GRAPH", "0") != "1"


def capture_after():
    """A plan's forward is captured into a CUDA graph on its N-th use (default 2): the first use of
    a batch composition runs eagerly, so compositions a serving run sees only once never pay for a
    capture + instantiation."""
    return int(os.environ.get("SDUSS_B200_CAPTURE_AFTER", "2"))


def run_plan(model, plan, prologue=None, run=None, state=None):
    """Runs prologue(plan) (per-step inputs: by-value kernel arguments that change every step, never
    captured) and then run(plan) (default model._run). All shapes and pointers of a plan are static,
    so the forward -- a few hundred to a few thousand launches -- is captured into a CUDA graph on
    the plan's second use and replayed from then on; per-launch profiling (ops.profile) forces the
    eager path. `state` holds graph / use counters (default: the plan itself; a plan that is run in
    more than one way, e.g. with and without the patch cache, passes one state object per way)."""
    global launch_count
    run = model._run if run is None else run
    st = plan if state is None else state
    if not getattr(model, "use_graphs", False) or profile is not None:
        _run_eager(model, plan, prologue, run)
        return
    st.uses = getattr(st, "uses", 0) + 1
    if getattr(st, "graph", None) is None:
        if st.uses < capture_after() or not getattr(st, "warm", False):
            _run_eager(model, plan, prologue, run)  # eager: allocates workspaces, encodes tensor maps
            st.warm = True
            if st.uses < capture_after():
                return
            torch.cuda.current_stream().synchronize()
            eager_done = True
        else:
            eager_done = False
            if prologue is not None:
                prologue(plan)
        n0 = launch_count
        st.graph = _capture(model, plan, run)
        st.graph_launches = launch_count - n0
        launch_count = n0
        if eager_done:
            return                            # the eager run already produced this call's result
        st.graph.replay()
        launch_count += st.graph_launches
        return
    if prologue is not None:
        prologue(plan)
    st.graph.replay()
    launch_count += st.graph_launches


def _capture(model, plan, run=None):
    """Stream capture of model._run(plan) without torch.cuda.graph()'s entry cost (it runs
    gc.collect(), a device-wide synchronize and empty_cache() on every capture: tens of ms, and the
    emptied allocator cache is paid again by the next steps). A warm plan allocates nothing while
    it runs, so none of that is needed. thread_local error mode: CUDA calls from other threads of
    the runner process stay legal during the capture window."""
    g = torch.cuda.CUDAGraph()
    cur = torch.cuda.current_stream()
    side = getattr(model, "_capture_stream", None)
    if side is None:
        side = model._capture_stream = torch.cuda.Stream(device=cur.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        g.capture_begin(capture_error_mode="thread_local")
        try:
            (run or model._run)(plan)
        finally:
            g.capture_end()
    cur.wait_stream(side)
    return g


class ArenaOverflow(RuntimeError):
    def __init__(self, need):
        super().__init__(f"arena too small: {need} bytes needed so far")
        self.need = need


def _run_eager(model, plan, prologue=None, run=None):
    """prologue(plan); model._run(plan). Plans that bump-allocate their workspaces from the model's
    Arena restart on a larger block when it overflows (a handful of times per process: the block
    doubles); the prologue is re-run because it fills buffers of the block."""
    while True:
        try:
            if prologue is not None:
                prologue(plan)
            return (run or model._run)(plan)
        except ArenaOverflow as e:
            if torch.cuda.is_available():  # kernels of the partial run still use the old views
                torch.cuda.current_stream().synchronize()
            plan.reset_workspaces()
            model.arena.grow(e.need)
            if hasattr(model, "_plans"):
                model._plans._drop_stale()


class PlanCache:
    """Least-recently-used cache of per-composition plans (static workspaces, descriptor tables and
    the captured CUDA graph). A serving run sees many batch compositions (up to 12 requests over
    three resolutions in the reference's experiments) and a plan owns hundreds of MB to tens of GB
    of workspace, so the cache is bounded by device memory, not by count: when the plans together
    exceed `budget_bytes` the least recently used ones are dropped (their graph with them) and
    re-created on demand. The reference allocates its activations from torch's caching allocator
    every step instead; here pointers must be stable for graph replay."""

    def __init__(self, device, budget_fraction=0.35, arena=None):
        import collections
        self.plans = collections.OrderedDict()
        self.arena = arena          # plans carved from an older (smaller) block of it are stale
        self.stale_dropped = 0
        total = torch.cuda.get_device_properties(device).total_memory if torch.cuda.is_available() else 1 << 40
        self.budget_bytes = int(total * budget_fraction)
        self.evictions = 0

    def __len__(self):
        return len(self.plans)

    def values(self):
        return self.plans.values()

    def clear(self):
        self.plans.clear()

    @staticmethod
    def plan_bytes(plan, seen=None):
        """Bytes of device/host storage reachable from the plan; storages already in `seen` (e.g.
        a workspace arena shared by several plans) are not counted again."""
        seen, n = (set() if seen is None else seen), 0
        stack = list(vars(plan).values())
        while stack:
            v = stack.pop()
            if torch.is_tensor(v):
                st = v.untyped_storage()
                if st.data_ptr() not in seen:
                    seen.add(st.data_ptr())
                    n += st.nbytes()
            elif isinstance(v, dict):
                stack.extend(v.values())
            elif isinstance(v, (list, tuple)):
                stack.extend(v)
        return n

    def total_bytes(self):
        """Bytes the cached plans own themselves. The shared arena block is excluded: evicting a
        plan never frees it, so counting it would make every get() drop all plans (and their CUDA
        graphs) once the arena alone exceeds the budget."""
        seen = set()
        if self.arena is not None and self.arena.block is not None:
            seen.add(self.arena.block.untyped_storage().data_ptr())
        return sum(self.plan_bytes(p, seen) for p in self.plans.values())

    def _drop_stale(self):
        """When the shared arena has grown, plans (and CUDA graphs) that live on its previous block
        are dropped so that block is freed; they are rebuilt on the new one on demand. The arena
        doubles, so this happens a handful of times per process."""
        if self.arena is None or self.arena.block is None:
            return
        stale = [k for k, p in self.plans.items()
                 if getattr(p, "block", None) is not None and p.block is not self.arena.block]
        for k in stale:
            del self.plans[k]
        self.stale_dropped += len(stale)

    def get(self, key, factory):
        self._drop_stale()
        plan = self.plans.get(key)
        if plan is not None:
            self.plans.move_to_end(key)
            return plan
        # make room before the new plan allocates (sizes of lazily filled plans are read now)
        while self.plans and self.total_bytes() > self.budget_bytes:
            self.plans.popitem(last=False)
            self.evictions += 1
        plan = self.plans[key] = factory()
        return plan


class Arena:
    """One growing device allocation that every plan of a model carves its per-step workspaces
    from. Only one plan runs at a time and each of these buffers is fully rewritten by every
    forward before it is read, so plans can overlap in memory: a serving run with hundreds of
    batch compositions then needs the workspace of the largest one, not their sum. When a
    larger plan arrives a new block is allocated; older plans keep (a reference to) the block
    their CUDA graph was captured on."""
    ALIGN = 1024
    MIN_BLOCK = 1 << 30

    def __init__(self, device):
        self.device = device
        self.block = None

    def carve(self, specs):
        """specs: [(name, shape, dtype)] -> {name: tensor view}, plus the block they live in."""
        offs, total = [], 0
        for _, shape, dtype in specs:
            n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            offs.append(total)
            total += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if self.block is None or self.block.numel() < total:
            grow = 0 if self.block is None else 2 * self.block.numel()
            self.block = None
            self.block = torch.empty((max(total, grow),), dtype=torch.uint8, device=self.device)
        out = {}
        for (name, shape, dtype), off in zip(specs, offs):
            n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            out[name] = self.block[off:off + n].view(dtype).view(*shape)
        return out, self.block

    # -- bump allocation for plans that discover their workspaces while they run (UNet, VAE) --
    def alloc(self, plan, shape, dtype):
        """Next `shape` tensor of `plan` on the current block; ArenaOverflow if the block is
        missing, has been replaced since the plan started, or is full (see _run_eager)."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if getattr(plan, "block", None) is None:
            plan.block, plan.arena_off = self.block, 0
        end = plan.arena_off + (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if self.block is None or plan.block is not self.block or end > self.block.numel():
            raise ArenaOverflow(end)
        t = self.block[plan.arena_off:plan.arena_off + n].view(dtype).view(*shape)
        plan.arena_off = end
        return t

    def grow(self, need):
        have = 0 if self.block is None else self.block.numel()
        self.block = None  # plans captured on the old block keep it alive; nothing else does
        self.block = torch.empty((max(2 * need, 2 * have, self.MIN_BLOCK),), dtype=torch.uint8, device=self.device)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t, dtype=torch.bfloat16):
    assert t.is_cuda and t.dtype == dtype, (t.device, t.dtype)
    return t


def gemm(a, w, out=None, *, epi=EPI_BIAS, bias=None, resid=None, gate=None, row_group=None,
         rowvec=None, rms_wq=None, rms_wk=None, rms_q_cols=0, rms_k_cols=0, rms_eps=1e-6,
         q_scale=1.0, out_fp32=False, act=0, row_mask=None, row_mask_shift=8, ln_stats=None, ln_colsum=None,
         ln_rowpart=None, ln_eps=1e-5, rowpart_out=None, w_static=False):
    """out = epilogue(a @ w.T). a: [M, K] bf16 (row stride may exceed K), w: [N, K] bf16."""
    _req(a), _req(w)
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if epi == EPI_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device,
                          dtype=torch.float32 if out_fp32 else torch.bfloat16)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    d = EpilogueDesc()
    d.C, d.ldc, d.out_fp32 = _ptr(out), out.stride(0), int(out.dtype == torch.float32)
    d.bias = _ptr(bias)
    d.resid, d.ldr = _ptr(resid), (resid.stride(0) if resid is not None else 0)
    d.gate, d.ldg = _ptr(gate), (gate.stride(0) if gate is not None else 0)
    d.row_group = _ptr(row_group)
    d.rowvec, d.ldv = _ptr(rowvec), (rowvec.stride(0) if rowvec is not None else 0)
    d.rms_wq, d.rms_wk = _ptr(rms_wq), _ptr(rms_wk)
    d.rms_q_cols, d.rms_k_cols = rms_q_cols, rms_k_cols
    d.rms_eps, d.q_scale, d.act = rms_eps, q_scale, act
    if ln_colsum is not None:  # LayerNorm folded into this GEMM: a is the un-normalised activation
        assert ln_colsum.dtype == torch.float32 and ln_colsum.numel() == N
        d.ln_colsum = _ptr(ln_colsum)
        if ln_rowpart is not None:   # row statistics from the partial sums a's producer left behind
            assert ln_rowpart.dtype == torch.float32 and K % 64 == 0 and ln_rowpart.numel() >= M * (K // 64) * 2
            d.ln_rowpart, d.ln_nparts, d.ln_eps = _ptr(ln_rowpart), K // 64, ln_eps
        else:
            assert ln_stats is not None and ln_stats.dtype == torch.float32
            d.ln_stats = _ptr(ln_stats)
    if rowpart_out is not None:  # leave per-chunk row sums of the output for a LayerNorm folded into the next GEMM
        assert rowpart_out.dtype == torch.float32 and rowpart_out.numel() >= M * ((N + 63) // 64) * 2
        d.rowpart_out = _ptr(rowpart_out)
    if row_mask is not None:  # patch cache: M tiles of clean patches are skipped, their rows of out kept
        _req(row_mask, torch.int32)
        d.row_mask, d.row_mask_shift = _ptr(row_mask), row_mask_shift
    d.w_static = int(bool(w_static))  # w holds weights: its first tiles may be fetched under the previous kernel
    _ev = _count("b200_gemm_bf16", (M, N, K, epi))
    check(lib.b200_gemm_bf16(_ptr(a), a.stride(0), _ptr(w), w.stride(0), M, N, K, epi,
                             ctypes.byref(d), _stream()), "b200_gemm_bf16")
    if _ev is not None:
        _ev.record()
    return out


# ------------------------------------------------------------------ attention
def attn_source(q=None, q_col=0, k=None, k_col=0, v=None, v_col=0, out=None, o_col=0):
    """Describes one side (A or B) of the packed attention inputs. q/k/v/out are 2-D bf16
    buffers [rows, ld]; head h of q sits at columns [q_col + 64 h, q_col + 64 h + 64)."""
    s = AttnSource()
    for name, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        if t is not None:
            _req(t)
            assert t.dim() == 2 and t.stride(1) == 1
    s.q, s.ldq, s.q_col = _ptr(q), (q.stride(0) if q is not None else 0), q_col
    s.q_rows = q.shape[0] if q is not None else 0
    s.k, s.ldk, s.k_col = _ptr(k), (k.stride(0) if k is not None else 0), k_col
    s.v, s.ldv, s.v_col = _ptr(v), (v.stride(0) if v is not None else 0), v_col
    s.kv_rows = k.shape[0] if k is not None else 0
    s.out, s.ldo, s.o_col = _ptr(out), (out.stride(0) if out is not None else 0), o_col
    s._keep = (q, k, v, out)
    return s


ATTN_Q_TILE = lib.b200_attn_rows_per_item()  # query rows per CTA of attn_fwd_kernel


def attn_max_ctas():
    """0 = one CTA per SM (the library default). SDUSS_B200_ATTN_MAX_CTAS overrides it for A/B
    measurements (a huge value gives one unit per CTA, i.e. plain hardware dispatch)."""
    return int(os.environ.get("SDUSS_B200_ATTN_MAX_CTAS", "0"))


def build_attn_plan(seqs, device, n_heads, max_ctas=None):
    """seqs: list of 8-tuples (qa_row, qa_len, qb_row, qb_len, ka_row, ka_len, kb_row, kb_len).
    Returns (seq_table, work_units, n_units, sched_state, max_ctas) for attn_varlen: int32 tensors
    on `device`; the unit list (longest first) is built by the library's host function
    b200_attn_build_schedule; sched_state is the zeroed, self-resetting unit counter of the
    persistent kernel (one per plan: launches of one plan never overlap)."""
    table = np.ascontiguousarray(np.asarray(seqs, dtype=np.int32).reshape(-1, 8))
    n_units = ctypes.c_int(0)
    check(lib.b200_attn_build_schedule(table.ctypes.data, table.shape[0], n_heads, None,
                                       ctypes.byref(n_units)), "b200_attn_build_schedule")
    units = np.zeros((n_units.value, 4), dtype=np.int32)
    check(lib.b200_attn_build_schedule(table.ctypes.data, table.shape[0], n_heads,
                                       units.ctypes.data, ctypes.byref(n_units)),
          "b200_attn_build_schedule")
    return (torch.from_numpy(table).to(device), torch.from_numpy(units).to(device), n_units.value,
            torch.zeros(lib.b200_attn_workspace_bytes() // 4, dtype=torch.int32, device=device),
            attn_max_ctas() if max_ctas is None else max_ctas)


def attn_varlen(src_a, src_b, seq_table, work_units, n_units, sched_state, max_ctas, scale,
                causal=False, rel_bias=None, rel_len=0, q_mask=None, q_mask_shift=8, bounded=False):
    """rel_bias: fp32 [heads, >= 2 rel_len - 1], bias of key offset (k - q) at column k - q + rel_len - 1,
    already divided by `scale` (T5); causal: CLIP's mask. Both off on the denoising path.
    bounded: the caller guarantees |logit * scale * log2(e)| <= 64 (B200AttnExtra.bounded_logits)."""
    _ev = _count("b200_attn_varlen_bf16", (src_a.q_rows, src_a.kv_rows, n_units))
    extra = None
    if causal or rel_bias is not None or q_mask is not None or bounded:
        extra = _lib.AttnExtra()
        extra.causal = int(causal)
        extra.bounded_logits = int(bool(bounded))
        if q_mask is not None:  # patch cache: query tiles of clean patches (segment A) are skipped
            _req(q_mask, torch.int32)
            extra.q_mask, extra.q_mask_shift = _ptr(q_mask), q_mask_shift
        if rel_bias is not None:
            _req(rel_bias, torch.float32)
            assert rel_bias.dim() == 2 and rel_bias.stride(1) == 1
            extra.rel_bias, extra.rel_len, extra.rel_ld = _ptr(rel_bias), rel_len, rel_bias.stride(0)
    check(lib.b200_attn_varlen_ex(ctypes.byref(src_a),
                                  ctypes.byref(src_b) if src_b is not None else None,
                                  _ptr(seq_table), _ptr(work_units), n_units, _ptr(sched_state),
                                  max_ctas, ctypes.c_float(scale),
                                  ctypes.byref(extra) if extra is not None else None, _stream()),
          "b200_attn_varlen_ex")
    if _ev is not None:
        _ev.record()


ATTN_CROSS_SHORT_KEYS = lib.b200_attn_cross_short_max_keys()


def attn_cross_short(src_q, src_kv, seq_table, n_seq, n_heads, max_q_len, max_kv_len, scale, q_mask=None,
                     q_mask_shift=8):
    """Cross attention against at most ATTN_CROSS_SHORT_KEYS keys per sequence (SDXL: the 77 text tokens):
    Q = segment A rows of src_q, K / V = segment B rows of src_kv, same seq_table as attn_varlen."""
    _ev = _count("b200_attn_cross_short_bf16", (src_q.q_rows, max_kv_len, n_seq * n_heads))
    if q_mask is not None:
        _req(q_mask, torch.int32)
    check(lib.b200_attn_cross_short_bf16(ctypes.byref(src_q), ctypes.byref(src_kv), _ptr(seq_table), n_seq,
                                         n_heads, max_q_len, max_kv_len, ctypes.c_float(scale), _ptr(q_mask),
                                         q_mask_shift, _stream()), "b200_attn_cross_short_bf16")
    if _ev is not None:
        _ev.record()


class DeviceForest:
    """A RandomForest flattened into device arrays for b200_patch_mask_bf16. Build it from a fitted
    sklearn RandomForestClassifier (`from_sklearn`) or from explicit arrays; `threshold_rule(tau)` is
    the one-node forest "recompute iff MSE > tau"."""

    def __init__(self, feature, threshold, left, right, value, roots, device):
        i32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int32).to(device)
        f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(device)
        self.t = (i32(feature), f32(threshold), i32(left), i32(right), f32(value), i32(roots))
        self.c = _lib.Forest()
        (self.c.feature, self.c.threshold, self.c.left, self.c.right, self.c.value, self.c.roots) = \
            [_ptr(t) for t in self.t]
        self.c.n_trees = len(roots)

    @classmethod
    def from_sklearn(cls, rf, device):
        feature, threshold, left, right, value, roots = [], [], [], [], [], []
        one = list(rf.classes_).index(1) if 1 in list(rf.classes_) else None
        for est in rf.estimators_:
            t, base = est.tree_, len(feature)
            roots.append(base)
            for n in range(t.node_count):
                leaf = t.children_left[n] < 0
                feature.append(-1 if leaf else int(t.feature[n]))
                threshold.append(0.0 if leaf else float(t.threshold[n]))
                left.append(0 if leaf else base + int(t.children_left[n]))
                right.append(0 if leaf else base + int(t.children_right[n]))
                v = t.value[n][0]
                value.append(float(v[one] / v.sum()) if one is not None else 0.0)
        return cls(feature, threshold, left, right, value, roots, device)

    @classmethod
    def from_npz(cls, path, device):
        """A forest stored flattened (tools/patch_cache_study.py writes sduss_b200/data/patch_cache_*.npz)."""
        z = np.load(path)
        return cls(z["feature"], z["threshold"], z["left"], z["right"], z["value"], z["roots"], device)

    @classmethod
    def threshold_rule(cls, tau, device):
        return cls([2, -1, -1], [tau, 0.0, 0.0], [1, 0, 0], [2, 0, 0], [0.0, 0.0, 1.0], [0], device)


def patch_mask_workspace(n_patches, device):
    return torch.zeros(lib.b200_patch_mask_workspace_bytes(n_patches), dtype=torch.uint8, device=device)


def patch_mask(x, prev, patch_latent, latent_t, latent_valid, skipped, mask, forest, block_index,
               refresh, workspace, rows_per_patch=256, mse=None, extra_mse=None):
    """mask[p] = 1 where patch p (rows_per_patch rows of x) must be recomputed; prev <- x.
    extra_mse [n_extra, n_patches] fp32: further features of an SDXL up block (patch_mse of its skips)."""
    _req(x), _req(prev)
    n = mask.numel()
    if extra_mse is None:
        _ev = _count("b200_patch_mask_bf16")
        check(lib.b200_patch_mask_bf16(_ptr(x), x.stride(0), _ptr(prev), prev.stride(0), n, rows_per_patch,
                                       x.shape[1], _ptr(patch_latent), _ptr(latent_t), _ptr(latent_valid),
                                       _ptr(skipped), _ptr(mask), _ptr(mse), ctypes.byref(forest.c), block_index,
                                       refresh, _ptr(workspace), _stream()), "b200_patch_mask_bf16")
    else:
        _req(extra_mse, torch.float32)
        assert extra_mse.shape[1] == n and extra_mse.is_contiguous()
        _ev = _count("b200_patch_mask_ex")
        check(lib.b200_patch_mask_ex(_ptr(x), x.stride(0), _ptr(prev), prev.stride(0), n, rows_per_patch,
                                     x.shape[1], _ptr(patch_latent), _ptr(latent_t), _ptr(latent_valid),
                                     _ptr(skipped), _ptr(mask), _ptr(mse), ctypes.byref(forest.c), block_index,
                                     refresh, _ptr(extra_mse), extra_mse.shape[0], _ptr(workspace), _stream()),
              "b200_patch_mask_ex")
    if _ev is not None:
        _ev.record()
    return mask


def patch_mse(x, prev, patch_latent, latent_t, latent_valid, mse, workspace, rows_per_patch=256):
    """mse[p] = mean squared difference of patch p of x against prev (float(sys.maxsize) where the
    latent's kept copies are not valid); prev <- x. No decision (b200_patch_mask_ex, forest = NULL)."""
    _req(x), _req(prev), _req(mse, torch.float32)
    _ev = _count("b200_patch_mask_ex")
    check(lib.b200_patch_mask_ex(_ptr(x), x.stride(0), _ptr(prev), prev.stride(0), mse.numel(), rows_per_patch,
                                 x.shape[1], _ptr(patch_latent), _ptr(latent_t), _ptr(latent_valid),
                                 None, None, _ptr(mse), None, 0, 0, None, 0, _ptr(workspace), _stream()),
          "b200_patch_mask_ex")
    if _ev is not None:
        _ev.record()
    return mse


def embed_rows(ids, table, out, pos=None, seq_len=0):
    """out[i] = table[ids[i]] (+ pos[i % seq_len]); ids int32 on the device."""
    _req(ids, torch.int32), _req(table), _req(out)
    n, D = ids.numel(), table.shape[1]
    _ev = _count("b200_embed_rows_bf16")
    check(lib.b200_embed_rows_bf16(_ptr(ids), n, _ptr(table), table.shape[0], D, _ptr(pos), seq_len,
                                   _ptr(out), out.stride(0), _stream()), "b200_embed_rows_bf16")
    if _ev is not None:
        _ev.record()
    return out


def rmsnorm(x, weight, y, eps):
    _req(x), _req(weight), _req(y)
    T, D = x.shape
    _ev = _count("b200_rmsnorm_bf16")
    check(lib.b200_rmsnorm_bf16(_ptr(x), x.stride(0), T, D, ctypes.c_float(eps), _ptr(weight), _ptr(y),
                                y.stride(0), _stream()), "b200_rmsnorm_bf16")
    if _ev is not None:
        _ev.record()
    return y


# ------------------------------------------------------------------ HBM-bound kernels
def layernorm_mod(x, y, *, eps, gamma=None, beta=None, mod=None, row_group=None, shift_col=0,
                  scale_col=0, y2=None, shift2_col=0, scale2_col=0, row_mask=None, row_mask_shift=8):
    """y = LN(x)[*gamma+beta][*(1+mod[g,scale_col:])+mod[g,shift_col:]]; optional y2."""
    _req(x), _req(y)
    T, D = x.shape
    _ev = _count("b200_layernorm_mod_bf16")
    check(lib.b200_layernorm_mod_bf16(
        _ptr(x), x.stride(0), T, D, ctypes.c_float(eps), _ptr(gamma), _ptr(beta), _ptr(mod),
        (mod.stride(0) if mod is not None else 0), _ptr(row_group), shift_col, scale_col,
        _ptr(y), y.stride(0), shift2_col, scale2_col, _ptr(y2),
        (y2.stride(0) if y2 is not None else 0), _ptr(row_mask), row_mask_shift, _stream()),
        "b200_layernorm_mod_bf16")
    if _ev is not None:
        _ev.record()
    return y


def row_stats(x, stats, eps):
    """stats[row] = (mean, rstd) of x[row]: the row statistics of a LayerNorm folded into a GEMM."""
    _req(x), _req(stats, torch.float32)
    T, D = x.shape
    assert stats.numel() >= 2 * T
    _ev = _count("b200_row_stats_bf16")
    check(lib.b200_row_stats_bf16(_ptr(x), x.stride(0), T, D, ctypes.c_float(eps), _ptr(stats), _stream()),
          "b200_row_stats_bf16")
    if _ev is not None:
        _ev.record()
    return stats


def fold_layernorm(w, gamma, beta, bias=None):
    """Weights of a GEMM that absorbs the LayerNorm in front of it: (W o gamma as bf16, colsum fp32 of
    those bf16 weights, bias + beta W^T as bf16). LN(x) W^T + b = rstd (x W'^T - mean colsum) + b'."""
    wf = w.float() * gamma.float()[None, :]
    wq = wf.to(torch.bfloat16).contiguous()
    colsum = wq.float().sum(dim=1).contiguous()
    b = (beta.float()[None, :] @ w.float().t()).reshape(-1)
    if bias is not None:
        b = b + bias.float()
    return wq, colsum, b.to(torch.bfloat16).contiguous()


def silu(x, y=None):
    _req(x)
    assert x.is_contiguous()
    y = torch.empty_like(x) if y is None else y
    _ev = _count("b200_silu_bf16")
    check(lib.b200_silu_bf16(_ptr(x), _ptr(y), x.numel(), _stream()), "b200_silu_bf16")
    if _ev is not None:
        _ev.record()
    return y


def timestep_embedding(t, dim, out=None):
    """t: [n] fp32 -> [n, dim] bf16 = [cos | sin]."""
    _req(t, torch.float32)
    n = t.numel()
    out = torch.empty((n, dim), device=t.device, dtype=torch.bfloat16) if out is None else out
    _ev = _count("b200_timestep_embedding")
    check(lib.b200_timestep_embedding(_ptr(t), n, dim, _ptr(out), out.stride(0), _stream()), "b200_timestep_embedding")
    if _ev is not None:
        _ev.record()
    return out


def sd3_patchify(lat_ptr, desc, n_latents, max_tokens, C, p, tokens):
    _ev = _count("b200_sd3_patchify")
    check(lib.b200_sd3_patchify(_ptr(lat_ptr), _ptr(desc), n_latents, max_tokens, C, p,
                                _ptr(tokens), tokens.stride(0), _stream()), "b200_sd3_patchify")
    if _ev is not None:
        _ev.record()


def sd3_unpatchify(tokens, desc, n_latents, max_tokens, C, p, out_ptr):
    _ev = _count("b200_sd3_unpatchify")
    check(lib.b200_sd3_unpatchify(_ptr(tokens), tokens.stride(0), _ptr(desc), n_latents,
                                  max_tokens, C, p, _ptr(out_ptr), _stream()), "b200_sd3_unpatchify")
    if _ev is not None:
        _ev.record()


_DT = {torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16, torch.float32: _lib.DT_F32}


def dtype_code(dtype):
    try:
        return _DT[dtype]
    except KeyError:
        raise TypeError(f"latents must be bf16, fp16 or fp32, got {dtype}") from None


def latent_refs(rows):
    """rows: iterable of (src_ptr, dst_ptr, elems, off_a, off_b, sigma, sigma_next) -> ctypes array
    of B200LatentRef (host memory; it travels by value inside the kernel parameters)."""
    rows = list(rows)
    arr = (_lib.LatentRef * len(rows))()
    for a, (src, dst, n, oa, ob, s, sn) in zip(arr, rows):
        a.src, a.dst, a.elems, a.off_a, a.off_b, a.sigma, a.sigma_next = src, dst, n, oa, ob, s, sn
    return arr


def gather_latents(refs, dtype, staging, scale_input=False):
    """Latents of all requests -> the model's bf16 input staging buffer (CFG duplicate, optional
    Euler input scaling). refs: latent_refs(...) with off_a / off_b = element offsets in staging."""
    _req(staging)
    _ev = _count("b200_gather_latents")
    check(lib.b200_gather_latents(refs, len(refs), dtype_code(dtype), int(scale_input),
                                  _ptr(staging), _stream()), "b200_gather_latents")
    if _ev is not None:
        _ev.record()


def cfg_scheduler_step(eps, refs, latent_dtype, guidance, cfg, mode):
    """CFG combine + scheduler update of all requests in one launch. eps: the model's flat output
    (bf16, or fp32 for the scheduler mixins on fp32 model outputs); refs: latent_refs(...) with
    off_a / off_b = element offsets of the uncond / cond prediction in eps."""
    assert eps.is_cuda and eps.dtype in (torch.bfloat16, torch.float32)
    _ev = _count("b200_cfg_scheduler_step")
    check(lib.b200_cfg_scheduler_step(_ptr(eps), dtype_code(eps.dtype), refs, len(refs),
                                      dtype_code(latent_dtype), ctypes.c_float(guidance), int(cfg),
                                      mode, _stream()), "b200_cfg_scheduler_step")
    if _ev is not None:
        _ev.record()


def write_f32(dst, values):
    """dst[:len(values)] = values (fp32), passed by value: no host->device copy, no sync."""
    _req(dst, torch.float32)
    n = len(values)
    assert dst.numel() >= n
    arr = (ctypes.c_float * n)(*values)
    _ev = _count("b200_write_f32")
    check(lib.b200_write_f32(_ptr(dst), arr, n, _stream()), "b200_write_f32")
    if _ev is not None:
        _ev.record()


def gather_rows(dst, srcs, bytes_each=None):
    """dst[i] = srcs[i] for a list of equally sized contiguous device tensors (or raw pointers when
    bytes_each is given); dst: [n, ...] with any row stride."""
    n = len(srcs)
    if bytes_each is None:
        bytes_each = srcs[0].numel() * srcs[0].element_size()
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
    else:
        ptrs = (ctypes.c_void_p * n)(*srcs)
    stride = dst.stride(0) * dst.element_size()
    assert dst.shape[0] >= n and stride >= bytes_each
    _ev = _count("b200_gather_rows")
    check(lib.b200_gather_rows(_ptr(dst), stride, ptrs, n, bytes_each, _stream()), "b200_gather_rows")
    if _ev is not None:
        _ev.record()


# ------------------------------------------------------------------ SDXL path
def _epi_desc(out, bias=None, resid=None, gate=None, row_group=None, rowvec=None):
    d = EpilogueDesc()
    d.C, d.ldc, d.out_fp32 = _ptr(out), out.stride(0), 0
    d.bias = _ptr(bias)
    d.resid, d.ldr = _ptr(resid), (resid.stride(0) if resid is not None else 0)
    d.gate, d.ldg = _ptr(gate), (gate.stride(0) if gate is not None else 0)
    d.row_group = _ptr(row_group)
    d.rowvec, d.ldv = _ptr(rowvec), (rowvec.stride(0) if rowvec is not None else 0)
    d.rms_eps, d.q_scale = 1e-6, 1.0
    return d


def conv3x3_encode_maps(x, cin, in_desc_host, stride):
    """x: packed NHWC buffer [rows, ld] (only the first `cin` columns are read).
    in_desc_host: int32 numpy [n,4] = {input row offset, Hin, Win, 0}. Returns a device uint8
    buffer holding n CUtensorMaps (64-byte aligned)."""
    _req(x)
    n = in_desc_host.shape[0]
    host = np.zeros(lib.b200_conv3x3_maps_bytes(n), dtype=np.uint8)
    desc = np.ascontiguousarray(in_desc_host, dtype=np.int32)
    check(lib.b200_conv3x3_encode_maps(_ptr(x), x.stride(0), cin,
                                       desc.ctypes.data_as(ctypes.c_void_p), n, stride,
                                       host.ctypes.data_as(ctypes.c_void_p)),
          "b200_conv3x3_encode_maps")
    dev = torch.empty(n * 128 + 64, dtype=torch.uint8, device=x.device)
    off = (-dev.data_ptr()) % 64
    dev = dev[off:off + n * 128]
    dev.copy_(torch.from_numpy(host))
    return dev


def conv3x3(maps_dev, tiles, n_mtiles, out_lat, cin, cout, stride, weight, out, *, out_maps,
            resid_maps=None, epi=EPI_BIAS, bias=None, rowvec=None, row_group=None, stats_out=None,
            row_mask=None, row_mask_shift=8, row_mask_scale=0):
    """out[M_total, cout] = conv3x3(x) through the implicit-GEMM kernel (see conv_sm100.cu).
    out_maps / resid_maps: conv3x3_encode_maps(out or resid buffer, cout, out_desc_host, 1).
    row_mask (patch cache): int32 per 2^shift rows of the level the decision was taken on, `scale` =
    log2(rows there / output rows); pixel blocks whose 16 pixel rows touch only clean patches keep
    their previous output."""
    _req(weight), _req(out)
    d = _epi_desc(out, bias=bias, rowvec=rowvec, row_group=row_group)
    if row_mask is not None:
        _req(row_mask, torch.int32)
        d.row_mask, d.row_mask_shift, d.row_mask_scale = _ptr(row_mask), row_mask_shift, row_mask_scale
    if stats_out is not None:  # per (tile, half, channel) sums for the GroupNorm that follows
        assert stats_out.dtype == torch.float32 and stats_out.numel() >= n_mtiles * 2 * cout * 2
        d.stats_out = _ptr(stats_out)
    _ev = _count("b200_conv3x3_bf16", (out.shape[0], cout, 9 * cin, stride))
    check(lib.b200_conv3x3_bf16(_ptr(maps_dev), _ptr(out_maps), _ptr(resid_maps), _ptr(tiles),
                                n_mtiles, _ptr(out_lat), cin, cout,
                                stride, _ptr(weight), out.shape[0], epi, ctypes.byref(d),
                                _stream()), "b200_conv3x3_bf16")
    if _ev is not None:
        _ev.record()
    return out


def groupnorm_workspace(total_rows, n_latents, device):
    n = lib.b200_groupnorm_workspace_bytes(total_rows, n_latents)
    return torch.zeros(n, dtype=torch.uint8, device=device)   # zeroed once: epoch / flags of the grid barrier


def groupnorm_nhwc(x, y, gamma, beta, row_group, lat_chunks, n_latents, workspace, *, groups=32,
                   eps=1e-5, silu=False, channels=None):
    _req(x), _req(y)
    T = x.shape[0]
    C = channels if channels is not None else x.shape[1]
    _ev = _count("b200_groupnorm_nhwc_bf16")
    check(lib.b200_groupnorm_nhwc_bf16(_ptr(x), x.stride(0), T, C, groups, ctypes.c_float(eps),
                                       _ptr(gamma), _ptr(beta), _ptr(row_group), _ptr(lat_chunks),
                                       n_latents, int(silu), _ptr(y), y.stride(0),
                                       _ptr(workspace), _stream()), "b200_groupnorm_nhwc_bf16")
    if _ev is not None:
        _ev.record()
    return y


def conv_stats_buffer(pl, out, level, n_tiles, cout):
    """The fp32 buffer [n_tiles * 2, cout, 2] a convolution writing `out` leaves its per-tile GroupNorm
    partial sums in (one per (level, cout) of the plan, from its arena), tagged so that a GroupNorm
    reading `out` next can tell the statistics are those of `out`'s current contents."""
    name = f"convstats{level}_{cout}"
    t = pl.bufs.get(name)
    if t is None:
        t = pl.bufs[name] = pl.arena.alloc(pl, (n_tiles * 2 * cout * 2,), torch.float32)
    gen = pl.stats_gen[name] = pl.stats_gen.get(name, 0) + 1
    pl.stats_tag[out.data_ptr()] = (name, gen, cout)
    return t


def fresh_conv_stats(pl, x):
    """The statistics buffer of the convolution that last wrote x, if nothing has reused it since."""
    tag = pl.stats_tag.get(x.data_ptr())
    if tag is None or pl.stats_gen.get(tag[0]) != tag[1] or tag[2] != x.shape[1]:
        return None
    return pl.bufs[tag[0]]


def groupnorm_from_conv_stats(x, y, gamma, beta, row_group, conv_stats, lat_tiles, n_latents, workspace, *,
                              groups=32, eps=1e-5, silu=False):
    """GroupNorm of a tensor b200_conv3x3_bf16 has just written with stats_out: one read of x."""
    _req(x), _req(y)
    T, C = x.shape
    _ev = _count("b200_groupnorm_nhwc_bf16")
    check(lib.b200_groupnorm_from_conv_stats(_ptr(x), x.stride(0), T, C, groups, ctypes.c_float(eps),
                                             _ptr(gamma), _ptr(beta), _ptr(row_group), _ptr(conv_stats),
                                             _ptr(lat_tiles), n_latents, int(silu), _ptr(y), y.stride(0),
                                             _ptr(workspace), _stream()), "b200_groupnorm_from_conv_stats")
    if _ev is not None:
        _ev.record()
    return y


def pack_im2col3x3(lat_ptr, desc, n_latents, max_pixels, C, out):
    _ev = _count("b200_pack_im2col3x3")
    check(lib.b200_pack_im2col3x3(_ptr(lat_ptr), _ptr(desc), n_latents, max_pixels, C, _ptr(out),
                                  out.stride(0), _stream()), "b200_pack_im2col3x3")
    if _ev is not None:
        _ev.record()


def scatter_nchw(x, desc, n_latents, max_pixels, C, out_ptr):
    _ev = _count("b200_scatter_nchw")
    check(lib.b200_scatter_nchw(_ptr(x), x.stride(0), _ptr(desc), n_latents, max_pixels, C,
                                _ptr(out_ptr), _stream()), "b200_scatter_nchw")
    if _ev is not None:
        _ev.record()


def upsample2x(x, in_desc, out_desc, n_latents, max_out_pixels, C, y):
    _ev = _count("b200_upsample2x_nhwc")
    check(lib.b200_upsample2x_nhwc(_ptr(x), x.stride(0), _ptr(in_desc), _ptr(out_desc), n_latents,
                                   max_out_pixels, C, _ptr(y), y.stride(0), _stream()),
          "b200_upsample2x_nhwc")
    if _ev is not None:
        _ev.record()


def latent_affine(in_ptr, out_ptr, desc, n_latents, max_pixels, c_in, c_out, weight, bias):
    """out_l = weight @ in_l + bias per pixel on NCHW bf16 latents (un-scaling + post_quant_conv)."""
    _ev = _count("b200_latent_affine")
    check(lib.b200_latent_affine(_ptr(in_ptr), _ptr(out_ptr), _ptr(desc), n_latents, max_pixels,
                                 c_in, c_out, _ptr(weight), _ptr(bias), _stream()), "b200_latent_affine")
    if _ev is not None:
        _ev.record()


def softmax_rows(s, p, scale):
    """p = softmax(scale * s, dim=-1); s fp32 [rows, cols], p bf16 [rows, cols] (row strides free)."""
    assert s.dtype == torch.float32 and p.dtype == torch.bfloat16 and s.shape == p.shape
    assert s.stride(1) == 1 and p.stride(1) == 1
    _ev = _count("b200_softmax_rows")
    check(lib.b200_softmax_rows(_ptr(s), s.stride(0), s.shape[0], s.shape[1], ctypes.c_float(scale),
                                _ptr(p), p.stride(0), _stream()), "b200_softmax_rows")
    if _ev is not None:
        _ev.record()
    return p


def copy_cols(src, dst, cols):
    """dst[:, :cols] = src[:, :cols]; src / dst are 2-D views (any row stride)."""
    _req(src), _req(dst)
    _ev = _count("b200_copy_cols_bf16")
    check(lib.b200_copy_cols_bf16(_ptr(src), src.stride(0), _ptr(dst), dst.stride(0),
                                  src.shape[0], cols, _stream()), "b200_copy_cols_bf16")
    if _ev is not None:
        _ev.record()


def split_patches(lat_ptr, ldesc, pdesc, n_patches, C, ps, out):
    _ev = _count("b200_split_patches")
    check(lib.b200_split_patches(_ptr(lat_ptr), _ptr(ldesc), _ptr(pdesc), n_patches, C, ps,
                                 _ptr(out), _stream()), "b200_split_patches")
    if _ev is not None:
        _ev.record()


def concat_patches(patches, ldesc, pdesc, n_patches, C, ps, out_ptr):
    _ev = _count("b200_concat_patches")
    check(lib.b200_concat_patches(_ptr(patches), _ptr(ldesc), _ptr(pdesc), n_patches, C, ps,
                                  _ptr(out_ptr), _stream()), "b200_concat_patches")
    if _ev is not None:
        _ev.record()
