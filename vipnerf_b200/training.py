"""Training step of the render path: the train-mode forward of VipNeRF.render_rays and
its backward as one torch.autograd.Function around the C ABI (vipnerf_train_forward / vipnerf_train_backward).

What the reference does with its op graph (src/Trainer01.py:93-102: `model(batch)` in train mode ->
LossComputer.compute_losses -> `TotalLoss.backward()`), this module does with two library calls; the losses stay the
reference's (they run on the returned tensors and autograd hands their gradients to `backward` here).  PyTorch is
plumbing: device memory, the autograd edge and the CPU generator the reference draws its random numbers from.
There is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib, renderpath


def draw_training_randoms(n_rays: int, n_coarse: int, n_fine: int, chunk: int, netchunk: Optional[int], perturb: bool,
                          raw_noise_std: float, has_fine: bool) -> Dict[str, torch.Tensor]:
    """Draws from torch's global CPU generator in the order the reference's train-mode forward consumes it, so that a
    seeded run sees the same numbers: per `chunk` of rays (batchify_rays, VipNeRF01.py:54) torch.rand [r,Nc]
    (get_z_vals_coarse :200), one torch.randn [n,1] per `netchunk` slice of the flattened coarse points (batchify
    :305 -> get_view_independent_outputs :551), torch.rand [r,Nf] (sample_pdf :242), then the fine network's randn
    slices.  Returns whole-batch CPU tensors t_rand [R,Nc], u_rand [R,Nf], sigma_noise_coarse [R,Nc],
    sigma_noise_fine [R,Nc+Nf] (noise already scaled by raw_noise_std); absent key = source off."""
    parts: Dict[str, List[torch.Tensor]] = {'t_rand': [], 'u_rand': [], 'sigma_noise_coarse': [], 'sigma_noise_fine': []}

    def noise(r: int, s: int) -> torch.Tensor:
        n_points = r * s
        step = netchunk if netchunk else n_points
        pieces = [torch.randn(min(step, n_points - i), 1) * raw_noise_std for i in range(0, n_points, step)]
        return torch.cat(pieces, 0).reshape(r, s)

    for i in range(0, n_rays, chunk):
        r = min(chunk, n_rays - i)
        if perturb:
            parts['t_rand'].append(torch.rand(r, n_coarse))
        if raw_noise_std > 0:
            parts['sigma_noise_coarse'].append(noise(r, n_coarse))
        if has_fine:
            if perturb:
                parts['u_rand'].append(torch.rand(r, n_fine))
            if raw_noise_std > 0:
                parts['sigma_noise_fine'].append(noise(r, n_coarse + n_fine))
    return {k: torch.cat(v, 0) for k, v in parts.items() if v}


def _aligned_bytes(n_bytes: int, device) -> torch.Tensor:
    buf = torch.empty(n_bytes + 256, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 256
    return buf[off:off + n_bytes]


class TrainOutputs(dict):
    """The train-mode output dict of the plugin (a plain dict for every consumer) that also carries what
    LossComputerFused needs to fuse the reference's losses into the backward: `fused` = {'token', 'state', 'cfg_kwargs'}."""
    fused: Optional[dict] = None


class FusedLossState:
    """Shared between the forward's autograd node and LossComputerFused: the loss description the backward should
    differentiate inside the compositing-backward kernel (None = only the dense upstream gradients)."""

    def __init__(self):
        self.loss_spec = None        # _lib.LossSpec
        self.keepalive = None        # tensors the spec points into


class _FusedTotal(torch.autograd.Function):
    """TotalLoss of the fused losses: its value comes from vipnerf_fused_losses; its gradient travels to the render's
    autograd node as the gradient of a scalar token - the dense dLoss/dOutput tensors never exist."""

    @staticmethod
    def forward(ctx, token: torch.Tensor, losses_dev: torch.Tensor):
        return losses_dev[4].clone()

    @staticmethod
    def backward(ctx, g):
        return g.reshape(()).to(torch.float32), None


class _RenderTrain(torch.autograd.Function):
    """forward(spec, *params): params = the 24 tensors of the coarse MLP in renderpath.MLP_PARAM_ORDER, then the 24 of
    the fine MLP (if any).  Returns the output tensors in spec['out_names'] order, then a scalar token whose gradient
    is the upstream gradient of a fused TotalLoss (FusedLossState)."""

    @staticmethod
    def forward(ctx, spec: dict, *params: torch.Tensor):
        lib = _lib.load()
        batch = spec['batch']
        device = batch['rays_o'].device
        R = batch['rays_o'].shape[0]
        has_fine = spec['has_fine']
        n_coarse, n_fine, V = spec['n_coarse'], spec['n_fine'] if has_fine else 0, spec['n_sec_views']
        cfg = _lib.make_cfg(n_coarse=n_coarse, n_fine=n_fine, n_sec_views=V, ndc=spec['ndc'],
                            white_bkgd=spec['white_bkgd'], lindisp=spec['lindisp'], precision='fp32',
                            train_precision=spec['train_precision'])
        names = renderpath.MLP_PARAM_ORDER
        packed_c = renderpath.pack_mlp(dict(zip(names, [p.detach() for p in params[:24]])), 'fp32')
        packed_f = renderpath.pack_mlp(dict(zip(names, [p.detach() for p in params[24:48]])), 'fp32') if has_fine else None
        keep: list = []
        rays = renderpath._make_rays(batch, spec['ndc'], n_coarse, n_fine, V, keep)
        noise_c = renderpath._f32c(batch['sigma_noise_coarse'], 'sigma_noise_coarse') if 'sigma_noise_coarse' in batch else None
        noise_f = renderpath._f32c(batch['sigma_noise_fine'], 'sigma_noise_fine') if has_fine and 'sigma_noise_fine' in batch else None
        out = _lib.Out()
        keys = renderpath.pass_keys(spec['ndc'], True, V)
        tensors: Dict[str, torch.Tensor] = {}
        for tag, pass_out, S in (('coarse', out.coarse, n_coarse), ('fine', out.fine, n_coarse + n_fine)):
            if tag == 'fine' and not has_fine:
                continue
            for k, t in renderpath._alloc_pass(pass_out, keys, R, S, V, device).items():
                tensors[f'{k}_{tag}'] = t
        with torch.cuda.device(device):
            saved_bytes = lib.vipnerf_train_saved_bytes(ctypes.byref(cfg), R)
            ws_bytes = lib.vipnerf_train_workspace_bytes(ctypes.byref(cfg), R)
            if saved_bytes == 0 or ws_bytes == 0:
                _lib.check(lib.vipnerf_check_config(ctypes.byref(cfg)), 'vipnerf_check_config')
            saved = _aligned_bytes(saved_bytes, device)
            ws = renderpath._workspace(ws_bytes, device)
            _lib.check(lib.vipnerf_train_forward(
                ctypes.byref(cfg), ctypes.byref(rays), R, renderpath._ptr(noise_c), renderpath._ptr(noise_f),
                packed_c.data_ptr(), packed_f.data_ptr() if has_fine else None, ctypes.byref(out), saved.data_ptr(),
                saved_bytes, ws.data_ptr(), ws_bytes, renderpath._stream(device)), 'vipnerf_train_forward')
        out_names = spec['out_names']
        token = torch.zeros((), dtype=torch.float32, device=device)
        outputs = tuple(tensors[n] for n in out_names) + (token,)
        ctx.mark_non_differentiable(*[tensors[n] for n in out_names if n.startswith('z_vals_')])
        ctx.spec = spec
        ctx.n_params = len(params)
        ctx.param_shapes = [tuple(p.shape) for p in params]
        ctx.set_materialize_grads(False)
        # the ray tensors stay alive through spec['batch'] (a dict this module owns)
        ctx.save_for_backward(saved, packed_c, packed_f if has_fine else packed_c, *outputs[:-1])
        return outputs

    @staticmethod
    def backward(ctx, *grads: Optional[torch.Tensor]):
        lib = _lib.load()
        spec = ctx.spec
        sv = ctx.saved_tensors
        saved, packed_c, packed_f = sv[0], sv[1], sv[2]
        outputs = sv[3:]
        out_names = spec['out_names']
        device = saved.device
        has_fine = spec['has_fine']
        n_coarse, n_fine, V = spec['n_coarse'], spec['n_fine'] if has_fine else 0, spec['n_sec_views']
        R = outputs[0].shape[0]
        cfg = _lib.make_cfg(n_coarse=n_coarse, n_fine=n_fine, n_sec_views=V, ndc=spec['ndc'],
                            white_bkgd=spec['white_bkgd'], lindisp=spec['lindisp'], precision='fp32',
                            train_precision=spec['train_precision'])
        keep: list = []
        rays = renderpath._make_rays(spec['batch'], spec['ndc'], n_coarse, n_fine, V, keep)
        fwd, gout = _lib.Out(), _lib.Out()
        token_grad = grads[len(out_names)] if len(grads) > len(out_names) else None
        state = spec.get('fused_state')
        loss_spec = state.loss_spec if (state is not None and token_grad is not None) else None
        if token_grad is not None:
            token_grad = renderpath._f32c(token_grad, 'grad of the loss token')
            keep.append(token_grad)
        for name, t, g in zip(out_names, outputs, grads):
            key, tag = name.rsplit('_', 1)
            setattr(getattr(fwd, tag), key, t.data_ptr())
            if g is not None and key != 'z_vals':
                g = renderpath._f32c(g, f'grad of {name}')
                keep.append(g)
                setattr(getattr(gout, tag), key, g.data_ptr())
        param_grads = [torch.empty(shape, dtype=torch.float32, device=device) for shape in ctx.param_shapes]
        arr_c = (ctypes.c_void_p * 24)(*[t.data_ptr() for t in param_grads[:24]])
        arr_f = (ctypes.c_void_p * 24)(*[t.data_ptr() for t in param_grads[24:48]]) if has_fine else None
        with torch.cuda.device(device):
            saved_bytes = lib.vipnerf_train_saved_bytes(ctypes.byref(cfg), R)
            ws_bytes = lib.vipnerf_train_workspace_bytes(ctypes.byref(cfg), R)
            ws = renderpath._workspace(ws_bytes, device)
            _lib.check(lib.vipnerf_train_backward_fused(
                ctypes.byref(cfg), ctypes.byref(rays), R, packed_c.data_ptr(), packed_f.data_ptr() if has_fine else None,
                ctypes.byref(fwd), ctypes.byref(gout), ctypes.byref(loss_spec) if loss_spec is not None else None,
                token_grad.data_ptr() if loss_spec is not None else None, saved.data_ptr(), saved_bytes, arr_c, arr_f,
                ws.data_ptr(), ws_bytes, renderpath._stream(device)), 'vipnerf_train_backward_fused')
        return (None, *param_grads)


def render_rays_train(batch: Dict[str, torch.Tensor], params_coarse: Dict[str, torch.Tensor],
                      params_fine: Optional[Dict[str, torch.Tensor]], *, ndc: bool, n_coarse: int = 64,
                      n_fine: int = 128, n_sec_views: int = 0, white_bkgd: bool = False,
                      lindisp: bool = False, tf32: bool = False, train_precision: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Train-mode VipNeRF.render_rays (VipNeRF01.py:74-171 with self.training: retraw and sec_views_vis on, :40),
    differentiable w.r.t. the MLP parameters.  `batch` holds the ray tensors plus the random draws (`t_rand`, `u_rand`,
    `sigma_noise_coarse`, `sigma_noise_fine`; see draw_training_randoms) - a missing draw switches that source off.
    `params_*`: the reference's state_dict names of one MLP -> parameter tensors (CUDA, fp32).
    `train_precision`: 'fp32' = CUDA cores (the reference's arithmetic); 'tf32' = every 256-wide product of the step
    (forward chain, backward-data chain, parameter gradients) on the tensor cores (tcgen05 kind::tf32: operands rounded
    to tf32, fp32 accumulation); 'fp16' = the same products on tcgen05 kind::f16 with every saved activation and chain
    gradient stored as fp16 (the same 11-bit significand, half the HBM bytes; gradients carry per-array power-of-two
    scales).  `tf32=True` is the older spelling of train_precision='tf32'."""
    if train_precision is None:
        train_precision = 'tf32' if tf32 else 'fp32'
    if train_precision not in _lib.TRAIN_PRECISIONS:
        raise ValueError(f'train_precision = {train_precision!r}')
    renderpath._require_cuda(batch['rays_o'], 'rays_o')
    has_fine = params_fine is not None and n_fine > 0
    names = renderpath.MLP_PARAM_ORDER
    keys = renderpath.pass_keys(ndc, True, n_sec_views)
    out_names = [f'{k}_coarse' for k in keys] + ([f'{k}_fine' for k in keys] if has_fine else [])
    state = FusedLossState()
    spec = dict(batch=batch, ndc=ndc, n_coarse=n_coarse, n_fine=n_fine, n_sec_views=n_sec_views, white_bkgd=white_bkgd,
                lindisp=lindisp, has_fine=has_fine, out_names=out_names, train_precision=train_precision, fused_state=state)
    params = [params_coarse[k] for k in names] + ([params_fine[k] for k in names] if has_fine else [])
    outputs = _RenderTrain.apply(spec, *params)
    result = TrainOutputs(zip(out_names, outputs[:-1]))
    result.fused = {'token': outputs[-1], 'state': state,
                    'cfg_kwargs': dict(n_coarse=n_coarse, n_fine=n_fine if has_fine else 0, n_sec_views=n_sec_views, ndc=ndc,
                                       white_bkgd=white_bkgd, lindisp=lindisp, precision='fp32', train_precision=train_precision)}
    for tag in ('coarse', 'fine'):
        if f'raw_rgb_{tag}' in result:   # the reference returns the same tensor under both names (:531)
            result[f'raw_rgb_view_dependent_{tag}'] = result[f'raw_rgb_{tag}']
    return result


def volume_rendering_backward(batch: Dict[str, torch.Tensor], z_vals: torch.Tensor, sigma: torch.Tensor,
                              rgb: torch.Tensor, visibility: torch.Tensor, visibility2: Optional[torch.Tensor],
                              grads: Dict[str, torch.Tensor], *, ndc: bool, white_bkgd: bool = False):
    """Backward of volume_rendering alone (vipnerf_composite_backward): network outputs of one sample set + upstream
    gradients keyed like the outputs -> (d_sigma_logit [R,S], d_head_logits [R,S,1+V,4])."""
    lib = _lib.load()
    z_vals = renderpath._f32c(z_vals, 'z_vals')
    device = z_vals.device
    R, S = z_vals.shape
    sigma = renderpath._f32c(sigma, 'sigma').reshape(R, S)
    rgb = renderpath._f32c(rgb, 'rgb').reshape(R, S, 3)
    visibility = renderpath._f32c(visibility, 'visibility').reshape(R, S)
    V = 0
    if visibility2 is not None:
        visibility2 = renderpath._f32c(visibility2, 'visibility2').reshape(R, S, -1)
        V = visibility2.shape[-1]
    cfg = _lib.make_cfg(n_coarse=S, n_fine=0, n_sec_views=V, ndc=ndc, white_bkgd=white_bkgd, precision='fp32')
    keep: list = []
    rays = renderpath._make_rays(batch, ndc, S, 0, 0, keep)
    gout = _lib.PassOut()
    for k, g in grads.items():
        g = renderpath._f32c(g, k)
        keep.append(g)
        setattr(gout, k, g.data_ptr())
    d_sigma = torch.empty((R, S), dtype=torch.float32, device=device)
    d_logits = torch.empty((R, S, 1 + V, 4), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.vipnerf_composite_backward(
            ctypes.byref(cfg), ctypes.byref(rays), R, S, z_vals.data_ptr(), sigma.data_ptr(), rgb.data_ptr(),
            visibility.data_ptr(), renderpath._ptr(visibility2), ctypes.byref(gout), d_sigma.data_ptr(),
            d_logits.data_ptr(), renderpath._stream(device)), 'vipnerf_composite_backward')
    return d_sigma, d_logits


def param_gradient_gemm(dy: torch.Tensor, x: torch.Tensor, *, mode: int = 0, with_bias: bool = True):
    """dW = dy^T x and db = column sums of dy over all rows (vipnerf_param_gradient_gemm).  dy [P,M], x [P,N] CUDA;
    mode 0 = fp32 CUDA cores, 1 = tcgen05 tf32 (fp32 arrays, N in {32, 64, 256}), 2 = tcgen05 fp16 (the arrays are
    converted to fp16 here; N in {64, 256})."""
    lib = _lib.load()
    if mode == 2:
        renderpath._require_cuda(dy, 'dy')
        dy, x = dy.to(torch.float16).contiguous(), x.to(torch.float16).contiguous()
    else:
        dy, x = renderpath._f32c(dy, 'dy'), renderpath._f32c(x, 'x')
    P, M = dy.shape
    N = x.shape[1]
    device = dy.device
    dw = torch.empty((M, N), dtype=torch.float32, device=device)
    db = torch.empty((M,), dtype=torch.float32, device=device) if with_bias else None
    with torch.cuda.device(device):
        ws_bytes = lib.vipnerf_param_gradient_gemm_workspace_bytes()
        ws = renderpath._workspace(ws_bytes, device)
        _lib.check(lib.vipnerf_param_gradient_gemm(dy.data_ptr(), M, M, x.data_ptr(), N, N, P, dw.data_ptr(), N, N,
                                                   renderpath._ptr(db), mode, ws.data_ptr(), ws_bytes,
                                                   renderpath._stream(device)), 'vipnerf_param_gradient_gemm')
    return dw, db


class GraphedTrainStep:
    """One training iteration of the reference's Trainer (Trainer01.py:61-107: batch to the device, `zero_grad`,
    `model(batch)` in train mode, `compute_losses`, `TotalLoss.backward()`, `optimizer.step()`) captured ONCE as a CUDA
    graph and replayed with a single launch per iteration: the ~170 kernel launches, ~50 tensor-map encodes and the
    Python between them leave the critical path (the fp16 step is otherwise host-bound: ~11 ms of GPU work behind
    ~14 ms of host work).

        step = GraphedTrainStep(model, loss_computer, optimizer, example_batch)
        loss = step(batch)           # pinned-host copy of the batch -> replay -> device scalar TotalLoss
        step.loss_host               # pinned host copy of the same value, valid after step.synchronize()

    Requirements (checked): configs['model']['rng'] = 'device' (the reference's CPU draws cannot be captured), an
    optimizer built with `capturable=True`, a fixed ray count and fixed batch keys.  Non-tensor batch entries
    (`iter_num`, ...) and the loss weights derived from them are baked in at capture time: call `recapture()` when an
    `iter_weights` threshold of the loss configs is crossed.  One GPU: data-parallel training (one process per GPU with
    sharding.allreduce_gradients between backward and step) uses eager launches - capturing the NCCL all-reduce with the
    iteration hung in the 2-GPU trial of r02 and is not supported."""

    def __init__(self, model, loss_computer, optimizer, example_batch: Dict, device=None, warmup: int = 3):
        cfg = getattr(model, 'configs', {}).get('model', {})
        if cfg.get('rng', 'reference') != 'device':
            raise ValueError("GraphedTrainStep needs configs['model']['rng'] = 'device': draws from the CPU generator "
                             "cannot be part of a CUDA graph")
        if not optimizer.defaults.get('capturable', False):
            raise ValueError('GraphedTrainStep needs an optimizer built with capturable=True')
        self.model, self.computer, self.optimizer = model, loss_computer, optimizer
        self.device = torch.device(device if device is not None else next(model.parameters()).device)
        # CPU tensors are the per-iteration inputs (pinned staging copy + upload inside the graph); tensors that already
        # live on the device and non-tensor entries are used in place
        self.extra = {k: v for k, v in example_batch.items() if not (isinstance(v, torch.Tensor) and v.device.type == 'cpu')}
        self.host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in example_batch.items()
                     if isinstance(v, torch.Tensor) and v.device.type == 'cpu'}
        self.dev = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in self.host.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.stream = torch.cuda.Stream(self.device)
        self.warmup = warmup
        self.load(example_batch)
        self.recapture()

    def load(self, batch: Dict) -> None:
        for k, v in self.host.items():
            v.copy_(batch[k])

    def _iteration(self):  # noqa: D401 - the captured body
        batch = dict(self.extra)
        for k, v in self.host.items():
            self.dev[k].copy_(v, non_blocking=True)
        batch.update(self.dev)
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(batch)
        total = self.computer.compute_losses(batch, out)['TotalLoss']
        total.backward()
        self.optimizer.step()
        self.loss.copy_(total.detach())
        self.loss_host.copy_(self.loss, non_blocking=True)

    def recapture(self) -> None:
        """Warm-up iterations on the capture stream (module load, allocator pools, lazily created optimizer state - these
        are REAL training iterations), then the capture."""
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for _ in range(self.warmup):
                    self._iteration()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._iteration()
            torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def set_lr(self, lr: float) -> None:
        """The reference trainer writes the decayed rate into `param_group['lr']` every iteration
        (Trainer01.py:293-295).  A Python float there is baked into the captured step; build the optimizer with the rate as
        a device tensor (`Adam(..., lr=torch.tensor(lr0, device=...), capturable=True)`) and the replayed step reads it."""
        for group in self.optimizer.param_groups:
            if not isinstance(group['lr'], torch.Tensor):
                raise ValueError("set_lr needs an optimizer whose lr is a device tensor (lr=torch.tensor(..., device=...)); "
                                 "a float lr is part of the captured graph - recapture() after changing it")
            group['lr'].fill_(lr)

    def __call__(self, batch: Optional[Dict] = None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        return self.loss

    def synchronize(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
