"""Host <-> device plumbing for serving loops around the plugin: ONE pinned host buffer and ONE device buffer per
direction (every ray tensor / output map is a 16-byte aligned view of them), so that a render step is one H2D copy,
one kernel and one D2H copy - and, with `GraphedRender`, one `cudaGraphLaunch` from the host.

The reference's Tester moves every tensor of the batch separately (`CommonUtils.move_to_device`, src/Tester01.py:60)
and pulls every output map separately (`DataPreprocessor01.retrieve_inference_outputs` :866-894): 14 small copies of
3-50 KB around a 0.8 ms kernel, plus ~90 us of Python per call.  PyTorch stays plumbing: pinned memory, streams,
`torch.cuda.CUDAGraph`; the arithmetic is the library's.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch


def _layout(shapes: Dict[str, Tuple[int, ...]]) -> Tuple[Dict[str, Tuple[int, int]], int]:
    offsets, total = {}, 0
    for k, shape in shapes.items():
        n = 1
        for d in shape:
            n *= int(d)
        offsets[k] = (total, n)
        total += (n + 3) // 4 * 4          # every array starts on a 16-byte boundary (the library's vector loads)
    return offsets, max(total, 4)


class FlatBuffers:
    """fp32 arrays `shapes` as views of one pinned host buffer (`.host[k]`) and one device buffer (`.dev[k]`)."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device):
        self.shapes = {k: tuple(int(d) for d in v) for k, v in shapes.items()}
        offsets, total = _layout(self.shapes)
        self.host_flat = torch.empty(total, dtype=torch.float32).pin_memory()
        self.dev_flat = torch.empty(total, dtype=torch.float32, device=device)
        self.host = {k: self.host_flat[o:o + n].view(self.shapes[k]) for k, (o, n) in offsets.items()}
        self.dev = {k: self.dev_flat[o:o + n].view(self.shapes[k]) for k, (o, n) in offsets.items()}
        self.nbytes = total * 4

    def upload(self) -> Dict[str, torch.Tensor]:
        self.dev_flat.copy_(self.host_flat, non_blocking=True)
        return self.dev

    def download(self) -> Dict[str, torch.Tensor]:
        self.host_flat.copy_(self.dev_flat, non_blocking=True)
        return self.host


def output_shapes(keys: Iterable[str], n_rays: int, n_coarse: int = 64, n_fine: int = 128, n_sec_views: int = 0
                  ) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the reference's per-ray / per-sample outputs by name (`rgb_fine`, `depth_coarse`, `alpha_fine` ...)."""
    per_ray = {'rgb': (3,), 'acc': (), 'depth': (), 'depth_var': (), 'depth_ndc': (), 'depth_var_ndc': (),
               'visibility2': (n_sec_views,)}
    shapes = {}
    for name in keys:
        key, tag = name.rsplit('_', 1)
        s = n_coarse if tag == 'coarse' else n_coarse + n_fine
        if key in per_ray:
            shapes[name] = (n_rays,) + per_ray[key]
        elif key in ('alpha', 'z_vals', 'visibility', 'weights'):
            shapes[name] = (n_rays, s)
        else:
            raise KeyError(f'{name}: not a per-ray / per-sample map')
    return shapes


class GraphedRender:
    """`model(batch, ...)` for a fixed ray count as a CUDA graph: H2D of the rays, the fused render, D2H of the
    requested maps - replayed with one launch.

        g = GraphedRender(model, example_host_batch, out_keys=('rgb_fine', 'depth_fine'))
        g.inputs.host['rays_o'][...] = ...      # fill the pinned input views (or g.load(batch))
        maps = g()                              # dict of pinned host views; valid after g.synchronize()
    """

    def __init__(self, model, example_batch: Dict[str, torch.Tensor], out_keys: Iterable[str], *, retraw: bool = False,
                 sec_views_vis: bool = False, device=None, download: bool = True):
        device = torch.device(device if device is not None else next(model.parameters()).device)
        self.model, self.device, self.kwargs = model, device, dict(retraw=retraw, sec_views_vis=sec_views_vis)
        rays = {k: v for k, v in example_batch.items() if isinstance(v, torch.Tensor) and v.dtype == torch.float32}
        self.extra = {k: v for k, v in example_batch.items() if k not in rays}
        self.inputs = FlatBuffers({k: tuple(v.shape) for k, v in rays.items()}, device)
        n_rays = example_batch['rays_o'].shape[0]
        n_sec = rays['rays_o2'].shape[1] if (sec_views_vis and 'rays_o2' in rays) else 0
        self.outputs = FlatBuffers(output_shapes(out_keys, n_rays, n_sec_views=n_sec), device)
        self.download = download
        self.load(rays)
        self.stream = torch.cuda.Stream(device)
        self.h2d_bytes, self.d2h_bytes = self.inputs.nbytes, self.outputs.nbytes if download else 0
        self._capture()

    def _capture(self) -> None:
        """(Re-)captures the graph for the model's CURRENT weights: a captured launch points at the packed weight images
        of one weights_version(); they are kept alive here for as long as the graph may replay them."""
        device = self.device
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.device(device):
            self.stream.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(self.stream):
                for _ in range(2):          # warm-up outside the capture: module load, attributes, weight packing
                    self._step()
            torch.cuda.synchronize(device)
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._step()
        version = getattr(self.model, 'weights_version', None)
        self._version = version() if version is not None else None
        cache = getattr(self.model, '_pack_cache', None)
        self._packed_alive = [e[1] for e in cache.entries.values()] if cache is not None else []

    def _step(self):
        batch = dict(self.extra)
        batch.update(self.inputs.upload())
        self.model(batch, out=self.outputs.dev, **self.kwargs)
        if self.download:
            self.outputs.download()

    def load(self, batch: Dict[str, torch.Tensor]) -> None:
        for k, v in self.inputs.host.items():
            v.copy_(batch[k])

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        if batch is not None:
            self.load(batch)
        if self._version is not None and self.model.weights_version() != self._version:
            self._capture()       # the weights changed (optimizer step, load_state_dict): the old launch is stale
        self.graph.replay()
        return self.outputs.host if self.download else self.outputs.dev

    def synchronize(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
