"""Host-side mirror of the reference ``models.yolo.Model`` / ``IDetect`` for the deployed path.

Keeps the call signatures the reference's detect.py / test.py use:
    Model(cfg, ch, nc, anchors)  .fuse() -> self   .forward(x, augment=False, profile=False) -> (pred, [raw0, raw1, raw2])
    model.model[-1] : detect layer with .stride .nl .na .no .nc .anchors .anchor_grid and .fuseforward(list_of_maps)
    model.stride, model.names, model.yaml, model.save
Parameters carry the reference checkpoint's state_dict names, so ``load_state_dict(reference_model.state_dict())`` works.
Only the fused (deploy) forward exists, and it runs in the CUDA library -- there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as N
from . import arch, fold, planner


class _Node(nn.Module):
    """Generic container used to reproduce the reference's dotted parameter names."""

    def __getitem__(self, i):
        return getattr(self, str(i if i >= 0 else len(self._modules) + i))

    def __len__(self):
        return len(self._modules)


def _insert(root, name, tensor, is_buffer):
    *path, leaf = name.split('.')
    node = root
    for part in path:
        if part not in node._modules:
            node.add_module(part, _Node())
        node = node._modules[part]
    if is_buffer:
        node.register_buffer(leaf, tensor)
    else:
        node.register_parameter(leaf, nn.Parameter(tensor, requires_grad=False))


class NativeEngine:
    """One bound plan per input shape, on one device."""

    def __init__(self, plan_ir, nc, device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise N.NativeError('the native engine needs a CUDA device (no CPU fallback on this path)')
        self.nc = nc
        self.plan_ir = plan_ir
        P = self.plan_ir
        self._tensors = (N.TensorDesc * len(P.tensors))(*P.tensors)
        self._ops = (N.OpDesc * len(P.ops))(*P.ops)
        blob = P.blob.bytes()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().ry_plan_create(self._tensors, len(P.tensors), self._ops, len(P.ops), blob.ctypes.data, blob.nbytes, nc,
                                           self.device.index or 0, C.byref(handle)), 'ry_plan_create')
        self.handle = handle
        self.shape = None
        self.workspace = None
        self.n_cand = 0
        self._image_u8 = False
        # CUDA-graph replay of the forward pass (opt-in, Model.cuda_graph): graphs keyed by (input address, dtype, output slot,
        # filter threshold), static output slots handed out round-robin
        self._graphs, self._seen, self._slots, self._slot, self._slot_of = {}, set(), [], 0, {}

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                N.lib().ry_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def bind(self, B, H, W):
        if self.shape == (B, H, W):
            return
        if H % 32 or W % 32:
            raise ValueError(f'image size {H}x{W} must be a multiple of the max stride 32 (reference check_img_size)')
        nbytes = C.c_size_t()
        L = N.lib()
        N.check(L.ry_plan_workspace_bytes(self.handle, B, H, W, C.byref(nbytes)), 'ry_plan_workspace_bytes')
        if self.workspace is None or self.workspace.numel() < nbytes.value:
            self.workspace = None
            self.workspace = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            N.check(L.ry_plan_bind(self.handle, B, H, W, self.workspace.data_ptr(), self.workspace.numel()), 'ry_plan_bind')
        n = C.c_int()
        N.check(L.ry_plan_num_candidates(self.handle, C.byref(n)), 'ry_plan_num_candidates')
        self.n_cand, self.shape = n.value, (B, H, W)
        self._graphs, self._seen, self._slots, self._slot, self._slot_of = {}, set(), [], 0, {}      # graphs bake the old arena / tensor maps

    def launch_count(self):
        n = C.c_int()
        N.check(N.lib().ry_plan_launch_count(self.handle, C.byref(n)), 'ry_plan_launch_count')
        return n.value

    def set_profiling(self, on: bool):
        N.check(N.lib().ry_plan_set_profiling(self.handle, int(on)), 'ry_plan_set_profiling')

    def op_times_ms(self):
        """per-op elapsed milliseconds of the last profiled run (call after synchronising the stream)"""
        n = len(self.plan_ir.ops)
        buf = (C.c_float * n)()
        N.check(N.lib().ry_plan_op_times(self.handle, buf, n), 'ry_plan_op_times')
        return list(buf)

    def tensor(self, t):
        """torch view of plan tensor ``t`` inside the bound workspace: [B, h, w, C] (maps) or [B, C] (vectors)."""
        off, h, w, c, dt = C.c_size_t(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        N.check(N.lib().ry_plan_tensor_info(self.handle, t, C.byref(off), C.byref(h), C.byref(w), C.byref(c), C.byref(dt)),
                'ry_plan_tensor_info')
        dtype = torch.bfloat16 if dt.value == N.RY_BF16 else torch.float32
        esz = 2 if dt.value == N.RY_BF16 else 4
        B = self.shape[0]
        base = (-self.workspace.data_ptr()) % 1024 + off.value
        kind = self.plan_ir.tensors[t].kind
        shape = (B, h.value, w.value, c.value) if kind == N.T_MAP else (B, c.value)
        numel = math.prod(shape)
        return self.workspace[base:base + numel * esz].view(dtype).view(shape)

    def _outputs(self, B, H, W):
        no = self.nc + 5
        pred = torch.empty((B, self.n_cand, no), dtype=torch.float32, device=self.device)
        raws = [torch.empty((B, 3, H >> l, W >> l, no), dtype=torch.float32, device=self.device) for l in (3, 4, 5)]
        return pred, raws

    def _set_image_dtype(self, x):
        u8 = x.dtype == torch.uint8
        if u8 != self._image_u8:
            N.check(N.lib().ry_plan_set_image_dtype(self.handle, N.RY_U8 if u8 else N.RY_F32), 'ry_plan_set_image_dtype')
            self._image_u8 = u8

    def _launch(self, x, pred, raws, mask, conf_filter):
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            if conf_filter is None:
                N.check(N.lib().ry_forward(self.handle, x.data_ptr(), pred.data_ptr(), raws[0].data_ptr(), raws[1].data_ptr(),
                                           raws[2].data_ptr(), C.c_void_p(st)), 'ry_forward')
            else:
                N.check(N.lib().ry_decode_filter(self.handle, x.data_ptr(), C.c_float(float(conf_filter)), pred.data_ptr(),
                                                 raws[0].data_ptr(), raws[1].data_ptr(), raws[2].data_ptr(), mask.data_ptr(),
                                                 C.c_void_p(st)), 'ry_decode_filter')

    def forward(self, x, conf_filter=None, graph=False, graph_slots=2):
        """``conf_filter``: confidence threshold of the fused decode + filter (ry_decode_filter): the Detect epilogue also
        records ``obj > conf_filter`` per candidate as ballot words; ``pred`` then carries them (``pred._ry_cand``) and
        ``non_max_suppression(pred, conf_thres >= conf_filter, ...)`` compacts from the mask instead of re-reading ``pred``.
        ``graph``: replay the ~160 launches of the pass as ONE CUDA graph (PDL edges included).  A graph bakes addresses, so
        the outputs are ``graph_slots`` static buffer sets; every input address keeps the set it was given at first sight (round
        robin), so a returned ``pred`` stays valid until the next call with the same input buffer (or, with more distinct input
        addresses than slots, with one that shares its slot) -- and a graph is captured the second time an input address is seen (a caller that
        feeds a fixed staging buffer, like detect.py's dataloader loop or bench.py, replays from its third call on; fresh
        addresses every call simply run the eager launches)."""
        # the library reinterprets the buffer by element type: anything but contiguous fp32 / uint8 NCHW on this device
        # would be read as garbage (and past its end), so it is an error here, never a silent cast
        if x.dtype not in (torch.float32, torch.uint8):
            raise N.NativeError(f'NativeEngine.forward: image dtype must be float32 or uint8, got {x.dtype}')
        if x.dim() != 4 or x.shape[1] != 3 or not x.is_contiguous():
            raise N.NativeError('NativeEngine.forward: image must be a contiguous [B, 3, H, W] tensor')
        if x.device != self.device:
            raise N.NativeError(f'NativeEngine.forward: image on {x.device}, engine on {self.device}')
        B, _, H, W = x.shape
        self.bind(B, H, W)
        self._set_image_dtype(x)
        new_mask = lambda: torch.empty((B, (self.n_cand + 31) // 32), dtype=torch.int32, device=self.device)
        if not graph:
            pred, raws = self._outputs(B, H, W)
            mask = new_mask() if conf_filter is not None else None
            self._launch(x, pred, raws, mask, conf_filter)
        else:
            while len(self._slots) < graph_slots:
                p_, r_ = self._outputs(B, H, W)
                self._slots.append((p_, r_, new_mask()))
            # an input buffer keeps its output slot (a loop over k fixed staging buffers gets k output sets, no re-capture when
            # the order of the buffers shifts); new addresses take the slots round-robin
            slot = self._slot_of.get(x.data_ptr())
            if slot is None:
                if len(self._slot_of) >= 64:
                    self._slot_of.clear()
                slot = self._slot_of[x.data_ptr()] = self._slot = (self._slot + 1) % graph_slots
            pred, raws, mask = self._slots[slot]
            key = (x.data_ptr(), x.dtype, slot, None if conf_filter is None else float(conf_filter))
            g = self._graphs.get(key)
            if g is not None:
                g.replay()
            elif key in self._seen and len(self._graphs) < 16:
                cur = torch.cuda.current_stream(self.device)
                side = torch.cuda.Stream(self.device)
                side.wait_stream(cur)
                g = torch.cuda.CUDAGraph()
                # launches only (ry_forward neither allocates nor synchronises); thread-local capture: CUDA calls of other host
                # threads (a sampler, a data loader) must not invalidate it
                with torch.cuda.graph(g, stream=side, capture_error_mode='thread_local'):
                    self._launch(x, pred, raws, mask, conf_filter)
                cur.wait_stream(side)
                self._graphs[key] = g
                g.replay()
            else:
                self._seen.add(key)
                self._launch(x, pred, raws, mask, conf_filter)
        if conf_filter is not None:
            pred._ry_cand = (mask, float(conf_filter), pred.data_ptr())
        elif hasattr(pred, '_ry_cand'):
            del pred._ry_cand
        return pred, raws

    def run_ops(self, first, last, image=None, pred=None, raws=(None, None, None)):
        ptr = lambda t: t.data_ptr() if t is not None else None
        if image is not None:
            self._set_image_dtype(image)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            N.check(N.lib().ry_run_ops(self.handle, first, last, ptr(image), ptr(pred), ptr(raws[0]), ptr(raws[1]), ptr(raws[2]),
                                       C.c_void_p(st)), 'ry_run_ops')


class IDetect(_Node):
    """Detect-layer facade (reference models/yolo.py:93-199): attributes + fuseforward on the native head."""
    export = False
    end2end = False
    include_nms = False

    def _configure(self, owner, nc, anchors, ch):
        object.__setattr__(self, '_owner', owner)
        self.nc, self.no, self.nl, self.na = nc, nc + 5, len(anchors), len(anchors[0]) // 2
        self.stride = torch.tensor(arch.STRIDES)
        self.f, self.i, self.type = [62, 63, 64], 65, 'models.yolo.IDetect'

    def fuseforward(self, x):
        """x: list of nl NCHW feature maps.  Mutates the list in place like the reference (yolo.py:140-142)."""
        return self._owner._detect_only(x)

    def _package(self, pred, raws):
        """Output contracts of IDetect.fuseforward (models/yolo.py:158-166): export -> raw head maps, end2end -> pred only,
        include_nms -> convert() (box xyxy via the 4x4 matrix, score = cls * obj; yolo.py:189-199), else (pred, raw list)."""
        if self.export:
            return raws
        if self.end2end:
            return pred
        if self.include_nms:
            box, conf, score = pred[:, :, :4], pred[:, :, 4:5], pred[:, :, 5:]
            score = score * conf
            m = torch.tensor([[1, 0, 1, 0], [0, 1, 0, 1], [-0.5, 0, 0.5, 0], [0, -0.5, 0, 0.5]], dtype=torch.float32, device=pred.device)
            return ((box @ m, score),)
        return (pred, raws)

    forward = fuseforward


class Model(nn.Module):
    # opt-in fused decode + confidence filter (north_star (c)): set to the conf_thres the following non_max_suppression call
    # will use (or lower); pred is still fully materialised, detections are identical (see NativeEngine.forward)
    decode_filter = None
    # opt-in CUDA-graph replay of the forward pass (serving loops that feed a fixed input buffer): see NativeEngine.forward
    cuda_graph = False

    def __init__(self, cfg=None, ch=3, nc=None, anchors=None):
        super().__init__()
        self.traced = False
        if cfg is None:
            cfg = arch.rep_yolo_cfg()
        if not isinstance(cfg, dict):
            import yaml
            with open(cfg) as f:
                cfg = yaml.safe_load(f)
        self.yaml = cfg = dict(cfg)
        cfg['ch'] = cfg.get('ch', ch)
        if nc and nc != cfg['nc']:
            cfg['nc'] = nc
        if anchors:
            cfg['anchors'] = anchors
        self._layers, self.save = arch.parse(cfg, cfg['ch'])
        self.names = [str(i) for i in range(cfg['nc'])]
        self.model = _Node()
        det = self._layers[-1]
        for L in self._layers:                         # one node per reference layer, so model[i] / model[-1] index alike
            self.model.add_module(str(L.i), IDetect() if L.kind == 'IDetect' else _Node())
        g = torch.Generator().manual_seed(0)
        for name, shape in arch.state_shapes(self._layers).items():
            leaf = name.rsplit('.', 1)[-1]
            is_buf = leaf in ('running_mean', 'running_var', 'num_batches_tracked', 'anchors', 'anchor_grid')
            if leaf == 'num_batches_tracked':
                t = torch.zeros((), dtype=torch.long)
            elif leaf == 'running_var' or (leaf == 'weight' and len(shape) == 1):
                t = torch.ones(shape)
            elif leaf == 'anchor_grid':
                t = torch.tensor(cfg['anchors'], dtype=torch.float32).view(shape)
            elif leaf == 'anchors':
                t = torch.tensor(cfg['anchors'], dtype=torch.float32).view(shape) / torch.tensor(arch.STRIDES).view(-1, 1, 1)
            elif len(shape) == 4 and leaf == 'weight':
                bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
                t = torch.empty(shape).uniform_(-bound, bound, generator=g)
            elif leaf == 'implicit':
                t = torch.empty(shape).normal_(0.0, 0.02, generator=g)
            else:
                t = torch.zeros(shape)
            _insert(self.model, name[len('model.'):], t, is_buf)
        self.model[-1]._configure(self, cfg['nc'], cfg['anchors'], det.c1)
        self.stride = self.model[-1].stride
        self._fused = None
        self._engine = None
        self._engines = {}

    # ---- reference API ----
    def fuse(self):
        """Reference Model.fuse (models/yolo.py:681-704): fold RepConv / RepS_Block / Conv+BN / IDetect implicit layers."""
        if self._fused is None:
            with torch.no_grad():
                self._fused = fold.fold_state_dict(self.state_dict(), self._layers)
            self._engine = None
            self._engines = {}
        return self

    def _check_input(self, x):
        """Shared by the plain and the augmented forward: deployed path only, CUDA only, fp32 (any other float type is
        cast, e.g. the .half() image of test.py:104) or uint8 (detect.py:73; the /255 is fused in the stem)."""
        if self._fused is None:
            raise RuntimeError('only the deployed path is built: call .fuse() first (attempt_load does)')
        if not torch.is_tensor(x) or x.dim() != 4:
            raise ValueError('Model.forward: expected a [B, 3, H, W] tensor')
        if not x.is_cuda:
            raise N.NativeError('Model.forward: input must be a CUDA tensor (no CPU fallback on this path)')
        if x.dtype == torch.uint8:
            return x.contiguous()
        if x.dtype != torch.float32 or not x.is_contiguous():
            return x.float().contiguous()
        return x

    def forward(self, x, augment=False, profile=False):
        x = self._check_input(x)
        if augment:
            return self._forward_augment(x)
        pred, raws = self.engine(x.device, (x.shape[0], x.shape[2], x.shape[3])).forward(x, conf_filter=self.decode_filter, graph=self.cuda_graph)
        return self.model[-1]._package(pred, raws)

    def load_state_dict(self, state_dict, strict=True, **kw):
        """New weights invalidate the folded copy and every bound engine (they hold packed bf16 weights): .fuse() again."""
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._fused = None
        self._engine = None
        self._engines = {}
        return out

    def _forward_augment(self, x):
        """Test-time augmentation exactly as the reference (models/yolo.py:570-585, utils/torch_utils.py:247-257): scales
        1 / 0.83 / 0.67 (bilinear, padded with 0.447 to a multiple of the max stride), the middle one left-right flipped;
        boxes de-scaled / de-flipped, predictions concatenated.  The resize / flip of the INPUT is host-side torch plumbing;
        every forward runs in the native engine (one cached engine per shape)."""
        import torch.nn.functional as F
        if x.dtype == torch.uint8:
            x = x.float() / 255.0
        img_size = x.shape[-2:]
        gs = int(self.stride.max())
        y = []
        for si, fi in zip((1, 0.83, 0.67), (None, 3, None)):
            xi = x.flip(fi) if fi else x
            if si != 1:
                h, w = xi.shape[2:]
                s = (int(h * si), int(w * si))
                xi = F.interpolate(xi, size=s, mode='bilinear', align_corners=False)
                h, w = [math.ceil(v * si / gs) * gs for v in (h, w)]
                xi = F.pad(xi, [0, w - s[1], 0, h - s[0]], value=0.447)
            yi = self.engine(xi.device, (xi.shape[0], xi.shape[2], xi.shape[3])).forward(xi.float().contiguous())[0]
            yi[..., :4] /= si
            if fi == 3:
                yi[..., 0] = img_size[1] - yi[..., 0]
            y.append(yi)
        return torch.cat(y, 1), None

    def info(self, verbose=False, img_size=640):
        n_p = sum(p.numel() for p in self.parameters())
        print(f'Model Summary: {len(self._layers)} layers, {n_p} parameters (native B200 deploy path)')

    # ---- native plumbing ----
    def engine(self, device, shape=None):
        """The engine for `device`; with `shape` = (B, H, W) one engine per shape is cached (TTA / alternating shapes would
        otherwise re-bind arena and tensor maps on every call)."""
        device = torch.device(device)
        if self._engine is not None and self._engine.device == device and (shape is None or self._engine.shape in (None, shape)):
            return self._engine
        key = (device, shape)
        eng = self._engines.get(key) if shape is not None else None
        if eng is None:
            eng = NativeEngine(planner.lower(self._layers, self._fused, self.yaml['nc']), self.yaml['nc'], device)
            if shape is not None:
                if self._engine is not None and self._engine.shape is not None:
                    self._engines[(self._engine.device, self._engine.shape)] = self._engine
                while len(self._engines) >= 4:
                    self._engines.pop(next(iter(self._engines)))
                self._engines[key] = eng
        self._engine = eng
        return eng

    def _detect_only(self, xs):
        eng = self.engine(xs[0].device)
        B, H, W = xs[0].shape[0], xs[0].shape[2] * 8, xs[0].shape[3] * 8
        eng.bind(B, H, W)
        grp = eng.plan_ir.groups[-1]
        with torch.cuda.device(eng.device):
            st = torch.cuda.current_stream(eng.device).cuda_stream
            for (_, view), x in zip(grp.inputs, xs):
                t = eng.tensor(view[0])
                if not x.is_cuda or x.device != eng.device:
                    raise N.NativeError('IDetect.fuseforward: feature maps must be CUDA tensors on the engine device')
                x = x.detach()
                if x.dtype != torch.float32 or not x.is_contiguous():
                    x = x.float().contiguous()
                # fp32 NCHW (what the reference's callers hand over) -> the head's NHWC bf16 input view, native kernel
                N.check(N.lib().ry_nchw_to_nhwc_bf16(x.data_ptr(), int(x.shape[0]), int(x.shape[1]), int(x.shape[2]), int(x.shape[3]),
                                                      t.data_ptr(), int(t.shape[-1]), int(view[1]), C.c_void_p(st)), 'ry_nchw_to_nhwc_bf16')
        pred, raws = eng._outputs(B, H, W)
        eng.run_ops(grp.first_op, grp.last_op, pred=pred, raws=raws)
        for i in range(len(xs)):
            xs[i] = raws[i]
        return self.model[-1]._package(pred, xs)
