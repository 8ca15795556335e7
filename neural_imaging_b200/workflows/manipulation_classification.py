"""End-to-end acquisition / distribution workflow with manipulation classification on the B200 path:

    raw -> (nip) -> rgb -> (N manipulations) -> [(downsample) ->] (compression) -> (forensics) -> class probabilities

API mirror of reference workflows/manipulation_classification.py:14-328. The training step (reference :260-285,
tf.GradientTape + Keras Adam) is an explicit forward / backward kernel sequence + one fused Adam launch per model;
with torch.distributed initialised the batch is sharded by rank and the flat gradient buffers are all-reduced.
"""
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib, nn, ops
from ..models import forensics, jpeg, pipelines
from ..tensor import Workspace, as_device, empty, nvtx, ptr, stream, wrap, zeros


class ManipulationClassification(object):

    def __init__(self, nip_model, manipulations=None, distribution=None, fan_args=None, trainable=None,
                 raw_patch_size=128, loss_metric='L2', seed=None):
        if raw_patch_size < 16 or raw_patch_size > 512:
            raise ValueError('The patch size ({}) looks incorrect, typical values should be >= 16 and <= 512'.format(raw_patch_size))
        self._trainable = set() if trainable is None else set(trainable)
        self._trainable.add('fan')
        fan_args = dict(fan_args or {})        # the reference crashes on None (SURVEY 8b "footguns"); we accept it

        if distribution is None:
            self._distribution = {'downsampling': 'pool:2', 'compression': 'jpeg',
                                  'compression_params': {'quality': 50, 'codec': 'soft'}}
        else:
            self._distribution = {}
            self._distribution.update(distribution)

        if ':' in nip_model:
            nip_model, nip_pretrained_dirname = nip_model.split(':')
        else:
            nip_pretrained_dirname = None
        if not (hasattr(pipelines, nip_model) and isinstance(getattr(pipelines, nip_model), type)
                and issubclass(getattr(pipelines, nip_model), pipelines.NIPModel)):
            raise ValueError('Invalid NIP model ({})! Available NIPs: ({})'.format(nip_model, pipelines.supported_models))
        if loss_metric not in ['L2', 'L1', 'SSIM']:
            raise ValueError('Invalid loss metric ({})!'.format(loss_metric))

        self.nip = getattr(pipelines, nip_model)(loss_metric=loss_metric, patch_size=raw_patch_size, seed=seed)
        if nip_pretrained_dirname is not None:
            self.nip.load_model(nip_pretrained_dirname)

        manipulations = manipulations or ['sharpen', 'resample', 'gaussian', 'jpeg']
        self._strengths = {'sharpen': 1, 'resample': 50, 'gaussian': 0.83, 'jpeg': 80, 'awgn': 5.1, 'gamma': 3, 'median': 3}
        self._strengths_range = {'sharpen': (0.25, 1.5), 'resample': (40, 90), 'gaussian': (0.5, 7), 'jpeg': (50, 90),
                                 'awgn': (1, 5), 'gamma': (1, 5), 'median': (3, 9)}
        requested = set()
        for m in manipulations:
            spec = m.split(':')
            requested.add(spec[0])
            if len(spec) > 1:
                self._strengths[spec[0]] = float(spec[-1])
        if any(x not in self._strengths.keys() for x in requested):
            raise ValueError('Unsupported manipulation requested! Available: {}'.format(self._strengths.keys()))

        makers = OrderedDict([('sharpen', ops.SharpenOp), ('resample', ops.ResampleOp), ('gaussian', ops.GaussianOp),
                              ('jpeg', ops.JpegOp), ('awgn', ops.AwgnOp), ('gamma', ops.GammaOp), ('median', ops.MedianOp)])
        self._ops = OrderedDict()
        self._operations = OrderedDict()       # name -> callable(x, strength), as in the reference
        self._forensics_classes = ['native']
        for name, maker in makers.items():     # fixed order, as in the reference (:105-131)
            if name in requested:
                op = maker()
                self._ops[name] = op
                self._operations[name] = (lambda x, strength, _op=op: wrap(_op.forward(as_device(x), empty(as_device(x).shape), strength)))
                self._forensics_classes.append('{}:{}'.format(name, self._strengths[name]))
        assert len(self._forensics_classes) == self.n_classes

        comp = self._distribution['compression']
        if comp == 'jpeg':
            self.codec = jpeg.JPEG(**self._distribution['compression_params'])
        elif comp == 'dcn':
            from ..models import compression
            params = dict(self._distribution.get('compression_params') or {})
            if 'dirname' in params:
                self.codec = compression.TwitterDCN.restore(params['dirname'], key='codec') if params['dirname'] else compression.TwitterDCN()
            else:
                self.codec = compression.TwitterDCN(**params)
        elif comp == 'none':
            self.codec = None
        else:
            raise ValueError('Unsupported channel compression {}'.format(comp))
        if 'dcn' in self._trainable and (self.codec is None or len(self.codec.parameters) == 0):
            raise ValueError('The current codec does not appear to be trainable: {}!'.format(
                None if self.codec is None else self.codec.class_name))

        fan_input_patch = 2 * raw_patch_size // self.downsampling_factor
        self.fan = forensics.FAN(n_classes=self.n_classes, patch_size=fan_input_patch, seed=seed, **fan_args)

        self._stores = [self.fan._store]
        if 'nip' in self._trainable and self.nip._store.trainable:
            self._stores.append(self.nip._store)
        if 'dcn' in self._trainable:
            if comp == 'jpeg':
                # a JPEG codec with trainable tables passes the parameter check of the reference (:141-142), but its loss is
                # MeanSquaredError()(c, C, sample_weight = NaN entropy): every gradient is NaN and the reference's first
                # training step raises RuntimeError('∇ NaNs: ...') (:281-282). Same outcome here, reported where it is decided.
                raise RuntimeError('∇ NaNs: a trainable JPEG codec cannot be optimised through this workflow — its loss term is NaN in the '
                                   'reference (JPEG.loss receives the NaN entropy as sample_weight)')
            self._stores.append(self.codec._store)
        self._parameters = [p for s in self._stores for p in (q.value for q in s.trainable)]
        self._grad_arena = nn.unify_gradients(self._stores)        # one bucket for the data-parallel all-reduce
        self._optimizer = nn.AdamKeras()
        self._ws = Workspace()
        self._stack = ops.PooledStack(self._ops)
        self._fuse_pool = True            # tests switch it off to compare against the operator-by-operator path
        self._labels = {}
        self._use_graph = False
        self._graphs = {}

    # ------------------------------------------------------------------------------------------------ properties
    @property
    def n_classes(self):
        return len(self._operations) + 1

    @property
    def downsampling_factor(self):
        ds = self._distribution['downsampling']
        if ds == 'none':
            return 1
        if ':' in ds:
            return int(ds.split(':')[-1])
        return 2

    def is_trainable(self, model):
        return model in self._trainable

    @property
    def trainable_models(self):
        return tuple(x for x in self._trainable)

    # ------------------------------------------------------------------------------------------------ forward pieces
    def _draw_strengths(self, randomize, override=None):
        override = override if override is not None else self._strengths
        return OrderedDict((name, override[name] if not randomize else np.random.uniform(*self._strengths_range[name]))
                           for name in self._ops)

    def _manipulate(self, Y, strengths, out, training=False):
        """Class-major stack [Y, op1(Y), op2(Y), ...] written slot by slot into `out` (no concat copy)."""
        B = Y.shape[0]
        ops.copy_into(out[:B], Y)
        for i, (name, op) in enumerate(self._ops.items()):
            op.forward(Y, out[(i + 1) * B:(i + 2) * B], strengths[name], training=training)
        return out

    def run_manipulations(self, batch_y, randomize=False, override=None):
        Y = as_device(batch_y)
        out = empty((self.n_classes * Y.shape[0],) + tuple(Y.shape[1:]))
        return wrap(self._manipulate(Y, self._draw_strengths(randomize, override), out))

    def manipulations_timing(self, batch_y):
        Y = as_device(batch_y)
        times = {}
        for name, op in self._ops.items():
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            op.forward(Y, empty(Y.shape), self._strengths[name])
            end.record()
            end.synchronize()
            times[name] = start.elapsed_time(end) / 1000.0
        return times

    def _downsample(self, m, out=None):
        ds, factor = self._distribution['downsampling'], self.downsampling_factor
        if ds.startswith('pool'):
            return ops.avgpool_fwd(m, factor, out=out)
        if ds == 'bilinear':
            n, h, w = ops._nhw3(m)
            s = h // factor                      # the reference uses shape[1] for both dims (:238)
            y = empty((n, s, s, 3)) if out is None else out
            _lib.lib().ni_resize_bilinear_fwd(ptr(m), ptr(y), n, h, w, s, s, 0, stream())
            return y
        if ds == 'none':
            return m
        raise ValueError('Unsupported channel down-sampling {}'.format(ds))

    def run_downsampling(self, batch_y):
        return wrap(self._downsample(as_device(batch_y)))

    def run_compression(self, batch_y, return_entropy=False):
        comp = self._distribution['compression']
        if comp in ('jpeg', 'dcn'):
            return self.codec.process(batch_y, return_entropy=return_entropy)
        if comp == 'none':
            return batch_y
        raise ValueError('Unsupported channel compression {}'.format(comp))

    def run_workflow(self, batch_x, augment=False, training=False):
        """raw -> isp -> manipulations -> downsample -> compression -> fan. Returns (Y, c, C, entropy, probabilities)."""
        batch_Y = self.nip.process(batch_x)
        batch_m = self.run_manipulations(batch_Y, augment)
        batch_c = self.run_downsampling(batch_m)
        if self._distribution['compression'] == 'none':
            batch_C, entropy = batch_c, np.nan
        else:
            batch_C, entropy = self.run_compression(batch_c, True)
        probabilities = self.fan.process(batch_C)
        return batch_Y, batch_c, batch_C, entropy, probabilities

    def run_workflow_to_decisions(self, batch_x):
        return self.run_workflow(batch_x)[-1].numpy().argmax(axis=1)

    def run_rgb_to_fan(self, batch_Y):
        batch_C = self.run_compression(self.run_downsampling(self.run_manipulations(batch_Y)))
        return batch_C if isinstance(batch_C, np.ndarray) else batch_C.numpy()

    def run_rgb_to_probabilities(self, batch_Y):
        batch_C = self.run_compression(self.run_downsampling(self.run_manipulations(batch_Y)))
        return self.fan.process(batch_C).numpy()

    def _batch_labels(self, batch_size):
        return np.concatenate([x * np.ones((batch_size,), dtype=np.int32) for x in range(self.n_classes)])

    def _device_labels(self, batch_size):
        if batch_size not in self._labels:
            self._labels[batch_size] = as_device(self._batch_labels(batch_size), torch.int32)
        return self._labels[batch_size]

    # ------------------------------------------------------------------------------------------------ training step
    def training_step(self, batch_x, batch_y, lambda_nip=0, lambda_dcn=0, augment=False, learning_rate=1e-4):
        """One joint optimisation step. Returns (loss, {'ce','nip','dcn'}) like the reference (:260-285)."""
        loss, parts = self.training_step_device(batch_x, batch_y, lambda_nip, lambda_dcn, augment, learning_rate)
        if self._optimizer.nonfinite():
            raise RuntimeError('∇ NaNs: non-finite gradients detected by the fused Adam kernel')
        return loss, parts

    def release_workspaces(self):
        """Drop every persistent activation / gradient buffer and captured graph (they are re-created on the next step). Workspaces keep
        one buffer per (name, shape) alive so that captured graphs stay valid; call this between very different batch sizes."""
        self._graphs.clear()
        for obj in (self, self.nip, self.fan, self.codec):
            ws_ = getattr(obj, '_ws', None)
            if ws_ is not None:
                ws_.clear()
        torch.cuda.empty_cache()

    def enable_cuda_graph(self, enabled=True):
        """Replay the step as two captured CUDA graphs (forward + backward | Adam + loss scalars) instead of ~1100 separate
        launches. Used when the step is static: augment=False and a fixed codec quality; the first call of a given
        (shapes, lambdas) combination runs eagerly (it sizes the workspaces), the second captures, later calls replay."""
        self._use_graph = bool(enabled)
        if not enabled:
            self._graphs.clear()

    def training_step_device(self, batch_x, batch_y, lambda_nip=0, lambda_dcn=0, augment=False, learning_rate=1e-4,
                             grad_sync=None):
        """The step without the host-side NaN check (no device->host sync). grad_sync(stores) is called between
        backward and Adam (data-parallel all-reduce hook)."""
        world = getattr(grad_sync, 'world', 1)
        if self._use_graph and self._graphable(augment):
            return self._graph_step(batch_x, batch_y, lambda_nip, lambda_dcn, learning_rate, grad_sync)
        x = as_device(batch_x)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        ctx = self._forward_backward(x, as_device(batch_y), lambda_nip, lambda_dcn, augment, world)
        if grad_sync is not None:
            grad_sync(self._stores)
        self._optimizer.lr = float(learning_rate)
        return self._update(ctx, getattr(grad_sync, 'gscale', 1.0))

    # ------------------------------------------------------------------------------------------------ CUDA-graph replay
    def _graphable(self, augment):
        if augment:
            return False            # manipulation strengths are drawn on the host every step
        if self._distribution['compression'] == 'jpeg' and not jpeg.is_number(self.codec.quality):
            return False            # random codec quality per step
        return True

    def _graph_step(self, batch_x, batch_y, lambda_nip, lambda_dcn, learning_rate, grad_sync):
        xin = batch_x if isinstance(batch_x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(batch_x, dtype=np.float32))
        tin = batch_y if isinstance(batch_y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(batch_y, dtype=np.float32))
        if xin.dim() == 3:
            xin = xin.unsqueeze(0)
        world = getattr(grad_sync, 'world', 1)
        gscale = getattr(grad_sync, 'gscale', 1.0)
        key = (tuple(xin.shape), tuple(tin.shape), float(lambda_nip), float(lambda_dcn), world, float(gscale))
        st = self._graphs.get(key)
        if st is None:              # first sight of this configuration: eager step (allocates workspaces / scratch / Adam state)
            self._graphs[key] = 'warm'
            ctx = self._forward_backward(as_device(xin), as_device(tin), lambda_nip, lambda_dcn, False, world)
            if grad_sync is not None:
                grad_sync(self._stores)
            self._optimizer.lr = float(learning_rate)
            return self._update(ctx, gscale)
        if st == 'warm':
            st = self._capture(xin.shape, tin.shape, lambda_nip, lambda_dcn, world, gscale)
            self._graphs[key] = st
        st['x'].copy_(xin, non_blocking=True)
        st['t'].copy_(tin, non_blocking=True)
        opt = self._optimizer
        opt.lr = float(learning_rate)
        opt.iterations += 1
        st['fwd_bwd'].replay()
        if grad_sync is not None:
            grad_sync(self._stores)
        # the bias-corrected step size reaches the captured Adam through a device scalar written by a stream-ordered launch (its value
        # is a kernel argument): the host may run any number of steps ahead without racing a pinned staging buffer
        _lib.lib().ni_fill(ptr(st['lr_dev']), float(opt.step_size()), 1, stream())
        st['update'].replay()
        loss, parts = st['out']
        return wrap(loss.clone()), {k: (wrap(v.clone()) if isinstance(v, torch.Tensor) else v) for k, v in parts.items()}

    def _capture(self, x_shape, t_shape, lambda_nip, lambda_dcn, world, gscale):
        L = _lib.lib()
        st = {'x': empty(tuple(x_shape)), 't': empty(tuple(t_shape)), 'lr_dev': zeros((1,))}
        torch.cuda.synchronize()
        n0 = int(L.ni_launch_count())
        st['fwd_bwd'] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st['fwd_bwd'], capture_error_mode='thread_local'):
            ctx = self._forward_backward(st['x'], st['t'], lambda_nip, lambda_dcn, False, world)
        st['update'] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st['update'], pool=st['fwd_bwd'].pool(), capture_error_mode='thread_local'):
            st['out'] = self._update(ctx, gscale, lr_t_dev=st['lr_dev'])
        st['launches'] = int(L.ni_launch_count()) - n0
        st['ctx'] = ctx
        return st

    @property
    def graph_launches_per_step(self):
        """Kernels of this library inside one replay of the captured step (0 when nothing has been captured)."""
        return max([g['launches'] for g in self._graphs.values() if isinstance(g, dict)] or [0])

    # ------------------------------------------------------------------------------------------------ the step proper
    def _forward_backward(self, x, t, lambda_nip, lambda_dcn, augment, world):
        """Forward + backward of the joint graph; leaves the gradients in the flat buffers of self._stores."""
        L, ws, s = _lib.lib(), self._ws, stream()
        B = int(x.shape[0])
        train_nip = 'nip' in self._trainable and bool(self.nip._store.trainable)
        train_dcn = 'dcn' in self._trainable
        comp = self._distribution['compression']
        # sum-type loss terms (the codec's l2_loss and its global-histogram entropy) are scaled by the world size so that the
        # 1/world average applied to the all-reduced gradients leaves them as the reference's single-process sums
        lam_dcn = float(lambda_dcn) * world if train_dcn else 0.0

        # ---- forward
        with nvtx('nip.forward'):
            Y = self.nip._forward(x, save=train_nip)
        strengths = self._draw_strengths(augment)
        M = self.n_classes * B
        # manipulations + 2x2 average pooling in one pass (no full-resolution stack) whenever the channel allows it
        fused = self._fuse_pool and ops.PooledStack.applicable(self._distribution['downsampling'], Y.shape[1], Y.shape[2])
        _r = nvtx('manipulations+downsample'); _r.__enter__()
        if fused:
            plan = self._stack.plan(strengths, Y.shape[1])
            m = None
            scratch = ws.get('mscratch', Y.shape) if plan['rest'] else None
            c = self._stack.forward(Y, ws.get('c', (M, Y.shape[1] // 2, Y.shape[2] // 2, 3)), plan, strengths, scratch, training=train_nip,
                                    mask=ws.get('gmask', Y.shape[:3], torch.uint8) if train_nip else None)
        else:
            m = ws.get('m', (M,) + tuple(Y.shape[1:]))
            self._manipulate(Y, strengths, m, training=train_nip)
            c = self._downsample(m, out=ws.get('c', (M, -(-Y.shape[1] // self.downsampling_factor), -(-Y.shape[2] // self.downsampling_factor), 3))
                                 if self._distribution['downsampling'].startswith('pool') else None)
        _r.__exit__()
        entropy = acc_dcn = None
        _r = nvtx('codec.forward'); _r.__enter__()
        if comp == 'jpeg':
            C = ws.get('C', c.shape)
            quality = self.codec._draw_quality(None)
            self.codec._with_quality(quality, lambda: self.codec._model.forward_into(c, C))
        elif comp == 'dcn':
            quality = None
            self.codec.set_data_parallel(world)
            C, entropy = self.codec.forward(c, save=(train_nip or train_dcn))
            acc_dcn = ws.get('loss_dcn', (1,))
            L.ni_fill(ptr(acc_dcn), 0.0, 1, s)
            L.ni_image_loss(ptr(c), ptr(C), ptr(acc_dcn), c.numel(), 0, s)
        else:
            C, quality = c, None
        _r.__exit__()
        labels = self._device_labels(B)
        with nvtx('fan.forward+loss'):
            probs, loss_ce, dlogits = self.fan.forward_loss(C, labels)
        acc = ws.get('loss_nip', (1,))
        L.ni_fill(ptr(acc), 0.0, 1, s)
        self.nip.loss_forward(Y, t, acc, float(lambda_nip))

        # ---- backward
        codec_bwd = comp == 'dcn' and (train_nip or train_dcn)
        with nvtx('fan.backward'):
            dC = self.fan.backward(dlogits, need_dx=train_nip or codec_bwd)
        if codec_bwd:
            # d/dC of lambda_dcn * (l2_loss(c - C) + w * H); the entropy enters through the latent inside codec.backward
            if lam_dcn:
                L.ni_image_loss_grad(ptr(C), ptr(c), ptr(dC), C.numel(), 0, lam_dcn * C.numel() / (2.0 * 255.0 * 255.0), 1, s)
            dc = self.codec.backward(dC, lam_dcn * float(self.codec._h.entropy_weight), need_dx=train_nip, need_dw=train_dcn)
            if train_nip and lam_dcn:           # c is also the TARGET of the codec loss
                L.ni_image_loss_grad(ptr(c), ptr(C), ptr(dc), c.numel(), 0, lam_dcn * c.numel() / (2.0 * 255.0 * 255.0), 1, s)
        if train_nip:
            if comp == 'jpeg':
                dc = ws.get('dc', c.shape)
                self.codec._with_quality(quality, lambda: self.codec._model.backward(c, dC, dc))
            elif comp != 'dcn':
                dc = dC
            dY = ws.get('dY', Y.shape)
            # dY = lambda_nip * d(nip loss)/dY + native slot + manipulation branches
            self.nip.loss_backward(Y, t, dY, float(lambda_nip))
            if fused:
                self._stack.backward(Y, dc, dY, plan, strengths, scratch)
            else:
                dm = self._downsample_bwd(dc, m.shape, ws)
                L.ni_axpy(ptr(dY), ptr(dm[:B]), 1.0, dY.numel(), s)
                for i, (name, op) in enumerate(self._ops.items()):
                    if op.has_grad:
                        op.backward(Y, dm[(i + 1) * B:(i + 2) * B], dY, strengths[name])
            with nvtx('nip.backward'):
                self.nip._backward(dY)
        return {'M': M, 'Y_numel': Y.numel(), 'loss_ce': loss_ce, 'acc': acc, 'acc_dcn': acc_dcn, 'entropy': entropy,
                'comp': comp, 'train_dcn': train_dcn, 'lambda_nip': float(lambda_nip), 'lambda_dcn': float(lambda_dcn)}

    def _update(self, ctx, gscale=1.0, lr_t_dev=None):
        """Fused Adam over the flat buffers + the step's loss scalars (device tensors)."""
        with nvtx('adam'):
            self._optimizer.apply(self._stores, gscale=gscale, lr_t_dev=lr_t_dev)
        comp = ctx['comp']
        loss_ce_v = ctx['loss_ce'] / float(ctx['M'])
        loss_nip_v = ctx['acc'] / float(ctx['Y_numel'])
        loss = loss_ce_v + (ctx['lambda_nip'] * loss_nip_v if 'nip' in self._trainable else 0.0)
        if comp == 'dcn':
            loss_dcn = ctx['acc_dcn'] / (2.0 * 255.0 * 255.0) + float(self.codec._h.entropy_weight) * ctx['entropy']
            if ctx['train_dcn']:
                loss = loss + ctx['lambda_dcn'] * loss_dcn
            loss_dcn = wrap(loss_dcn.reshape(()))
        else:
            loss_dcn = float('nan') if comp == 'jpeg' else 0.0    # Keras MSE with sample_weight = NaN entropy
        return wrap(loss.reshape(())), {'ce': wrap(loss_ce_v.reshape(())), 'nip': wrap(loss_nip_v.reshape(())), 'dcn': loss_dcn}

    def _downsample_bwd(self, dc, m_shape, ws):
        ds, factor = self._distribution['downsampling'], self.downsampling_factor
        if ds.startswith('pool'):
            return ops.avgpool_bwd(dc, m_shape, factor, out=ws.get('dm', m_shape))
        if ds == 'bilinear':
            dm = ws.get('dm', m_shape)
            L = _lib.lib()
            L.ni_fill(ptr(dm), 0.0, dm.numel(), stream())
            L.ni_resize_bilinear_bwd(ptr(dc), ptr(dm), m_shape[0], m_shape[1], m_shape[2], dc.shape[1], dc.shape[2], 1.0, stream())
            return dm
        return dc

    # ------------------------------------------------------------------------------------------------ summaries
    def summary_compact(self):
        return '{class_name}[{trainables}]: {nip} -> [{manips}] {pool}{codec}-> {fan}'.format(
            class_name=type(self).__name__, nip=self.nip.class_name, manips=''.join([x[0] for x in self._forensics_classes]),
            trainables=''.join([x[0] for x in self.trainable_models]),
            pool='' if self._distribution['downsampling'] == 'none' else '-> {} '.format(self._distribution['downsampling']),
            codec='' if self.codec is None else '-> {} '.format(self.codec.summary_compact()), fan='FAN')

    def summary(self):
        return '{class_name}[opt={trainables}]: {input} -> {nip} -> {n_ops} manipulations [{manips}] {pool}{codec}-> {fan}'.format(
            class_name=type(self).__name__, input='(rgb)' if self.nip.x.shape[-1] == 3 else '(raw)', nip=self.nip.class_name,
            n_ops=self.n_classes - 1, manips=''.join([x[0] for x in self._forensics_classes]),
            trainables=''.join([x[0] for x in self.trainable_models]),
            pool='' if self._distribution['downsampling'] == 'none' else '-> {} '.format(self._distribution['downsampling']),
            codec='' if self.codec is None else '-> {} '.format(self.codec.summary_compact()),
            fan='FAN -> (prob. {} classes)'.format(self.n_classes))

    def details(self):
        out = [self.summary()]
        out.append('Input         : {} {}'.format(self.nip.x.shape, '(rgb)' if self.nip.x.shape[-1] == 3 else '(raw)'))
        out.append('Camera ISP    : {}'.format(self.nip.summary()))
        out.append('Manipulations : {} -> {}'.format(self.n_classes, self._forensics_classes))
        out.append('Downsampling  : {}'.format(self._distribution['downsampling']))
        out.append('Codec         : {}'.format('' if self.codec is None else self.codec.summary()))
        out.append('Forensics     : {}'.format(self.fan.summary()))
        out.append('Output        : {}'.format(self.fan.y.shape))
        return '\n'.join(out)
