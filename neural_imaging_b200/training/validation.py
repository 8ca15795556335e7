"""Validation helpers of the manipulation-classification loop on the B200 path (API mirror of reference
training/validation.py:163-203 `validate_fan`, :96-160 `validate_nip`, :19-41 `validate_jpeg`, :44-93 `validate_dcn`, without the figure rendering).

The reference runs batches of 10 through `flow.run_workflow_to_decisions` and builds the confusion matrix on the host with
an n_classes^2 Python loop per batch; here whole groups of batches go through the device in one pass, the decisions and the matrix
are accumulated by a kernel, and the host reads the result once. PSNR / loss of the NIP are computed on the host from the developed images exactly as the reference does; SSIM
(skimage structural_similarity in the reference, helpers/metrics.py:9-26) runs in the fused device kernel ni_ssim.
"""
import json
import os
from collections import OrderedDict

import numpy as np

from ..helpers import metrics


GROUP_IMAGES = 160          # validation images per device pass (x n_classes manipulated copies each): bounds the activation memory


def validate_fan(flow, data, get_labels=False):
    """Confusion matrix of the FAN inside a ManipulationClassification workflow. Returns (accuracy, conf[, labels]).

    Same samples, same numbers as the reference (batches of 10, `count_validation // 10` of them, matrix normalised by the number of
    validation images, accuracy = mean of the per-batch accuracies) — but the batches are stacked into groups of up to GROUP_IMAGES
    images per device pass, decisions and the int32 confusion matrix stay on the device (`ni_confusion_accumulate`), and the host reads
    the matrix (and, on request, the decisions) once at the end instead of once per batch."""
    import torch
    from .. import _lib
    from ..tensor import as_device, ptr, stream, zeros
    batch_size = int(np.minimum(10, data.count_validation))
    n_batches = data.count_validation // batch_size
    n_classes = flow.n_classes
    L = _lib.lib()
    conf_dev = zeros((n_classes, n_classes), torch.int32)
    per_group = max(1, GROUP_IMAGES // batch_size)
    preds = []
    for first in range(0, n_batches, per_group):
        ids = range(first, min(first + per_group, n_batches))
        parts = []
        for b in ids:
            bx = data.next_validation_batch(b, batch_size)
            parts.append(np.asarray(bx[0] if isinstance(bx, tuple) else bx, dtype=np.float32))
        k = len(parts)
        probs = flow.run_workflow(np.concatenate(parts, axis=0))[-1]            # class-major over the whole group: (n_classes * k * bs, n_classes)
        labels = as_device(np.repeat(np.arange(n_classes, dtype=np.int32), k * batch_size), torch.int32)
        pred = zeros((n_classes * k * batch_size,), torch.int32) if get_labels else None
        L.ni_confusion_accumulate(ptr(probs), ptr(labels), ptr(conf_dev), ptr(pred), n_classes * k * batch_size, n_classes, stream())
        if get_labels:
            preds.append((pred, k))
    counts = conf_dev.cpu().numpy().astype(np.float64)                         # the one device -> host read
    total = n_batches * batch_size
    conf = counts / total
    accuracy = float(np.trace(counts) / (n_classes * total))                    # equal-size batches: mean of batch accuracies = overall accuracy
    if get_labels:
        out_labels = []
        for pred, k in preds:                                                  # back to the reference's order: batch by batch, class-major inside
            p = pred.cpu().numpy().reshape(n_classes, k, batch_size)
            out_labels += [int(v) for b in range(k) for v in p[:, b, :].reshape(-1)]
        return accuracy, conf, out_labels
    return accuracy, conf


NIP_GROUP = 16      # validation images developed per device pass (the reference develops them one by one and renders a figure)


def validate_nip(model, data, save_dir=None, epoch=0, show_ref=False, loss_type='L2'):
    """Per-image (ssim, psnr, loss) lists of a NIP on the validation set (reference training/validation.py:96-160 without the figure).
    The images go through the ISP in groups; squared / absolute error and SSIM are reduced per image on the device, one host read per group."""
    if loss_type not in ('L1', 'L2'):
        raise ValueError('Invalid loss! Use either L1 or L2.')
    import torch
    from ..tensor import as_device
    n = int(data.count_validation)
    ssims, psnrs, losss = [], [], []

    def run(batch_id, size):
        example_x, example_y = data.next_validation_batch(batch_id, size)
        developed = as_device(model.process(example_x)).clamp(0.0, 1.0)
        reference = as_device(np.asarray(example_y, dtype=np.float32))
        if reference.dim() == 3:
            reference = reference.unsqueeze(0)
        diff = (reference - developed).double()
        stats = torch.stack((diff.pow(2).mean(dim=(1, 2, 3)), diff.abs().mean(dim=(1, 2, 3))), dim=1).cpu().numpy()
        ss = np.atleast_1d(metrics.ssim(reference, developed) if size > 1 else metrics.ssim(reference[0], developed[0]))
        for i in range(size):
            mse = float(stats[i, 0])
            psnrs.append(float(10.0 * np.log10(1.0 / mse)) if mse > 0 else float('inf'))
            ssims.append(float(ss[i]))
            losss.append(mse if loss_type == 'L2' else float(stats[i, 1]))

    full = n // NIP_GROUP
    for g in range(full):
        run(g, NIP_GROUP)
    for b in range(full * NIP_GROUP, n):
        run(b, 1)
    return ssims, psnrs, losss


def validate_jpeg(jpeg, data, batch_size=1):
    """Mean psnr / ssim / entropy of a JPEG codec model on the validation set (reference training/validation.py:19-41)."""
    from ..models.jpeg import JPEG
    if not isinstance(jpeg, JPEG):
        raise ValueError('Codec needs to be as instance of {} but is {}'.format(JPEG, getattr(jpeg, 'class_name', type(jpeg).__name__)))
    batch_size = int(np.minimum(batch_size, data.count_validation))
    n_batches = data.count_validation // batch_size
    results = {k: [] for k in ('psnr', 'ssim', 'entropy')}
    for batch_id in range(n_batches):
        batch_x = data.next_validation_batch(batch_id, batch_size)
        if isinstance(batch_x, tuple):
            batch_x = batch_x[-1]
        batch_y, entropy = jpeg.process(batch_x, return_entropy=True)
        batch_y = batch_y.numpy() if hasattr(batch_y, 'numpy') else np.asarray(batch_y)
        results['ssim'].append(metrics.batch(batch_x, batch_y, metrics.ssim))
        results['psnr'].append(metrics.batch(batch_x, batch_y, metrics.psnr))
        results['entropy'].append(entropy)
    return {k: float(np.mean(v)) for k, v in results.items()}


def validate_dcn(dcn, data, save_dir=None, epoch=0, show_ref=False):
    """Validation metrics of a learned codec: {'ssim', 'psnr', 'loss', 'entropy'} over the whole validation set in one batch
    (reference training/validation.py:44-93 without the figure; returns None for anything that is not a DCN, like the reference)."""
    from ..models.compression import DCN
    if not isinstance(dcn, DCN):
        return None
    batch_x = data.next_validation_batch(0, data.count_validation)
    if isinstance(batch_x, tuple):
        batch_x = batch_x[-1]
    batch_x = np.asarray(batch_x)
    batch_y, entropy = dcn.process(batch_x, return_entropy=True)
    entropy = float(np.asarray(entropy.numpy()).reshape(-1)[0])
    out = batch_y.numpy()
    ssim = np.atleast_1d(metrics.ssim(batch_x, out)).tolist()
    psnr = np.atleast_1d(metrics.psnr(batch_x, out)).tolist()
    loss = float(dcn.loss(batch_x, batch_y, entropy).numpy())
    return {'ssim': float(np.mean(ssim)), 'psnr': float(np.mean(psnr)), 'loss': loss, 'entropy': entropy}


def save_training_progress(training_summary, flow, root_dir, quiet=False):
    """Write `<root_dir>/training.json`: summary, channel, classes and per-model (nip / forensics / codec) constructor + hyper-parameters +
    performance logs — the file the reference's result analysis and resume logic read (reference training/validation.py:301-352)."""
    training = OrderedDict()
    training['summary'] = training_summary
    training['distribution'] = flow._distribution
    training['manipulations'] = flow._forensics_classes
    training['nip'] = OrderedDict()
    training['nip']['model'] = flow.nip.class_name
    training['nip']['init'] = repr(flow.nip)
    training['nip']['args'] = flow.nip._h.to_json() if hasattr(flow.nip, '_h') else {}
    training['nip']['performance'] = flow.nip.performance
    training['forensics'] = OrderedDict()
    training['forensics']['model'] = flow.fan.class_name
    training['forensics']['init'] = repr(flow.fan)
    training['forensics']['args'] = flow.fan._h.to_json()
    training['forensics']['performance'] = flow.fan.performance
    codec = getattr(flow, 'codec', None)
    if codec is not None:
        training['codec'] = OrderedDict()
        training['codec']['model'] = codec.class_name
        training['codec']['init'] = repr(codec)
        if hasattr(codec, '_h'):
            training['codec']['args'] = codec._h.to_json()
        if hasattr(codec, 'performance'):
            training['codec']['performance'] = codec.performance
    os.makedirs(root_dir, exist_ok=True)
    filename = os.path.join(root_dir, 'training.json')
    if not quiet:
        print('> Training progress --> {}'.format(filename))
    with open(filename, 'w') as f:
        json.dump(training, f, indent=4, default=str)       # anything exotic in a user-supplied channel description is logged as text
    return filename
