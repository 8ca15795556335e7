"""Op-for-op PyTorch-CPU restatement of the reference operators (TEST INFRASTRUCTURE — see oracle/__init__.py).

Every function cites the reference file:line it follows. Tensors are NHWC like TensorFlow's. `dtype` selects float32
(reference behaviour) or float64 (truth for tolerance budgeting). Gradients come from torch autograd, with
stop_gradient -> .detach().
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

TWO_PI = 2 * np.pi


# ------------------------------------------------------------------------------------------------ layout helpers
def space_to_depth(x, b):
    """tf.nn.space_to_depth, NHWC, block-major channel order."""
    n, h, w, c = x.shape
    x = x.reshape(n, h // b, b, w // b, b, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, h // b, w // b, b * b * c)


def depth_to_space(x, b):
    n, h, w, c = x.shape
    co = c // (b * b)
    x = x.reshape(n, h, w, b, b, co).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, h * b, w * b, co)


def tf_pad(x, p, mode):
    """tf.pad on H and W of an NHWC tensor. mode: 'REFLECT' | 'SYMMETRIC' | 'CONSTANT'."""
    if p == 0:
        return x
    xc = x.permute(0, 3, 1, 2)
    if mode == 'REFLECT':
        y = F.pad(xc, (p, p, p, p), mode='reflect')
    elif mode == 'SYMMETRIC':
        idx_h = torch.tensor([(-i - 1 if i < 0 else (2 * x.shape[1] - 1 - i if i >= x.shape[1] else i)) for i in range(-p, x.shape[1] + p)])
        idx_w = torch.tensor([(-i - 1 if i < 0 else (2 * x.shape[2] - 1 - i if i >= x.shape[2] else i)) for i in range(-p, x.shape[2] + p)])
        y = xc[:, :, idx_h][:, :, :, idx_w]
    else:
        y = F.pad(xc, (p, p, p, p))
    return y.permute(0, 2, 3, 1)


def same_pads(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d(x, w, b=None, stride=1, padding='SAME'):
    """tf.nn.conv2d / Keras Conv2D: NHWC input, HWIO kernel, TF SAME rule (asymmetric, extra on bottom/right)."""
    xc = x.permute(0, 3, 1, 2)
    if padding == 'SAME':
        pt, pb = same_pads(x.shape[1], w.shape[0], stride)
        pl, pr = same_pads(x.shape[2], w.shape[1], stride)
        xc = F.pad(xc, (pl, pr, pt, pb))
    y = F.conv2d(xc, w.permute(3, 2, 0, 1), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_2x2(x, w1x1, b):
    """Keras Conv2DTranspose(filters, 2, 2, 'SAME'): out[2i+a,2j+b,f] = sum_c in[i,j,c] K[a,b,f,c] + bias[f].
    w1x1 is the product's storage layout (1,1,cin,4*cout) with column (a*2+b)*cout+f."""
    cin = w1x1.shape[2]
    cout = w1x1.shape[3] // 4
    wt = w1x1.reshape(cin, 2, 2, cout).permute(0, 3, 1, 2)           # torch layout (cin, cout, kh, kw)
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, b, stride=2)
    return y.permute(0, 2, 3, 1)


def max_pool(x, same):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2, ceil_mode=bool(same)).permute(0, 2, 3, 1)


def avg_pool(x, k):
    """tf.nn.avg_pool k x k stride k SAME (workflows/manipulation_classification.py:235): padded cells are excluded."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), k, k, ceil_mode=True, count_include_pad=False).permute(0, 2, 3, 1)


def leaky_relu(x):
    return F.leaky_relu(x, 0.2)       # helpers/tf_helpers.py:23


ACT = {'leaky_relu': leaky_relu, 'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid, None: (lambda v: v)}


def ste_clip(y, lo=0.0, hi=1.0):
    """y = stop_gradient(clip(y) - y) + y (models/pipelines.py:223)."""
    return (y.clamp(lo, hi) - y).detach() + y


# ------------------------------------------------------------------------------------------------ JPEG tables
def jpeg_qtable(quality, channel=0):
    """compression/jpeg_helpers.py:264-305."""
    quality = np.maximum(np.minimum(100, quality), 1)
    quality = 5000 / quality if quality < 50 else 200 - quality * 2
    if channel == 0:
        t = np.array([[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55], [14, 13, 16, 24, 40, 57, 69, 56],
                      [14, 17, 22, 29, 51, 87, 80, 62], [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
                      [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]], np.float32)
    else:
        t = np.array([[17, 18, 24, 47, 99, 99, 99, 99], [18, 21, 26, 66, 99, 99, 99, 99], [24, 26, 56, 99, 99, 99, 99, 99],
                      [47, 66, 99, 99, 99, 99, 99, 99]] + [[99] * 8] * 4, np.float32)
    t = np.floor((t * quality + 50) / 100)
    t[t < 1] = 1
    t[t > 255] = 255
    return t


def zigzag(n):
    """compression/jpeg_helpers.py:253-261 (sort cells by anti-diagonal, alternating direction)."""
    def key(xy):
        x, y = xy
        return (x + y, -y if (x + y) % 2 else y)
    zz = np.zeros((n, n), dtype=np.uint16)
    for i, (x, y) in enumerate(sorted(((x, y) for x in range(n) for y in range(n)), key=key)):
        zz[x, y] = i
    return zz


def jpeg_qf_estimation(q_mtx, channel=0):
    return int(np.argmin([np.mean(np.abs(jpeg_qtable(qf, channel) - q_mtx)) for qf in range(1, 101)])) + 1


DCT_F = np.array([[0.3536, 0.3536, 0.3536, 0.3536, 0.3536, 0.3536, 0.3536, 0.3536],
                  [0.4904, 0.4157, 0.2778, 0.0975, -0.0975, -0.2778, -0.4157, -0.4904],
                  [0.4619, 0.1913, -0.1913, -0.4619, -0.4619, -0.1913, 0.1913, 0.4619],
                  [0.4157, -0.0975, -0.4904, -0.2778, 0.2778, 0.4904, 0.0975, -0.4157],
                  [0.3536, -0.3536, -0.3536, 0.3536, 0.3536, -0.3536, -0.3536, 0.3536],
                  [0.2778, -0.4904, 0.0975, 0.4157, -0.4157, -0.0975, 0.4904, -0.2778],
                  [0.1913, -0.4619, 0.4619, -0.1913, -0.1913, 0.4619, -0.4619, 0.1913],
                  [0.0975, -0.2778, 0.4157, -0.4904, 0.4904, -0.4157, 0.2778, -0.0975]], dtype=np.float32)   # models/jpeg.py:78-85
COLOR_F = np.array([[0, 0.299, 0.587, 0.114], [128, -0.168736, -0.331264, 0.5], [128, 0.5, -0.418688, -0.081312]], dtype=np.float32)
COLOR_I = np.array([[-1.402 * 128, 1, 0, 1.402], [1.058272 * 128, 1, -0.344136, -0.714136], [-1.772 * 128, 1, 1.772, 0]], dtype=np.float32)


# ------------------------------------------------------------------------------------------------ quantisation
def quantization(x, rounding):
    """models/layers.py:118-136 ('harmonic': taylor_terms receives 1 positionally, so only the first term is active)."""
    c = torch.tensor(TWO_PI, dtype=x.dtype)     # TF casts the python scalar to the tensor dtype (float32 in the reference; exact in the float64 truth run)
    if rounding == 'round':
        return torch.round(x)
    if rounding == 'sin':
        return x - torch.sin(c * x) / c
    if rounding == 'soft':
        x_ = x - torch.sin(c * x) / c
        return (torch.round(x) - x_).detach() + x_
    if rounding == 'harmonic':
        return x - torch.sin(c * x) / (c / 2)
    if rounding == 'identity':
        return x
    raise ValueError('Unsupported quantization: {}'.format(rounding))


def soft_quantization(x, alpha=255):
    """helpers/tf_helpers.py:271-277."""
    c = torch.tensor(TWO_PI, dtype=x.dtype)
    x = alpha * x
    x_ = x - torch.sin(c * x) / c
    return ((torch.round(x) - x_).detach() + x_) / alpha


# ------------------------------------------------------------------------------------------------ dJPEG
def djpeg(inputs, q_luma, q_chroma, rounding='soft'):
    """DifferentiableJPEG.call, models/jpeg.py:91-159, statement by statement. Returns (y, X)."""
    dt = inputs.dtype
    n, h, w, _ = inputs.shape
    bs = 8
    cF = torch.tensor(COLOR_F, dtype=dt)
    cI = torch.tensor(COLOR_I, dtype=dt)
    dF = torch.tensor(DCT_F, dtype=dt)
    dI = dF.t()
    xc = torch.cat((torch.ones(n, h, w, 1, dtype=dt), 255.0 * inputs), dim=3)                     # :99
    ycbcr = xc @ cF.t()                                                                          # :100 (1x1 conv)
    p = (ycbcr - 127).permute(0, 3, 1, 2)                                                        # :105
    p = p.reshape(-1, h, w).unsqueeze(3)                                                         # :106-107
    p = space_to_depth(p, bs)                                                                    # :108 -> (3n, h/8, w/8, 64)
    p = p.permute(0, 3, 1, 2)                                                                    # :109
    p = p.reshape(-1, bs, bs, p.shape[2] * p.shape[3])                                           # :110
    r = p.permute(0, 3, 1, 2)                                                                    # :113
    r = r.reshape(-1, bs, bs)                                                                    # :114
    X = dF.unsqueeze(0) @ r                                                                      # :118
    X = X @ dI.unsqueeze(0)                                                                      # :119
    nb = p.shape[-1]
    # tables: numpy arrays (fixed codec) or tensors carrying gradient (trainable=True: the tables are weights, :58-62)
    as_t = lambda q: q.to(dt) if isinstance(q, torch.Tensor) else torch.tensor(np.asarray(q, np.float32), dtype=dt)
    Ql = as_t(q_luma).unsqueeze(0).repeat(nb, 1, 1)                                              # :125
    Qc = as_t(q_chroma).unsqueeze(0).repeat(2 * nb, 1, 1)
    Q = torch.cat((Ql, Qc), dim=0).repeat(n, 1, 1)                                               # :127-128
    X = X / Q
    X = quantization(X, rounding)
    X = X * Q                                                                                    # :129-131
    xi = dI.unsqueeze(0) @ X
    xi = xi @ dF.unsqueeze(0)                                                                    # :135-136
    xi = xi.reshape(3 * n, -1, bs, bs).permute(0, 2, 3, 1)                                       # :140-141
    q = xi.reshape(-1, bs * bs, h // bs, w // bs).permute(0, 2, 3, 1)                            # :145-147
    q = depth_to_space(q, bs)                                                                    # :148
    q = q.reshape(-1, 3, h, w).permute(0, 2, 3, 1)                                               # :149-150
    qc = torch.cat((torch.ones(n, h, w, 1, dtype=dt), q + 127), dim=3)                           # :154
    y = (qc @ cI.t()) / 255.0                                                                    # :155-156
    return y.clamp(0, 1), X                                                                      # :157 (real clip: zero grad outside)


# ------------------------------------------------------------------------------------------------ manipulations
def gkern(kernlen=5, std=0.83):
    """helpers/kernels.py:94-98 (scipy.signal.gaussian == exp(-n^2 / 2 sigma^2), symmetric window)."""
    n = np.arange(0, kernlen) - (kernlen - 1.0) / 2.0
    g = np.exp(-n ** 2 / (2 * std * std))
    g2 = np.outer(g, g)
    return g2 / g2.sum()


def repeat_2dfilter(f, channels=3):
    rf = np.zeros((f.shape[0], f.shape[1], channels, channels))
    for r in range(channels):
        rf[:, :, r, r] = f
    return rf


def rgb_to_hsv(x):
    """tensorflow/core/kernels/colorspace_op.h (RGBToHSV functor), element for element."""
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    v = torch.maximum(r, torch.maximum(g, b))
    rng = v - torch.minimum(r, torch.minimum(g, b))
    s = torch.where(v > 0, rng / v, torch.zeros_like(v))
    norm = (1.0 / rng) * (1.0 / 6.0)
    hh = torch.where(r == v, norm * (g - b), torch.where(g == v, norm * (b - r) + 2.0 / 6.0, norm * (r - g) + 4.0 / 6.0))
    hh = torch.where(rng > 0, hh, torch.zeros_like(hh))
    hh = torch.where(hh < 0, hh + 1, hh)
    return torch.stack((hh, s, v), dim=-1)


def hsv_to_rgb(x):
    h, s, v = x[..., 0], x[..., 1], x[..., 2]
    dh = h * 6
    dr = ((dh - 3).abs() - 1).clamp(0, 1)
    dg = (-(dh - 2).abs() + 2).clamp(0, 1)
    db = (-(dh - 4).abs() + 2).clamp(0, 1)
    one_s = -s + 1
    return torch.stack(((one_s + s * dr) * v, (one_s + s * dg) * v, (one_s + s * db) * v), dim=-1)


def manipulation_sharpen(x, strength=1, hsv=True, tf_version_21=True):
    """helpers/tf_helpers.py:156-184. With tf_version_21 the HSV conversions are NotDifferentiable (TF 2.1
    python/ops/image_ops_impl.py), i.e. no gradient reaches x through this branch."""
    gk = np.array([[-0.0833, -0.1667, -0.0833], [-0.1667, 0, -0.1667], [-0.0833, -0.1667, -0.0833]])
    gk = strength * gk / np.abs(gk.sum())
    gk[1, 1] = strength + 1
    gf = repeat_2dfilter(gk, 3)
    if hsv:
        gf[:, :, 1:2, 1:2] = 0
        gf[2, 2, 1:2, 1:2] = 1
    gkk = torch.tensor(gf.astype(np.float32), dtype=x.dtype)
    y = tf_pad(x, 1, 'SYMMETRIC')
    if hsv:
        y = rgb_to_hsv(y)
        if tf_version_21:
            y = y.detach()
    y = conv2d(y, gkk, padding='VALID')
    if hsv:
        y = hsv_to_rgb(y)
        if tf_version_21:
            y = y.detach()
    return y.clamp(0, 1)


def resize_bilinear(x, oh, ow):
    """tf.image.resize(method='bilinear') TF2: half-pixel centres, no antialias (core/kernels/image_resizer_state.h)."""
    n, ih, iw, c = x.shape

    def weights(o, i):
        scale = np.float32(i) / np.float32(o)
        src = (np.arange(o, dtype=np.float32) + np.float32(0.5)) * scale - np.float32(0.5)
        fl = np.floor(src)
        lo = np.maximum(fl.astype(np.int64), 0)
        hi = np.minimum(np.ceil(src).astype(np.int64), i - 1)
        return torch.tensor(lo), torch.tensor(hi), torch.tensor((src - fl).astype(np.float32)).to(x.dtype)
    y0, y1, ly = weights(oh, ih)
    x0, x1, lx = weights(ow, iw)
    top, bot = x[:, y0], x[:, y1]
    lx = lx.view(1, 1, -1, 1)
    ly = ly.view(1, -1, 1, 1)
    t = top[:, :, x0] + (top[:, :, x1] - top[:, :, x0]) * lx
    b = bot[:, :, x0] + (bot[:, :, x1] - bot[:, :, x0]) * lx
    return t + (b - t) * ly


def manipulation_resample(x, factor=50):
    """helpers/tf_helpers.py:68-76 (shape[1] is used for both dimensions)."""
    if 0 < factor <= 1:
        factor = 100 * factor
    s = x.shape[1] * int(factor) // 100
    return resize_bilinear(resize_bilinear(x, s, s), x.shape[1], x.shape[1])


def manipulation_gaussian(x, kernel, std, skip_clip=False):
    """helpers/tf_helpers.py:113-125."""
    kernel = int(kernel)
    gf = np.zeros((kernel, kernel, 3, 3))
    gk = gkern(kernel, std)
    for r in range(3):
        gf[:, :, r, r] = gk
    y = conv2d(tf_pad(x, kernel // 2, 'REFLECT'), torch.tensor(gf.astype(np.float32), dtype=x.dtype), padding='VALID')
    return y if skip_clip else y.clamp(0, 1)


def manipulation_awgn(x, strength, noise):
    """helpers/tf_helpers.py:79-82 with the N(0,1) tensor injected (tf.random.normal is not reproducible)."""
    return soft_quantization(x + strength * noise).clamp(0, 1)


def manipulation_gamma(x, strength=2.0):
    """helpers/tf_helpers.py:85-88."""
    return torch.pow(soft_quantization(torch.pow(x, strength)).clamp(1.0 / 255, 1), 1 / strength)


def manipulation_median(x, kernel=3):
    """helpers/tf_helpers.py:91-110: REFLECT pad, k*k patches, top_k, element (area+1)//2 - 1."""
    kernel = int(kernel)
    if kernel % 2 == 0:
        kernel += 1
    kernel = max(kernel, 1)
    xp = tf_pad(x, kernel // 2, 'REFLECT')
    n, h, w, c = x.shape
    patches = xp.permute(0, 3, 1, 2).unfold(2, kernel, 1).unfold(3, kernel, 1)          # n, c, h, w, k, k
    patches = patches.reshape(n, c, h, w, kernel * kernel).permute(0, 2, 3, 1, 4)
    area = kernel ** 2
    floor = (area + 1) // 2
    ceil = area // 2 + 1
    top = torch.topk(patches, ceil, dim=-1, sorted=True).values
    return top[..., floor - 1]


def jpeg_manipulation(x, quality):
    """models.jpeg.differentiable_jpeg: shared JPEG(None, 'soft') instance (models/jpeg.py:38-42)."""
    q = int(quality)
    return djpeg(x, jpeg_qtable(q, 0), jpeg_qtable(q, 1), 'soft')[0]


# ------------------------------------------------------------------------------------------------ losses / optimizer
def mse(a, b):
    return torch.mean(torch.pow(255 * a - 255 * b, 2.0))       # helpers/tf_helpers.py:31-32


def mae(a, b):
    return torch.mean(torch.abs(255 * a - 255 * b))


def sparse_categorical_crossentropy(labels, probs):
    """Keras SparseCategoricalCrossentropy() on probabilities, eager path (backend.sparse_categorical_crossentropy,
    from_logits=False): clip to [1e-7, 1-1e-7], log, then sparse softmax CE on those 'logits'; mean over samples."""
    eps = 1e-7
    logits = torch.log(probs.clamp(eps, 1 - eps))
    return F.cross_entropy(logits, torch.as_tensor(labels, dtype=torch.long), reduction='mean')


def adam_keras_step(params, grads, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam (non-amsgrad) dense update, in place on the lists of tensors; t = iterations + 1."""
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    for p, g, mi, vi in zip(params, grads, m, v):
        mi += (g - mi) * (1 - beta1)
        vi += (g * g - vi) * (1 - beta2)
        p -= lr_t * mi / (torch.sqrt(vi) + eps)


# ------------------------------------------------------------------------------------------------ learned codec (DCN)
def _codebook_weights(values, codebook, v=50, gamma=25):
    """float64 kernel weights of every value w.r.t. every code word, normalised over the code book
    (models/layers.py:139-156 and helpers/tf_helpers.py:308-321 are the same computation)."""
    eps = 1e-72
    dff = values.reshape(-1, 1).double() - codebook.reshape(1, -1).double()
    if v <= 0:
        weights = torch.exp(-gamma * dff ** 2)
    else:
        weights = (1 + (gamma * dff) ** 2 / v) ** (-(v + 1) / 2)
    return (weights + eps) / (weights + eps).sum(dim=1, keepdim=True)


def soft_codebook_quantization(x, codebook, v=50, gamma=25):
    """Quantization('soft-codebook').call, models/layers.py:139-170: soft value through the gradient, hard value forward."""
    weights = _codebook_weights(x, codebook, v, gamma)
    soft = (weights @ codebook.reshape(-1, 1).double()).mean(dim=1).to(torch.float32).reshape(x.shape)
    hard = codebook.reshape(-1)[weights.argmax(dim=1)].reshape(x.shape).to(torch.float32)
    return (hard - soft).detach() + soft


def entropy(values, codebook, v=50, gamma=25):
    """helpers/tf_helpers.py:290-333: entropy of the soft histogram (returned as float32 like the reference)."""
    weights = _codebook_weights(values, codebook, v, gamma)
    histogram = weights.mean(dim=0).clamp(1e-9, float(np.finfo(np.float32).max))
    histogram = histogram / histogram.sum()
    return (-(histogram * torch.log(histogram)).sum() / 0.6931).to(torch.float32), histogram


def discrete_latent(z, scale, codebook, v=50, gamma=25, rounding='soft-codebook'):
    """DiscreteLatent.call, models/layers.py:195-203: the entropy is estimated on the QUANTISED latent. rounding: the Quantization
    mode the codec was built with (models/compression.py:66: 'soft-codebook' | 'sin' | 'soft' | 'identity')."""
    latent = z * scale.to(z.dtype) if scale is not None else z
    if rounding != 'soft-codebook':
        q = quantization(latent, rounding)
        return q, entropy(q, codebook, v, gamma)[0]
    if z.dtype == torch.float32:
        q = soft_codebook_quantization(latent, codebook, v, gamma)
    else:       # float64 truth run: keep the soft value in float64
        weights = _codebook_weights(latent, codebook, v, gamma)
        soft = (weights @ codebook.reshape(-1, 1).double()).reshape(z.shape)
        hard = codebook.reshape(-1).double()[weights.argmax(dim=1)].reshape(z.shape)
        q = (hard - soft).detach() + soft
    return q, entropy(q, codebook, v, gamma)[0]


def l2_loss(t):
    """tf.nn.l2_loss."""
    return (t * t).sum() / 2


# ---------------------------------------------------------------------------------------------------------------- SSIM
def _window_moments(x, y, win):
    """VALID separable filtering of x, y, x^2, y^2, xy with the 1-D window `win` along H and W (float64; arrays (H, W, C))."""
    import scipy.ndimage as ndi
    k = len(win)
    lo, hi = k // 2, k - 1 - k // 2

    def f(t):
        t = ndi.correlate1d(t, win, axis=0, mode='constant')
        t = ndi.correlate1d(t, win, axis=1, mode='constant')
        return t[lo:t.shape[0] - hi, lo:t.shape[1] - hi]
    return f(x), f(y), f(x * x), f(y * y), f(x * y)


def ssim_tf(a, b, max_val=1.0):
    """tf.image.ssim(a, b, max_val) restated (TF 2.1 python/ops/image_ops_impl.py: _fspecial_gauss(11, 1.5), VALID depthwise conv,
    luminance * contrast-structure, mean over space then over channels), as called by models/compression.py:89. (N,H,W,C) -> (N,)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    x = np.arange(11, dtype=np.float64) - 5.0
    g = np.exp(-0.5 * x * x / 1.5 ** 2)
    g /= g.sum()
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    out = []
    for i in range(a.shape[0]):
        m0, m1, e00, e11, e01 = _window_moments(a[i], b[i], g)
        lum = (2 * m0 * m1 + c1) / (m0 * m0 + m1 * m1 + c1)
        cs = (2 * e01 - 2 * m0 * m1 + c2) / (e00 + e11 - m0 * m0 - m1 * m1 + c2)
        out.append(np.mean(np.mean(lum * cs, axis=(0, 1))))
    return np.array(out)


def ssim_skimage(a, b):
    """skimage.metrics.structural_similarity(a, b, multichannel=True, data_range=1) restated (scikit-image 0.16: 7 x 7 uniform filter,
    use_sample_covariance -> N / (N - 1), K1 = 0.01, K2 = 0.03, 3-pixel border cropped before the mean), as wrapped by
    helpers/metrics.py:9-26. (H,W,C) -> float."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    win = np.full((7,), 1.0 / 7.0)
    ux, uy, uxx, uyy, uxy = _window_moments(a, b, win)
    cov = 49.0 / 48.0
    vx, vy, vxy = cov * (uxx - ux * ux), cov * (uyy - uy * uy), cov * (uxy - ux * uy)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return float(np.mean(s))


# ------------------------------------------------------------------------------------------------ SSIM / MS-SSIM losses (autograd)
MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _ssim_per_channel_t(a, b, max_val=1.0):
    """TF 2.1 image_ops_impl._ssim_per_channel restated on torch tensors (NHWC): Gaussian 11 x 11 (sigma 1.5) VALID depthwise filter,
    (luminance * cs, cs) averaged over space -> two (N, C) tensors. Differentiable."""
    x = torch.arange(11, dtype=a.dtype) - 5.0
    g = torch.exp(-0.5 * x * x / 1.5 ** 2)
    g = g / g.sum()
    c = a.shape[-1]
    w = (g[:, None] * g[None, :])[None, None].repeat(c, 1, 1, 1)

    def red(t):
        return F.conv2d(t.permute(0, 3, 1, 2), w, groups=c)
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    m0, m1 = red(a), red(b)
    num0 = m0 * m1 * 2.0
    den0 = m0 * m0 + m1 * m1
    lum = (num0 + c1) / (den0 + c1)
    num1 = red(a * b) * 2.0
    den1 = red(a * a + b * b)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).mean(dim=(2, 3)), cs.mean(dim=(2, 3))


def ssim_loss(a, b):
    """helpers/tf_helpers.py:39-40: mean(255 * (1 - tf.image.ssim(a, b, 1.0)))."""
    s, _ = _ssim_per_channel_t(a, b)
    return (255.0 * (1.0 - s.mean(dim=-1))).mean()


def msssim_loss(a, b):
    """helpers/tf_helpers.py:43-44: mean(255 * (1 - tf.image.ssim_multiscale(a, b, 1.0))) — TF 2.1 ssim_multiscale: five scales
    (2 x 2 average pooling between them; even sizes assumed, so TF's SYMMETRIC padding of odd sizes never triggers), relu(cs) of
    scales 0..3 and relu(ssim) of scale 4, weighted geometric mean, mean over channels."""
    vals = []
    for k in range(len(MSSSIM_WEIGHTS)):
        if k > 0:
            a, b = avg_pool(a, 2), avg_pool(b, 2)
        s, cs = _ssim_per_channel_t(a, b)
        vals.append(torch.relu(cs))
    vals[-1] = torch.relu(s)
    ms = torch.ones_like(vals[0])
    for v, p in zip(vals, MSSSIM_WEIGHTS):
        ms = ms * v ** p
    return (255.0 * (1.0 - ms.mean(dim=-1))).mean()
