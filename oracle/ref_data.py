"""CPU restatement of the reference's training-batch assembly — TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows `helpers/loading.py:132-211` (sample_patch: even top-left coordinates for Bayer alignment, the three discard policies with
their panic counter / best-candidate memory) and `helpers/dataset.py:89-131` (next_training_batch: per image one patch, RAW patch at
half the coordinates, uint16 / 65535 and uint8 / 255 through float64 into a float32 batch). Patch statistics are computed the way the
reference does (np.mean / np.var over the float64 patch), NOT with the integral images of the product path.
"""
import numpy as np


def _stats(img, x, y, p):
    patch = img[y:y + p, x:x + p].astype(np.float64) / 255
    return float(np.mean(patch)), float(np.var(patch))


def sample_patch(rgb, p=128, discard=None, max_attempts=25):
    max_x, max_y = rgb.shape[1] - p, rgb.shape[0] - p
    xx = yy = 0
    if not (max_x > 0 or max_y > 0):
        return xx, yy
    panic, best = max_attempts, None
    found = False
    while not found:
        xx = 2 * (np.random.randint(0, max_x) // 2) if max_x > 0 else 0
        yy = 2 * (np.random.randint(0, max_y) // 2) if max_y > 0 else 0
        if not discard:
            break
        mean, var = _stats(rgb, xx, yy, p)
        if discard == 'flat':
            if var < 0.005:
                panic -= 1
                found = panic <= 0
            elif var < 0.01:
                found = np.random.uniform() > 0.5
            else:
                found = True
        elif discard == 'flat-aggressive':
            if var < 0.02:
                if panic == max_attempts or var > best[2]:
                    best = (xx, yy, var)
                panic -= 1
                found = panic <= 0
                if found:
                    xx, yy = best[0], best[1]
            else:
                found = True
        elif discard == 'dark-n-textured':
            if 0 < var < 0.005 and 0.35 < mean < 0.99:
                found = True
            else:
                if panic == max_attempts or (var < 2 * best[3] and mean > 1.1 * best[2]):
                    best = (xx, yy, mean, var)
                panic -= 1
                found = panic <= 0
                if found:
                    xx, yy = best[0], best[1]
        else:
            raise ValueError('Unrecognized discard mode: {}'.format(discard))
    return xx, yy


def next_training_batch(x_u16, y_u8, batch_id, batch_size, rgb_patch_size, discard='flat', max_attempts=25):
    rp = rgb_patch_size // 2
    bx = np.zeros((batch_size, rp, rp, 4), dtype=np.float32)
    by = np.zeros((batch_size, rgb_patch_size, rgb_patch_size, 3), dtype=np.float32)
    pos = []
    for b in range(batch_size):
        bid = batch_id * batch_size + b
        xx, yy = sample_patch(y_u8[bid], rgb_patch_size, discard, max_attempts)
        pos.append((bid, yy, xx))
        bx[b] = x_u16[bid][yy // 2:yy // 2 + rp, xx // 2:xx // 2 + rp].astype(np.float64) / (2 ** 16 - 1)
        by[b] = y_u8[bid][yy:yy + rgb_patch_size, xx:xx + rgb_patch_size].astype(np.float64) / (2 ** 8 - 1)
    return bx, by, np.array(pos, dtype=np.int32)
