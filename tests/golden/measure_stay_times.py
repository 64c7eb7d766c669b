"""Length of stay of a trip: the reference's numpy draw against the oracle / kernel form.

Reference (base_model.py:426-432 with params.get_gamma_shape_scale, params.py:576-580, quirk Q7):
    hours = np.random.gamma(shape = mean**2 / std, scale = std / mean)
Oracle / kernel (oracle/abm_oracle.py:245-251, csrc/abm_kernels.cuh stay_return_tick): Wilson-Hilferty transform
    hours = mean * (1 - c + sqrt(c) * z)**3,  c = 1 / (9 * shape),  z = table normal quantile of a 32-bit draw.
Over the (mean, std) range of the stay table (mean 24..48 h, std 10..60, +0.001 each), one million draws
per cell: mean, standard deviation and quantiles in hours.    python tests/golden/measure_stay_times.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'covid19-agent-based-model_b200')):
    sys.path.insert(0, p)
from abm_b200 import tables                      # noqa: E402
from oracle.abm_oracle import sample_quantile     # noqa: E402

N = 1_000_000
rng = np.random.default_rng(7)
znorm = tables.znorm_table()
qs = (1, 5, 25, 50, 75, 95, 99)
print(f'{N} draws per cell; hours.  WH = Wilson-Hilferty form of the kernel, numpy = np.random.gamma (legacy RandomState)')
print(f'{"mean":>6} {"std":>6} {"shape":>7} | {"mean np":>8} {"mean WH":>8} {"sd np":>7} {"sd WH":>7} | quantiles ' + ' '.join(f'{q:>3d}%' for q in qs) + '  (WH - numpy, hours)')
worst = 0.0
for mean, std in ((24.001, 60.001), (24.001, 30.001), (24.001, 10.001), (36.001, 60.001), (36.001, 30.001), (48.001, 60.001),
                  (48.001, 10.001), (12.0, 60.0), (6.0, 30.0)):
    shape, scale = mean ** 2 / std, std / mean
    np.random.seed(11)
    ref = np.random.gamma(shape, scale, N)
    xs = rng.integers(0, 2 ** 32, N, dtype=np.uint64).astype(np.uint32)
    mag = sample_quantile(znorm, (xs.astype(np.uint64) << np.uint64(1)) & np.uint64(0xFFFFFFFF))
    z = np.where(xs >> np.uint32(31) == 1, -mag, mag).astype(np.float64) * (1.0 / 65536.0)
    c9 = 1.0 / (9.0 * shape)
    t = (1.0 - c9) + np.sqrt(c9) * z
    wh = np.clip(mean * ((t * t) * t), 0.0, None)
    dq = np.percentile(wh, qs) - np.percentile(ref, qs)
    worst = max(worst, abs(wh.mean() / ref.mean() - 1), abs(wh.std() / ref.std() - 1))
    print(f'{mean:6.1f} {std:6.1f} {shape:7.1f} | {ref.mean():8.3f} {wh.mean():8.3f} {ref.std():7.3f} {wh.std():7.3f} | ' + ' '.join(f'{v:+5.2f}' for v in dq))
print(f'largest relative difference of a mean or standard deviation: {100 * worst:.2f} %  (the last two rows lie outside the table\'s range: shape 2.4 and 1.2)')
