"""Writes tests/golden/record-ref.pth with the UNMODIFIED reference LossRecorder (utils/save_load/recorders.py imported
from /root/reference): three batches (the last one short) of two keys.  Run in the build container."""
import os
import sys
import types

import torch

for name in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, '/root/reference')
from utils.save_load.recorders import LossRecorder      # noqa: E402

torch.manual_seed(0)
r = LossRecorder(4)
for n in (4, 4, 3):
    r.append_batch(total=torch.arange(5 * n, dtype=torch.float32).view(5, n) + 100 * len(r), y_true=torch.arange(n) + 10 * len(r))
r.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'record-ref.pth'))
print(r, len(r), r.recorded_samples)
