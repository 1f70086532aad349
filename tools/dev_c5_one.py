"""Developer probe: one C5-band render (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict
from eradiate_b200.kernel._render import _device_scene
import torch
sc = mi_load_dict(scenes.config_c5())
dev = _device_scene(sc)
acc = torch.zeros(7, dtype=torch.float64, device="cuda")
for i in range(2):
    dev.render_device(0, 5 + i, 1 << 23, 0, acc.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(acc.sum()))
