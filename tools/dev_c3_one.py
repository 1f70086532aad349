"""Developer probe: one C3 render (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict
from eradiate_b200.kernel._render import _device_scene
import torch
sc = mi_load_dict(scenes.config_c3(spp=16))
dev = _device_scene(sc)
acc = torch.zeros(3 * 1024, dtype=torch.float64, device="cuda")
for i in range(2):
    dev.render_device(0, 5 + i, 1 << 15, 0, acc.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(acc.sum()))
