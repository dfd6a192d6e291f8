"""K1 preprocess on resampling geometries (raw size != network size): minmax + resample kernels, CUDA-event timed through the
device-pointer C ABI.

    python tools/k1_resample.py
"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import medseg_b200 as ms
from tools.stage_roofline import timed
eng = ms.Engine(None)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for (w, h) in [(512, 512), (1024, 1024), (768, 600), (256, 256)]:
    for batch in (32,):
        src = torch.randint(0, 65535, (batch, h, w), dtype=torch.int32, device='cuda').to(torch.uint16)
        out = torch.empty((batch, 512, 512), dtype=torch.uint8, device='cuda')
        t = timed(lambda: eng.preprocess_dev(src.data_ptr(), w, h, batch, out.data_ptr(), 0, st))
        by = batch * (2 * w * h + 2 * 512 * 512)
        print(f"{w}x{h} batch {batch}: {t*1e3:.1f} us  {by/t/1e6:.0f} GB/s algorithmic")
eng.cleanup()
