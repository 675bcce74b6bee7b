import sys, torch
sys.path.insert(0, ".")
from hicdiff_b200 import ops
g = torch.Generator().manual_seed(5)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.randn(B, 64, 64, 256, generator=g).to(torch.bfloat16).cuda()
dy = (torch.randn(B, 64, 64, 256, generator=g) * 0.1).to(torch.bfloat16).cuda()
import time
t0 = time.time()
try:
    dw = ops.conv3x3_wgrad_nhwc(x, dy)
    torch.cuda.synchronize()
    print("wgrad ok", float(dw.abs().mean()))
except Exception as e:
    print("wgrad failed:", e)
print('elapsed %.2f s' % (time.time() - t0))
