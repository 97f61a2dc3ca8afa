import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from edsnet_b200 import BatchPlan, DSNet
dev = torch.device("cuda:0")
torch.manual_seed(bench.SEED)
T, scales = 16384, [4, 8, 16, 32]
m = DSNet("nystromformer", 1024, 128, scales, 8, fc_depth=5, pooling_type="roi").to(dev).eval()
x = bench.synth_features_device(T, dev, 1)
db = BatchPlan.build([T]).to(dev)
with torch.no_grad():
    cls, loc = m._forward_nograd(x, db)
    for _ in range(3):
        r = m.nms_packed(cls, loc, db, 0.5)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    r = m.nms_packed(cls, loc, db, 0.5)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print(int(r["keep_count"].cpu()[0]))
