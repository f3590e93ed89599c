"""Dev: perform kernel on the tensor cores (tcgen05 kind::tf32) vs the fp32 FFMA2 kernel: error and time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rex_gym_b200.agents import ForwardGaussianPolicy
for O, A, n in ((4, 2, 4096), (4, 2, 65536), (16, 4, 65536)):
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.normal(0, 1, (n, O)).astype(np.float32)).cuda()
    res = {}
    for tcore in (False, True):
        net = ForwardGaussianPolicy(O, A, seed=1, tensor_cores=tcore)
        w = net.get_weights()
        for k in w: w[k] = (rng.normal(0, 0.15, w[k].shape)).astype(np.float32) if k != "logstd" else w[k]
        rng = np.random.default_rng(0); _ = rng.normal(0, 1, (n, O))
        net.set_weights(w)
        out = net.perform(x, training=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5): net.perform(x, training=True, seed=1, step=0)
        a.record()
        for k in range(50): net.perform(x, training=True, seed=1, step=k)
        b.record(); torch.cuda.synchronize()
        res[tcore] = (out["mean"].cpu().numpy(), out["value"].cpu().numpy(), a.elapsed_time(b) / 50)
        net.close()
    print("O=%d A=%d n=%d  fp32 %.4f ms  tf32-tcgen05 %.4f ms  |dmean| %.2e |dvalue| %.2e" % (
        O, A, n, res[False][2], res[True][2], np.abs(res[False][0] - res[True][0]).max(), np.abs(res[False][1] - res[True][1]).max()), flush=True)
