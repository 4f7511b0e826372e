"""gsr_adam_step alone at cfg3 size (61 floats x 6M): ms and GB/s of the 28 B/parameter it must move."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
H.pkg()
optim = importlib.import_module(H.PKG_NAME + ".optim"); mv = importlib.import_module(H.PKG_NAME + ".multiview")
P = 6_000_000
params = optim.FlatParameters(P, "cuda"); grads = mv.FlatGradients(P, "cuda", split_sh=True)
params.buffer.normal_(); grads.buffer.normal_()
opt = optim.FusedAdam(params, grads, {"xyz": 1e-6, "f_dc": 1e-6, "f_rest": 1e-6, "opacity": 1e-6, "segment": 1e-6, "scaling": 1e-6, "rotation": 1e-6})
for _ in range(5): opt.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): opt.step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print("variant %s: %.3f ms  %.0f GB/s" % (os.environ.get("GSR_ADAM_VARIANT", "0"), ms, 28 * params.buffer.numel() / ms / 1e6))
