import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from tests import test_model as T
g, m, loss = T._train_step(0)
params = dict(m.named_parameters())
worst = []
for key in [k[3:] for k in g.files if k.startswith("gs_")]:
    grad = params[key].grad.double().cpu()
    samp = grad.flatten()[:: max(1, grad.numel() // 64)][:64].numpy()
    ref = g["gs_" + key]
    worst.append((np.abs(samp - ref).max() / np.abs(ref).max(), key))
worst.sort(reverse=True)
print("loss rel err", abs(loss - float(g["loss"])) / abs(float(g["loss"])))
print(worst[:6])
