"""Host-side cost of queuing one DA step: run it on tiny images so that the GPU is never the limit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, train as T
from dasemanticsegmentationaml_b200.model import BiSeNet, FCDiscriminator
build.build()
dev = torch.device("cuda", 0)
model = BiSeNet("STDCNet813", 19).to(dev)
disc = FCDiscriminator(19).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4, fused=True)
opt_d = torch.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.9, 0.99), fused=True)
x = torch.randn(2, 3, 128, 256, device=dev); xt = torch.randn(2, 3, 128, 256, device=dev)
lab = torch.randint(0, 19, (2, 128, 256), device=dev)
for _ in range(5):
    T.train_da_step(model, disc, opt, opt_d, x, lab, xt)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    T.train_da_step(model, disc, opt, opt_d, x, lab, xt)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step, incl. drain %.2f ms/step" % ((t1 - t0) * 50, (t2 - t0) * 50))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    T.train_da_step(model, disc, opt, opt_d, x, lab, xt)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
