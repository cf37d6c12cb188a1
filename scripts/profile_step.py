"""One DA train step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, train as T
from dasemanticsegmentationaml_b200.model import BiSeNet, FCDiscriminator

build.build()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
nb = int(os.environ.get("B200_BATCH", "8"))
model = BiSeNet("STDCNet813", 19).to(dev)
disc = FCDiscriminator(19).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
opt_d = torch.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.9, 0.99))
g = torch.Generator().manual_seed(1)
x = torch.randn(nb, 3, 512, 1024, generator=g).to(dev)
xt = torch.randn(nb, 3, 512, 1024, generator=g).to(dev)
lab = torch.randint(0, 19, (nb, 512, 1024), generator=g).to(dev)
for _ in range(3):
    T.train_da_step(model, disc, opt, opt_d, x, lab, xt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
T.train_da_step(model, disc, opt, opt_d, x, lab, xt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
