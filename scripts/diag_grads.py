"""GPU diagnostic: per-parameter gradient cosine of the product BiSeNet vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import segnet_oracle as O
from tests.helpers import cosine, load_oracle_state, rel_l2, to_device
from dasemanticsegmentationaml_b200.model import BiSeNet
from dasemanticsegmentationaml_b200 import train as T, build
build.build()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
sd = O.make_bisenet_state(seed=11, randomize_bn=True)
m = load_oracle_state(BiSeNet("STDCNet813", 19), sd).to(DEV)
size = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 256)
g = torch.Generator().manual_seed(6)
x = torch.randn(2, 3, *size, generator=g).to(DEV)
labels = torch.randint(0, 20, (2, *size), generator=g)
labels[labels == 19] = 255
labels = labels.to(DEV)
# eval
m.eval()
with torch.no_grad():
    out = m(x)
    ref = O.bisenet_forward(to_device(sd, DEV), x, training=False)
    # torch's own bf16 autocast as the yard-stick
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref16 = O.bisenet_forward(to_device(sd, DEV), x, training=False)
for i in range(3):
    print("eval head", i, "rel-L2 ours", rel_l2(out[i], ref[i]), "torch-bf16", rel_l2(ref16[i], ref[i]),
          "argmax ours", (out[i].argmax(1) == ref[i].argmax(1)).float().mean().item(),
          "torch-bf16", (ref16[i].float().argmax(1) == ref[i].argmax(1)).float().mean().item())
# train
m.train()
m.zero_grad()
loss, lr = T.supervised_loss(m, x, labels)
loss.backward()
osd = to_device(sd, DEV, True)
loss_o, outs_o = O.supervised_loss(osd, x, labels, training=True)
loss_o.backward()
osd2 = to_device(sd, DEV, True)
with torch.autocast("cuda", dtype=torch.bfloat16):
    loss_b, _ = O.supervised_loss(osd2, x, labels, training=True)
loss_b.backward()
print("loss ours", loss.item(), "oracle", loss_o.item(), "torch-bf16", loss_b.item())
names = dict(m.named_parameters())
for k, v in osd.items():
    if v.requires_grad and v.grad is not None:
        pg = names[k].grad
        print("%-55s ours %.4f  torch-bf16 %.4f  |g| %.3e" % (k, cosine(pg, v.grad) if pg is not None else float('nan'),
              cosine(osd2[k].grad, v.grad), v.grad.norm().item()))
