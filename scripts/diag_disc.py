"""GPU diagnostic: per-parameter gradient cosine of the discriminators vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import segnet_oracle as O
from tests.helpers import bf16_round, cosine, load_oracle_state, rel_l2, to_device
from dasemanticsegmentationaml_b200 import losses as L, build
from dasemanticsegmentationaml_b200.model import FCDiscriminator, DepthWiseSepFCDiscriminator, DepthWiseSepBNFCDiscriminator
build.build()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
for kind, cls in (("dwsep_bn", DepthWiseSepBNFCDiscriminator), ("dwsep", DepthWiseSepFCDiscriminator)):
    sd = O.make_discriminator_state(kind, seed=3)
    d = load_oracle_state(cls(19), sd).to(DEV).train()
    g = torch.Generator().manual_seed(7)
    p = torch.softmax(2 * torch.randn(2, 19, 128, 256, generator=g), dim=1).to(DEV)
    pq = bf16_round(p)
    pin = pq.clone().requires_grad_(True)
    y = d(pin.to(torch.bfloat16))
    loss = L.bce_with_logits_const(y, 0.0)
    loss.backward()
    osd = to_device(sd, DEV, True)
    po = pq.clone().requires_grad_(True)
    yo = O.discriminator_forward(kind, osd, po, training=True)
    loss_o = O.bce_with_logits_const(yo, 0.0)
    loss_o.backward()
    osd2 = to_device(sd, DEV, True)
    po2 = pq.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yo2 = O.discriminator_forward(kind, osd2, po2, training=True)
        loss2 = O.bce_with_logits_const(yo2.float(), 0.0)
    loss2.backward()
    print(kind, "out rel-L2", rel_l2(y, yo), "torch-bf16", rel_l2(yo2, yo))
    for k, v in osd.items():
        if v.requires_grad:
            pg = dict(d.named_parameters())[k].grad
            print("  %-22s ours %.4f torch-bf16 %.4f |g| %.3e ours|g| %.3e" % (k, cosine(pg, v.grad), cosine(osd2[k].grad, v.grad), v.grad.norm().item(), pg.norm().item()))
    print("  dinput ours %.4f torch-bf16 %.4f" % (cosine(pin.grad, po.grad), cosine(po2.grad, po.grad)))
