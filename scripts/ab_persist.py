"""A/B of the persistent implicit-GEMM kernel against the one-tile-per-CTA kernel on every conv
shape of the DA step: output equality and time (CUDA events, best of 2 x 10 launches)."""
import itertools
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, kernels as K, train as T
from dasemanticsegmentationaml_b200.model import BiSeNet, FCDiscriminator, DepthWiseSepBNFCDiscriminator

build.build()
dev = torch.device("cuda", 0)
BF = torch.bfloat16
tuned = dict(K.TUNED)
K.RECORD = []
torch.manual_seed(0)
nb = 8
model = BiSeNet("STDCNet813", 19).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
g = torch.Generator().manual_seed(1)
x = torch.randn(nb, 3, 512, 1024, generator=g).to(dev)
xt = torch.randn(nb, 3, 512, 1024, generator=g).to(dev)
lab = torch.randint(0, 19, (nb, 512, 1024), generator=g).to(dev)
for cls in (FCDiscriminator, DepthWiseSepBNFCDiscriminator):
    disc = cls(19).to(dev)
    opt_d = torch.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.9, 0.99))
    T.train_da_step(model, disc, opt, opt_d, x, lab, xt)
    del disc, opt_d
torch.cuda.synchronize()
records = {}
counts = {}
for kind, key, meta in K.RECORD:
    if kind == "conv":
        records.setdefault(key, meta)
        counts[key] = counts.get(key, 0) + 1
K.RECORD = None
del model, opt, x, xt, lab
torch.cuda.empty_cache()


def timeit(fn, reps=10):
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


out_table = {}
tot_a = tot_b = tot_best = 0.0
for key, m in sorted(records.items()):
    geom = m["geom"]
    xin = torch.randn(m["n"], m["hin"], m["win"], m["in_ld"], device=dev).to(BF)[..., :m["cin"]]
    filt = (torch.randn(m["rows"], m["n_slabs"], m["cin_pad"], device=dev) * 0.05).to(BF)
    dt = torch.float32 if m["f32"] else BF
    o1 = torch.zeros((m["n"], geom.Hout, geom.Wout, m["out_ld"]), device=dev, dtype=dt)[..., :m["rows"]]
    o2 = torch.zeros((m["n"], geom.Hout, geom.Wout, m["out_ld"]), device=dev, dtype=dt)[..., :m["rows"]]
    o3 = torch.zeros((m["n"], geom.Hout, geom.Wout, m["out_ld"]), device=dev, dtype=dt)[..., :m["rows"]]
    bias = torch.randn(m["rows"], device=dev) if m["bias"] else None
    base_tune = tuned.get(key, 0)
    res = {}
    base_tune &= ~(3 << 20)
    for name, tune, out in (("tile", base_tune, o1), ("persist", base_tune | (1 << 20), o2), ("persistB", base_tune | (3 << 20), o3)):
        st = torch.zeros(2, m["rows"], device=dev) if m["stats"] else None

        def run():
            K.conv_igemm(xin, filt, out, geom, bias=bias, act=2 if bias is not None else 0, slope=0.2, stats=st, bn_tile=tune)
        run()
        torch.cuda.synchronize()
        stats_once = None
        if st is not None:
            st.zero_()
            run()
            torch.cuda.synchronize()
            stats_once = st.clone()
        res[name] = (timeit(run), stats_once)
    same = torch.equal(o1, o2) and torch.equal(o1, o3)
    st_ok = True
    if m["stats"]:
        a, b = res["tile"][1], res["persist"][1]
        st_ok = ((a - b).abs().max() / (a.abs().max() + 1e-6)).item() < 1e-4
    ta, tb = res["tile"][0], min(res["persist"][0], res["persistB"][0])
    n = counts[key]
    tot_a += ta * n
    tot_b += tb * n
    tot_best += min(ta, tb) * n
    if tb < 0.97 * ta:
        out_table[key] = base_tune | (1 << 20)
    print("%-56s x%2d tile %7.1f us  persist %7.1f  persistB %7.1f us  %s%s" % (key, n, ta, res["persist"][0], res["persistB"][0], "same" if same else "OUTPUT DIFFERS",
                                                               "" if st_ok else " STATS DIFFER"), flush=True)
print("per-step sum: tile %.1f us, persist %.1f us, best-of %.1f us" % (tot_a, tot_b, tot_best))
merged = dict(tuned)
merged.update(out_table)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(merged, open("gpurun_out/tuned_tiles_persist.json", "w"), indent=0, sort_keys=True)
print("persistent wins on %d of %d shapes" % (len(out_table), len(records)))
