#!/usr/bin/env python
"""Benchmark of the north-star hot path: the adversarial domain-adaptation train step
(BiSeNet-STDC813 + discriminator, forward/backward/update) on synthetic Cityscapes/GTAV-shaped data.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

Prints ONE JSON line (rank 0).  `value` is whole-job source images per second with the inputs
resident in HBM; `e2e` is the same step driven from pinned HOST buffers through the public API
(H2D copies and the D2H read of the losses inside the timed region).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, BATCH, NCLS = 512, 1024, 8, 19
WORKLOADS = {
    "da_dense": "config[2]: adversarial DA train step, FCDiscriminator, 512x1024, batch 8/GPU",
    "da_dwsep": "config[3]: DA train step, DepthWiseSepFCDiscriminator, 512x1024, batch 8/GPU",
    "da_dwsep_bn": "config[3]: DA train step, DepthWiseSepBNFCDiscriminator, 512x1024, batch 8/GPU",
    "supervised": "config[1]: supervised train step (3x CE), 720x1280, batch 8",
    "supervised_ohem": "config[1]: supervised train step (3x OHEM CE, radix select), 720x1280, batch 8",
    "eval": "config[4]: full-resolution eval (forward + fused argmax + 19x19 confusion matrix), 1024x2048",
}
# algorithmic dense-conv FLOPs per (source, target) pair / image
GFLOP_PER_UNIT = {"da_dense": 444.87, "da_dwsep": 226.0, "da_dwsep_bn": 226.0, "supervised": 186.14,
                  "supervised_ohem": 186.14, "eval": 141.04}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="da_dense", choices=sorted(WORKLOADS))
    ap.add_argument("--torch-optim", action="store_true", help="torch's fused SGD/Adam instead of optim.FusedSGD/FusedAdam")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel time table here")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------- CPU reference arm
def cpu_da_step_rate(workload, batch, steps, warmup, threads=None):
    """Oracle port of the reference step (fp32, CPU, all host threads): images per second."""
    import torch
    from oracle import segnet_oracle as O
    # every host core this process may use (torchrun exports OMP_NUM_THREADS=1: override it)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(threads or avail)
    kind = {"da_dense": "dense", "da_dwsep": "dwsep", "da_dwsep_bn": "dwsep_bn"}.get(workload)
    h, w = (720, 1280) if workload == "supervised" else (H, W)
    g = torch.Generator().manual_seed(0)
    seg = O.clone_state(O.make_bisenet_state(seed=0), requires_grad=True)
    opt = torch.optim.SGD([v for v in seg.values() if v.requires_grad], lr=0.01, momentum=0.9, weight_decay=5e-4)
    x = torch.randn(batch, 3, h, w, generator=g)
    xt = torch.randn(batch, 3, h, w, generator=g)
    labels = torch.randint(0, NCLS, (batch, h, w), generator=g)
    if kind:
        dsd = O.clone_state(O.make_discriminator_state(kind, seed=1), requires_grad=True)
        opt_d = torch.optim.Adam([v for v in dsd.values() if v.requires_grad], lr=1e-3, betas=(0.9, 0.99))

    def step():
        if kind:
            O.da_step(seg, dsd, kind, x, labels, xt, opt, opt_d)
        else:
            opt.zero_grad()
            loss, _ = O.supervised_loss(seg, x, labels, True)
            loss.backward()
            opt.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.steps or 2
    warmup = args.warmup if args.warmup is not None else 1
    batch = 2  # bounded sample: BatchNorm on the 1x1 pooled maps needs >= 2 images per step
    rate, ms, threads = cpu_da_step_rate(args.workload, batch, steps, warmup)
    sample = "%d-image %s step(s) of the oracle port (fp32, torch CPU), %d warm-up" % (batch, args.workload, warmup)
    line = {
        "impl": "reference", "metric": "da_train_step_img_per_s" if args.workload != "supervised" else "train_step_img_per_s",
        "value": rate, "unit": "img/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "batch_per_step": batch, "height": H, "width": W},
        "cpu_baseline": {"value": rate, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons, sampled every 100 ms from before the warm-up; stop()
    summarises the samples whose timestamps fall inside the timed window."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self, t0, t1):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]),
                             [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 - 0.05 <= r[0] <= t1 + 0.05]
        window = "timed region"
        if not inside and rows:  # region shorter than the sampling period: take the closest samples
            inside = sorted(rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:3]
            window = "nearest samples (timed region shorter than the 100 ms period)"
        sm = sorted(r[1] for r in inside)
        reasons = sorted(set(n for r in inside for n in r[4]))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max([r[2] for r in inside]) if inside else None,
                "power_w_max": max([r[3] for r in inside]) if inside else None, "samples": len(inside),
                "window": window, "reasons": reasons}


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from dasemanticsegmentationaml_b200 import _lib, build
    from dasemanticsegmentationaml_b200 import train as T
    from dasemanticsegmentationaml_b200.model import (BiSeNet, FCDiscriminator, DepthWiseSepFCDiscriminator,
                                                      DepthWiseSepBNFCDiscriminator)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    _lib.ensure_device(local)
    _lib.set_device_index(local)

    steps = args.steps or 20
    warmup = max(3, args.warmup if args.warmup is not None else 5)
    nb = args.batch
    h, w = {"supervised": (720, 1280), "supervised_ohem": (720, 1280), "eval": (1024, 2048)}.get(args.workload, (H, W))

    torch.manual_seed(0)  # identical initial weights on every rank
    model = BiSeNet("STDCNet813", NCLS).to(dev)
    # same optimizers and hyper-parameters as train.py:170-172; fused=True only changes how torch
    # launches the update (one multi-tensor kernel per step)
    from dasemanticsegmentationaml_b200 import optim as B200Optim
    if args.torch_optim:
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4, fused=True)
    else:   # one launch per optimizer step; same arithmetic as torch.optim.SGD / Adam
        opt = B200Optim.FusedSGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    model_d = opt_d = None
    if args.workload.startswith("da_"):
        cls = {"da_dense": FCDiscriminator, "da_dwsep": DepthWiseSepFCDiscriminator,
               "da_dwsep_bn": DepthWiseSepBNFCDiscriminator}[args.workload]
        model_d = cls(NCLS).to(dev)
        if args.torch_optim:
            opt_d = torch.optim.Adam(model_d.parameters(), lr=1e-3, betas=(0.9, 0.99), fused=True, capturable=True)
        else:
            opt_d = B200Optim.FusedAdam(model_d.parameters(), lr=1e-3, betas=(0.9, 0.99))

    g = torch.Generator().manual_seed(100 + rank)  # rank-dependent data
    host = {
        "images": torch.randn(nb, 3, h, w, generator=g).pin_memory(),
        "labels": torch.randint(0, NCLS + 1, (nb, h, w), generator=g).pin_memory(),
        "images_t": torch.randn(nb, 3, h, w, generator=g).pin_memory(),
    }
    if args.workload == "supervised_ohem":
        host["labels"].clamp_(max=NCLS - 1)  # the reference's OHEM has no ignore_index
    else:
        host["labels"][host["labels"] == NCLS] = 255
    devbuf = {k: v.to(dev) for k, v in host.items()}
    eval_hist = [None]

    def step(buf):
        if args.workload == "eval":
            eval_hist[0], _ = T.eval_batch(model, buf["images"], buf["labels"], NCLS, eval_hist[0])
            return (eval_hist[0].sum().float(),)
        if args.workload == "supervised_ohem":
            return (T.train_step(model, opt, buf["images"], buf["labels"], loss="ohem"),)
        if model_d is None:
            return (T.train_step(model, opt, buf["images"], buf["labels"]),)
        return T.train_da_step(model, model_d, opt, opt_d, buf["images"], buf["labels"], buf["images_t"])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # eager warm-up (also fills lazy state), then capture the whole step into one CUDA graph
    eager_step = step
    for _ in range(2):
        eager_step(devbuf)
    graphed = None
    graph_note = "eager (--no-graph)"
    if not args.no_graph and args.workload != "eval":
        try:
            names_g = [k for k in devbuf if not (model_d is None and k == "images_t")]
            graphed = T.GraphedStep(lambda **kw: eager_step(dict(devbuf, **kw)), {k: devbuf[k] for k in names_g})
            graph_note = "whole step replayed as one CUDA graph"

            def step(buf):  # noqa: F811
                if buf is devbuf:  # already in the graph's static buffers
                    return graphed()
                return graphed(**{k: buf[k] for k in names_g})
        except Exception as ex:  # capture is an optimisation: report and fall back to eager launches
            graphed = None
            graph_note = "eager (graph capture failed: %r)" % (ex,)
            torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(warmup):
        step(devbuf)
    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    sync_all()
    t_wall0 = time.time()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(steps):
        losses = step(devbuf)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / steps  # host time to queue one step
    e1.record()
    sync_all()
    t_wall1 = time.time()
    launches = _lib.launch_count - l0
    if graphed is not None:  # replays do not pass through the binding: count one eager step's launches
        l1 = _lib.launch_count
        eager_step(devbuf)
        torch.cuda.synchronize()
        launches = (_lib.launch_count - l1) * steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    loss_vals = [float(v) for v in losses]

    # ---- timed region 2: end to end from pinned host buffers -------------------------------------
    # Every step copies ITS inputs host->device (pinned memory, a dedicated copy stream, double
    # buffered so that the copy of step i+1 overlaps the compute of step i) and reads the step's
    # losses back to the host (one D2H + sync per step, as train.py's .item() calls do).
    names = [k for k in devbuf if not (model_d is None and k == "images_t")]
    stage = [{k: torch.empty_like(devbuf[k]) for k in names} for _ in range(2)]
    for sbuf in stage:
        if "images_t" not in sbuf:
            sbuf["images_t"] = devbuf["images_t"]
    h2d = sum(host[k].numel() * host[k].element_size() for k in names)
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the step that last used this slot is done
            for k in names:
                stage[slot][k].copy_(host[k], non_blocking=True)
            ready[slot].record(copy_stream)

    sync_all()
    for ev in consumed:
        ev.record()
    d2h_stream = torch.cuda.Stream()
    host_buf = torch.empty(8, dtype=torch.float32).pin_memory()

    def read_back(item):
        """Blocking D2H read of one step's stacked losses on a side stream (waits for THAT step only)."""
        stacked, done = item
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done)
            host_buf[:stacked.numel()].copy_(stacked, non_blocking=True)
        d2h_stream.synchronize()
        return host_buf[:stacked.numel()].clone()

    e0.record()
    issue_copy(0)
    pending = None
    for i in range(steps):
        slot = i & 1
        if i + 1 < steps:
            issue_copy(slot ^ 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        out = step(stage[slot])
        consumed[slot].record()
        stacked = torch.stack([o.float() for o in out])
        done = torch.cuda.Event()
        done.record()
        # every step's losses are read back to the host; the read of step i happens after step i+1
        # has been queued, so the host-side sync never leaves the GPU idle
        if pending is not None:
            host_losses = read_back(pending)
        pending = (stacked, done)
    host_losses = read_back(pending)
    e1.record()
    sync_all()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item())

    # ---- per-kernel profile of ONE step (CUDA events around every launch of ours) -----------------
    prof = None
    if not args.no_profile:
        sync_all()
        _lib.profile_start()
        eager_step(devbuf)
        prof = _lib.profile_stop()

    if world > 1:
        dist.barrier()
    if rank != 0:
        _finish(world)
        return

    peaks = measured_peaks()
    unit_per_step = nb * world
    value = unit_per_step * steps / (ms_total / 1e3)
    line = {
        "metric": "da_train_step_img_per_s" if model_d is not None else ("eval_img_per_s" if args.workload == "eval" else "train_step_img_per_s"),
        "value": value, "unit": "img/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "batch_per_gpu": nb, "height": h, "width": w,
                   "classes": NCLS, "parallelism": "dp%d" % world,
                   "l2": "inputs (134 MB/step) and activations (>1 GB/step) exceed the 126 MB L2; no explicit flush",
                   "pairs_per_s": value if model_d is not None else None,
                   "losses_last_step": loss_vals, "host_enqueue_ms_per_step": host_enqueue_ms,
                   "launch_mode": graph_note},
        "e2e": {"value": unit_per_step * steps / (e2e_ms / 1e3), "unit": "img/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * len(out), "ms_per_step": e2e_ms / steps},
        "gpu_launches": launches,
        "clocks": clk,
    }
    step_tf = GFLOP_PER_UNIT[args.workload] * nb / 1e3 / (ms_total / steps / 1e3)  # TFLOP/s per GPU, whole step
    line["config"]["step_algorithmic_tflops_per_gpu"] = step_tf
    if prof:
        conv = [v for k, v in prof.items() if k in ("b200_conv_igemm",)]
        tot_ms = sum(v["ms"] for v in prof.values())
        table = sorted(((k, v["calls"], v["ms"], v["flops"]) for k, v in prof.items()), key=lambda r: -r[2])
        if conv and conv[0]["ms"] > 0:
            c = conv[0]
            ach = c["flops"] / (c["ms"] / 1e3) / 1e12
            line["roofline"] = {"bound": "tensor", "kernel": "conv_igemm_kernel + conv_igemm_persistent_kernel (forward + data-gradient launches)",
                                "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                                "frac": ach / peaks["tf_sustained"], "traffic": conv_traffic(args.workload),
                                "traffic_source": "profiles/r1_conv_traffic.json (ncu dram__bytes_read+write per launch, da_dense step)",
                                "peak_source": peaks["source"] + " bf16_tflops_sustained",
                                "launches_per_step": c["calls"], "avg_launch_ms": c["ms"] / c["calls"],
                                "share_of_kernel_time": c["ms"] / tot_ms}
        line["kernel_time_ms_per_step"] = {k: round(ms_, 3) for k, _, ms_, _ in table}
        if "b200_conv_wgrad" in prof and prof["b200_conv_wgrad"]["ms"] > 0:
            wg = prof["b200_conv_wgrad"]
            line["wgrad_tflops"] = wg["flops"] / (wg["ms"] / 1e3) / 1e12
        if args.profile_json:
            with open(args.profile_json, "w") as f:
                json.dump({"by_kernel": prof, "by_shape": _lib.last_profile_detail}, f, indent=1)
    if not args.no_cpu_baseline and world == 1 and (args.workload.startswith("da_") or args.workload == "supervised"):
        try:
            cb_batch, cb_steps = 4, 4
            rate, cms, threads = cpu_da_step_rate(args.workload, cb_batch, cb_steps, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
                                    "sample": "%d steps of %d source + %d target images (same %dx%d workload, reduced batch) "
                                              "of the oracle port, fp32 torch CPU, after one warm-up step; %.1f s per step"
                                              % (cb_steps, cb_batch, cb_batch, h, w, cms / 1e3)}
        except Exception as ex:  # the baseline is informational; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "img/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %r" % (ex,)}
    emit(line)
    _finish(world)


def _finish(world):
    """Leave without tearing NCCL down: destroying a communicator that a live CUDA graph still
    references can block forever, and the process is exiting anyway."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def conv_traffic(workload):
    """DRAM bytes per conv launch from the committed ncu capture (only measured for the default workload)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_conv_traffic.json")
    if workload != "da_dense" or not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get("traffic_bytes_per_launch")


_REAL_STDOUT = sys.stdout


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else any library or
    module prints while the bench runs (e.g. STDCNet813's 'use pretrain model' notice, kept for parity
    with stdcnet.py:139) is routed to stderr."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    args = parse()
    sys.stdout = sys.stderr
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
