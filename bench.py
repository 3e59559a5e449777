#!/usr/bin/env python
"""bench.py -- federated-round throughput (images/s, whole job) of the B200-native PriMIA hot path T, with the
encrypted-inference (path E) latency reported beside it.

    python bench.py --gpus N --steps K --warmup W                (N=1 directly; N>1 under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path, timed on host cores)

Workload (BASELINE.json configs[1]): ResNet-18 3-class, 224x224x3 synthetic chest-X-ray-shaped batches, one hospital
(PySyft VirtualWorker) per GPU, batch 64 per hospital, Adam(lr 1e-4, betas (0.5,0.99), wd 5e-4), FedAvg after every
local step (sync_every_n_batch = 1) with optimizer reset, bf16 tensor-core throughput mode (fp32 master weights).
A "step" = one federated round = local step on every hospital (concurrently, one per GPU) + FedAvg all-reduce.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMAGE = 10.645  # fwd + bwd conv/fc work, SURVEY.md section 8d (5.3227 GMAC)
METRIC = "federated_round_images_per_sec"


def conv_traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum over the conv launches of one step, from the committed ncu launch list of
    this very command (profiles/r02_launches_one_step.csv, written by scripts/ncu_r02.sh + scripts/ncu_extract.py)"""
    import csv

    p = os.path.join(ROOT, "profiles", "r02_launches_one_step.csv")
    if not os.path.exists(p):
        return None, "no ncu launch list committed"
    tot, n = 0.0, 0
    for r in csv.DictReader(open(p)):
        if any(k in r["kernel"] for k in ("conv_halo_kernel", "conv_tma_kernel", "wgrad_", "stem::stem_kernel", "dgrad_s2")):
            tot += float(r["dram_read_MB"]) + float(r["dram_write_MB"])
            n += 1
    return tot * 1e6, f"profiles/r02_launches_one_step.csv: {n} conv launches of one step (B = 64); ncu does not attribute the write-back of a kernel's outputs to it, so this is mostly operand reads; algorithmic operand + output bytes ~1.6 GB"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """cores this process may actually use (affinity mask and cgroup cpu quota), not the machine's core count"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path (oracle restatement of torchlib/utils.py:1108-1233 +
    torchlib/models.py; the PySyft stack itself cannot be imported on this Python/torch -- SURVEY.md section 8c)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import train_oracle as O

    cores = host_cores()
    torch.set_num_threads(cores)
    n_workers = args.gpus
    dp = args.config == "C3"
    if dp:
        from oracle import dp_oracle as D

        def step(model, opt, loss_fn, x, y):  # the straightforward per-sample DP-SGD step (oracle/dp_oracle.py), fresh noise
            noise = {n: torch.randn_like(p) for n, p in model.named_parameters()}
            return D.dp_step(model, opt, loss_fn, x, y, noise, args.dp_sigma, 1.0)[0]
    else:
        step = O.local_step
    # bounded sample: size the per-hospital batch so that (steps + warmup) rounds fit in ~150 s of host time
    probe = O.resnet18(seed=42)
    px, py = torch.randn(4, 3, 224, 224), torch.randint(0, 3, (4,))
    popt = O.make_optimizer(probe)
    step(probe, popt, O.make_loss(), px, py)
    t0 = time.perf_counter()
    step(probe, popt, O.make_loss(), px, py)
    per_img = (time.perf_counter() - t0) / 4
    budget = 150.0 / max(1, (args.steps + args.warmup) * n_workers)
    sample = int(max(2, min(args.ref_batch or args.batch, budget / max(per_img, 1e-6))))
    del probe, popt
    ids = [f"hospital{i}" for i in range(n_workers)]
    base = O.resnet18(seed=42)
    models = {w: O.clone_model(base) for w in ids}
    local = O.clone_model(base)
    loss_fn = O.make_loss()
    g = torch.Generator().manual_seed(42)
    data = {w: (torch.randn(sample, 3, 224, 224, generator=g), torch.randint(0, 3, (sample,), generator=g)) for w in ids}

    def one_round():
        opts = {w: O.make_optimizer(models[w]) for w in ids}  # re-created after every aggregation (utils.py:1209-1218)
        for w in ids:  # hospitals sequentially, as utils.py:1160
            step(models[w], opts[w], loss_fn, *data[w])
        O.aggregation(local, models, ids)
        O.send_new_models(local, models, ids)

    for _ in range(args.warmup):
        one_round()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_round()
    dt = time.perf_counter() - t0
    value = n_workers * sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: ResNet-18 federated round, 224x224x3, FedAvg every step, Adam" + (", DP-SGD" if dp else ""),
                   "workers": n_workers, "batch_per_worker": args.batch, "parallelism": f"{n_workers} hospitals sequential on host cores"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images per hospital per step instead of {args.batch} (bounded CPU sample); torch-CPU fp32 oracle"
                                   + (" with one backward pass per sample (DP-SGD per-sample gradients)" if dp else "") +
                                   ", omits PySyft per-op msgpack round-trips => lower bound on the reference's time"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------- our arm
def cpu_baseline_sample(seconds_cap=25.0):
    """torch-CPU fp32 oracle local step on a bounded sample, rank 0 / N=1 only."""
    import torch

    from oracle import train_oracle as O

    cores = host_cores()
    torch.set_num_threads(cores)
    m = O.resnet18(seed=42)
    opt = O.make_optimizer(m)
    loss_fn = O.make_loss()
    Bs = 64   # the workload's own batch (C2: 64 images per hospital per step)
    g = torch.Generator().manual_seed(42)
    x, y = torch.randn(Bs, 3, 224, 224, generator=g), torch.randint(0, 3, (Bs,), generator=g)
    O.local_step(m, opt, loss_fn, x, y)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        O.local_step(m, opt, loss_fn, x, y)
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds_cap or (n >= 8 and dt > 10.0) or n >= 40:   # ~10-25 s of CPU work
            break
    return {"value": Bs * n / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{n} local steps of {Bs} images (fwd+bwd+Adam) with the torch-CPU fp32 oracle, {torch.get_num_threads()} threads"}


def encrypted_cpu_sample(size=224, pf=16):
    """CPU arm of path E: the oracle restatement of inference.py:279-321 on ONE full-size image -- torch-CPU int64 Beaver convs,
    the 80-step Newton BatchNorms, and the FSS comparisons through the oracle's C twin (SHA-512, OpenMP over the host cores;
    the reference's own path is single-threaded at B = 1: spdz.py:95-107 leaves one non-empty slice).  Primitive generation
    (offline) is timed separately from the online protocol."""
    import torch

    from oracle import fss_oracle_c
    from oracle import ring_oracle as R
    from oracle import train_oracle as O

    R.FSS = fss_oracle_c
    try:
        torch.manual_seed(42)
        model = O.ResNet18(input_size=size).eval()
        tape = R.GeneratingTape(1, 21 * 10 ** pf)
        sh = lambda q: [(s0 := tape._r(q.shape)), q - s0]
        P = {k: sh(R.encode(v.float().contiguous(), 10, pf)) for k, v in model.state_dict().items() if not k.endswith("num_batches_tracked")}
        x = sh(R.encode(torch.randn(1, 3, size, size) * 0.1, 10, pf))
        t0 = time.perf_counter()
        R.resnet18_forward_shared(P, x, tape, 10, pf, size)
        total = time.perf_counter() - t0
    finally:
        R.FSS = None
    return {"online_ms": (total - tape.gen_seconds) * 1e3, "offline_ms": tape.gen_seconds * 1e3, "comparisons": tape.n_fss,
            "beaver_products": tape.n_triples}


CMP_PER_IMAGE = 64 * 56 * 56 * 8 + 64 * 56 * 56 + 4 * 64 * 56 * 56 + 4 * 128 * 28 * 28 + 4 * 256 * 14 * 14 + 4 * 512 * 7 * 7   # 3 311 616
INT64_MAC_PER_IMAGE = 1.81356288e9      # forward MACs of the 20 convs + fc
FSS_INSTR_PER_HASH = 3728               # SASS instructions pm_fss_dif_eval executes per SHA-512 (profiles/README.md)
ALU_WARP_INSTR_PER_CLK_SM = 1.94        # measured issue rate of the integer ALU pipe (scripts/ubench/int_pipes.cu)


def fss_roofline(dev, sm_mhz=1965.0):
    """the dominant kernel of path E timed alone: pm_fss_dif_eval on the largest layer's instance count (64 x 112 x 112)"""
    import torch

    from primia_b200 import ring

    n = 64 * 112 * 112
    keys = ring.fss.build_fss_keys(n, dev, 1, 1)
    x = ring.ops.random_i64((n,), 2, 2, dev)
    win = keys[0].window(n)
    with torch.cuda.device(dev):
        for _ in range(2):
            ring.fss.dif_eval(0, x, win)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            ring.fss.dif_eval(0, x, win)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    achieved = 32 * n / ms / 1e6                                                       # G SHA-512 / s
    peak = ALU_WARP_INSTR_PER_CLK_SM * 32 * 148 * sm_mhz * 1e6 / FSS_INSTR_PER_HASH / 1e9
    key_bytes = sum(t.numel() * t.element_size() for t in keys[0].tensors()) + 16 * n
    return {"bound": "int-alu", "kernel": "fss::fss_dif_eval_kernel (DIF.eval: 32 SHA-512 compressions per comparison; thread = instance)",
            "achieved": achieved, "peak": peak, "unit": "G SHA-512/s", "frac": achieved / peak, "traffic": None,
            "algorithmic": f"32 hashes x {n} comparisons per launch = {32 * n / 1e6:.1f} M hashes, {FSS_INSTR_PER_HASH} integer instructions each",
            "launch_ms": ms, "hbm_GBps": key_bytes / ms / 1e6,
            "peak_source": "integer-ALU issue bound: 1.94 warp-instr/clk/SM measured by scripts/ubench/int_pipes.cu x 32 lanes x 148 SMs "
                           f"x {sm_mhz:.0f} MHz / {FSS_INSTR_PER_HASH} instr per hash (IADD3 partly dual-issues on the FMA pipe, so slightly "
                           "above 1.0 is possible); neither the HBM nor the tensor roofline applies: keys are 1.2 KB per comparison read once"}


def encrypted_inference_block(steps=3, devs=None, label=None):
    """Path E beside the headline: the reference's encrypted inference of ONE image (inference.py:292-317) end to end on
    shares -- 20 convs + fc (Beaver matmuls on the int8 tensor cores), 20 BatchNorms (80-step Newton inverse sqrt), 17 ReLUs
    and the 3x3 max-pool (FSS comparisons: 32 SHA-512 per element per party), avg-pool.  ``devs`` = (model_owner, data_owner,
    crypto_provider) devices: all on one GPU, or the SURVEY.md section 8e placement on three GPUs (openings = peer reads over
    NVLink, one multi-device CUDA graph).  Offline (triples + FSS keys) and online phases are timed separately."""
    import torch
    import torchvision

    from primia_b200 import _lib, ring
    from primia_b200.ring.resnet import EncryptedInferenceGraph

    cur = "cuda:%d" % torch.cuda.current_device()
    devs = devs or [cur, cur, cur]
    dev = devs[0]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    sync = lambda: [torch.cuda.synchronize(d) for d in set(devs)]
    parties = [ring.Party("model_owner", devs[0]), ring.Party("data_owner", devs[1])]
    prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", devs[2]), seed=42)
    torch.manual_seed(42)
    sd = torchvision.models.resnet18(num_classes=3).state_dict()   # key-compatible with torchlib/models.py resnet18
    net = ring.EncryptedResNet18.from_state_dict(sd, parties, prov, 10, 16, input_size=224)
    himg = (torch.randn(1, 3, 224, 224) * 0.1).pin_memory()
    eg = EncryptedInferenceGraph(net, himg)                         # warm-up, primitive schedule, capture of the online phase
    off, on = [], []
    hout = torch.zeros(1, 3).pin_memory()
    with torch.cuda.device(dev):
        for it in range(steps + 2):
            sync()
            e = [ev() for _ in range(3)]
            e[0].record()
            eg.offline()                                            # fresh triples + FSS keys into the static buffers
            for d in set(devs) - {dev}:
                torch.cuda.current_stream(dev).wait_stream(torch.cuda.current_stream(d))
            e[1].record()
            out, _pred = eg.online(himg.to(dev, non_blocking=True)) # H2D image, encode+share, graph replay (forward, reconstruct, decode)
            hout.copy_(out)                                         # D2H logits
            e[2].record()
            sync()
            if it >= 2:                                             # two warm-up images (allocator growth)
                off.append(e[0].elapsed_time(e[1]))
                on.append(e[1].elapsed_time(e[2]))
    gb = eg.bytes_per_image / 1e9
    launches = eg.kernels_in_graph
    del eg, net, parties, prov
    torch.cuda.empty_cache()
    on_ms, off_ms = sum(on) / len(on), sum(off) / len(off)
    n_gpus = len(set(devs))
    return {"placement": label or f"{n_gpus} GPU", "devices": devs, "online_ms": on_ms, "offline_ms": off_ms, "gpu_launches": launches,
            "primitives_GB_per_image": gb}


def linear_layers_block(steps=3):
    """the 21 linear layers alone (the Beaver matmuls named in north_star) on synthetic activation shares, as one CUDA graph"""
    import torch

    from primia_b200 import ring
    from primia_b200.ring.resnet import EncryptedLinearGraph, SharedLinearLayers

    dev = "cuda:%d" % torch.cuda.current_device()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    parties = [ring.Party("model_owner", dev), ring.Party("data_owner", dev)]
    prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", dev), seed=42)
    from primia_b200.ring import resnet as _rr

    lin = SharedLinearLayers(parties, prov, 10, 16)
    eg = EncryptedLinearGraph(lin, lin.make_inputs(1), 1)
    launches = eg.launches
    offl, onl = [], []
    for it in range(steps + 1):
        torch.cuda.synchronize()
        e = [ev() for _ in range(4)]
        e[0].record()
        eg.offline()
        e[1].record()
        torch.cuda.synchronize()
        e[2].record()
        eg.online()
        e[3].record()
        torch.cuda.synchronize()
        if it:
            offl.append(e[0].elapsed_time(e[1]))
            onl.append(e[2].elapsed_time(e[3]))
    lon = sum(onl) / len(onl)
    tri_bytes = 226733592
    pk = peaks()
    # per party and image: 2 GEMMs (delta@(b+eps) and a@eps fused as [delta|a]@[b+eps;eps]) = 2 x 1.81 G int64-MAC = 72 x that in
    # int8 MACs (36 limb pairs); bytes: triples read + delta/eps written, exchanged and read back ~ 3 x 227 MB per party
    int8_tops = 2 * 2 * INT64_MAC_PER_IMAGE * 36 * 2 / (lon * 1e-3) / 1e12
    hbm = 2 * 3 * tri_bytes / (lon * 1e-3) / 1e9
    return {"online_ms": lon, "offline_triple_gen_ms": sum(offl) / len(offl),
            "int64_gmac_per_s_per_party": 2 * INT64_MAC_PER_IMAGE / (lon * 1e-3) / 1e9,
            "scope": "20 convs + fc Beaver protocol only, CUDA graph; " + (
                "online = mask the activation, open it into limb planes, 2-segment GEMM on the int8 tensor cores, truncate -- the weight "
                "half (mask + open w, planes of a, b + eps, eps) runs in the offline phase" if _rr.HOIST_WEIGHT_SIDE else
                "online = the whole protocol per layer (mask, open, planarise four operands, GEMM, truncate)"),
            "online_launches": launches,
            "triple_bytes_per_party": tri_bytes,
            "roofline": {"bound": "hbm", "kernel": f"ring_i8::ring_gemm_i8_kernel + mask / open+planarize / truncate kernels ({launches} launches)",
                         "achieved": hbm, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm / pk["hbm_gbs"], "traffic": None,
                         "algorithmic": "per party: 226.7 MB of triples read + 206.9 MB of masked operands written and opened + outputs ~ 3 x "
                                        "226.7 MB; both parties on this GPU",
                         "int8_tensor": {"achieved_TOPs": int8_tops, "peak_TOPs": 2 * pk["bf16_tflops"],
                                         "frac": int8_tops / (2 * pk["bf16_tflops"]),
                                         "note": "36 u8 x u8 limb-pair products per int64 MAC, 2 ops per MAC; int8 dense peak taken as 2 x the "
                                                 "measured bf16 peak"}}}


def encrypted_inference_report(steps=3, cpu=True, enc_gpus=None):
    """the ``encrypted_inference`` object of the bench line: placements, rooflines, CPU baseline at the same 224 x 224 size"""
    import torch

    cur = torch.cuda.current_device()
    placements = [encrypted_inference_block(steps, None, "1 GPU: both parties and the crypto provider time-share it")]
    want3 = (enc_gpus or 0) >= 3 or (enc_gpus is None and torch.cuda.device_count() >= 3)
    if want3 and torch.cuda.device_count() >= 3:
        trio = [f"cuda:{(cur + i) % torch.cuda.device_count()}" for i in range(3)]
        try:
            placements.append(encrypted_inference_block(steps, trio, "3 GPUs: model_owner / data_owner / crypto_provider, openings over "
                                                                     "NVLink peer reads, one multi-device CUDA graph"))
        except Exception as exc:
            placements.append({"placement": "3 GPUs", "error": repr(exc)})
    best = min((p for p in placements if "online_ms" in p), key=lambda p: p["online_ms"])
    one = placements[0]
    out = {"metric": "encrypted_inference_ms_per_image", "value": best["online_ms"], "unit": "ms/image", "higher_is_better": False,
           "dtype": "int64", "online_ms": best["online_ms"], "offline_ms": best["offline_ms"], "gpu_launches": best["gpu_launches"],
           "n_gpus": len(set(best["devices"])), "placements": placements,
           "primitives_GB_per_image": one["primitives_GB_per_image"],
           "e2e": {"value": best["online_ms"], "unit": "ms/image", "h2d_bytes_per_step": 3 * 224 * 224 * 4, "d2h_bytes_per_step": 12,
                   "note": "host image -> H2D -> encode+share -> forward on shares -> reconstruct -> decode -> D2H logits"},
           "config": {"workload": "C4: SPDZ 2-party + crypto provider ResNet-18, base 10 pf 16 int64 ring, one 224x224x3 image, "
                                  "protocol fss, online phase = one CUDA-graph replay; value = the best placement measured"}}
    out["roofline"] = fss_roofline(f"cuda:{cur}")
    out["fss"] = {"comparisons_per_image": CMP_PER_IMAGE, "sha512_per_image_online": CMP_PER_IMAGE * 64,
                  "eval_ms_per_image_at_roofline_rate": CMP_PER_IMAGE * 64 / (out["roofline"]["achieved"] * 1e6),
                  "note": "both parties evaluate every comparison: 2 x 32 hashes; on one GPU the two parties' evaluations share the ALU pipe, "
                          "on separate GPUs they run concurrently"}
    out["linear_layers"] = linear_layers_block(steps)
    if cpu:
        c = encrypted_cpu_sample(224, 16)
        out["cpu_baseline"] = {"kind": "port", "cores": host_cores(), "unit": "ms/image", "value": c["online_ms"],
                               "offline_ms": c["offline_ms"],
                               "sample": "the SAME workload: one 224x224 image, pf 16, 3 311 616 comparisons -- oracle restatement, torch-CPU "
                                         "int64 matmuls + C/OpenMP SHA-512 on all host cores (the reference's own B=1 path is single-"
                                         "threaded and pays PySyft message overhead on top: this is a lower bound on its time)"}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from primia_b200.train import HospitalWorker, ResNet18Engine, aggregation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        group = dist.group.WORLD
    B = args.batch
    eng = ResNet18Engine(B, 3, 3, 224, "max", dev, args.mode)
    eng.init_random(seed=42)
    dp = {"noise_multiplier": args.dp_sigma, "max_grad_norm": 1.0} if args.config == "C3" else None
    worker = HospitalWorker(f"hospital{rank}", eng, dp=dp)
    nbuf = 4  # rotate 4 input batches (154 MB fp32) > L2; the activation working set (~1 GB/step) is >> L2 anyway
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    xs = [torch.randn(B, 3, 224, 224, device=dev, generator=g) for _ in range(nbuf)]
    ys = [torch.randint(0, 3, (B,), device=dev, generator=g) for _ in range(nbuf)]
    # host staging of the e2e leg: the loader's pinned batches.  bf16 mode stages bf16 pixels (half the PCIe bytes): the stem
    # rounds every pixel to bf16 before its MMA anyway, so the step is bit-identical to shipping fp32
    # (tests/test_train_graph_gpu.py::test_bf16_staged_batch_gives_the_identical_step)
    stage_bf16 = args.mode == "bf16" and args.e2e_dtype == "bf16"
    if stage_bf16:
        xs = [x.bfloat16().float() for x in xs]      # the device-resident leg runs on the very same pixel values
    hx = [(x.bfloat16() if stage_bf16 else x).cpu().pin_memory() for x in xs]
    hy = [y.cpu().pin_memory() for y in ys]

    overlap = bool(args.overlap and world > 1 and dp is None and args.graph)

    def fed_round(i):
        if overlap and eng._graph2 is not None:   # FedAvg of the layer4 bucket overlaps the backward of layers 3..1 (two-graph step)
            worker.local_step_and_fedavg(xs[i % nbuf], ys[i % nbuf], group)
        else:
            worker.local_step(xs[i % nbuf], ys[i % nbuf])
            aggregation([worker], None, group)
        eng.reset_optimizer()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(max(args.warmup, 3)):
        fed_round(i)
    if overlap:
        eng.capture_graph_overlap(xs[0], ys[0])
    elif args.graph and dp is None:
        eng.capture_graph(xs[0], ys[0])
    elif args.graph:                # the DP step's Philox counter lives in device memory: every replay draws fresh noise
        from primia_b200.train.dp import capture_dp_graph

        capture_dp_graph(eng, xs[0], ys[0], **dp)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(fed_round, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # the bench's own clock record: the timed region is ~40 ms at 20 steps (2 nvidia-smi samples), so the same load runs on
    # for another ~1 s, untimed, while the sampler keeps going
    if rank == 0 and clocks is not None and (clocks.get("samples") or 0) < 8:
        sampler = ClockSampler(local_rank)
        sampler.start()
    t_ext = time.perf_counter()
    i_ext = 0
    while time.perf_counter() - t_ext < 1.0 and world == 1:   # single-rank only: ranks must issue identical collectives
        fed_round(i_ext)
        i_ext += 1
        if i_ext % 50 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    if rank == 0 and world == 1 and (clocks.get("samples") or 0) < 8:
        ext = sampler.stop()
        ext["note"] = f"timed region ({clocks.get('samples')} samples) + {i_ext} more identical steps, untimed, for the clock record"
        if not clocks.get("reasons"):
            clocks = ext

    # end to end through the public API with HOST (pinned) buffers: H2D of the batch + D2H of the loss every step.  The loss of
    # step k is copied into a pinned ring and READ at step k + LAG (the values the reference's loss.item() returns, a few steps
    # late): the host never blocks on the step it has just enqueued, so the next batch's copy and launch overlap it.
    LAG = 3
    ring = torch.zeros(LAG + 1, 1).pin_memory()
    ring_ev = [None] * (LAG + 1)
    seen = []

    def fed_round_host(i):
        if overlap:
            loss = worker.local_step_and_fedavg(hx[i % nbuf], hy[i % nbuf], group, host=True)
        else:
            loss = worker.local_step_host(hx[i % nbuf], hy[i % nbuf])       # H2D of this step's batch (pinned -> device)
        worker.prefetch_host(hx[(i + 1) % nbuf], hy[(i + 1) % nbuf])        # loader look-ahead: next batch's H2D overlaps
        slot = i % (LAG + 1)
        ring[slot].copy_(loss, non_blocking=True)                           # D2H of this step's loss
        ev = torch.cuda.Event()
        ev.record()
        ring_ev[slot] = ev
        old = (i - LAG) % (LAG + 1)
        if i >= LAG and ring_ev[old] is not None:
            ring_ev[old].synchronize()                                      # step i - LAG has long finished
            seen.append(float(ring[old]))
        if not overlap:
            aggregation([worker], None, group)
        eng.reset_optimizer()

    for i in range(LAG + 1):
        fed_round_host(i)
    ms_e2e = timed(fed_round_host, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    assert all(v == v and v > 0 for v in seen), "loss readback returned garbage"

    # roofline of the dominant kernel family (tensor-core implicit-GEMM convs), measured live with CUDA events
    conv_ms, launches = eng.profile_conv_time(xs[0], ys[0], steps=2)
    pk = peaks()
    achieved = GFLOP_PER_IMAGE * B / conv_ms  # GFLOP / ms == TFLOP/s
    roof = {"bound": "tensor", "kernel": "conv family, tcgen05/TMEM implicit GEMMs: halo::conv_halo_kernel (3x3/s1 fwd + dgrad, one TMA strip "
                                         "serves all 9 taps), stem::stem_kernel (7x7/s2 fwd + wgrad straight from the fp32 NCHW batch), "
                                         "tma::conv_tma_kernel / s2p::dgrad_s2_kernel (1x1 and stride-2), wgh::wgrad_halo_kernel + "
                                         "tma::wgrad_tma_kernel: 20 fwd + 19 dgrad + 20 wgrad launches per step",
            "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
            "frac_of_burst_peak": achieved / pk["bf16_tflops"],
            "traffic": conv_traffic_from_profile()[0] if args.mode == "bf16" and B == 64 else None,
            "traffic_note": conv_traffic_from_profile()[1],
            "conv_ms_per_step": conv_ms, "conv_ms_by_kind": getattr(eng, "conv_ms_by_kind", None),
            "step_ms_by_family": getattr(eng, "step_ms_by_family", None),
            "step_share": conv_ms / (ms / args.steps),
            "timing": "CUDA events on the launching stream around each of the 59 conv launches of an eager step (a spin kernel queued "
                      "before each bracket keeps host launch latency out of it); in the timed CUDA-graph step the 20 wgrad launches run "
                      "on a side stream, so step_share is an upper bound of the family's share of the critical path",
            "algorithmic": f"{GFLOP_PER_IMAGE} GFLOP/image x {B} images per step", "peak_source": pk["source"] + ", sustained bf16 GEMM"}
    if args.mode != "bf16":
        roof["note"] = "fp32 parity mode runs on the FP32 pipe; fraction is still quoted against the bf16 tensor peak"

    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": (f"{args.config}: ResNet-18 3-class federated round, 224x224x3, batch {B} per hospital, FedAvg (NCCL all-reduce "
                                "of the 44.75 MB flat state) + optimizer reset after every local step, Adam"
                                + (f", DP-SGD per-sample clipping C=1.0 + Gaussian noise sigma={args.dp_sigma} (BatchNorm frozen in the DP step)"
                                   if dp else "")), "workers": world,
                   "batch_per_worker": B, "global_batch": world * B, "parallelism": f"fed{world} (one hospital per GPU)",
                   "l2": "4 rotating input batches (154 MB) and ~1 GB of activations per step: working set >> 126 MB L2",
                   "cuda_graph": bool(args.graph), "mode": args.mode,
                   "fedavg": ("two buckets (layer4+fc+BN statistics 33.6 MB, rest 11.1 MB), the first all-reduce overlapped with the backward "
                              "of layers 3..1, ncclAvg") if overlap else "one all-reduce of the flat state after the step (ncclAvg)"},
        "clocks": clocks, "gpu_launches": launches * args.steps,
        "e2e": {"value": e2e_value, "unit": "images/s",
                "h2d_bytes_per_step": world * (hx[0].numel() * hx[0].element_size() + hy[0].numel() * 8),
                "host_batch_dtype": str(hx[0].dtype),
                "d2h_bytes_per_step": world * 4, "ms_per_step": ms_e2e / args.steps,
                "note": f"pinned {'bf16' if stage_bf16 else 'fp32'} batch -> H2D on a copy stream (one batch of look-ahead) -> step -> loss D2H "
                        f"into a pinned ring, read {LAG} steps later"
                        + ("; bf16 staging is bit-identical for this mode (the stem rounds pixels to bf16 before its MMA)" if stage_bf16 else "")},
        "roofline": roof,
    }
    par = os.path.join(ROOT, "profiles", "r02_bf16_c2_errors.json")
    if args.mode == "bf16" and os.path.exists(par):
        d = json.load(open(par))
        line["bf16_parity"] = {k: d[k] for k in ("config", "loss_rel", "logits_rel", "grad_cos_min", "grad_rel_max", "grad_rel_median",
                                                 "mean_cos_deficit_vs_autocast") if k in d}
        line["bf16_parity"]["note"] = ("this mode's measured error against the torch-CPU fp32 oracle at this very configuration (frozen as the "
                                       "gate of tests/test_train_graph_gpu.py); the 1e-5 gate runs in mode f32 (tests/test_train_gpu.py)")
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample()
            try:
                line["encrypted_inference"] = encrypted_inference_report(cpu=True, enc_gpus=args.enc_gpus)
            except Exception as exc:  # the headline must still print
                import traceback

                line["encrypted_inference"] = {"error": repr(exc), "trace": traceback.format_exc()[-1500:]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: everything else this process or its libraries print (NCCL's version
    banner, warnings) is routed to stderr by pointing fd 1 at fd 2 for the duration of the run."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--config", default="C2", choices=["C2", "C3"], help="BASELINE.json configs: C2 = B 64/hospital, FedAvg every "
                    "step; C3 = B 128/hospital, FedAvg + DP-SGD (sigma 1.0, C 1.0)")
    ap.add_argument("--dp-sigma", type=float, default=1.0)
    ap.add_argument("--ref-batch", type=int, default=None, help="images per hospital per step for the bounded CPU reference sample")
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--e2e-dtype", default="bf16", choices=["bf16", "f32"], help="dtype of the pinned host batches of the e2e leg (bf16 mode)")
    ap.add_argument("--overlap", type=int, default=1, help="N > 1: all-reduce of the layer4 bucket overlapped with the rest of the backward")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--enc-gpus", type=int, default=None, help="path E placement: 1, or 3 = parties on two GPUs + provider on a third "
                                                                "(default: 3 when three GPUs are visible)")
    ap.add_argument("--path", default="T", choices=["T", "E"], help="E: print the encrypted-inference line alone")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 128 if args.config == "C3" else 64
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.path == "E":
        import torch

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if int(os.environ.get("RANK", "0")) == 0:
            sampler = ClockSampler(torch.cuda.current_device())
            sampler.start()
            line = encrypted_inference_report(steps=max(args.steps, 3) if args.steps < 10 else 5, cpu=not args.no_cpu, enc_gpus=args.enc_gpus)
            line["clocks"] = sampler.stop()
            line.update({"steps": args.steps, "warmup": 2, "data": "synthetic", "vs_baseline": None, "scaling": "replicas only"})
            emit(line)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
