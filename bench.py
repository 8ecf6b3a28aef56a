#!/usr/bin/env python
"""Benchmark of the graph-propagation + ranking hot path (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scale C2]

One "step" = one training batch (B = 512) of the CLUSSL model on the synthetic Allrecipes-scale
graph C2: full propagation over the four normalised graphs (forward), fused BPR / regulariser /
distance-correlation losses, backward through every propagation, Adam step -- the per-batch body of
the reference's `Trainer._train_epoch` (FoodRec/common/trainer.py:177-224).  The headline metric is
train epochs/s = 1 / (ceil(n_train / B) * seconds per step).  `value` is measured with the batch
indices already resident in HBM; `e2e` goes through the public model API with pinned HOST batches
(H2D inside the timed region) and reads every loss term back (D2H), like the reference trainer.

`--impl reference` times the CPU restatement of the reference path (oracle/, torch-CPU ops -- the
same library ops the reference calls) on the host cores, same workload and metric.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 512
METRIC, UNIT = "train_epochs_per_s", "epochs/s"


class Cfg(dict):
    def __getitem__(self, k):
        return self.get(k)


def model_cfg(ds, device):
    # configs/model/PRICAI_ModelX.yaml (allrecipe block) + configs/overall.yaml
    return Cfg(device=device, embedding_size=64, train_batch_size=BATCH, is_multimodal_model=True, end2end=False,
               use_health_level_multi_hot=True, n_ri_layers=2, n_mm_layers=1, n_ui_layers=1, reg_weight=0.01,
               loss_cl=0.1, n_cluster=ds.cfg.n_cluster, knn_k=10, mm_image_weight=0.1, learning_rate=0.002)


def workload_name(scale, ds):
    return (f"{scale}: CLUSSL (PRICAI_ModelX) train step, B={BATCH}, {ds.n_users} users / {ds.n_items} items / "
            f"{ds.n_train} train interactions, {ds.cfg.n_cluster} clusters x2 (from {ds.cfg.dv}-d image / "
            f"{ds.cfg.dt}-d text features), {ds.num_ingredients} ingredients, d=64, "
            f"2 item-side layers x3 graphs + 1 user-item layer, fwd+bwd+Adam")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ------------------------------------------------------------------------------- reference (CPU) arm
class OracleClussl:
    """CPU restatement of the reference's CLUSSL train step (oracle/, torch-CPU), used as the
    `cpu_baseline` leg and as `--impl reference`."""

    def __init__(self, ds, state_dict, lr):
        from oracle import adjacency
        self.ds = ds
        self.S_ui = adjacency.norm_adj_user_item(ds.train_coo_matrix, ds.n_users, ds.n_items)
        self.S_g = adjacency.norm_adj_item_side(ds.rIngre_triples, ds.n_items, ds.num_ingredients)
        self.S_v = adjacency.norm_adj_item_side(ds.image_cluster_triples, ds.n_items, ds.cfg.n_cluster)
        self.S_t = adjacency.norm_adj_item_side(ds.text_cluster_triples, ds.n_items, ds.cfg.n_cluster)
        names = ("user_embedding.weight", "item_embedding.weight", "ingre_embedding.weight",
                 "image_prototype_embedding.weight", "text_prototype_embedding.weight")
        self.P = {k: state_dict[k].detach().cpu().clone().requires_grad_(True) for k in names}
        self.opt = torch.optim.Adam(list(self.P.values()), lr=lr)

    def step(self, batch):
        from oracle import losses, propagation
        ds, P = self.ds, self.P
        self.opt.zero_grad()
        out = propagation.clussl_forward(
            self.S_ui, self.S_g, self.S_v, self.S_t, P["user_embedding.weight"], P["item_embedding.weight"],
            P["ingre_embedding.weight"], P["image_prototype_embedding.weight"], P["text_prototype_embedding.weight"],
            ds.n_users, ds.n_items, ds.num_ingredients, ds.cfg.n_cluster, 2, 1)
        u, p, n = (torch.from_numpy(batch[k]) for k in ("u_id", "pos_i_id", "neg_i_id"))
        terms = losses.clussl_loss(out, P["user_embedding.weight"], P["item_embedding.weight"], u, p, n, 0.01, 0.1)
        sum(terms).sum().backward()
        self.opt.step()
        return [float(t) for t in terms]


def time_cpu(oracle, batches, steps, warmup):
    for i in range(warmup):
        oracle.step(batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(steps):
        oracle.step(batches[(warmup + i) % len(batches)])
    return (time.perf_counter() - t0) / steps


# ------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", default="C2")
    ap.add_argument("--cpu-steps", type=int, default=3, help="steps of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-schgn", action="store_true", help="skip the SCHGN side measurement")
    ap.add_argument("--eager", action="store_true", help="drop-in eager step (no CUDA graph)")
    ap.add_argument("--foreach-adam", action="store_true", help="torch's default foreach Adam instead of fused=True")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import foodrec_b200  # noqa: F401
    from foodrec_b200.synth import make_dataset, sample_train_batches

    if args.impl == "reference":
        if rank != 0:
            return 0
        torch.set_num_threads(os.cpu_count() or 1)
        ds = make_dataset(args.scale)
        steps_per_epoch = math.ceil(ds.n_train / BATCH)
        torch.manual_seed(999)
        sd = _init_state_dict(ds)
        oracle = OracleClussl(ds, sd, 0.002)
        batches = sample_train_batches(ds, BATCH, min(8, args.steps + args.warmup), seed=7)
        timed = min(args.steps, 100)          # one step = one batch (~0.45 s on 16 cores): keep the arm within minutes
        s = time_cpu(oracle, batches, timed, min(args.warmup, 3))
        v = 1.0 / (steps_per_epoch * s)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.scale, ds), "steps_per_epoch": steps_per_epoch},
            "timed_steps": timed,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{timed} train batches of {BATCH} (full-graph fwd+bwd+Adam each), "
                                       f"extrapolated to {steps_per_epoch} batches/epoch"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ B200 arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from foodrec_b200 import _lib, ops
    from foodrec_b200.models.pricai_modelx import PRICAI_ModelX

    ds = make_dataset(args.scale)
    steps_per_epoch = math.ceil(ds.n_train / BATCH)
    cfg = model_cfg(ds, str(dev))
    torch.manual_seed(999)
    model = PRICAI_ModelX(cfg, ds)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    model.train()
    # trainer.py:142-143 builds optim.Adam(params, lr, weight_decay).  Same optimizer, torch's fused
    # multi-tensor implementation (one launch instead of ~15), capturable so the step stays on the device.
    opt = torch.optim.Adam(model.parameters(), lr=cfg["learning_rate"], weight_decay=0.0,
                           capturable=not args.eager, fused=not args.foreach_adam)
    n_b = args.steps + args.warmup
    # weak scaling: every rank trains on its own batches of the replicated graph (multi-GPU)
    host_batches = sample_train_batches(ds, BATCH, min(n_b, 64), seed=7 + rank)
    keys = ("u_id", "pos_i_id", "neg_i_id")
    pinned = [{k: torch.from_numpy(b[k]).pin_memory() for k in keys} for b in host_batches]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in pinned]

    from foodrec_b200.train import GraphedTrainStep

    def eager(batch):
        opt.zero_grad()
        losses = model.calculate_loss(batch)
        loss = sum(losses)
        loss.backward()
        if world > 1 and hook is None:
            _allreduce_grads(model, world)
        elif world > 1:
            hook()
        opt.step()
        return losses

    hook = None
    step_mode = "eager"
    if args.eager:
        step = eager
    else:
        from foodrec_b200.train import OverlappedGradAllReduce
        try:   # N > 1: per-gradient NCCL all-reduces on a side stream, captured inside the graph
            hook = OverlappedGradAllReduce(model) if world > 1 else None
            step = GraphedTrainStep(model, opt, resident[0], keys=keys, grad_hook=hook)
            step_mode = "cuda_graph_replay"
        except Exception as e:  # noqa: BLE001  (capture of a collective refused: stay eager, say so)
            print(f"[bench] graph capture failed, running eager: {e}", file=sys.stderr)
            if world > 1 and hook is not None:
                hook.remove()
                hook = None
            step = eager

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    sampler = ClockSampler(local_rank) if rank == 0 else None   # nvidia-smi needs ~0.3 s to start sampling
    for i in range(args.warmup):
        step(resident[i % len(resident)])
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(resident[(args.warmup + i) % len(resident)])
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    launches = _lib.launch_count() - l0

    # ---- e2e: pinned host batches in, loss terms out, through the public model API
    h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
    for i in range(2):
        step({k: v.to(dev, non_blocking=True) for k, v in pinned[i % len(pinned)].items()})
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for i in range(args.steps):
        b = {k: v.to(dev, non_blocking=True) for k, v in pinned[(args.warmup + i) % len(pinned)].items()}
        losses = step(b)
        vals = [float(x.item()) for x in losses]  # trainer.py:186: per-term .item()
        d2h = 4 * len(vals)
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel (propagation SpMM), CUDA events around every launch
    prof = []
    model.fork_streams = False      # serialise the sub-graphs so each launch is timed alone on its stream
    ops.PROFILE = prof
    for i in range(3):
        eager(resident[i % len(resident)])  # same kernels as the graph replays; events need eager launches
    torch.cuda.synchronize()
    ops.PROFILE = None
    model.fork_streams = True
    l1 = _lib.launch_count()
    eager(resident[0])
    launches_per_step_eager = _lib.launch_count() - l1  # kernels of this library per step (graph replays the same)
    # Each of the step's propagation launches (3 eager steps recorded) is re-issued 20x back to back with the
    # same plan and operand shapes, bracketed by one CUDA-event pair on the launch stream: the stream stays
    # busy, so the figure is kernel time, not Python launch latency.  Operands stay L2-warm, as inside the step.
    tot_ms, tot_bytes, n_launch, REP = 0.0, 0.0, len(prof), 20
    probe_ms, gathered = 0.0, 0.0          # the same launches' row gathers alone (`fr_probe_gather`): the gather roofline
    probe_out = torch.empty(148 * 32 * 32, device=dev)
    bufs = {}
    for _, _, nb, g, has_z in prof:
        key = (g.n_rows, g.n_cols)
        if key not in bufs:
            bufs[key] = (torch.randn(g.n_cols, 64, device=dev), torch.randn(g.n_rows, 64, device=dev),
                         torch.empty(g.n_rows, 64, device=dev))
        X, Z, Y = bufs[key]
        ops.spmm(g, X, Z=Z if has_z else None, alpha=0.5, beta=0.5, out=Y)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(REP):
            ops.spmm(g, X, Z=Z if has_z else None, alpha=0.5, beta=0.5, out=Y)
        b.record()
        torch.cuda.synchronize()
        tot_ms += a.elapsed_time(b) / REP
        tot_bytes += nb
        a.record()
        for _ in range(REP):
            _lib.check(_lib.lib.fr_probe_gather(X.data_ptr(), 64, g.col.data_ptr(), g.nnz, 8, 148 * 32,
                                                probe_out.data_ptr(), _lib.stream_ptr()), "fr_probe_gather")
        b.record()
        torch.cuda.synchronize()
        probe_ms += a.elapsed_time(b) / REP
        gathered += g.nnz * 256.0
    del bufs
    peaks = _peaks()
    achieved = tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0

    ev = _bench_eval(model, ds, dev, rank, world, barrier)
    times = torch.tensor([ms_dev, ms_e2e, ev["ms_dev"], ev["ms_e2e"], ev["kernel_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(times[0]), float(times[1])
    ev.update(ms_dev=float(times[2]), ms_e2e=float(times[3]), kernel_ms=float(times[4]))
    if rank != 0:
        return 0

    value = world / (steps_per_epoch * ms_dev * 1e-3)
    e2e = world / (steps_per_epoch * ms_e2e * 1e-3)
    ws_mb = _working_set_mb(model, opt)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.scale, ds), "steps_per_epoch": steps_per_epoch,
                   "l2": f"no explicit flush: step working set {ws_mb:.0f} MB (params+grads+Adam state+graphs+"
                         f"activations) exceeds the 126 MB L2",
                   "multi_gpu": "replicated graph, per-rank batches, NCCL all-reduce of the dense gradients inside the "
                                "captured step; evaluation sharded by user" if world > 1 else "single GPU"},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches) if step_mode == "eager" else int(launches_per_step_eager * args.steps),
        "step_mode": step_mode,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "spmm_group_kernel<64,8,4>", "achieved": achieved, "peak": peaks[0],
                     "unit": "GB/s", "frac": achieved / peaks[0], "traffic": _spmm_traffic(), "peak_source": peaks[1],
                     "traffic_source": "profiles/r1_spmm_step_traffic.json: mean dram read+write bytes per launch, ncu --set "
                                       "full over the 14 propagation launches of one step (below the algorithmic bytes: "
                                       "the step's tables stay in the 126 MB L2)",
                     "algorithmic_bytes_per_launch": tot_bytes / max(n_launch, 1),
                     "launches_per_step": n_launch / 3, "avg_launch_us": tot_ms * 1e3 / max(n_launch, 1),
                     "kernel_share_of_step": tot_ms / 3 / ms_dev,
                     "gather_bound": {
                         "note": "the tables are L2-resident at this scale, so the binding resource is the 256-byte row "
                                 "gather (L1/L2 -> SM), not HBM: `fr_probe_gather` issues only the gathers of the same "
                                 "launches (same column indices, nothing else) and is the ceiling for any gather-based SpMM",
                         "gathered_GBs_in_kernel": gathered / (tot_ms * 1e-3) / 1e9,
                         "gather_only_probe_GBs": gathered / (probe_ms * 1e-3) / 1e9,
                         "kernel_time_over_gather_only_time": tot_ms / probe_ms, "frac_of_gather_roofline": probe_ms / tot_ms}},
    }
    tpeak = _tensor_peak()
    n_eval = ev["n_users"]
    line["eval"] = {
        "metric": "full_sort_eval_users_per_s", "unit": "users/s",
        "workload": f"{n_eval} users x {ds.n_items} items, d=64, top-20, training-history mask, "
                    f"bf16 tcgen05 scores + fp32 re-score of 32 candidates, users sharded over {world} GPU(s)",
        "value": n_eval / (ev["ms_dev"] * 1e-3), "ms": ev["ms_dev"],
        "e2e": {"value": n_eval / (ev["ms_e2e"] * 1e-3), "ms": ev["ms_e2e"], "h2d_bytes": ev["h2d"],
                "d2h_bytes": ev["d2h"]},
        "roofline": {"bound": "tensor", "kernel": "gemm_topk_kernel_v2 (two sweeps: bounding + collection; FLOPs counted once)", "achieved": ev["flops"] / (ev["kernel_ms"] * 1e-3) / 1e12,
                     "peak": tpeak[0], "unit": "TFLOP/s", "frac": ev["flops"] / (ev["kernel_ms"] * 1e-3) / 1e12 / tpeak[0],
                     "traffic": None, "peak_source": tpeak[1]},
        "metrics_vs_oracle": ev.get("metrics"),
    }
    if world == 1:
        line["eval_c4_slice"] = _bench_eval_c4(dev, tpeak)
    if world == 1 and not args.no_schgn:
        torch.set_num_threads(os.cpu_count() or 1)
        line["schgn"] = _bench_schgn(ds, dev, steps_per_epoch, not args.no_cpu_baseline)
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        oracle = OracleClussl(ds, sd0, cfg["learning_rate"])
        s = time_cpu(oracle, host_batches, args.cpu_steps, 1)
        line["cpu_baseline"] = {"value": 1.0 / (steps_per_epoch * s), "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "ms_per_step": s * 1e3,
                                "sample": f"{args.cpu_steps} train batches of {BATCH} on the same graph "
                                          f"(full fwd+bwd+Adam each), extrapolated to {steps_per_epoch} batches/epoch"}
    print(json.dumps(line))
    return 0


SCHGN_CFG = dict(embedding_size=64, train_batch_size=BATCH, is_multimodal_model=True, end2end=False,
                 num_attention_heads=2, num_hidden_layers=2, hidden_act="gelu", inner_size=256, hidden_dropout_prob=0.5,
                 attention_probs_dropout_prob=0.5, regs=0.01, reg_image=1, reg_w=0.05, reg_g=0.01, reg_health=0.01,
                 ssl=0.008, SCHGN_ssl=True, neg_sample_num=4, learning_rate=0.0005)  # configs/model/SCHGN.yaml


def _bench_schgn(ds, dev, steps_per_epoch, cpu_baseline):
    """BASELINE.json configs[1] names SCHGN on this graph: its train step (GCN on the propagation kernel,
    kernel gathers, batch-sized scorer in torch; one CUDA-graph replay per batch) and its full sort (fused pair
    scorer, 256 users) -- reported beside the headline, each with the CPU restatement timed on the host."""
    from foodrec_b200 import _lib
    from foodrec_b200.models.schgn import SCHGN
    from foodrec_b200.synth import sample_train_batches
    from foodrec_b200.train import GraphedTrainStep
    cfg = Cfg({**SCHGN_CFG, "device": str(dev)})
    torch.manual_seed(999)
    model = SCHGN(cfg, ds)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(dev).train()
    host = sample_train_batches(ds, BATCH, 4, seed=21, schgn=True)
    batches = [{k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in b.items()} for b in host]
    opt = torch.optim.Adam(model.parameters(), lr=cfg["learning_rate"], fused=True, capturable=True)
    l0 = _lib.launch_count()
    step = GraphedTrainStep(model, opt, batches[0], keys=tuple(batches[0].keys()), warmup=3)
    launches_per_step = (_lib.launch_count() - l0) / 4          # 3 eager warm-up steps + the capture pass
    for i in range(5):
        step(batches[i % 4])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    n_steps = 50
    for i in range(n_steps):
        step(batches[i % 4])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n_steps
    n_nodes = ds.n_users + ds.n_items + ds.num_ingredients + ds.num_calories_level
    out = {"workload": f"SCHGN on the same graph: GCN over {n_nodes} nodes, B={BATCH}, {ds.image_size}-d image rows, "
                       f"masked-ingredient task, Adam; dropout on",
           "train_ms_per_step": ms, "train_epochs_per_s": 1.0 / (steps_per_epoch * ms * 1e-3),
           "library_launches_per_step": launches_per_step}
    model.eval()
    users = torch.arange(256, device=dev)
    with torch.no_grad():
        model.full_sort_topk(users, 20)
        torch.cuda.synchronize()
        a.record()
        for _ in range(2):
            model.full_sort_topk(users, 20)
        b.record()
        torch.cuda.synchronize()
    fs_ms = a.elapsed_time(b) / 2
    out["full_sort"] = {"workload": f"256 users x {ds.n_items} items through the fused pair scorer + top-20",
                        "ms": fs_ms, "users_per_s": 256 / (fs_ms * 1e-3), "pairs_per_s": 256 * ds.n_items / (fs_ms * 1e-3)}
    if cpu_baseline:
        from oracle import schgn as O
        P = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and k != "ingre_embed_second")
             for k, v in sd0.items()}
        popt = torch.optim.Adam([v for v in P.values() if v.requires_grad], lr=cfg["learning_rate"])
        ei = O.schgn_edge_index(ds)
        sizes = (ds.n_users, ds.n_items, ds.num_ingredients, ds.num_calories_level)

        def cpu_step(hb):
            popt.zero_grad()
            sum(O.calculate_loss(P, {k: torch.from_numpy(np.asarray(v)) for k, v in hb.items()}, cfg, ei, sizes)).backward()
            popt.step()
        cpu_step(host[0])
        t0 = time.perf_counter()
        for hb in host[1:3]:
            cpu_step(hb)
        cpu_ms = (time.perf_counter() - t0) / 2 * 1e3
        with torch.no_grad():
            Pd = {k: v.detach() for k, v in P.items()}
            t0 = time.perf_counter()
            O.full_sort_scores(Pd, ds, 3, ei, sizes)
            cpu_fs = time.perf_counter() - t0
        out["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "train_ms_per_step": cpu_ms,
                               "full_sort_users_per_s": 1.0 / cpu_fs,
                               "sample": "2 train batches (dropout = identity); 1 user against all items"}
    return out


def _bench_eval(model, ds, dev, rank, world, barrier):
    """Full-sort evaluation of every user (sharded by user over the ranks), k = 20, history mask."""
    from foodrec_b200 import evaluation as E
    model.eval()
    with torch.no_grad():
        user_all, item_all = model._tables()
        user_all, item_all = user_all.contiguous(), item_all.contiguous()
    hist = E.HistoryCSR(ds.train_coo_matrix, ds.n_users, dev)
    per = -(-ds.n_users // world)
    lo, hi = rank * per, min((rank + 1) * per, ds.n_users)
    users_host = torch.arange(lo, hi, dtype=torch.int64).pin_memory()
    users_dev = users_host.to(dev)

    def run(users):
        with torch.no_grad():
            return E.full_sort_topk(user_all, item_all, users, 20, hist=hist)[1]
    for _ in range(3):
        run(users_dev)
    barrier()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(users_dev)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        top = run(users_host.to(dev, non_blocking=True)).cpu()     # ids in from pinned host, top-K indices back out
    ms_e2e = (time.perf_counter() - t0) * 1e3 / 3
    prof = []
    E.PROFILE = prof
    run(users_dev)
    torch.cuda.synchronize()
    E.PROFILE = None
    out = {"ms_dev": sorted(ts)[len(ts) // 2], "ms_e2e": ms_e2e, "kernel_ms": prof[0][0].elapsed_time(prof[0][1]),
           "flops": prof[0][2] * world, "n_users": ds.n_users, "h2d": users_host.numel() * 8, "d2h": top.numel() * 8}
    if rank == 0 and world == 1:
        # Recall/NDCG of the fused path vs the fp32 oracle ranking on a 2048-user sample (4 d.p. equality)
        from foodrec_b200 import metrics as Mx
        sample = torch.arange(0, ds.n_users, max(1, ds.n_users // 2048))[:2048]
        S = (user_all[sample.to(dev)] @ item_all.t()).cpu()
        for r, u in enumerate(sample.tolist()):
            S[r, hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]].astype(np.int64)] = -float("inf")
        ref = torch.topk(S, 20, dim=-1)[1].numpy()
        pos = [ds.testRatings[u] for u in sample.tolist()]
        a = Mx.topk_metrics(top[sample - lo].numpy(), pos, metrics=("recall", "ndcg"), topk=(10, 20))
        b = Mx.topk_metrics(ref, pos, metrics=("recall", "ndcg"), topk=(10, 20))
        out["metrics"] = {"fused": a, "fp32_topk": b, "equal_4dp": a == b,
                          "index_mismatches": int((top[sample - lo].numpy() != ref).sum())}
    model.train()
    return out


def _spmm_traffic():
    p = os.path.join(ROOT, "profiles", "r1_spmm_step_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch_mean"])
    except (OSError, KeyError, ValueError):
        return None


def _bench_eval_c4(dev, tpeak):
    """BASELINE.json configs[3] shape (1 M users x 500 k items, d = 64, top-20, history mask), measured on a
    75 776-user slice (4 full waves of 148 row blocks); users are independent, so users/s carries over."""
    from foodrec_b200 import evaluation as E
    import scipy.sparse as sp
    M, N, K, k = 148 * 128 * 4, 500_000, 64, 20
    g = torch.Generator(device=dev).manual_seed(4)
    U = torch.randn(M, K, device=dev, generator=g) * 0.1
    I = torch.randn(N, K, device=dev, generator=g) * 0.1
    rng = np.random.default_rng(4)
    rows = np.repeat(np.arange(M), 20)
    hist = E.HistoryCSR(sp.coo_matrix((np.ones(rows.size, np.float32), (rows, rng.integers(0, N, size=rows.size))),
                                      shape=(M, N)), M, dev)
    Ib = E.to_bf16(I)
    run = lambda: E.gemm_topk(U, I, k, hist=hist, B_bf16=Ib)   # item table converted once per evaluation
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    prof, ts = [], []
    E.PROFILE = prof
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        val, idx = run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    E.PROFILE = None
    kms = sorted(x[0].elapsed_time(x[1]) for x in prof)[1]
    ms = sorted(ts)[1]
    # exactness on 256 users of the slice against the dense fp32 ranking
    sub = torch.arange(0, M, M // 256, device=dev)[:256]
    S = (U[sub] @ I.t()).cpu()
    for r, u in enumerate(sub.tolist()):
        S[r, hist.idx_host[hist.ptr_host[u]:hist.ptr_host[u + 1]].astype(np.int64)] = -float("inf")
    ref = torch.topk(S, k, dim=-1)[1]
    tfl = 2.0 * M * N * K / (kms * 1e-3) / 1e12
    return {"workload": f"{M} users x {N} items, d=64, top-20, 20-item history mask per user (slice of the 1M x 500k sweep)",
            "users_per_s": M / (ms * 1e-3), "ms": ms, "kernel_ms": kms,
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": tpeak[0], "unit": "TFLOP/s", "frac": tfl / tpeak[0],
                         "note": "algorithmic FLOPs 2MNK counted once; the kernel runs two sweeps (bounding + collection)"},
            "index_mismatches_vs_fp32_topk": int((idx[sub].cpu() != ref).sum())}


def _tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops burst (kernel timed alone; of measured)"
    return 1590.0, "B200_PROFILING.md fallback 1.59 PFLOP/s (of fallback)"


def _init_state_dict(ds):
    """Initial parameters exactly as the model constructor draws them (CPU, no library needed)."""
    import torch.nn as nn
    d, out = 64, {}
    for name, n in (("user_embedding", ds.n_users), ("item_embedding", ds.n_items),
                    ("ingre_embedding", ds.num_ingredients + 1), ("image_prototype_embedding", ds.cfg.n_cluster),
                    ("text_prototype_embedding", ds.cfg.n_cluster)):
        w = torch.empty(n, d)
        nn.init.xavier_uniform_(w)
        out[name + ".weight"] = w
    return out


def _allreduce_grads(model, world):
    import torch.distributed as dist
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat.div_(world)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def _working_set_mb(model, opt):
    n = sum(p.numel() for p in model.parameters()) * 4 * 4  # param, grad, exp_avg, exp_avg_sq
    for g in (model.g_ui, model.g_image, model.g_text, model.g_ingre):
        n += g.nnz * 8 + g.n_seg * 16 + 3 * g.n_rows * 64 * 4
    return n / 1e6


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs_sustained"] if "hbm_gbs_sustained" in json.load(open(p)) else json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


if __name__ == "__main__":
    sys.exit(main())
