#!/usr/bin/env python
"""bench.py — generated slates/sec of the PivotCVAE hot path on B200.

One "step" = one batch of users through recommend(return_item=True) (prior MLP ->
reparameterise -> PSM MLP -> pivot pick over the catalog -> SCM MLP -> per-slot
arg-max over the catalog) + the response-model score of the generated slates.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c4] [--mode greedy|sampled|list]
    python bench.py --impl reference ...      # CPU arm: the oracle port of the reference on the host cores

N>1 is launched by the driver through torch.distributed.run (one rank per GPU).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_items, n_users, L, D, Z, hidden, prior_hidden, no_user, batch)
    "c1": dict(n_items=3707, n_users=6041, L=5, D=8, Z=16, H=256, PH=128, no_user=False, B=64,
               desc="C1 ML-1M shape: 3707 items, 6041 users, slate 5, dim 8, with user, B=64"),
    "c2": dict(n_items=50000, n_users=1, L=10, D=8, Z=16, H=256, PH=128, no_user=True, B=1024,
               desc="C2 Yoochoose shape: 50k items, slate 10, dim 8, nouser, B=1024, PivotCVAE gt_pi + response MLP"),
    "c3": dict(n_items=100000, n_users=6041, L=5, D=8, Z=16, H=256, PH=128, no_user=False, B=4096,
               desc="C3: PivotCVAE gt training, fused full-catalog soft-max CE + KL, 100k items, slate 5, dim 8, B=4096"),
    "c4": dict(n_items=1000000, n_users=6041, L=5, D=8, Z=16, H=256, PH=128, no_user=False, B=4096,
               desc="C4: 1M items, slate 5, dim 8, with user, B=4096 (replicated table, batch data-parallel)"),
}


def structs(w):
    L, D, Z, H, PH = w["L"], w["D"], w["Z"], w["H"], w["PH"]
    C, ud = L + 1, (0 if w["no_user"] else D)
    return dict(enc=[L * D + C + ud, H, H], psm=[Z + C + ud, H, H, D], scm=[Z + C + D + ud, H, H, (L - 1) * D],
                dec=[Z + C + ud, H, H, L * D], prior=[C + ud, PH, PH], resp=[(L + (0 if w["no_user"] else 1)) * D, H, H, L])


def make_weights(w, model_kind):
    """Synthetic weights with the reference's initialisers, built on the CPU so both arms share them:
    tables uniform(-a, a), a = sqrt(2/D) (env/response_model.py:29-37), then row-normalised by the
    model (cvae.py:31); hidden layers kaiming_uniform_, heads default nn.Linear init (pivotcvae.py:108-152)."""
    g = torch.Generator().manual_seed(20211)
    a = (2.0 / w["D"]) ** 0.5
    doc = (torch.rand(w["n_items"], w["D"], generator=g) * 2 - 1) * a
    usr = (torch.rand(w["n_users"], w["D"], generator=g) * 2 - 1) * a
    torch.manual_seed(0)
    st = structs(w)
    sd, env_sd = {}, {}

    def mlp(prefix, dims, out):
        for i in range(len(dims) - 1):
            lin = torch.nn.Linear(dims[i], dims[i + 1])
            torch.nn.init.kaiming_uniform_(lin.weight)
            out["%s_%d.weight" % (prefix, i + 1)] = lin.weight.detach().numpy().copy()
            out["%s_%d.bias" % (prefix, i + 1)] = lin.bias.detach().numpy().copy()

    def head(name, n_in, out):
        lin = torch.nn.Linear(n_in, w["Z"])
        out[name + ".weight"], out[name + ".bias"] = lin.weight.detach().numpy().copy(), lin.bias.detach().numpy().copy()

    mlp("enc", st["enc"], sd)
    head("encmu", st["enc"][-1], sd)
    head("enclogvar", st["enc"][-1], sd)
    if model_kind == "list":
        mlp("dec", st["dec"], sd)
    else:
        mlp("psm", st["psm"], sd)
        mlp("scm", st["scm"], sd)
    mlp("prior", st["prior"], sd)
    head("priorMu", st["prior"][-1], sd)
    head("priorLogvar", st["prior"][-1], sd)
    nd = torch.nn.functional.normalize(doc, p=2, dim=1)
    sd["docEmbed.weight"] = nd.numpy().copy()
    env_sd["docEmbed.weight"] = doc.numpy().copy()
    if not w["no_user"]:
        sd["userEmbed.weight"] = torch.nn.functional.normalize(usr, p=2, dim=1).numpy().copy()
        env_sd["userEmbed.weight"] = usr.numpy().copy()
    mlp("mlp", st["resp"], env_sd)
    return sd, env_sd


def make_inputs(w, B, step, seed=1234):
    """Eval-loop inputs (train_generative.py:177-184): uniform users, context = first k responses set."""
    g = torch.Generator().manual_seed(seed + step)
    users = torch.randint(0, w["n_users"], (B,), generator=g)
    k = step % w["L"] + 1
    ctx = torch.zeros(B, w["L"])
    ctx[:, :k] = 1
    return ctx, users


class Emb:
    def __init__(self, wt):
        self.weight = torch.from_numpy(wt)


def build_gpu(w, sd, env_sd, mode, device):
    from pivotcvae_b200.env.response_model import UserResponseModel_MLP
    from pivotcvae_b200.models.listcvae import UserListCVAEWithPrior
    from pivotcvae_b200.models.pivotcvae import PIVOTCVAE_MODELS
    st = structs(w)
    L, D, Z = w["L"], w["D"], w["Z"]
    uemb = None if w["no_user"] else Emb(env_sd["userEmbed.weight"])
    if mode == "list":
        m = UserListCVAEWithPrior(Emb(env_sd["docEmbed.weight"]), uemb, L, D, Z, L + 1, st["enc"], st["dec"], st["prior"],
                                  w["no_user"], device)
    else:
        key = "pivotcvae_gt_spi" if mode == "sampled" else "pivotcvae_gt_pi"
        m = PIVOTCVAE_MODELS[key](Emb(env_sd["docEmbed.weight"]), uemb, L, D, Z, L + 1, st["enc"], st["psm"], st["scm"],
                                  st["prior"], w["no_user"], device)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    env = UserResponseModel_MLP(w["n_items"] - 1, w["n_users"] - 1, D, L, st["resp"], device, w["no_user"])
    env.load_state_dict({k: torch.from_numpy(v) for k, v in env_sd.items()})
    env.to(device)
    return m, env


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle port of the reference)
def cpu_step(oracle, w, sd, env_sd, mode, ctx, users, eps, noise=None):
    if mode == "list":
        out = oracle.list_recommend(sd, ctx, users, eps, w["no_user"])
    else:
        out = oracle.pivot_recommend(sd, ctx, users, eps, w["no_user"], "sample" if mode == "sampled" else "max", noise)
    resp = oracle.resp_mlp(env_sd, out["items"].reshape(ctx.shape[0], -1), users, w["no_user"])
    return out["items"], resp


def time_cpu(w, sd, env_sd, mode, B, steps, warmup, min_seconds=0.0):
    import oracle
    threads = os.cpu_count() or 1
    oracle.set_threads(threads)
    rng = np.random.default_rng(0)
    times = []
    i = 0
    t_total = 0.0
    while i < warmup + steps or t_total < min_seconds:
        ctx, users = make_inputs(w, B, i)
        eps = rng.standard_normal((B, w["Z"])).astype(np.float32)
        noise = rng.exponential(size=(B, w["n_items"])).astype(np.float32) if mode == "sampled" else None
        t0 = time.perf_counter()
        cpu_step(oracle, w, sd, env_sd, mode, ctx.numpy(), users.numpy(), eps, noise)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            t_total += dt
        i += 1
        if len(times) >= 2000:
            break
    return float(np.mean(times)), threads, len(times)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mode = args.mode
    sd, env_sd = make_weights(w, "list" if mode == "list" else "pivot")
    B = min(w["B"], args.cpu_batch)
    sec, threads, n = time_cpu(w, sd, env_sd, mode, B, args.steps, args.warmup)
    val = B / sec
    line = {"impl": "reference", "metric": "generated slates/sec (%s)" % mode, "value": val, "unit": "slates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "batch_per_step": B, "mode": mode},
            "cpu_baseline": {"value": val, "unit": "slates/s", "cores": threads, "kind": "port",
                             "sample": "%d steps of %d slates (oracle/pcv_oracle.c, pthreads x%d)" % (n, B, threads)},
            "e2e": {"value": val, "unit": "slates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def make_train_batch(w, B, step, seed=4321):
    """Synthetic (slate, user, response) rows: uniform items/users, Bernoulli(0.5) responses."""
    g = torch.Generator().manual_seed(seed + step)
    slates = torch.randint(0, w["n_items"], (B, w["L"]), generator=g)
    users = torch.randint(0, w["n_users"], (B, 1), generator=g)
    resp = (torch.rand(B, w["L"], generator=g) < 0.5).float()
    return {"slates": slates, "users": users, "responses": resp}


def run_train(args, w):
    """train samples/sec: get_gen_loss (prior + encoder + decoder + fused catalog CE + KL) -> backward -> Adam.step,
    per step one batch; n_neg = --n-neg (N = full-catalog soft-max, the C3 config; 1000 = the reference default)."""
    import torch.distributed as dist
    from pivotcvae_b200 import ops
    from pivotcvae_b200.train_generative import get_gen_loss
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    B, K, W = (args.batch or w["B"]), args.steps, args.warmup
    sd, env_sd = make_weights(w, "pivot")
    model, _ = build_gpu(w, sd, env_sd, "greedy", device)
    model.noise.reseed(99 + rank)
    model.ce_engine = args.ce_engine
    # the reduced-precision training config (tf32 CE engine) also lets cuBLAS run the MLP backward GEMMs
    # on the tensor cores in tf32; the exact engine keeps fp32 SIMT GEMMs
    n_neg = w["n_items"] if args.n_neg <= 0 else args.n_neg
    tf32_bwd = args.ce_engine == "tf32" and n_neg >= w["n_items"]
    torch.backends.cuda.matmul.allow_tf32 = tf32_bwd
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=(world == 1))
    params = [p for p in model.parameters() if p.requires_grad]
    batches = [make_train_batch(w, B, i, seed=4321 + 7919 * rank) for i in range(K + W)]
    dev_batches = [{k: v.to(device) for k, v in b.items()} for b in batches]
    pin_batches = [{k: v.pin_memory() for k, v in b.items()} for b in batches]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()

    def step(batch):
        opt.zero_grad(set_to_none=True)
        loss, rec, kld = get_gen_loss(batch, model, None, 0.001, n_neg=n_neg)
        loss.backward()
        if world > 1:      # data-parallel: one all-reduce of the (~270k fp32) MLP gradients per step
            flat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
            dist.all_reduce(flat)
            flat /= world
            o = 0
            for p in params:
                if p.grad is not None:
                    n = p.grad.numel()
                    p.grad.copy_(flat[o:o + n].view_as(p.grad))
                    o += n
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = None
    if world == 1:     # single GPU: the whole step is one CUDA graph (the DP arm keeps the eager NCCL all-reduce)
        from pivotcvae_b200.graphs import GraphedTrainStep
        graphed = GraphedTrainStep(model, opt, B, 0.001, n_neg)
        run = lambda b: graphed(b)[0]
    else:
        run = step
    for i in range(W):
        run(dev_batches[i])
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ops.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        run(dev_batches[W + i])
        ev[i][1].record()
    barrier()
    launches = graphed.launches_per_step * K if graphed else ops.launch_count() - l0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.zero_()
        ev2[i][0].record()
        if graphed:
            loss = run(pin_batches[W + i])
        else:
            loss = run({k: v.to(device, non_blocking=True) for k, v in pin_batches[W + i].items()})
        loss_h.copy_(loss.detach(), non_blocking=True)
        ev2[i][1].record()
    barrier()
    # per-kernel CUDA-event timing of the same step launched eagerly (roofline of the dominant kernel)
    with ops.KernelTimer() as kt:
        for i in range(min(K, 10)):
            flush.zero_()
            if graphed:
                for k_, v_ in graphed.static.items():
                    v_.copy_(dev_batches[W + i][k_].reshape(v_.shape))
                graphed._step()
            else:
                step(dev_batches[W + i])
    barrier()
    ksum = kt.summary()
    clk = clocks.stop() if rank == 0 else None
    ms_e2e = sum(a.elapsed_time(b) for a, b in ev2)
    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    if rank != 0:
        return
    ms, ms_e2e = float(t[0]), float(t[1])
    total = B * world * K
    L_, D, N = w["L"], w["D"], w["n_items"]
    dom = ksum.get("ce_fwd_bwd", {"ms_avg": float("nan"), "calls": 0})
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    keep = n_neg / N
    flops = (2.0 * D * N * B * L_ + 2.0 * D * N * B * (L_ - 1)) * keep   # logits + dq accumulation over kept items
    ach = flops / (dom["ms_avg"] * 1e-3) / 1e12
    line = {"metric": "train samples/sec (PivotCVAE gt, fused catalog CE + KL, fwd+bwd+Adam)", "value": total / (ms / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": ("tf32 logits and MLP-gradient GEMMs, f32 accumulate" if tf32_bwd else "f32"),
            "data": "synthetic",
            "config": {"workload": w["desc"], "batch_per_gpu": B, "n_neg": n_neg, "beta": 0.001, "ce_engine": args.ce_engine,
                       "mlp_backward_gemm": "cuBLAS tf32" if tf32_bwd else "cuBLAS fp32",
                       "parallelism": "dp%d (replicated table, grad all-reduce)" % world,
                       "l2": "flushed between steps (256 MiB write); per-step CUDA-event pairs summed"},
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": "samples/s",
                    "h2d_bytes_per_step": int(B * L_ * 8 + B * 8 + B * L_ * 4), "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "ce_kernel<D=%d> (+finalize), M=%d rows x N=%d items, keep=%.4f" % (D, B * L_, N, keep),
                         "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                         "ms_avg_launch": dom["ms_avg"], "share_of_step": dom["ms_avg"] / (ms / K) if dom["calls"] else None,
                         "per_call_ms": {k: round(v["ms_avg"], 4) for k, v in ksum.items()}},
            "cpu_baseline": None, "clocks": clk}
    print(json.dumps(line))


def run_ours(args, w):
    import torch.distributed as dist
    from pivotcvae_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    ops.device_ok(local)
    mode = args.mode
    B, K, W = (args.batch or w["B"]), args.steps, args.warmup
    sd, env_sd = make_weights(w, "list" if mode == "list" else "pivot")
    model, env = build_gpu(w, sd, env_sd, mode, device)
    vp = args.parallel == "vp" and world > 1
    model.noise.reseed(1234 + (0 if vp else rank))   # vocab-parallel ranks replicate inputs AND noise
    model.select_engine = args.engine
    if vp:
        model.enable_vocab_parallel()
    no_user = w["no_user"]

    # synthetic inputs: one distinct batch per step, resident in HBM for `value`
    n_in = K + W
    ctxs, userss = zip(*[make_inputs(w, B, i, seed=1234 + (0 if vp else 7919 * rank)) for i in range(n_in)])
    ctx_d = [c.to(device) for c in ctxs]
    usr_d = [u.to(device) for u in userss]
    ctx_h = [c.pin_memory() for c in ctxs]
    usr_h = [u.pin_memory() for u in userss]
    items_h = torch.empty(B * w["L"], dtype=torch.int64).pin_memory()
    resp_h = torch.empty(B, w["L"], dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    from pivotcvae_b200.graphs import GraphedSlateGenerator

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the whole step (recommend + response score) is one CUDA graph over static buffers;
    # the Philox row counter lives on the device, so every replay draws fresh noise
    if vp:
        # vocab-parallel: the per-step NCCL all-gathers stay eager (not captured)
        class _Eager:
            launches_per_step = 0

            def __init__(self):
                self.ctx = torch.zeros(B, w["L"], device=device)
                self.users = torch.zeros(B, dtype=torch.int64, device=device)

            def load_inputs(self, c, u=None):
                self.ctx.copy_(c, non_blocking=True)
                if u is not None:
                    self.users.copy_(u, non_blocking=True)

            def _step(self):
                items, _ = model.recommend(self.ctx, None if no_user else self.users, return_item=True)
                return items, None, env(items.view(B, -1), self.users)

            def __call__(self, c, u=None):
                self.load_inputs(c, u)
                l0_ = ops.launch_count()
                items, _, resp = self._step()
                self.launches_per_step = ops.launch_count() - l0_
                return items, resp

        gen = _Eager()
    else:
        gen = GraphedSlateGenerator(model, env, B, warmup=3)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(W):
        gen(ctx_d[i], usr_d[i])
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        gen(ctx_d[W + i], usr_d[W + i])      # inputs already resident in HBM
        ev[i][1].record()
    barrier()
    launches = gen.launches_per_step * K
    ms = sum(a.elapsed_time(b) for a, b in ev)

    # end-to-end: host buffers in, host results out, copies inside the timed region
    for i in range(min(W, 3)):
        gen(ctx_d[i], usr_d[i])
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.zero_()
        ev2[i][0].record()
        items, resp = gen(ctx_h[W + i], usr_h[W + i])   # pinned host -> static device buffers
        items_h.copy_(items, non_blocking=True)
        resp_h.copy_(resp, non_blocking=True)
        ev2[i][1].record()
    barrier()
    ms_e2e = sum(a.elapsed_time(b) for a, b in ev2)

    # per-kernel CUDA-event timing of the same step, launched eagerly on the same stream
    # (events cannot be placed inside a graph replay); feeds the roofline of the dominant kernel
    KP = min(K, 50)
    with ops.KernelTimer() as kt:
        for i in range(KP):
            flush.zero_()
            gen.load_inputs(ctx_d[W + i], usr_d[W + i])
            gen._step()
    barrier()
    ksum = kt.summary()
    # the dominant call alone, as a CUDA-graph replay on the decoder's own queries (no host launch gaps
    # inside the event pair, same kernels as in the timed region), L2 flushed before every replay
    dom_graph_ms = None
    if not vp:
        seen = {}
        orig_select = model._select

        def spy(q, mode="greedy", **kw):
            if mode == "greedy" and q.shape[0] == B * w["L"]:
                seen["q"] = q.detach().clone()
            return orig_select(q, mode, **kw)
        model._select = spy
        try:
            gen._step()
        finally:
            model._select = orig_select
        if "q" in seen:
            torch.cuda.synchronize(device)
            sg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(sg):
                seen["out"] = orig_select(seen["q"])
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KP)]
            for a, b in evs:
                flush.zero_()
                a.record()
                sg.replay()
                b.record()
            torch.cuda.synchronize(device)
            dom_graph_ms = sum(a.elapsed_time(b) for a, b in evs) / KP
    clk = clocks.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total = B * (1 if vp else world) * K      # vocab-parallel: all ranks work on the SAME batch (strong scaling)
    value = total / (ms / 1e3)
    e2e = total / (ms_e2e / 1e3)
    # ---- roofline of the dominant kernel: the per-slot score+select over the catalog
    L_, D, N = w["L"], w["D"], w["n_items"]
    if vp:
        N = model.item_table().n_rows      # the local shard this rank scores
    key = "score_select_greedy_M%d" % (B * L_)
    dom = dict(ksum.get(key, {"ms_avg": float("nan"), "calls": 0}))
    eager_ms = dom["ms_avg"]
    if dom_graph_ms is not None:
        dom["ms_avg"], dom["calls"] = dom_graph_ms, KP
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get("bytes")
    except Exception:
        pass
    flops = 2.0 * D * N * (B * L_)
    ach = flops / (dom["ms_avg"] * 1e-3) / 1e12
    tc = args.engine in ("auto", "tcgen05") and D == 8 and N >= 2048 and mode != "sampled_all"
    kname = ("score_select_tc_kernel (tcgen05 tf32 filter) + tc_refine (exact fp32)" if tc
             else "score_select_kernel<D=%d> (exact fp32 SIMT) + finalize" % D)
    roofline = {"bound": "tensor", "kernel": "%s, M=%d rows x N=%d items" % (kname, B * L_, N),
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590 (B200_PROFILING.md)",
                "traffic": traffic if not vp else None, "algorithmic_bytes": N * D * 4 + B * L_ * (D * 4 + 8),
                "ms_avg_launch": dom["ms_avg"],
                "epilogue_roofline": {"note": "at D=8 the consumer of the logits bounds the kernel (SURVEY H2): one ALU-pipe max slot "
                                              "per logit = 64 logits/cycle/SM", "peak_logits_per_s": 64 * 148 * 1.965e9,
                                      "frac": (B * L_) * N / (dom["ms_avg"] * 1e-3) / (64 * 148 * 1.965e9)},
                "logits_per_s": (B * L_) * N / (dom["ms_avg"] * 1e-3),
                "engine": args.engine, "share_of_step": dom["ms_avg"] / (ms / K) if dom["calls"] else None,
                "timing": ("CUDA events around a graph replay of the call alone (its 2 kernels) on the decoder's queries, L2 flushed, "
                           "%d replays after the timed region; eager launch of the same call: %.4f ms" % (KP, eager_ms)) if dom_graph_ms is not None
                          else "CUDA events around the eager launch of the same call, %d steps after the timed region" % KP,
                "per_call_ms": {k: round(v["ms_avg"], 4) for k, v in ksum.items()}}
    cpu = None
    if world == 1 and not args.no_cpu:
        cb = min(B, args.cpu_batch)
        sec, threads, n = time_cpu(w, sd, env_sd, mode, cb, 3, 1, min_seconds=10.0)
        cpu = {"value": cb / sec, "unit": "slates/s", "cores": threads, "kind": "port",
               "sample": "%d steps of %d slates of the same workload (oracle/pcv_oracle.c, pthreads x%d)" % (n, cb, threads)}
    line = {"metric": "generated slates/sec (%s)" % mode, "value": value, "unit": "slates/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if vp else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "batch_per_gpu": B, "mode": mode,
                       "parallelism": ("vp%d (catalog sharded, 1 all-gather per scoring step)" % world) if vp else "dp%d (replicated table)" % world,
                       "l2": "flushed between steps (256 MiB write); per-step CUDA-event pairs summed",
                       "launch": ("eager launches + NCCL all-gathers (%d kernels of libpcv_b200 per step)" if vp else
                                  "one CUDA graph replay per step (%d kernels of libpcv_b200)") % gen.launches_per_step},
            "e2e": {"value": e2e, "unit": "slates/s", "h2d_bytes_per_step": int(B * w["L"] * 4 + B * 8),
                    "d2h_bytes_per_step": int(B * w["L"] * 8 + B * w["L"] * 4), "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="greedy", choices=["greedy", "sampled", "list", "train"])
    ap.add_argument("--n-neg", type=int, default=0, help="train mode: negatives per row (0 = the whole catalog)")
    ap.add_argument("--ce-engine", default="tf32", choices=["exact", "tf32"],
                    help="train mode, full catalog: CE logits in exact fp32 (SIMT) or tf32 on the tensor cores (C3 is a reduced-precision config)")
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--parallel", default="dp", choices=["dp", "vp"], help="N>1: batch data-parallel (default) or vocab-parallel")
    ap.add_argument("--cpu-batch", type=int, default=256, help="slates per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    elif args.mode == "train":
        run_train(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
