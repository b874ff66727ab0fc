#!/usr/bin/env python
"""MC-NeRF hot-path benchmark (BASELINE.json metric: train rays/s at 64+128 samples/ray).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

One "step" = one full train step of the reference's GLOBAL_OPTIM stage through the drop-in API exactly as
reference main.py:79-85 drives it: optimizer.zero_grad -> MC_Model(data, epoch, stage, ratio) -> MC_NeRF_Loss
-> backward -> RAdam.step, on the BASELINE configs[1] workload (110 cameras, 800x800 images, 4096 rays/batch,
64 coarse + 128 fine samples, both MLPs 8x256, random-init weights, synthetic Ball rig).

 value : rays/s with the step's inputs already resident in HBM (CUDA-event timed, max over ranks)
 e2e   : the same steps with HOST (pinned) inputs: H2D of the image/calibration tensors and the D2H read of
         the loss are inside the timed region
 N > 1 : one process per GPU (torchrun).  Default = STRONG scaling, BASELINE configs[2] / north_star: ONE 4096-ray batch
         split into equal ray slices, MLP + camera gradients all-reduced with NCCL (parallel.GradSync: the fine
         network's flat gradient buffer is reduced on a communication stream inside backward, the rest right after;
         1/N folded into RAdam).  The WEAK mode (the reference's DDP semantic: every rank its own camera's 4096 rays)
         is measured in the same run and reported under `other_scaling`.  After the timed steps all ranks are checked
         to hold bit-identical parameters (`ranks_identical`).
 --impl reference : the UNMODIFIED reference (baseline/_ref) on the host CPUs, all threads, the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train_rays_per_sec_64c_128f"
UNIT = "rays/s"
N_CAM, IMG, RAYS, SC, SCALE = 110, 800, 4096, 64, 2
MACS_PER_EVAL = 629248          # GEMM MACs per MLP evaluation (8x256, skip[4], sigma + SH-27 heads)
# Bytes the weight-gradient kernel must read per MLP evaluation: every DISTINCT operand tile once, bf16, per 128-row
# tile: ten dY tiles and ten activation tiles of 64 KB, the encoding tile (16 KB) and the head-gradient tile (8 KB)
# = 1304 KB / 128 rows.  (Its 13 jobs fetch 1456 KB: dY of the skip layer, the last trunk activation, the encoding and
# the head tile are each used by two jobs - the second fetch is an L2 hit when the jobs run in step.)  DESIGN.md §3.4
WGRAD_BYTES_PER_EVAL = 1304 * 1024 // 128
STAGE, RATIO = "GLOBAL_OPTIM_EPOCH", 0.5


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured (sustained)")
    return dict(hbm=6650.0, tensor=1590.0, src="fallback")


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md) through NVML in a
    background thread (an `nvidia-smi -lms` subprocess stalls kernel launches for ~100 ms per query)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.05):
        self.index, self.period, self.rows, self.stop_flag, self.t = index, period, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception:
            self.t = None

    def _loop(self):
        while not self.stop_flag:
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                why = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((mhz, why))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if self.t is None:
            return None
        self.stop_flag = True
        self.t.join(timeout=1)
        if not self.rows:
            return None
        reasons = sorted({name for _, why in self.rows for bit, name in self.REASONS.items() if why & bit})
        return dict(sm_mhz=statistics.median(r[0] for r in self.rows), sm_max_mhz=self.max_mhz, reasons=reasons,
                    samples=len(self.rows))


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------- ours
def build_workload(device, rank, precision, rays=RAYS, img=IMG):
    from mc_nerf_b200 import synthetic as syn
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss, RAdam
    sp = syn.make_sys_param(n_cam=N_CAM, img_h=img, img_w=img, batch=rays, samples=SC, scale=SCALE, device=device,
                            with_images=False)
    sp["mlp_precision"] = precision
    sp["pixel_sampler"] = "device"                    # the package default (synthetic.py pins "randperm" for replay tests)
    torch.manual_seed(42 + rank)                      # reference main.py:273-277
    model = MC_Model(sp).to(device)
    with torch.no_grad():
        for k, v in syn.init_camera_weights(sp).items():
            getattr(model, k).copy_(v)
    loss_fn = MC_NeRF_Loss(sp)
    # reference main.py:191-195 (the GLOBAL_OPTIM optimiser: every parameter, stage-2 lr, weight decay)
    opt = RAdam([p for p in model.parameters()], lr=5e-4, eps=1e-8, weight_decay=4e-4)
    batch = syn.make_train_batch(sp, img_id=(3 + 7 * rank) % N_CAM, seed=11 + rank)
    return sp, model, loss_fn, opt, batch


def measure(args, strong, rank, world, local, device, full):
    """Time the train step in one sharding mode.  strong: ONE --rays batch of one camera split into equal ray slices
    (BASELINE configs[2], north_star); weak: every rank renders its own camera's --rays batch (the reference's DDP
    semantic, ref: main.py:60-62).  full: also the e2e leg, the per-kernel roofline and the collective's exposed time."""
    from mc_nerf_b200._lib import lib
    from mc_nerf_b200 import parallel
    rays_rank = args.rays // world if strong else args.rays
    sp, model, loss_fn, opt, batch = build_workload(device, 0 if strong else rank, args.precision, rays_rank, args.img)
    if strong:
        torch.manual_seed(4242 + rank)           # same weights / image everywhere, different pixels per rank
    parallel.broadcast_parameters(model)         # what DDP does at construction (ref: main.py:61)
    net, sync = model, None
    if world > 1:
        if args.allreduce == "ddp":     # exactly the reference's wrapper (main.py:61)
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
        else:                           # all-reduce of the flat gradient buffers overlapped with the backward pass
            sync = parallel.GradSync(model, overlap=args.allreduce != "flat",
                                     transport="p2p" if args.allreduce == "p2p" else "nccl").install()
            opt.grad_scale = 1.0 / world
    dev_batch = tuple(t.to(device) for t in batch)
    host_batch = tuple(t.pin_memory() for t in batch)

    class LossReader:
        """D2H read of every step's loss without stalling the launch queue: step k's loss is copied to pinned host
        memory asynchronously and READ while step k+1 runs (the reference reads loss.item() every step for its progress
        bar, main.py:86); `drain()` reads the last one - inside the timed region."""

        def __init__(self):
            self.buf = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
            self.ev = [torch.cuda.Event(), torch.cuda.Event()]
            self.k, self.values = 0, []

        def push(self, loss):
            i = self.k % 2
            self.buf[i].copy_(loss.detach().reshape(1), non_blocking=True)
            self.ev[i].record()
            if self.k > 0:
                self._read((self.k - 1) % 2)
            self.k += 1

        def _read(self, i):
            self.ev[i].synchronize()
            self.values.append(float(self.buf[i][0]))

        def drain(self):
            if self.k > 0 and len(self.values) < self.k:
                self._read((self.k - 1) % 2)
            return self.values[-1] if self.values else None

    reader = LossReader()

    def eager_step(data, read_loss, do_sync=True):
        opt.zero_grad()
        loss_dict, _, _, _ = net(data, 25, STAGE, RATIO)
        loss = loss_fn(loss_dict, STAGE)
        loss.backward()
        if sync is not None and do_sync:
            sync.finish()
        opt.step()
        if read_loss:
            reader.push(loss)
        return loss

    use_graph = args.graph and not (world > 1 and args.allreduce == "ddp")
    if use_graph:      # forward + loss + backward (+ the gradient all-reduce) replayed from one CUDA graph
        from mc_nerf_b200.graph import GraphedTrainStep
        gstep = GraphedTrainStep(model, loss_fn)
        if sync is not None:
            gstep.after_backward.append(sync.finish)

        def step(data, read_loss):
            loss = gstep(data, 25, STAGE, RATIO)      # (`gstep` is rebound below for the exchange-free timing)
            if not data[0].is_cuda:
                gstep.prefetch(data)          # e2e: the next step's H2D upload overlaps this step's graph
            opt.step()
            if read_loss:
                reader.push(loss)
            return loss
    else:
        step = eager_step

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, data, read_loss, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            last = fn(data, read_loss)
        if read_loss:
            reader.drain()                 # the last step's loss is read inside the timed region too
        timed.host_ms = (time.perf_counter() - t0) * 1e3 / steps      # CPU time to ISSUE a step (no sync inside)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    for _ in range(max(args.warmup, 3)):
        step(dev_batch, False)
    sampler = None
    if full and rank == 0 and os.environ.get("MCNERF_NO_CLOCKS") is None:
        sampler = ClockSampler(local, period=float(os.environ.get("MCNERF_CLOCK_PERIOD", "0.05")))
        sampler.start()
    ms, last = timed(step, dev_batch, False, args.steps)
    res = dict(strong=strong, rays_rank=rays_rank, ms=ms, host_issue_ms=timed.host_ms, use_graph=use_graph,
               clocks=sampler.stop() if sampler else None, loss=float(last.item()) if torch.is_tensor(last) else float(last))
    rays_step = rays_rank * world
    res["value"] = rays_step * args.steps / (ms / 1e3)
    res["ranks_identical"] = parallel.parameters_identical(model) if world > 1 else None
    for _ in range(2):
        step(host_batch, True)
    reader.drain()
    n_read0 = len(reader.values)
    ms_e2e, _ = timed(step, host_batch, True, args.steps)
    res["losses_read_e2e"] = len(reader.values) - n_read0          # == steps: every step's loss reached the host
    res["ms_e2e"] = ms_e2e
    res["e2e"] = rays_step * args.steps / (ms_e2e / 1e3)
    res["h2d"] = sum(t.numel() * t.element_size() for t in batch)
    res["collectives_per_step"] = None
    if sync is not None:
        sync.n_collectives = 0
        ms2, _ = timed(step, dev_batch, False, 5)
        res["collectives_per_step"] = sync.n_collectives / 5
    if not full:
        if sync is not None:
            sync.uninstall()
        return res

    # exposed (non-overlapped) time of the gradient exchange: the same graph-replayed steps captured once more WITHOUT
    # it (ranks then diverge: timing only, done after every parity-relevant measurement)
    if sync is not None and use_graph:
        ms_with, _ = timed(step, dev_batch, False, args.steps)
        sync.uninstall()
        gstep = GraphedTrainStep(model, loss_fn)          # `step` picks up the new (exchange-free) graph
        for _ in range(3):
            step(dev_batch, False)
        ms_without, _ = timed(step, dev_batch, False, args.steps)
        res["allreduce_exposed_us"] = round((ms_with - ms_without) / args.steps * 1e3, 1)
        sync.install()
    L = lib()
    if rank == 0:
        L.profile_begin()
    n0 = L.launch_count()
    for _ in range(3):                 # every rank steps (the step contains the gradient all-reduce)
        eager_step(dev_batch, False)
    barrier()
    res["launches"] = (L.launch_count() - n0) // 3 * args.steps     # the graph replays exactly the eager step's kernels
    res["prof"] = L.profile_end() if rank == 0 else None
    if rank == 0:
        from mc_nerf_b200 import render
        res["n_fine"] = int(render.LAST["n_rows_dev"].item()) if render.LAST.get("n_rows_dev") is not None else render.LAST["n_rows"]
    if sync is not None:
        sync.uninstall()
    return res


def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the MC-NeRF hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    scaling = args.scaling or ("strong" if world > 1 else "weak")
    strong = scaling == "strong" and world > 1
    other = None
    if world > 1 and not args.single_mode:        # the other sharding mode rides along as an extra key of the same line
        other = measure(args, not strong, rank, world, local, device, full=False)
    m = measure(args, strong, rank, world, local, device, full=True)
    ms, rays_rank = m["ms"], m["rays_rank"]
    roof = None
    if rank == 0:
        prof, n_fine = m["prof"], m["n_fine"]
        evals = rays_rank * SC + n_fine
        mlp_ms = sum(v for k, v in prof.items() if k.startswith("mcnerf_mlp_")) / 3
        comp_ms = sum(v for k, v in prof.items() if k.startswith(("mcnerf_composite", "mcnerf_sigma2w", "mcnerf_coarse_tail",
                                                                  "mcnerf_fine_tail"))) / 3
        peaks = load_peaks()
        flops = 6.0 * MACS_PER_EVAL * evals          # fwd + dgrad + wgrad
        ach = flops / (mlp_ms / 1e3) / 1e12
        traffic, traffic_src, ncu = None, None, {}
        for name in ("r02_ncu_summary.json", "r01_ncu_summary.json"):
            try:   # DRAM bytes per MLP evaluation from the committed ncu --set full capture (dram__bytes_read+write)
                ncu = json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"]
                per_eval = sum(k["dram_bytes_per_eval"] for n, k in ncu.items() if n != "mlp_tc_fwd_k<0>")   # <0> = inference variant
                traffic = round(per_eval * evals / 1e9, 3)
                traffic_src = f"GB per step = ncu dram bytes/evaluation (fwd+bwd-chain+wgrad, profiles/{name}) x evaluations"
                break
            except Exception:
                continue
        per_kernel = []
        for label, ncu_name, bound in (("mcnerf_mlp_tc_fwd", "mlp_tc_fwd_k<1>", "tensor"),
                                       ("mcnerf_mlp_tc_bwd[chain]", "mlp_tc_bwd_k", "tensor"),
                                       ("mcnerf_mlp_tc_bwd[wgrad]", "mlp_tc_wgrad_k", "hbm")):
            if label not in prof:
                continue
            k_ms = prof[label] / 3
            if bound == "tensor":
                a_k, pk, unit = 2.0 * MACS_PER_EVAL * evals / (k_ms / 1e3) / 1e12, peaks["tensor"], "TFLOP/s"
            else:
                a_k, pk, unit = WGRAD_BYTES_PER_EVAL * evals / (k_ms / 1e3) / 1e9, peaks["hbm"], "GB/s"
            tr = ncu.get(ncu_name, {}).get("dram_bytes_per_eval")
            per_kernel.append(dict(kernel=ncu_name, entry=label, bound=bound, achieved=round(a_k, 1), peak=pk, unit=unit,
                                   frac=round(a_k / pk, 4), ms_per_step=round(k_ms, 3), launches_per_step=2,
                                   traffic=round(tr * evals / 1e9, 3) if tr else None, traffic_unit="GB per step",
                                   share_of_step=round(k_ms / (ms / args.steps), 3)))
        roof = dict(bound="tensor", achieved=round(ach, 2), peak=peaks["tensor"], unit="TFLOP/s", kernels=per_kernel,
                    frac=round(ach / peaks["tensor"], 4), traffic=traffic, traffic_unit="GB", traffic_source=traffic_src,
                    peak_source=peaks["src"],
                    kernel="mcnerf_mlp_* (fwd+bwd, coarse+fine)", kernel_ms_per_step=round(mlp_ms, 3),
                    mlp_evals_per_step=evals, fine_selected_frac=round(n_fine / (rays_rank * SC * SCALE), 4),
                    kernel_ms_by_name={k: round(v / 3, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:12]})
        # compositing kernels against the HBM roofline (SURVEY section 8d algorithmic bytes: 3348 B fwd + 6156 B bwd per ray)
        if comp_ms > 0:
            gbs = 9504.0 * rays_rank / (comp_ms / 1e3) / 1e9
            roof["compositing"] = dict(bound="hbm", achieved=round(gbs, 1), peak=peaks["hbm"], unit="GB/s",
                                       frac=round(gbs / peaks["hbm"], 4), kernel_ms_per_step=round(comp_ms, 4))
    cpu, ref_gpu = None, None
    if rank == 0 and world == 1 and roof is not None:
        roof["dense"] = dense_leg(args, device)
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference(steps=2, warmup=1, rays=args.rays)
        ref_gpu = reference_gpu_eager(rays=args.rays)
    if rank == 0:
        def side(r):
            return dict(scaling="strong" if r["strong"] else "weak", value=round(r["value"], 1), unit=UNIT,
                        ms_per_step=round(r["ms"] / args.steps, 3), rays_per_step_per_gpu=r["rays_rank"],
                        e2e=round(r["e2e"], 1), ranks_identical=r["ranks_identical"],
                        collectives_per_step=r["collectives_per_step"])
        line = dict(metric=METRIC, value=round(m["value"], 1), unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=round(ms / args.steps, 3), higher_is_better=True,
                    scaling="strong" if strong else "weak", vs_baseline=None,
                    dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
                    config=workload_config(rays_rank, args.img, world, strong, args.rays),
                    impl_notes=dict(precision=args.precision,
                                    issue=("CUDA-graph replay of forward+loss+backward (+ gradient all-reduce) per step "
                                           "(GraphedTrainStep); RAdam.step launched eagerly") if m["use_graph"] else "eager launches",
                                    allreduce=(f"{args.allreduce}: fine network's flat gradient buffer all-reduced on a "
                                               "communication stream inside backward, coarse + camera gradients after it "
                                               + ("(mcnerf_allreduce_p2p: one two-shot kernel per buffer over NVLink-mapped "
                                                  "symmetric memory); " if args.allreduce == "p2p" else "(NCCL); ")
                                               + "1/N folded into RAdam") if world > 1 and args.allreduce != "ddp" else
                                              ("DistributedDataParallel" if world > 1 else None)),
                    e2e=dict(value=round(m["e2e"], 1), unit=UNIT, h2d_bytes_per_step=m["h2d"], d2h_bytes_per_step=4,
                             ms_per_step=round(m["ms_e2e"] / args.steps, 3), losses_read=m["losses_read_e2e"],
                             how="pinned host inputs uploaded per step (next step's upload overlaps the running graph); "
                                 "every step's loss copied to pinned memory and read on the host one step later"),
                    allreduce_collectives_per_step=m["collectives_per_step"],
                    allreduce_exposed_us=m.get("allreduce_exposed_us"), ranks_identical=m["ranks_identical"],
                    other_scaling=side(other) if other else None,
                    gpu_launches=int(m["launches"]), host_issue_ms_per_step=round(m["host_issue_ms"], 3), clocks=m["clocks"],
                    roofline=roof, cpu_baseline=cpu, reference_gpu_eager=ref_gpu, loss=m["loss"])
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def dense_leg(args, device, steps=10, warmup=3):
    """at random init only ~77 % of the fine samples pass the selection threshold, so the roofline is also
    reported for the DENSE step - the same workload with sample_weight_thresh = 0 (every one of the 64 + 128 samples of
    every ray goes through both networks: 724.9 MFLOP per ray).  Eager launches, device-resident inputs."""
    from mc_nerf_b200 import synthetic as syn
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss, RAdam
    from mc_nerf_b200._lib import lib
    sp = syn.make_sys_param(n_cam=N_CAM, img_h=args.img, img_w=args.img, batch=args.rays, samples=SC, scale=SCALE, device=device,
                            with_images=False)
    sp["mlp_precision"] = args.precision
    sp["pixel_sampler"] = "device"
    sp["sample_weight_thresh"] = 0.0
    torch.manual_seed(42)
    model = MC_Model(sp).to(device)
    with torch.no_grad():
        for k, v in syn.init_camera_weights(sp).items():
            getattr(model, k).copy_(v)
    loss_fn = MC_NeRF_Loss(sp)
    opt = RAdam([p for p in model.parameters()], lr=5e-4, eps=1e-8, weight_decay=4e-4)
    batch = tuple(t.to(device) for t in syn.make_train_batch(sp, img_id=3, seed=11))

    def step():
        opt.zero_grad()
        loss_dict, _, _, _ = model(batch, 25, STAGE, RATIO)
        loss_fn(loss_dict, STAGE).backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    L = lib()
    L.profile_begin()
    step()
    torch.cuda.synchronize()
    prof = L.profile_end()
    mlp_ms = sum(v for k, v in prof.items() if k.startswith("mcnerf_mlp_tc_fwd") or k.startswith("mcnerf_mlp_tc_bwd"))
    evals = args.rays * (SC + SC * SCALE)
    peaks = load_peaks()
    ach = 6.0 * MACS_PER_EVAL * evals / (mlp_ms / 1e3) / 1e12
    return dict(what="same step with sample_weight_thresh = 0: all 64 + 128 samples of every ray evaluated", value=round(args.rays / (ms / 1e3), 1),
                unit=UNIT, ms_per_step=round(ms, 3), steps=steps, mlp_evals_per_step=evals, kernel_ms_per_step=round(mlp_ms, 3),
                achieved=round(ach, 2), peak=peaks["tensor"], frac=round(ach / peaks["tensor"], 4), issue="eager launches")


# --------------------------------------------------------------------------------------------- other BASELINE configs
DEMO_METRIC, STRESS_METRIC = "demo_render_rays_per_sec_64c_128f", "train_rays_per_sec_64c_256f"
DEMO_CHUNK, STRESS_RAYS, STRESS_MICRO, STRESS_SCALE = 65536, 65536, 8, 4


def demo_config(img, chunk):
    return dict(workload="BASELINE configs[3]: --demo render of whole 800x800 test views of the 110-camera rig, inference-only "
                         "sampling + coarse/fine 8x256 MLPs + compositing (64+128 samples), one step = one view",
                rays_per_step_per_gpu=img * img, img=img, chunk_rays=chunk, parallelism="single device",
                l2="every chunk's activations stay on chip; ray / output tensors of a view (26 MB) are streamed once")


def stress_config():
    return dict(workload="BASELINE configs[4]: Room-style stress, 65536 rays per optimiser step x (64 coarse + 256 fine) samples, "
                         "coarse+fine 8x256 MLPs, GLOBAL_OPTIM stage, one step = fwd+loss+bwd over 8 accumulation "
                         "micro-batches of 8192 rays + RAdam",
                rays_per_step_per_gpu=STRESS_RAYS, img=IMG, micro_batches=STRESS_MICRO, parallelism="single device",
                l2="per-step working set (activation stash > 3 GB per micro-batch) exceeds the 126 MB L2; no flush needed")


def run_demo(args):
    """BASELINE configs[3] through the drop-in API: MC_Model(mode=1)(img_idx) (ref: model/mc_nerf.py:106-122).
    value: views rendered with rays and outputs resident on the device; e2e: the public call, whose three outputs
    arrive in pinned host memory (chunk-wise D2H on a side stream, inside the timed region)."""
    import tempfile
    from mc_nerf_b200 import synthetic as syn, render
    from mc_nerf_b200.model import MC_Model
    from mc_nerf_b200._lib import lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the MC-NeRF hot path has no CPU fallback)")
    # N > 1 (torchrun): the 110 views are independent units - every rank renders its own views with its own replica of
    # the networks, no data-path collective (SURVEY 8e / "replicas"); the process group only serves the barrier and the
    # max-over-ranks time
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    kw = dict(n_cam=N_CAM, img_h=args.img, img_w=args.img, batch=DEMO_CHUNK, samples=SC, scale=SCALE, with_images=False)
    sp = syn.make_sys_param(device=dev, **kw)
    sp["mlp_precision"] = args.precision
    torch.manual_seed(42)
    m = MC_Model(sp).to(dev)
    with torch.no_grad():
        for k, v in syn.init_camera_weights(sp).items():
            getattr(m, k).copy_(v)
    tmp = tempfile.mkdtemp()
    m.nerf.weights_pth = tmp
    m.nerf.save_model(m, 0)
    sp2 = syn.make_sys_param(device=dev, mode=1, **kw)
    sp2["mlp_precision"] = args.precision
    sp2["demo_ckpt"] = m.nerf.file_path
    demo = MC_Model(sp2).to(dev).eval()
    n = args.img * args.img

    def view_resident(v):
        rays_d, rays_o = demo.get_rays(demo.test_pose, v, demo.intr_test_inv.to(dev))
        outs = []
        for ii in range(0, n, DEMO_CHUNK):
            outs.append(demo.nerf(rays_d[ii:ii + DEMO_CHUNK], rays_o[ii:ii + DEMO_CHUNK]))
        return outs

    def timed(fn, steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn((1 + i * world + rank) % N_CAM)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            view_resident((i * world + rank) % N_CAM)
        sampler = ClockSampler(local)
        sampler.start()
        ms = timed(view_resident, args.steps)
        clocks = sampler.stop()
        for i in range(2):
            demo(torch.tensor([i]))
        ms_e2e = timed(lambda v: demo(torch.tensor([v])), args.steps)
        L = lib()
        L.profile_begin()
        n0 = L.launch_count()
        view_resident(7)
        torch.cuda.synchronize()
        evals, fine_rows = 0, []
        launches = L.launch_count() - n0
        prof = L.profile_end()
        # MLP evaluations of one view: coarse grid + the fine samples selected on the device
        rays_d, rays_o = demo.get_rays(demo.test_pose, 7, demo.intr_test_inv.to(dev))
        for ii in range(0, n, DEMO_CHUNK):
            demo.nerf(rays_d[ii:ii + DEMO_CHUNK], rays_o[ii:ii + DEMO_CHUNK])
            nd = render.LAST.get("n_rows_dev")
            fine_rows.append(int(nd.item()) if nd is not None else int(render.LAST["n_rows"]))
        evals = n * SC + sum(fine_rows)
    peaks = load_peaks()
    mlp_ms = sum(v for k, v in prof.items() if k.startswith("mcnerf_mlp_"))
    ach = 2.0 * MACS_PER_EVAL * evals / (mlp_ms / 1e3) / 1e12
    roof = dict(bound="tensor", achieved=round(ach, 2), peak=peaks["tensor"], unit="TFLOP/s", frac=round(ach / peaks["tensor"], 4),
                traffic=None, peak_source=peaks["src"], kernel="mlp_tc_fwd_k<0> (inference variant, coarse+fine)",
                kernel_ms_per_step=round(mlp_ms, 3), mlp_evals_per_step=evals,
                fine_selected_frac=round(sum(fine_rows) / (n * SC * SCALE), 4),
                kernel_ms_by_name={k: round(v, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8]})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    cpu = None if (args.no_cpu or world > 1) else cpu_reference_demo(steps=1, warmup=1, rays=8192)
    n = n * world                                   # rays of one "step" = one view per rank
    line = dict(metric=DEMO_METRIC, value=round(n * args.steps / (ms / 1e3), 1), unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=round(ms / args.steps, 3), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
                config=dict(demo_config(args.img, DEMO_CHUNK), parallelism="single device" if world == 1 else
                            f"{world} replicas, one view per rank and step, no collective"),
                e2e=dict(value=round(n * args.steps / (ms_e2e / 1e3), 1), unit=UNIT, h2d_bytes_per_step=8,
                         d2h_bytes_per_step=n * 5 * 4, ms_per_step=round(ms_e2e / args.steps, 3)),
                gpu_launches=int(launches) * args.steps * world, clocks=clocks, roofline=roof, cpu_baseline=cpu)
    print(json.dumps(line), flush=True)


def cpu_reference_demo(steps, warmup, rays):
    """The reference's test render (ref: model/mc_nerf.py:648-680) on the host CPUs, a bounded sample of `rays` rays."""
    from mc_nerf_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = reference_modules()
    sp = syn.make_sys_param(n_cam=N_CAM, img_h=IMG, img_w=IMG, batch=rays, samples=SC, scale=SCALE, with_images=False)
    g = torch.Generator().manual_seed(5)
    rd = torch.nn.functional.normalize(torch.randn(rays, 3, generator=g), dim=-1)
    ro = torch.randn(rays, 3, generator=g) * 0.3
    torch.manual_seed(42)
    if ref is not None:
        nerf = ref[1](sp)

        def one():
            with torch.no_grad():
                nerf.render_rays_test(rd, ro, nerf.nerf_coarse, nerf.nerf_fine)
    else:
        from oracle import mcnerf_oracle as orc
        cfg = orc.cfg_from_sys_param(sp)
        pc, pf = orc.init_mlp_params(*cfg["coarse"], seed=42), orc.init_mlp_params(*cfg["fine"], seed=43)

        def one():
            rng = syn.draw_step_rng(sp, rays, seed=1, train=False)
            with torch.no_grad():
                orc.render_rays(pc, pf, cfg, rd, ro, rng, train=False)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    sec = (time.perf_counter() - t0) / steps
    return dict(value=round(rays / sec, 2), unit=UNIT, cores=cores, kind="reference" if ref is not None else "port",
                rays_per_step=rays, sec_per_step=round(sec, 3),
                sample=f"{steps} timed test renders (after {warmup} warm-up) of {rays} rays (64+128 samples, 8x256 MLPs), torch "
                       f"CPU fp32, {cores} threads; " + ("unmodified reference NeRF_Model.render_rays_test (baseline/_ref)"
                                                          if ref is not None else "oracle port"))


def run_stress(args):
    """BASELINE configs[4]: 65536 rays per optimiser step, 64 + 256 samples, as 8 accumulation micro-batches (the
    activation + gradient stashes of one 65536-ray batch, ~190 GB, exceed the 180 GB of HBM; the reference would need the
    same accumulation).  With scale 4 the reference's train-only 128-per-ray cap is active (model/mc_nerf.py:630-632)."""
    from mc_nerf_b200 import synthetic as syn, render
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss, RAdam
    from mc_nerf_b200._lib import lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the MC-NeRF hot path has no CPU fallback)")
    dev = "cuda:0"
    torch.cuda.set_device(0)
    rays = STRESS_RAYS // STRESS_MICRO
    main = _stress_mode(args, dev, rays, cap_off=False, steps=args.steps, full=True)
    # the reference's hard-coded 128-per-ray cap truncates a 256-fine run; the same step with the cap
    # switched off (sys_param fine_sample_cap = 256: every selected sample is evaluated) is reported next to it
    main["cap_off"] = _stress_mode(args, dev, rays, cap_off=True, steps=max(1, args.steps // 2), full=False)
    print(json.dumps(main), flush=True)


def _stress_mode(args, dev, rays, cap_off, steps, full):
    from mc_nerf_b200 import synthetic as syn, render
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss, RAdam
    from mc_nerf_b200._lib import lib
    sp = syn.make_sys_param(n_cam=N_CAM, img_h=args.img, img_w=args.img, batch=rays, samples=SC, scale=STRESS_SCALE,
                            device=dev, with_images=False)
    sp["mlp_precision"] = args.precision
    sp["pixel_sampler"] = "device"
    if cap_off:
        sp["fine_sample_cap"] = SC * STRESS_SCALE
    torch.manual_seed(42)
    model = MC_Model(sp).to(dev)
    with torch.no_grad():
        for k, v in syn.init_camera_weights(sp).items():
            getattr(model, k).copy_(v)
    loss_fn = MC_NeRF_Loss(sp)
    opt = RAdam(list(model.parameters()), lr=5e-4, eps=1e-8, weight_decay=4e-4)
    host = [tuple(t.pin_memory() for t in syn.make_train_batch(sp, img_id=(3 + 7 * i) % N_CAM, seed=11 + i))
            for i in range(STRESS_MICRO)]
    devb = [tuple(t.to(dev) for t in b) for b in host]

    def step(batches, read_loss=False, count=False):
        opt.zero_grad()
        evals, loss = 0, None
        for b in batches:
            loss_dict, _, _, _ = model(b, 25, STAGE, 0.8)
            loss = loss_fn(loss_dict, STAGE) / STRESS_MICRO
            loss.backward()
            if count:      # (one host sync per micro-batch: only in the untimed warm-up)
                nd = render.LAST.get("n_rows_dev")
                evals += rays * SC + (int(nd.item()) if nd is not None else int(render.LAST["n_rows"]))
        opt.step()
        if read_loss:
            loss.item()
        return evals

    def timed(batches, read_loss, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(batches, read_loss)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(max(args.warmup, 3) - 1):
        step(devb)
    evals = step(devb, count=True)
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(devb, False, steps)
    clocks = sampler.stop()
    if not full:
        return dict(value=round(STRESS_RAYS * steps / (ms / 1e3), 1), unit=UNIT, steps=steps, ms_per_step=round(ms / steps, 3),
                    mlp_evals_per_step=evals, fine_sample_cap=SC * STRESS_SCALE, clocks=clocks,
                    peak_mem_gb=round(torch.cuda.max_memory_allocated() / 1e9, 1))
    step(host, True)
    ms_e2e = timed(host, True, steps)
    L = lib()
    L.profile_begin()
    n0 = L.launch_count()
    step(devb)
    torch.cuda.synchronize()
    launches = L.launch_count() - n0
    prof = L.profile_end()
    peaks = load_peaks()
    mlp_ms = sum(v for k, v in prof.items() if k.startswith("mcnerf_mlp_"))
    ach = 6.0 * MACS_PER_EVAL * evals / (mlp_ms / 1e3) / 1e12
    roof = dict(bound="tensor", achieved=round(ach, 2), peak=peaks["tensor"], unit="TFLOP/s", frac=round(ach / peaks["tensor"], 4),
                traffic=None, peak_source=peaks["src"], kernel="mcnerf_mlp_* (fwd+bwd, coarse+fine)",
                kernel_ms_per_step=round(mlp_ms, 3), mlp_evals_per_step=evals,
                kernel_ms_by_name={k: round(v, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8]})
    h2d = sum(t.numel() * t.element_size() for b in host for t in b)
    line = dict(metric=STRESS_METRIC, value=round(STRESS_RAYS * steps / (ms / 1e3), 1), unit=UNIT, n_gpus=1,
                steps=steps, warmup=max(args.warmup, 3), ms_per_step=round(ms / steps, 3), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
                config=stress_config(),
                e2e=dict(value=round(STRESS_RAYS * steps / (ms_e2e / 1e3), 1), unit=UNIT, h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=4, ms_per_step=round(ms_e2e / steps, 3)),
                gpu_launches=int(launches) * steps, clocks=clocks, roofline=roof, cpu_baseline=None,
                peak_mem_gb=round(torch.cuda.max_memory_allocated() / 1e9, 1))
    return line


# --------------------------------------------------------------------------------------------- reference arm
def workload_config(rays_rank, img, world, strong, total_rays):
    """`config` of the JSON line: names the workload only, identical for the GPU arm and the reference arm."""
    if world > 1:
        par = (f"dp{world}: one camera's {total_rays}-ray batch split into {rays_rank}-ray slices" if strong
               else f"dp{world}: one camera's {total_rays}-ray batch per rank") + ", all-reduce of MLP+camera gradients"
    else:
        par = "single device"
    return dict(workload="BASELINE configs[1]: 110 cameras, 800x800, 4096 rays/batch, 64 coarse + 128 fine samples, "
                         "coarse+fine 8x256 MLPs, GLOBAL_OPTIM stage, one step = fwd+loss+bwd+RAdam",
                rays_per_step_per_gpu=rays_rank, img=img, parallelism=par,
                l2="per-step working set (activation stash > 3 GB) exceeds the 126 MB L2; no flush needed")


def reference_modules():
    """The UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py) or None."""
    try:
        from baseline import ref_loader
        if ref_loader.reference_root() is None:
            return None
        return ref_loader.import_reference()
    except Exception as e:      # noqa: BLE001 - an unimportable reference falls back to the port, loudly
        sys.stderr.write(f"bench.py: reference modules not importable ({e!r}); timing the oracle port instead\n")
        return None


def cpu_reference(steps, warmup, rays, budget_s=None):
    """The reference's own CPU implementation of one train step, all host threads.

    kind "reference": the unmodified modules from baseline/_ref - MC_Model.forward -> MC_NeRF_Loss -> backward ->
    RAdam.step exactly as reference main.py:79-85 (incl. its two full-image get_rays, the 110-iteration inverse loop and
    randperm(H*W)); kind "port": oracle/mcnerf_oracle.py, only when baseline/_ref is absent.
    Times EXACTLY `steps` steps after `warmup`.  budget_s: if the first warm-up step predicts a longer run, the rays
    per step are reduced (a bounded sample of the same workload; reported)."""
    from mc_nerf_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = reference_modules()

    def build(n_rays):
        sp = syn.make_sys_param(n_cam=N_CAM, img_h=IMG, img_w=IMG, batch=n_rays, samples=SC, scale=SCALE,
                                with_images=False)
        batch = syn.make_train_batch(sp, img_id=3)
        torch.manual_seed(42)
        if ref is not None:
            MC_Model, _, MC_NeRF_Loss, _, net_utils = ref
            m = MC_Model(sp)
            with torch.no_grad():
                for k, v in syn.init_camera_weights(sp).items():
                    getattr(m, k).copy_(v)
            loss_fn = MC_NeRF_Loss(sp)
            opt = net_utils.RAdam([p for p in m.parameters()], lr=5e-4, eps=1e-8, weight_decay=4e-4)

            def one(i):
                opt.zero_grad()
                loss_dict, _, _, _ = m(batch, 25, STAGE, RATIO)
                loss = loss_fn(loss_dict, STAGE)
                loss.backward()
                opt.step()
                return float(loss.item())
            return one
        from oracle import mcnerf_oracle as orc
        cfg = orc.cfg_from_sys_param(sp)
        cam = {k: v.clone().requires_grad_(True) for k, v in syn.init_camera_weights(sp).items()}
        pc = {k: v.clone().requires_grad_(True) for k, v in orc.init_mlp_params(*cfg["coarse"], seed=42).items()}
        pf = {k: v.clone().requires_grad_(True) for k, v in orc.init_mlp_params(*cfg["fine"], seed=43).items()}

        def one(i):
            rng = syn.draw_step_rng(sp, n_rays, seed=100 + i)
            for d in (cam, pc, pf):
                for v in d.values():
                    v.grad = None
            return float(orc.train_step(cam, pc, pf, cfg, batch, rng, step_r=RATIO, stage=STAGE)["loss"])
        return one

    one = build(rays)
    t0 = time.perf_counter()
    one(0)
    first = time.perf_counter() - t0
    if budget_s is not None and first * (steps + warmup) > budget_s and rays > 512:
        shrink = max(512, int(rays * budget_s / (first * (steps + warmup))) // 512 * 512)
        sys.stderr.write(f"bench.py: reference step takes {first:.1f} s at {rays} rays; sampling {shrink} rays/step\n")
        rays = shrink
        one = build(rays)
        one(0)
    for i in range(1, warmup):
        one(i)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        one(warmup + i)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    kind = "reference" if ref is not None else "port"
    what = ("unmodified reference modules (baseline/_ref): MC_Model.forward -> MC_NeRF_Loss -> backward -> RAdam.step"
            if ref is not None else "oracle port (oracle/mcnerf_oracle.py; baseline/_ref absent): fwd+loss+bwd, no optimiser")
    return dict(value=round(rays / sec, 2), unit=UNIT, cores=cores, kind=kind, rays_per_step=rays,
                sample=f"{steps} timed train steps (after {warmup} warm-up) of {rays} rays of the same workload "
                       f"(110 cameras, 800x800, 64+128 samples, 8x256 MLPs), torch CPU fp32, {cores} threads; {what}",
                sec_per_step=round(sec, 3))


def reference_gpu_eager(steps=5, warmup=3, rays=RAYS):
    """Context only: the UNMODIFIED reference modules with device_type="cuda" - PyTorch eager on
    the same B200, same workload and step (forward -> loss -> backward -> its own RAdam).  The reference ships no CUDA
    kernels; this is the only "existing GPU implementation" of the path."""
    ref = reference_modules()
    if ref is None:
        return dict(unavailable="baseline/_ref not installed")
    try:
        from mc_nerf_b200 import synthetic as syn
        dev = "cuda:0"
        MC_Model, _, MC_NeRF_Loss, _, net_utils = ref
        sp = syn.make_sys_param(n_cam=N_CAM, img_h=IMG, img_w=IMG, batch=rays, samples=SC, scale=SCALE, device=dev,
                                with_images=False)
        torch.manual_seed(42)
        m = MC_Model(sp).to(dev)
        with torch.no_grad():
            for k, v in syn.init_camera_weights(sp).items():
                getattr(m, k).copy_(v)
        loss_fn = MC_NeRF_Loss(sp)
        opt = net_utils.RAdam([p for p in m.parameters()], lr=5e-4, eps=1e-8, weight_decay=4e-4)
        batch = tuple(t.to(dev) for t in syn.make_train_batch(sp, img_id=3))

        def one():
            opt.zero_grad()
            loss_dict, _, _, _ = m(batch, 25, STAGE, RATIO)
            loss = loss_fn(loss_dict, STAGE)
            loss.backward()
            opt.step()
            return loss
        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = dict(value=round(rays / (ms / 1e3), 1), unit=UNIT, ms_per_step=round(ms, 2), steps=steps,
                   what="unmodified reference modules (baseline/_ref), device_type=cuda, PyTorch eager fp32, same step")
        del m, opt
        torch.cuda.empty_cache()
        return out
    except Exception as e:      # noqa: BLE001 - context measurement only
        return dict(unavailable=repr(e)[:200])


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cpu = cpu_reference(steps=args.steps, warmup=max(1, args.warmup), rays=args.rays, budget_s=280.0)
    cfg = workload_config(args.rays, args.img, 1, False, args.rays)
    if cpu["rays_per_step"] != args.rays:
        cfg["bounded_sample"] = f"{cpu['rays_per_step']} rays/step of the {args.rays}-ray batch (CPU time budget)"
    line = dict(metric=METRIC, value=cpu["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=max(1, args.warmup), ms_per_step=round(cpu["sec_per_step"] * 1e3, 1), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference", config=cfg,
                cpu_baseline=cpu, e2e=dict(value=cpu["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MCNERF_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--rays", type=int, default=RAYS)
    ap.add_argument("--img", type=int, default=IMG)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--allreduce", default=os.environ.get("MCNERF_ALLREDUCE", "p2p"), choices=["p2p", "overlap", "flat", "ddp"],
                    help="p2p (default): libmcnerf's two-shot NVLink all-reduce kernel on symmetric gradient buffers, on a "
                         "communication stream inside backward; overlap: the same schedule with NCCL all-reduces; flat: "
                         "NCCL on the main stream; ddp: the reference's DistributedDataParallel wrapper")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1 default: strong = ONE --rays batch split into equal ray slices across the ranks (BASELINE "
                         "configs[2], north_star); weak (the reference's DDP semantic): every rank renders its own "
                         "camera's --rays batch.  The other mode is measured too and reported under `other_scaling`.")
    ap.add_argument("--single-mode", action="store_true", help="N > 1: measure only the selected scaling mode")
    ap.add_argument("--no-graph", dest="graph", action="store_false", default=os.environ.get("MCNERF_BENCH_GRAPH", "1") != "0",
                    help="issue every step with eager launches instead of replaying the captured CUDA graph")
    ap.add_argument("--workload", default="train", choices=["train", "demo", "stress"],
                    help="train (default): BASELINE configs[1]/[2], the headline metric; demo: configs[3], whole-view "
                         "inference renders; stress: configs[4], 65536 rays x (64+256) samples per optimiser step")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "demo":
            rank, _, _ = dist_env()
            if rank == 0:
                cpu = cpu_reference_demo(steps=args.steps, warmup=max(1, args.warmup), rays=8192)
                cfg = demo_config(IMG, DEMO_CHUNK)
                cfg["bounded_sample"] = "8192 rays per step of the 640000-ray view (CPU time budget)"
                print(json.dumps(dict(metric=DEMO_METRIC, value=cpu["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                                      warmup=max(1, args.warmup), ms_per_step=round(cpu["sec_per_step"] * 1e3, 1),
                                      higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                                      impl="reference", config=cfg, cpu_baseline=cpu, gpu_launches=0,
                                      e2e=dict(value=cpu["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))),
                      flush=True)
        else:
            run_reference(args)
    elif args.workload == "demo":
        run_demo(args)
    elif args.workload == "stress":
        run_stress(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
