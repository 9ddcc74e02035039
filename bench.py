#!/usr/bin/env python
"""Benchmark of the render() hot path (BASELINE.json: rays/s for render() forward, 1200x680 image in
1024-ray chunks, 64+64 samples, shipped network shapes, random-init synthetic model).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16|bf16x3]

A "step" is one full-image render (816 000 rays) per GPU.  One JSON line is printed by rank 0:
  value    -- rays/s with rays and uniform draws already resident in HBM (device-timed, max over ranks)
  e2e      -- the same metric through VectorFieldNerf.render() the way evaluation/methods.py:516-530 calls
              it: per 1024-ray chunk, pinned-host uv/pose/intrinsics -> device, CPU-generator draws -> device,
              render, rgb/depth -> host; all inside the timed region
  roofline -- the dominant kernel (VF MLP chain) timed alone with CUDA events, algorithmic FLOPs / time
  cpu_baseline -- the oracle port of the reference's CPU path on this box's host cores (rank 0, N=1)
`--impl reference` times that CPU path alone (bounded: one 1024-ray chunk per step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W_IMG, FOCAL = 680, 1200, 600.0
N_COARSE, N_FINE = 64, 64
F_VF, F_RN = 1_050_112, 542_720                    # algorithmic FLOP / point
A_FWD = (N_COARSE + N_FINE) * (F_VF + F_RN)        # FLOP / ray, unique points only
CASE = dict(seed=0, vf_hidden=(256,) * 8, feat=256, rn_hidden=(256,) * 4, n_coarse=N_COARSE, n_fine=N_FINE,
            max_samples=100, perturb=False, near=0.0, far=6.0, fine_range=0.3, window=11,
            dir_to_normal_th=-0.2, vf_gain=2.0)     # evaluation settings: evaluate.py:30,32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VFNERF_PRECISION", "bf16"))
    ap.add_argument("--chunk", type=int, default=0, help="rays per render() call of the device-resident leg")
    ap.add_argument("--rays", type=int, default=H * W_IMG, help="rays per step (default: the full image)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    return ap.parse_args()


def ncu_traffic(points_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused RENDER launch from the committed `ncu --set full` capture
    (profiles/run_render_points.py, same points per launch); None if the capture does not match this launch size."""
    path = os.path.join(ROOT, "profiles", "r01_prof_render_final_raw.csv")
    if points_per_launch != 65536 * (N_COARSE + N_FINE) or not os.path.exists(path):
        return None, None
    import csv
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(vals[i]) * scale[units[i]]
    return tot, "profiles/r01_prof_render_final_raw.csv (ncu --set full, one launch of the same size)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_pack():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vfn_testutil as U
    return U


def cpu_reference_rays_per_s(n_steps, n_warm, threads=None):
    """The reference's CPU path (oracle port, torch-CPU fp32 eager, all host threads): one 1024-ray chunk
    per step, deterministic sampling -- the `--gpu cpu` configuration of BASELINE config 1."""
    U = oracle_pack()
    # torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use every host core it is allowed to run on
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    torch.set_num_threads(threads or ncores)
    st = U.S.synthetic_state(CASE["seed"], vf_gain=CASE["vf_gain"])
    R = 1024
    uv, pose, K = U.S.synthetic_rays(R, seed=0, start=0, stride=797)
    _, _, U3 = U.S.synthetic_draws(R, N_COARSE, N_FINE, seed=1)
    t_vals = torch.linspace(0., 1., N_COARSE)
    cfg = U.oracle_cfg(CASE)
    times = []
    with torch.no_grad():
        for i in range(n_warm + n_steps):
            t0 = time.perf_counter()
            U.O.render(st["vf_net"], st["rendering_net"], st["density"], cfg, uv, pose, K, t_vals, None, None, U3)
            dt = time.perf_counter() - t0
            if i >= n_warm:
                times.append(dt)
    med = statistics.median(times)
    return R / med, med, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rps, med, threads = cpu_reference_rays_per_s(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "render_fwd_rays_per_sec", "value": rps, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "render() forward, Replica-shaped 1200x680 image in 1024-ray chunks, 64+64 samples "
                               "(reference arm: one 1024-ray chunk per step on the host CPU)",
                   "rays_per_step": 1024, "n_coarse": N_COARSE, "n_fine": N_FINE},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": "one 1024-ray chunk per step, torch-CPU fp32 eager restatement of the reference "
                                   "(oracle/render_oracle.py); the reference itself is Python and cannot travel"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL banners, library warnings printed
    with printf from native code) was diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # native-code prints (e.g. "NCCL version ...") must not pollute the JSON line
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (the product has no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import types
    from vfnerf_b200 import synthetic as _S
    U = types.SimpleNamespace(S=_S, make_model=_S.make_model)    # this arm imports nothing from oracle/ or tests/
    from vfnerf_b200 import _lib
    _lib.build()
    L = _lib.lib()
    st = U.S.synthetic_state(CASE["seed"], vf_gain=CASE["vf_gain"])
    model = U.make_model(CASE, st, dev, precision=args.precision)
    model.return_ray_dirs = False       # the evaluation caller reads rgb/depth only (methods.py:529-530)
    R = args.rays
    chunk = args.chunk or (4096 if args.precision == "fp32" else 65536)

    # ---- inputs: one synthetic camera per rank (weak scaling: every GPU renders a full image)
    pose1, K1 = U.S.synthetic_camera(seed=rank, height=H, width=W_IMG, focal=FOCAL)
    uv_h = U.S.pixel_grid(H, W_IMG)[:R].contiguous().pin_memory()
    pose_h = pose1.repeat(R, 1, 1).contiguous().pin_memory()
    K_h = K1.repeat(R, 1, 1).contiguous().pin_memory()
    uv_d, pose_d, K_d = uv_h.to(dev), pose_h.to(dev), K_h.to(dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    U3_d = torch.rand(R, N_FINE, device=dev, generator=gen)
    rgb_img = torch.empty(R, 3, device=dev)
    dep_img = torch.empty(R, 1, device=dev)

    def step_resident():
        with torch.no_grad():
            for a in range(0, R, chunk):
                b = min(R, a + chunk)
                out = model.render(pose_d[a:b], uv_d[a:b], K_d[a:b], 0, draws=(None, None, U3_d[a:b]))
                rgb_img[a:b] = out.coarse_rgb_values
                dep_img[a:b] = out.coarse_depth_map
        if world > 1:    # final gather of rgb+depth to rank 0 (16 B/ray), SURVEY.md §8e
            both = torch.cat([rgb_img, dep_img], dim=1)
            gl = [torch.empty_like(both) for _ in range(world)] if rank == 0 else None
            dist.gather(both, gl, dst=0)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    sync_all()
    clocks = ClockSampler(local)
    clocks.start()
    n0 = L.vfnerf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    sync_all()
    launches = L.vfnerf_launch_count() - n0
    clk = clocks.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    value = world * R * args.steps / (ms_total * 1e-3)

    # ---- e2e: the reference-facing call render(pose, pixels, intrinsics, epoch) with HOST buffers, chunk by chunk like
    # evaluation/methods.py:516-530: per chunk H2D of uv/pose/K from pinned memory, the sampler draws on the CPU
    # generator + H2D (like the reference), D2H of rgb/depth into pinned host images; one synchronize at the end of
    # the image.  Timed at the chunk size this path is built for (the headline) and at the reference's 1024-ray chunks,
    # where Python launch overhead, not the GPU, is the limit.
    e2e = None
    if not args.no_e2e:
        rgb_h = torch.empty(R, 3).pin_memory()
        dep_h = torch.empty(R, 1).pin_memory()

        copy_stream = torch.cuda.Stream(device=dev)

        def stage(a, b):
            # the caller's H2D of one chunk (train/vector_field_nerf_train.py:172-174 does it inline); here the next
            # chunk's copy runs on a side stream while the current chunk renders
            with torch.cuda.stream(copy_stream):
                t = (uv_h[a:b].to(dev, non_blocking=True), pose_h[a:b].to(dev, non_blocking=True),
                     K_h[a:b].to(dev, non_blocking=True))
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return t, ev

        def step_e2e(ck):
            h2d = d2h = 0
            main = torch.cuda.current_stream(dev)
            with torch.no_grad():
                nxt = stage(0, min(R, ck))
                for a in range(0, R, ck):
                    b = min(R, a + ck)
                    (px, po, ki), ev = nxt
                    if b < R:
                        nxt = stage(b, min(R, b + ck))
                    main.wait_event(ev)
                    for t in (px, po, ki):
                        t.record_stream(main)
                    out = model.render(po, px, ki, 0)          # draws U3 on the CPU generator + H2D, like the reference
                    rgb_h[a:b].copy_(out.coarse_rgb_values, non_blocking=True)
                    dep_h[a:b].copy_(out.coarse_depth_map, non_blocking=True)
                    h2d += (px.numel() + po.numel() + ki.numel() + (b - a) * N_FINE) * 4
                    d2h += (b - a) * 4 * 4
            return h2d, d2h

        def time_e2e(ck):
            step_e2e(ck)
            sync_all()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            h2d, d2h = step_e2e(ck)
            s1.record()
            sync_all()
            ms2 = torch.tensor([s0.elapsed_time(s1)], device=dev)
            if world > 1:
                dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
            return world * R / (ms2.item() * 1e-3), h2d, d2h
        v_big, h2d, d2h = time_e2e(chunk)
        v_1k, _, _ = time_e2e(1024)
        model.draws_on_device = True
        v_big_dev, _, _ = time_e2e(chunk)
        v_1k_dev, _, _ = time_e2e(1024)
        model.draws_on_device = False
        e2e = {"value": v_big, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "chunk": chunk,
               "draws_on_device": {"value": v_big_dev, "unit": "rays/s", "note": "opt-in (model.draws_on_device): the uniform "
                                   "draws are made on the device instead of 52 M numbers per image from the reference's CPU "
                                   "generator, which is what bounds the default e2e once the kernels are this fast"},
               "chunk_1024": {"value": v_1k, "unit": "rays/s", "note": "the reference's evaluation chunk size; bounded by "
                              "Python launch overhead per render() call, not by the GPU",
                              "draws_on_device": {"value": v_1k_dev, "unit": "rays/s", "note": "opt-in: uniform draws made on "
                                                  "the device instead of the reference's CPU generator + H2D copy"}}}

    # ---- training step (BASELINE config 3): 1024-ray batch, render -> VFLoss terms -> backward -> clip -> Adam,
    # the sequence of train/vector_field_nerf_train.py:177-260.  Timed on the bench precision (bf16: fused tcgen05
    # forward with activation stash, fused tcgen05 dgrad chain, MN-major tcgen05 weight-gradient GEMMs) and, for
    # comparison, on the fp32 CUDA-core path.
    train = None
    if not args.no_train:
        Rt = 1024
        uvt, poset, Kt = uv_d[:Rt], pose_d[:Rt], K_d[:Rt]
        g2 = torch.Generator(device=dev).manual_seed(7 + rank)
        draws_t = (torch.rand(Rt, N_COARSE, device=dev, generator=g2), torch.rand(Rt, N_FINE, device=dev, generator=g2),
                   torch.rand(Rt, N_FINE, device=dev, generator=g2))
        rgb_gt = torch.rand(Rt, 3, device=dev, generator=g2)
        dep_gt = torch.rand(Rt, 1, device=dev, generator=g2) * CASE["far"]

        # the reference's VFLoss with the shipped weights (confs/vf_nerf.conf:77-91), epoch 0, as ONE fused launch
        # (vfnerf_b200/losses.py); sync=False: the per-term values stay on the device instead of six .item() calls
        import types
        from vfnerf_b200.losses import VFLoss
        loss_mod = VFLoss(types.SimpleNamespace(norm_smaller_than_one_start=11000, depth_loss_clamp=0.5,
                                                directional_derivatives_start=100),
                          types.SimpleNamespace(rgb=2.0, depth=0.5, unit_norm=0.1, supervision=1.0,
                                                norm_smaller_than_one=0.1, directional_derivatives=0.0), sync=False)

        def vf_loss(out, rgb_t, dep_t):
            return loss_mod({"rgb": out.coarse_rgb_values, "depth": out.coarse_depth_map,
                             "normals": out.coarse_normals.reshape(-1, 3), "supervised_normals": None,
                             "directional_derivatives": None}, {"rgb": rgb_t, "depth": dep_t}, 0)[0]

        def time_train(prec, n_tr, arena=False):
            tm = U.make_model(dict(CASE, perturb=True, dir_to_normal_th=-2.0), st, dev, precision=prec)
            if arena:
                from vfnerf_b200 import optim as voptim
                voptim.use_arena_optimizer(tm, max_norm=0.5)

            def train_step():
                out = tm.render(poset, uvt, Kt, 0, draws=draws_t)
                loss = vf_loss(out, rgb_gt, dep_gt)
                tm.optimizer.zero_grad()
                loss.backward()
                if world > 1:
                    from vfnerf_b200 import dist as vd
                    vd.allreduce_gradients(tm)
                if not arena:
                    torch.nn.utils.clip_grad_norm_(tm.parameters(), 0.5)
                tm.optimizer.step()
            for _ in range(3):
                train_step()
            sync_all()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for _ in range(n_tr):
                train_step()
            t1e.record()
            sync_all()
            mst = torch.tensor([t0e.elapsed_time(t1e)], device=dev)
            if world > 1:
                dist.all_reduce(mst, op=dist.ReduceOp.MAX)
            del tm
            torch.cuda.empty_cache()
            return mst.item() / n_tr
        ms_tr = time_train(args.precision, 10 if args.precision == "bf16" else 5)
        train = {"value": world * Rt / (ms_tr * 1e-3), "unit": "rays/s", "rays_per_step_per_gpu": Rt,
                 "ms_per_step": ms_tr, "precision": args.precision,
                 "includes": "eager: render fwd + fused VFLoss + backward + (allreduce) + clip_grad_norm_ + Adam"}
        if args.precision != "fp32":
            train["fp32_ms_per_step"] = time_train("fp32", 3)
        ms_ar = time_train(args.precision, 10, arena=True)
        train["arena_adam_eager"] = {"ms_per_step": ms_ar, "value": world * Rt / (ms_ar * 1e-3), "unit": "rays/s",
                                     "includes": "eager: render + fused VFLoss + backward + (allreduce of the flat gradient "
                                                 "arenas) + ArenaAdam (clip + Adam)"}

        # the same sequence captured once as a CUDA graph and replayed (vfnerf_b200/graphed.py), and -- SURVEY.md §8(d)
        # training protocol (i) -- the kernels alone: fwd + loss gradient + bwd into the flat gradient buffers,
        # graph replay, median of 50.  Single-GPU legs (the allreduce of a multi-GPU step is not captured).
        if world == 1:
            from vfnerf_b200 import graphed

            def loss_fn(out, rgb_gt, depth_gt):
                return vf_loss(out, rgb_gt, depth_gt)

            def time_graphed(n_rays, full, reps=50, arena=False):
                tm = U.make_model(dict(CASE, perturb=True, dir_to_normal_th=-2.0), st, dev, precision=args.precision)
                if arena:
                    from vfnerf_b200 import optim as voptim
                    voptim.use_arena_optimizer(tm)
                else:
                    graphed.make_capturable(tm)
                uvg, poseg, Kg = uv_d[:n_rays], pose_d[:n_rays], K_d[:n_rays]
                tg = dict(rgb_gt=torch.rand(n_rays, 3, device=dev), depth_gt=torch.rand(n_rays, 1, device=dev) * CASE["far"])
                step = graphed.GraphedTrainStep(tm, loss_fn, n_rays, tg, clip_norm=0.5 if full else None, optimizer_step=full)
                for _ in range(3):
                    step(poseg, uvg, Kg, **tg)
                torch.cuda.synchronize()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
                for a, b in ev:
                    a.record()
                    if full:
                        step(poseg, uvg, Kg, **tg)
                    else:
                        step.graph.replay()
                    b.record()
                torch.cuda.synchronize()
                ts = sorted(a.elapsed_time(b) for a, b in ev)
                del step, tm
                torch.cuda.empty_cache()
                return ts[len(ts) // 2]
            A_train = 3 * (N_COARSE + N_FINE) * (F_VF + F_RN)
            pk_t = peaks()[0]
            ms_g = time_graphed(Rt, True)
            train["graphed"] = {"ms_per_step": ms_g, "value": Rt / (ms_g * 1e-3), "unit": "rays/s",
                                "includes": "CUDA-graph replay of render + loss + backward + clip + Adam, inputs copied per step"}
            ms_a = time_graphed(Rt, True, arena=True)
            train["graphed_arena_adam"] = {"ms_per_step": ms_a, "value": Rt / (ms_a * 1e-3), "unit": "rays/s",
                                           "includes": "CUDA-graph replay of render + fused VFLoss + backward + ArenaAdam "
                                                       "(clip + Adam on the flat arenas, 2 launches per network)"}
            ko = {}
            for n_r in (Rt, 8192):
                ms_k = time_graphed(n_r, False)
                tf = A_train * n_r / (ms_k * 1e-3) / 1e12
                ko[str(n_r)] = {"ms_per_step": ms_k, "rays_per_s": n_r / (ms_k * 1e-3), "algorithmic_tflops": tf,
                                "frac_of_sustained_bf16_peak": tf / pk_t.get("bf16_tflops_sustained", pk_t["bf16_tflops"])}
            train["kernels_only"] = dict(ko, note="fwd + loss gradient + bwd into the flat gradient buffers, CUDA-graph replay, "
                                         "median of 50; algorithmic FLOP = 3 * 203.88 MFLOP/ray (SURVEY.md 8d)")

    # ---- VF-only grid query (BASELINE config 5 shape: a marching-cubes grid, coordinates generated in-kernel); each rank
    # takes a z-slab of the 256^3 grid.  `value` with the result left on the device, `e2e` through the
    # reference-shaped get_set_predictions (CPU samples in, CPU vectors out, mc_utils.py:88-104).
    gridq = None
    if not args.no_train:
        from vfnerf_b200 import grid_query as gq
        res = 256
        n_all = res ** 3
        lo, hi = n_all * rank // world, n_all * (rank + 1) // world
        with torch.no_grad():
            gq.grid_query(model.vector_field_network, res, i0=lo, n_points=hi - lo, chunk=1 << 23)
            sync_all()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                gq.grid_query(model.vector_field_network, res, i0=lo, n_points=hi - lo, chunk=1 << 23)
            g1.record()
            sync_all()
        msg = torch.tensor([g0.elapsed_time(g1) / 3], device=dev)
        if world > 1:
            dist.all_reduce(msg, op=dist.ReduceOp.MAX)
        gridq = {"value": n_all / (msg.item() * 1e-3), "unit": "points/s", "resolution": res, "ms": msg.item(),
                 "algorithmic_tflops": F_VF * n_all / (msg.item() * 1e-3) / 1e12}
        if world == 1:
            samples = torch.rand(4 * 1024 * 1024, 3)
            gq.get_set_predictions(model.vector_field_network, samples[:1 << 20], 1 << 20, torch.device(dev))
            t0g = time.perf_counter()
            gq.get_set_predictions(model.vector_field_network, samples, 1 << 20, torch.device(dev))
            dtg = time.perf_counter() - t0g
            gridq["e2e"] = {"value": samples.shape[0] / dtg, "unit": "points/s", "points": samples.shape[0],
                            "note": "get_set_predictions: pageable CPU samples in, CPU vectors out (24 B/point over PCIe)"}

    # ---- HBM-bound kernels of the path (sampler, density/transmittance scan, compositor), each timed alone with CUDA
    # events through its stage entry point of the C ABI on 262 144 rays (working sets of 0.2-0.9 GB, far beyond L2):
    # achieved = algorithmic bytes / time, against the measured HBM copy bandwidth.
    hbm = None
    if not args.no_train and world == 1:
        import ctypes as C
        Rh, Nc, Nf = 262144, N_COARSE, N_FINE
        N = Nc + Nf
        sp = torch.cuda.current_stream().cuda_stream
        f = lambda *sh: torch.empty(*sh, device=dev)
        dirs_h = torch.nn.functional.normalize(torch.randn(Rh, 3, device=dev), dim=1)
        cam_h = torch.randn(Rh, 3, device=dev)
        tv = torch.linspace(0., 1., Nc).to(dev)
        U3h = torch.rand(Rh, Nf, device=dev)
        z_c, pts_c, w_c = f(Rh, Nc), f(Rh, Nc, 3), torch.rand(Rh, Nc, device=dev)
        z_m, pts_m = f(Rh, N), f(Rh, N, 3)
        nrm = torch.tanh(torch.randn(Rh, N, 3, device=dev))
        col = torch.rand(Rh, N, 3, device=dev)
        wts, rgb_o, dep_o = f(Rh, N), f(Rh, 3), f(Rh, 1)
        cfg_h = model._render_cfg(Rh, False)
        dpar = model.density.flat()
        near, far, fr = float(model.ray_sampler.near), float(model.ray_sampler.far), float(model.fine_sampler.range)
        kernels = {
            "coarse_sample_kernel": (lambda: L.vfnerf_coarse_sample(Rh, Nc, near, far, 0, tv.data_ptr(), None, dirs_h.data_ptr(),
                                                                    cam_h.data_ptr(), z_c.data_ptr(), pts_c.data_ptr(), sp),
                                     Rh * (16 * Nc + 24)),
            "fine_sample_kernel": (lambda: L.vfnerf_fine_sample(Rh, Nc, Nf, near, far, fr, 0, z_c.data_ptr(), w_c.data_ptr(), None,
                                                                U3h.data_ptr(), dirs_h.data_ptr(), cam_h.data_ptr(), z_m.data_ptr(),
                                                                pts_m.data_ptr(), sp),
                                   Rh * (8 * Nc + 4 * Nf + 16 * N + 24)),
            "density_weights_kernel": (lambda: L.vfnerf_density_weights(C.byref(cfg_h), N, dpar.data_ptr(), nrm.data_ptr(), 3,
                                                                        dirs_h.data_ptr(), z_m.data_ptr(), None, None,
                                                                        wts.data_ptr(), sp),
                                       Rh * (20 * N + 12)),
            "composite_kernel": (lambda: L.vfnerf_composite(Rh, N, wts.data_ptr(), col.data_ptr(), z_m.data_ptr(), rgb_o.data_ptr(),
                                                            dep_o.data_ptr(), sp),
                                 Rh * (20 * N + 16)),
        }
        # marching-cubes preprocessing of a 256^3 grid (count pass: 12 B/point in, 1 B/cell out)
        Ng = 256
        from vfnerf_b200 import mc_utils as vmc
        pred_g = U.S.synthetic_vector_grid(Ng, seed=1).to(dev)
        keep_g = torch.empty(Ng ** 3, dtype=torch.uint8, device=dev)
        cnt_g = torch.zeros((Ng ** 3 + 255) // 256, dtype=torch.int32, device=dev)
        kernels["mc_count_kernel"] = (lambda: L.vfnerf_mc_count(pred_g.data_ptr(), Ng, keep_g.data_ptr(), cnt_g.data_ptr(), None,
                                                                None, None, sp), Ng ** 3 * 13)
        hbm_peak = peaks()[0]["hbm_gbs"]
        hbm = {"rays": Rh, "n_coarse": Nc, "n_fine": Nf, "peak_gbs": hbm_peak, "kernels": {}}
        for name, (fn, nbytes) in kernels.items():
            for _ in range(3):
                assert fn() == 0, name
            torch.cuda.synchronize()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(20):
                fn()
            h1.record()
            torch.cuda.synchronize()
            t_h = h0.elapsed_time(h1) * 1e-3 / 20
            hbm["kernels"][name] = {"ms": t_h * 1e3, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / t_h / 1e9,
                                    "frac_of_measured_hbm_peak": nbytes / t_h / 1e9 / hbm_peak}
        del z_c, pts_c, w_c, z_m, pts_m, nrm, col, wts
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel, timed alone with CUDA events on the launching stream:
    #   bf16: the fused tcgen05 launch (VF + colour MLPs, RENDER program) on one chunk of merged points;
    #   fp32: the CUDA-core VF MLP chain (9 GEMM launches) on one chunk.
    from vfnerf_b200 import ops
    pk, pk_src = peaks()
    P = min(R, chunk) * (N_COARSE + N_FINE)
    pts = (torch.rand(P, 3, device=dev) - 0.5) * 6
    dirs = torch.nn.functional.normalize(torch.randn(P // (N_COARSE + N_FINE), 3, device=dev), dim=1)
    reps = 5
    with torch.no_grad():
        if args.precision == "bf16":
            _, _, ws_k = ops.mlp_points(model.vector_field_network, model.rendering_network, pts, dirs, N_COARSE + N_FINE)
            run = lambda: ops.mlp_points(model.vector_field_network, model.rendering_network, pts, dirs,
                                         N_COARSE + N_FINE, workspace=ws_k, repack=False)
            flop_pt, kname = F_VF + F_RN, "vfn::mlp_tc_kernel, RENDER program (VF + colour MLPs fused, one launch)"
        else:
            run = lambda: ops.vf_query(model.vector_field_network, pts)
            flop_pt, kname = F_VF, "vfn::gemm_kernel x9 (fp32 CUDA-core VF MLP chain)"
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(reps):
            run()
        r1.record()
        torch.cuda.synchronize()
    t_k = r0.elapsed_time(r1) * 1e-3 / reps
    ach = flop_pt * P / t_k / 1e12
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    traffic, traffic_src = ncu_traffic(P) if args.precision == "bf16" else (None, None)
    roofline = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": P * 36 + (P // (N_COARSE + N_FINE)) * 12, "kernel": kname, "points_per_launch": P,
                "algorithmic_flop_per_launch": flop_pt * P, "ms_per_launch": t_k * 1e3,
                "peak_source": pk_src + ", sustained figure (kernel timed in a loop)",
                "whole_path_algorithmic_frac": (A_FWD * value / world) / 1e12 / peak_tf}

    line = {
        "metric": "render_fwd_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "bf16x3": "bf16x3"}[args.precision], "data": "synthetic",
        "config": {"workload": "render() forward, Replica-shaped 1200x680 image (816000 rays) per GPU, 64+64 samples, "
                               "shipped VF(39-256x8-259)+colour(289-256x4-3) nets, deterministic sampling",
                   "rays_per_step_per_gpu": R, "chunk_rays": chunk, "n_coarse": N_COARSE, "n_fine": N_FINE,
                   "l2": "working set per step (workspace + outputs) exceeds the 126 MB L2 many times over; no explicit flush",
                   "parallelism": f"rays sharded, {world} process(es), final gather to rank 0"},
        "clocks": clk, "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if e2e:
        line["e2e"] = e2e
    if train:
        line["train_step"] = train
    if gridq:
        line["grid_query"] = gridq
    if hbm:
        line["hbm_kernels"] = hbm
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rps, med, threads = cpu_reference_rays_per_s(3, 1)
        line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": "3 timed renders of one 1024-ray chunk (median) after 1 warm-up, oracle port on host cores"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
