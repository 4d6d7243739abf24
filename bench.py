#!/usr/bin/env python
"""bench.py — keyframe-BA-updates/s (+ rendered Mrays/s) on synthetic GO-SLAM workloads.

Workload (BASELINE.json configs[1]): Replica room0 RGB-D shapes — 8-keyframe local window,
1/8-resolution 40x80, 36 edges (|i-j| <= 3).  One STEP = one keyframe-BA-update:
    all-pairs correlation build + 4-level pyramid for the window's 36 edges (tcgen05)
  + reprojection of every edge
  + fused 4-level radius-3 lookup
  + dense bundle adjustment, 3 Gauss-Newton iterations (RGB-D prior on)
`value` is measured with inputs resident in HBM; `e2e` runs the same step through the public
Python API from pinned HOST buffers (H2D of the step's inputs and D2H of the updated state
inside the timed region).  The renderer (2^18-ray batches x 72 samples through the fused
marcher) is reported in the same line under "render".

    python bench.py --gpus N --steps K --warmup W [--impl reference]
N > 1: one rank per GPU under torchrun; every rank owns an independent keyframe window and
ray batch (weak scaling, no data-path collective — SURVEY §8e), time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

NUM_KF, HT, WD, BA_ITERS = 8, 40, 80, 3
RAYS, SAMPLES = 1 << 18, 72
METRIC = "keyframe-BA-updates/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML (nvidia_ml_py) in a thread at
    ~1 kHz (the timed region of the default run is ~10 ms), falling back to nvidia-smi polling."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop, self.max_mhz = index, [], False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if self.nvml is not None:
            nv = self.nvml
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self.stop:
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = int(get_reasons(self.h))
                    self.samples.append([mhz, self.max_mhz] + [("Active" if (r & m) else "Not Active") for m in bits.values()])
                except Exception:
                    pass
                time.sleep(0.001)
            return
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(str(s[2 + i]).lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------ workload
def make_window(seed):
    from goslam_b200 import synthetic
    sc, g = synthetic.make_scene(num_kf=NUM_KF, ht=HT, wd=WD, seed=seed, rgbd=True)
    coords = synthetic.true_reprojection(sc)          # plain-torch setup helper (targets = reprojection + noise)
    targets, weights, eta = synthetic.make_update(sc, coords[0], g, noise=0.5)
    sc.update(targets=targets, weights=weights, eta=eta)
    return sc


class Window:
    """device-resident state of one 8-keyframe window + the public-API step."""

    def __init__(self, sc, dev):
        self.dev = dev
        self.host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in sc.items()}
        self.d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
        self.poses0 = self.d["poses"].clone()
        self.disps0 = self.d["disps"].clone()
        # the factor graph's correlation slot pool, allocated once (FactorGraph.max_factors slots)
        from goslam_b200.modules.corr import CorrPool
        self.pool = CorrPool(int(sc["ii"].numel()), HT, WD, device=dev, layout="tiled")
        self.corr = None

    def build(self, km):
        from goslam_b200.modules import CorrBlock
        if self.corr is not None:
            self.corr.free()                    # rm_factors: the slots go back to the pool
        self.corr = CorrBlock.from_video(km, self.d["ii"], self.d["jj"], HT, WD, pool=self.pool)
        return self.corr

    def step(self, reset=True):
        from goslam_b200 import droid_backends
        from goslam_b200.modules import CorrBlock
        from goslam_b200.modules.corr import fmaps_to_kmajor
        d = self.d
        if reset:                               # device-resident loop: restart from the same state every step
            d["poses"].copy_(self.poses0)       # (the end-to-end loop gets fresh inputs from the host instead)
            d["disps"].copy_(self.disps0)
        ii, jj = d["ii"], d["jj"]
        # FactorGraph.add_factors' volume, video-level: K-major re-layout of the window's feature
        # maps (per keyframe, redone every step here) + on-device edge -> frame indexing
        km = fmaps_to_kmajor(d["fmaps"][:NUM_KF])
        corr = self.build(km)
        coords, _ = droid_backends.reproject(d["poses"], d["disps"], d["intrinsics"], ii, jj, want_valid=False)
        feat = corr(coords)
        droid_backends.ba(d["poses"], d["disps"], d["intrinsics"][0], d["disps_sens"], d["targets"],
                          d["weights"], d["eta"], ii, jj, 1, NUM_KF, BA_ITERS, 1e-4, 0.1, False)
        d["disps"].clamp_(min=0.001)            # src/depth_video.py:269
        return feat

    E2E_KEYS = ("fmaps", "poses", "disps", "disps_sens", "intrinsics", "targets", "weights", "eta", "ii", "jj")

    def step_e2e(self, out_pinned):
        """one update through the public API from pinned HOST buffers: H2D of every input of the
        step, the step, D2H of the updated state.  Copies run on a side stream into the other half
        of a double buffer, so the transfer of update i+1 overlaps the kernels of update i (each
        step still waits for ITS inputs and its result is read back)."""
        if not hasattr(self, "_e2e"):
            # all inputs of a step live in ONE pinned host block and one device block per buffer half:
            # one H2D copy per step instead of ten (the end-to-end loop is otherwise host-bound)
            offs, total = {}, 0
            for k in self.E2E_KEYS:
                offs[k] = total
                total += (self.host[k].numel() * self.host[k].element_size() + 255) // 256 * 256
            packed = torch.empty(total, dtype=torch.uint8).pin_memory()

            def views(block):
                return {k: block[offs[k]:offs[k] + self.host[k].numel() * self.host[k].element_size()]
                        .view(self.host[k].dtype).view(self.host[k].shape) for k in self.E2E_KEYS}
            hv = views(packed)
            for k in self.E2E_KEYS:
                hv[k].copy_(self.host[k])
            self._e2e = dict(copy=torch.cuda.Stream(self.dev), bufs=[None, None], blocks=[None, None], packed=packed,
                             ready=[None, None], done=[None, None], parity=0)
            for b in range(2):
                self._e2e["blocks"][b] = torch.empty(total, dtype=torch.uint8, device=self.dev)
                self._e2e["bufs"][b] = views(self._e2e["blocks"][b])
        st = self._e2e
        cur = st["parity"]
        main = torch.cuda.current_stream(self.dev)
        if st["ready"][cur] is None:                      # first call: nothing prefetched yet
            self._prefetch(cur)
        main.wait_event(st["ready"][cur])
        for k in self.E2E_KEYS:
            self.d[k] = st["bufs"][cur][k]
        self._prefetch(cur ^ 1)                           # next update's inputs, overlapped
        self.step(reset=False)
        out_pinned[0].copy_(self.d["poses"], non_blocking=True)
        out_pinned[1].copy_(self.d["disps"], non_blocking=True)
        st["done"][cur] = torch.cuda.Event()
        st["done"][cur].record(main)
        st["parity"] = cur ^ 1

    def _prefetch(self, b):
        st = self._e2e
        with torch.cuda.stream(st["copy"]):
            if st["done"][b] is not None:                 # buffer b must not be in use by an older step
                st["copy"].wait_event(st["done"][b])
            st["blocks"][b].copy_(st["packed"], non_blocking=True)
            st["ready"][b] = torch.cuda.Event()
            st["ready"][b].record(st["copy"])

    def h2d_bytes(self):
        return sum(self.host[k].numel() * self.host[k].element_size() for k in
                   ("fmaps", "poses", "disps", "disps_sens", "intrinsics", "targets", "weights", "eta", "ii", "jj"))

    def d2h_bytes(self):
        return self.d["poses"].numel() * 4 + self.d["disps"].numel() * 4


def make_renderer(dev, seed):
    from goslam_b200 import neus, synthetic
    offs, ress, _, total = neus.hashgrid_layout()
    w = synthetic.make_neus_weights(seed=seed, total_grid_params=total, layout=(offs, ress))
    net = neus.InstantNeuS(synthetic.NEUS_CFG, [[-2.0, 2.0]] * 3)
    with torch.no_grad():
        net.sdf_network.encoding.encoding.params.copy_(w["grid"])
        net.sdf_network.sdf_layer.weight.copy_(w["sdf_w"])
        net.sdf_network.sdf_layer.bias.copy_(w["sdf_b"])
        net.color_network._B.copy_(w["color_B"])
        net.color_network.network.params.copy_(w["mlp"])
    net = net.to(dev)
    rays = synthetic.make_rays(RAYS, S=SAMPLES, seed=seed)
    return net, rays, w


def time_gpu(fn, steps, warmup, dist_barrier):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist_barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist_barrier()
    return e0.elapsed_time(e1) / steps          # ms per step


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(sc, n_corr_edges=4, ba_iters=1):
    """Reference algorithm on the host cores (oracle port): correlation build + pyramid + lookup
    on `n_corr_edges` of the 36 edges and `ba_iters` BA iteration(s) on the full graph; returns
    the time extrapolated to the full step (36 edges, 3 iterations)."""
    from oracle import ba_oracle, corr_oracle, geom_oracle
    ii, jj = sc["ii"], sc["jj"]
    N = ii.numel()
    t0 = time.perf_counter()
    pyr = corr_oracle.corr_build(sc["fmaps"][ii[:n_corr_edges], 0].float(), sc["fmaps"][jj[:n_corr_edges], 0].float(), 4)
    coords, _ = geom_oracle.reproject(sc["poses"].numpy(), sc["disps"].numpy(), sc["intrinsics"].numpy(),
                                      ii[:n_corr_edges].numpy(), jj[:n_corr_edges].numpy())
    corr_oracle.corr_pyramid_lookup([p.numpy() for p in pyr], coords[0], 3)
    t1 = time.perf_counter()
    ba_oracle.ba(sc["poses"].numpy(), sc["disps"].numpy(), sc["intrinsics"][0].numpy(), sc["disps_sens"].numpy(),
                 sc["targets"].numpy(), sc["weights"].numpy(), sc["eta"].numpy(), ii.numpy(), jj.numpy(),
                 1, NUM_KF, ba_iters, 1e-4, 0.1, False)
    t2 = time.perf_counter()
    full = (t1 - t0) * N / n_corr_edges + (t2 - t1) * BA_ITERS / ba_iters
    return full, (t2 - t0)


def cpu_sample_plan(sc, budget_s):
    """(edges, BA iterations) of the CPU sample: the whole step when it fits `budget_s` seconds on this
    host, otherwise as many of the 36 edges (+ 1 of the 3 iterations) as fit; probed warm."""
    cpu_reference_step(sc, 2, 1)                     # thread pools, allocator
    est_full, _ = cpu_reference_step(sc, 2, 1)
    if est_full <= budget_s:
        return 36, BA_ITERS
    return max(2, int(36 * budget_s / est_full)), 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sc = make_window(43)
    n_e, n_it = cpu_sample_plan(sc, 150.0 / max(1, args.steps + args.warmup))
    for _ in range(max(args.warmup, 0)):
        cpu_reference_step(sc, n_e, n_it)
    ts = []
    for _ in range(args.steps):
        full, _ = cpu_reference_step(sc, n_e, n_it)
        ts.append(full)
    ms = 1e3 * float(np.mean(ts))
    val = 1e3 / ms                # the host's cores are the same whatever N is: its N windows run one after another
    sample = ("per step: corr build+pyramid+lookup on %d/36 edges and %d/3 BA iterations (numpy/torch CPU port of the "
              "reference kernels)%s" % (n_e, n_it, "" if n_e == 36 else ", extrapolated linearly to the full step"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(world):
    return {"workload": "configs[1]: Replica room0 RGB-D shapes, 8-keyframe window, 40x80 @1/8, 36 edges, "
                        "corr build + 4-level r=3 lookup + 3 BA iters per update",
            "keyframes": NUM_KF, "grid": [HT, WD], "edges": 36, "ba_iters": BA_ITERS,
            "windows_per_gpu": 1, "parallelism": "window-per-gpu x%d (no collective)" % world,
            "l2": "each step writes a 0.98 GB correlation pyramid (> 126 MB L2) before it is read back, no explicit flush needed",
            "corr_layout": "tiled slot pool (CorrPool)",
            "render": {"rays": RAYS, "samples_per_ray": SAMPLES}}


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        barrier = lambda: dist.barrier()   # noqa: E731
    else:
        barrier = lambda: None             # noqa: E731

    from goslam_b200 import _lib
    _lib.load()
    pk = peaks()
    sc = make_window(43 + rank)
    win = Window(sc, dev)
    warm = max(args.warmup, 3)

    with ClockSampler(local) as clk:
        ms_step = time_gpu(win.step, args.steps, warm, barrier)
    ms_step = max_over_ranks(ms_step, world)
    clocks = clk.summary()

    # ---- end to end through the public API from pinned host buffers
    outp = (torch.empty_like(sc["poses"]).pin_memory(), torch.empty_like(sc["disps"]).pin_memory())
    ms_e2e = max_over_ranks(time_gpu(lambda: win.step_e2e(outp), args.steps, warm, barrier), world)

    # ---- dominant kernel: correlation build (tcgen05) timed alone on this stream
    from goslam_b200.modules import CorrBlock
    d = win.d
    from goslam_b200.modules.corr import fmaps_to_kmajor
    km = fmaps_to_kmajor(d["fmaps"][:NUM_KF])
    ms_build = time_gpu(lambda: win.build(km), max(args.steps, 10), warm, lambda: None)
    N, hw = 36, HT * WD
    lvl = sum((HT >> i) * (WD >> i) for i in range(4))
    build_bytes = N * (2 * 128 * hw * 2 + hw * lvl * 2)
    build_flops = N * 2.0 * 128 * hw * hw
    ach = build_bytes / (ms_build * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):      # dram__bytes_read+write of one `ncu --set full` capture of this kernel/workload
        traffic = json.load(open(tp)).get("corr_build_tc_kernel", {}).get("dram_bytes_per_launch")
    roof = {"kernel": "corr_build_tc_kernel", "bound": "hbm", "achieved": ach,
            "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": traffic,
            "algorithmic_bytes": build_bytes,
            "peak_source": pk["src"] + " (burst copy bandwidth)", "ms_per_launch": ms_build,
            "share_of_step": ms_build / ms_step,
            "tensor_tflops": build_flops / (ms_build * 1e-3) / 1e12,
            "tensor_frac_of_measured_burst": build_flops / (ms_build * 1e-3) / 1e12 / pk["tf_burst"]}

    line = {"metric": METRIC, "value": world * 1e3 / ms_step, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 corr (f32 accumulate) / f32 BA (f64 solve)",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": world * 1e3 / ms_e2e, "unit": "updates/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": win.h2d_bytes(), "d2h_bytes_per_step": win.d2h_bytes()},
            "gpu_launches": 6 * args.steps,   # kmajor, corr build, reproject, lookup, ba_prep, ba (all iterations in one cooperative kernel)
            "roofline": roof}

    if not args.no_render:
        net, rays, _ = make_renderer(dev, 43 + rank)
        rd = [r.to(dev) for r in rays]
        ms_r = max_over_ranks(time_gpu(lambda: net(*rd), max(5, args.steps // 2), 3, barrier), world)
        # end to end = the reference's public call, Renderer.render_batch_ray(rays_o, rays_d, net,
        # gt_depth): rays and sensor depth come from pinned host memory, z-sampling (one launch)
        # and the marcher run on the device, colour + depth go back to the host.
        from goslam_b200 import render as render_mod
        import types
        rcfg = {"rendering": {"lindisp": False, "perturb": 1.0, "N_samples": 24, "N_surface": SAMPLES - 24}}
        renderer = render_mod.Renderer(rcfg, None, types.SimpleNamespace(H=512, W=512, fx=460.8, fy=460.8, cx=256.0, cy=256.0))
        gt_depth = 0.5 + 2.5 * torch.rand(RAYS, generator=torch.Generator().manual_seed(43 + rank))
        hp = [rays[0].pin_memory(), rays[1].pin_memory(), gt_depth.pin_memory()]
        keep = {"color": torch.empty(RAYS, 3).pin_memory(), "depth": torch.empty(RAYS, 1).pin_memory()}

        def render_e2e():
            ro, rdir, gd = [x.to(dev, non_blocking=True) for x in hp]
            out = renderer.render_batch_ray(ro, rdir, net, None, device=dev, gt_depth=gd)
            for k in ("color", "depth"):
                keep[k].copy_(out[k].reshape(keep[k].shape), non_blocking=True)
        ms_re = max_over_ranks(time_gpu(render_e2e, max(5, args.steps // 2), 3, barrier), world)
        rbytes = RAYS * (512.0 * SAMPLES + 1216.0)
        line["render"] = {"metric": "rendered Mrays/s", "value": world * RAYS / ms_r / 1e3, "unit": "Mrays/s",
                          "ms_per_batch": ms_r,
                          "e2e": {"value": world * RAYS / ms_re / 1e3, "unit": "Mrays/s", "ms_per_batch": ms_re,
                                  "call": "Renderer.render_batch_ray (z-sampling + marcher), host rays in, colour+depth out",
                                  "h2d_bytes_per_batch": RAYS * 7 * 4, "d2h_bytes_per_batch": RAYS * 4 * 4},
                          "roofline": {"kernel": "neus_forward_kernel", "bound": "hbm",
                                       "achieved": rbytes / (ms_r * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                       "frac": rbytes / (ms_r * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                                       "note": "algorithmic bytes (38,080 B/ray); the 25 MB table is L2-resident"}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        # the whole step (36 edges, 3 iterations, nothing extrapolated) when the host does it in <= 20 s,
        # else a proportional sample
        n_e, n_it = cpu_sample_plan(sc, 20.0)
        full, spent = cpu_reference_step(sc, n_e, n_it)
        line["cpu_baseline"] = {"value": 1.0 / full, "unit": "updates/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "corr build+pyramid+lookup on %d/36 edges + %d/3 BA iterations "
                                          "(%.1f s of CPU work)%s" % (n_e, n_it, spent, "" if n_e == 36 else
                                                                      ", extrapolated linearly to the full step")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
