"""Benchmark of the supervised-gan hot path: one G+D training step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|eager_cuda] [--batch B]
                    [--precision fp32|tf32] [--config fcgan|cgan|twostage] [--pool-size P]

Default workload (BASELINE.json configs[3] at N=1 -- the metric "fcgan 512^2 G+D train steps/sec" is quoted on the fcgan
config; configs[0] is the same nets at batch 1): deconv G (n_layers_G 5, ngf 32, noise 8x8x8) + 3-scale n_layers D (ndf 32,
scale 1/2/4), instance norm, BCE, 512x512, 2 channels, batch 8 per GPU, synthetic data, random-init weights, the reference's
default history buffer (pool_size 50), n_update_D = n_update_G = 1.

Prints ONE JSON line (contract in the task statement):
  value         images/s of the whole job, step replayed from a CUDA graph, inputs resident in HBM
  e2e           same metric through the public API (set_input + optimize_parameters + get_current_errors), with the
                pinned-host -> device copy of every batch and the device -> host loss read inside the timed region
  roofline      dominant kernel: algorithmic FLOPs / gap-free per-launch time (per-layer CUDA-graph replays, CUDA events on the
                launching stream, back to back); `bandwidth` holds
                achieved GB/s against ALGORITHMIC bytes for the HBM-bound kernels; `peak` for tf32 is measured live
  cpu_baseline  the reference step on this box's host cores (N = 1 only), all host threads
  eager_cuda    (N = 1) the UNMODIFIED reference modules moved to the GPU: PyTorch eager + cuDNN (TF32 convs), same workload

`--impl reference` times the reference's own CPU implementation: the unmodified FCGANModel from baseline/_ref (a byte copy
made by oracle/vendor_ref.py; kind "reference"), else the oracle port (oracle/nets.FcganStep; kind "port").
`--impl eager_cuda` prints the eager-cuDNN leg alone.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "images/s"
WORKLOADS = {
    "fcgan": dict(
        metric="fcgan 512x512 G+D train throughput (steps/s x global batch)",
        workload=("fcgan 512x512 G+D step (BASELINE configs[3]): deconv G n_layers 5 ngf 32 noise 8x8x8 + "
                  "3-scale n_layers D ndf 32 scale 1/2/4, instance norm, BCE, 2 channels"),
        flop_per_sample_step=85.81e9),   # useful conv FLOPs per sample-step, SURVEY.md 8(d)
    "cgan": dict(
        metric="cgan unet_256 512x512 G+D train throughput (steps/s x global batch)",
        workload=("cgan 512x512 G+D step (BASELINE configs[1]): unet_256 G ngf 64 + 2-scale n_layers D ndf 64 n_layers 3/4, "
                  "instance norm, BCE + weighted L1 (2 4), label 2 ch -> image 1 ch"),
        flop_per_sample_step=600.5e9),
    "twostage": dict(
        metric="twostage_cycle DSGAN 512x512 G+D train throughput (steps/s x global batch)",
        workload=("twostage_cycle 512x512 step (BASELINE configs[2], README.md:18): fcgan G1 + CRN G2 ngf2 64 bilinear + "
                  "unet_128 F2 nff2 32 + 2-scale D1 + 4-scale D2 ndf2 64, instance norm, BCE"),
        flop_per_sample_step=986.5e9),
}
L2_NOTE = "inputs larger than L2: the step streams > 1 GB of activations per replay (126 MB L2), no flush"


def fcgan_opt(batch, gpu, pool_size):
    return argparse.Namespace(
        isTrain=True, gpu_ids=[gpu], checkpoints_dir="/tmp/sgk_ckpt", name="bench", pretrained_model_dir="",
        which_channel="rg", batchSize=batch, output_nc=2, input_nc=2, fineSize=512, noise_nc=8, noiseSize=8, ngf=32,
        which_model_netG="fcgan", norm="instance", no_dropout=True, n_layers_G=5, use_residual=False,
        add_gaussian_noise=False, gaussian_sigma=0.1, upsample_mode="convt", n_layers_CRN_block=1,
        no_share_label_block_weights=False, no_lsgan=True, scale_factor=[1, 2, 4], lambda_D=[0.5, 0.4, 0.1],
        n_layers_D=[3, 3, 3], ndf=32, which_model_netD="n_layers", continue_train=False, which_epoch="latest",
        pool_size=pool_size, lr=2e-4, beta1=0.5, which_direction="A", n_update_D=1, n_update_G=1, no_logD_trick=False,
        niter_decay=100)


def cond_opt(config, batch, gpu, pool_size):
    """Options of the conditional / two-stage drivers at BASELINE configs[1] / configs[2] (README.md:38 / :18 recipes without
    the random in-network noise)."""
    d = dict(isTrain=True, gpu_ids=[gpu], checkpoints_dir="/tmp/sgk_ckpt", name="bench", pretrained_model_dir="",
             batchSize=batch, norm="instance", add_gaussian_noise=False, gaussian_sigma=0.1, continue_train=False,
             which_epoch="latest", pool_size=pool_size, lr=2e-4, beta1=0.5, no_logD_trick=False, niter_decay=100,
             no_cgan=False, dataset_mode="single", which_direction="AtoB", lambda_A=10.0, which_channel="rg_b", fineSize=512,
             input_nc=2, output_nc=1)
    if config == "cgan":
        d.update(noise_nc=8, noiseSize=4, ngf=64, ndf=64, which_model_netG="unet_256", which_model_netD="n_layers",
                 no_dropout=True, n_layers_G=5, use_residual=False, upsample_mode="convt", n_layers_CRN_block=1,
                 no_share_label_block_weights=False, n_layers_G_skip=-1, no_lsgan=True, scale_factor=[1, 1],
                 n_layers_D=[3, 4], lambda_D=[0.5, 0.5], weights=[2.0, 4.0], n_update_D=1, n_update_G=1,
                 transform_1to2="none")
    else:
        d.update(noise_nc1=8, noiseSize1=4, noise_nc2=8, noiseSize2=8, ngf1=32, ngf2=64, nff2=32, ndf1=32, ndf2=64,
                 which_model_netG1="fcgan", which_model_netG2="crn", which_model_netF2="unet_128",
                 which_model_netD1="n_layers", which_model_netD2="n_layers", which_model_netD="n_layers", n_layers_G1=5,
                 n_layers_G2=5, n_layers_F2=5, no_dropout1=True, no_dropout2=True, use_residual2=False,
                 upsample_mode1="convt", upsample_mode2="bilinear", n_layers_CRN_block1=1, n_layers_CRN_block2=2,
                 no_share_label_block_weights1=False, no_share_label_block_weights2=False, transform_1to2="bilinear_2",
                 scale_factor1=[1, 2], lambda_D1=[0.5, 0.4], n_layers_D1=[3, 3], scale_factor2=[1, 1, 2, 2],
                 lambda_D2=[0.3, 0.3, 0.2, 0.2], n_layers_D2=[3, 4, 3, 4], no_lsgan1=True, no_lsgan2=True,
                 use_multi_class_GAN=False, use_fixed_noise1=False, sequential_train=False, lr1=2e-4, lr2=2e-4,
                 n_update_D1=1, n_update_D2=1, n_update_G=1, detach_G1_from_G2_x=False, detach_G1_from_G2_y=False,
                 GAN_losses_D2=["real_fake"], GAN_losses_G2=["real_fake"], lambda_B=10.0, lambda_A_cycle=5.0,
                 lambda_fake_cycle=1.0, weights=None)
    return argparse.Namespace(**d)


def line_config(args, world):
    """Identical for every arm of one workload, so that the driver can pair the lines."""
    return {"workload": WORKLOADS[args.config]["workload"] + ", pool_size %d" % args.pool_size,
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "parallelism": "dp%d" % world, "l2": L2_NOTE}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING a timed region.  NVML (nvidia_ml_py) is polled every 5 ms from a
    thread, on rank 0 only -- the timed region of the default run is only ~80 ms, shorter than `nvidia-smi -lms` needs to start -- with the
    nvidia-smi loop as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, gpu, enabled=True, period=0.005):
        self.gpu, self.proc, self.lines = gpu, None, []
        self.enabled, self.period = enabled, period        # one sampling rank per job: 8 polling threads on 16 host cores cost steps
        self.nvml, self.handle, self.thread, self.stop_flag, self.samples = None, None, None, False, []

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, h

    def _poll(self):
        nv, h = self.nvml, self.handle
        while True:
            try:
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(mask)))
            except Exception:
                pass
            if self.stop_flag:
                return
            time.sleep(self.period)

    def start(self):
        if not self.enabled:
            return
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.smax = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.stop_flag, self.samples = False, []
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.enabled:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "not sampled on this rank"}
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(float(x[0]) for x in self.samples)
            reasons = set()
            for _, _, mask in self.samples:
                for bit, name in self.BITS.items():
                    if mask & bit:
                        reasons.add(name)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": max((x[1] for x in self.samples), default=None), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1]); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ reference legs
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm is a CPU job and gets every host core."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _reference_model(batch, pool_size):
    """The UNMODIFIED reference FCGANModel (baseline/_ref or /root/reference), built on the CPU; None if not available."""
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    Model = ref_loader.load_model_class("fcgan")
    m = Model()
    with contextlib.redirect_stdout(sys.stderr):      # the reference prints its networks from initialize()
        m.initialize(ref_loader.fcgan_opt(batchSize=batch, pool_size=pool_size, name="bench_ref"))
    return m


def cpu_reference_throughput(batch, steps, warmup, pool_size, seed=0):
    """images/s of the reference step on this box's host cores.  -> (ips, s/step, threads, kind)"""
    import torch
    threads = _use_all_host_threads()
    gen = torch.Generator().manual_seed(seed)
    data = torch.rand(batch, 3, 512, 512, generator=gen) * 2 - 1
    torch.manual_seed(seed)
    m = _reference_model(batch, pool_size)
    if m is not None:
        kind = "reference"

        def step():
            m.set_input({"A": data, "A_paths": ["synthetic"]})
            m.optimize_parameters()
    else:
        from oracle import nets as ON
        kind = "port"
        sdG = ON.init_fcgan_generator(gen, 8, 2, 32, 5)
        sdDs = [ON.init_nlayer_discriminator(gen, 2, 32, 3, s) for s in (1, 2, 4)]
        st = ON.FcganStep(sdG, sdDs, pool_size=pool_size)
        real = data[:, :2].contiguous()

        def step():
            st.step(real, torch.randn(batch, 8, 8, 8, generator=gen))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads, kind


def eager_cuda_throughput(batch, steps, warmup, pool_size, gpu=0, seed=0):
    """The unmodified reference modules moved to the GPU (SURVEY 8c recipe: nets built with gpu_ids=[] then .cuda()):
    PyTorch eager, cuDNN with benchmark autotuning, TF32 convolutions allowed.  -> dict"""
    import torch
    m = _reference_model(batch, pool_size)
    kind = "reference modules (.cuda())"
    dev = torch.device("cuda", gpu)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    gen = torch.Generator().manual_seed(seed)
    data = (torch.rand(batch, 3, 512, 512, generator=gen) * 2 - 1).pin_memory()
    if m is not None:
        m.netG.cuda(dev)
        for d in m.netD:
            d.cuda(dev)
        m.input, m.noise_ = m.input.cuda(dev), m.noise_.cuda(dev)
        m.Tensor = m.criterionGAN.Tensor = torch.cuda.FloatTensor     # what BaseModel.initialize sets for gpu_ids != []

        def step():
            m.set_input({"A": data, "A_paths": ["synthetic"]})
            m.optimize_parameters()
            return float(m.loss_G)
    else:
        from oracle import nets as ON
        kind = "oracle port on CUDA (reference tree absent)"
        sdG = {k: v.to(dev) for k, v in ON.init_fcgan_generator(gen, 8, 2, 32, 5).items()}
        sdDs = [{k: v.to(dev) for k, v in ON.init_nlayer_discriminator(gen, 2, 32, 3, s).items()} for s in (1, 2, 4)]
        st = ON.FcganStep(sdG, sdDs, pool_size=pool_size)

        def step():
            real = data[:, :2].to(dev, non_blocking=True)
            return st.step(real, torch.randn(batch, 8, 8, 8, device=dev))[0]
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    return {"value": batch * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "kind": kind,
            "what": "PyTorch %s eager, cuDNN benchmark, TF32 convs, H2D of the batch + loss read-back per step" % torch.__version__}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    total = args.steps + args.warmup
    # bound the CPU work to a few minutes: ~0.1 s per sample-step on 32 cores, ~0.55 s on 8
    per_sample = 0.6 * 8.0 / max(os.cpu_count() or 8, 8)
    b = max(1, min(args.batch, int(150.0 / (per_sample * max(total, 1)))))
    ips, s_per_step, threads, kind = cpu_reference_throughput(b, args.steps, args.warmup, args.pool_size)
    W = WORKLOADS["fcgan"]
    sample = ("%s FCGANModel.optimize_parameters on the host CPU, batch %d of the %d-per-GPU workload, %d warm-up + %d timed "
              "steps, %.2f s/step" % ("unmodified reference" if kind == "reference" else "oracle port of", b, args.batch,
                                      args.warmup, args.steps, s_per_step))
    line = {"impl": "reference", "metric": W["metric"], "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": line_config(args, world),
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_eager_cuda(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    r = eager_cuda_throughput(args.batch, args.steps, args.warmup, args.pool_size)
    W = WORKLOADS["fcgan"]
    line = {"impl": "eager_cuda", "metric": W["metric"], "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic", "config": line_config(args, world), "eager_cuda": r}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ peaks
def probe_tf32_peak(seconds=1.0):
    """Dense tf32 tensor-core throughput of this GPU, measured live with a library GEMM (a PEAK PROBE, not the product
    path): torch.matmul 8192^3 with TF32 allowed, best of 10 (burst) and back to back for `seconds` (sustained)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); e1.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        reps = max(10, int(seconds / (2.0 * n ** 3 / (best * 1e12))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record(); e1.synchronize()
        sustained = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
        return {"burst": best, "sustained": sustained}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------------------------ our arm
def build_model(args, local, world):
    from supervised_gan_b200 import dist as sdist
    B = args.batch
    if args.config == "fcgan":
        from supervised_gan_b200.fcgan_model import FCGANModel
        opt = fcgan_opt(B, local, args.pool_size)
        m = FCGANModel()
    elif args.config == "cgan":
        from supervised_gan_b200.cgan_model import CGANModel
        opt = cond_opt("cgan", B, local, args.pool_size)
        m = CGANModel()
    else:
        from supervised_gan_b200.twostage_cycle_model import TwoStageCycleModel
        opt = cond_opt("twostage", B, local, args.pool_size)
        m = TwoStageCycleModel()
    opt.grad_scale = 1.0 / world
    opt.cuda_graph = not args.no_graph
    opt.graph_warmup = max(3, args.warmup)
    m.initialize(opt)
    if world > 1:
        sdist.broadcast_parameters([t for n in model_nets(m) for t in list(n.parameters()) + list(n.buffers())])
        if args.config == "fcgan" and os.environ.get("SGK_OVERLAP_COMM", "1") != "0":
            # one bucket per discriminator scale (their backward passes run one after the other), two for the generator
            m.grad_sync = sdist.OverlappedGradSync(world, {"D": [list(d.model.parameters()) for d in m.netD],
                                                           "G": sdist.size_split(list(m.netG.parameters()))})
        else:
            m.grad_sync = sdist.GradSync(world)
    return m


def model_nets(m):
    nets = []
    for name in ("netG", "netG1", "netG2", "netF2"):
        if hasattr(m, name):
            nets.append(getattr(m, name))
    for name in ("netD", "netD1", "netD2"):
        nets += list(getattr(m, name, []))
    return nets


def run_ours(args):
    import torch
    import torch.distributed as dist
    import supervised_gan_b200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S.set_precision(args.precision)
    lib = S._lib.load()
    B = args.batch
    W = WORKLOADS[args.config]
    torch.manual_seed(1234 + rank)
    m = build_model(args, local, world)

    gen = torch.Generator().manual_seed(99 + rank)
    host_batches = [(torch.rand(B, 3, 512, 512, generator=gen) * 2 - 1).pin_memory() for _ in range(2)]
    m.set_input({"A": host_batches[0], "A_paths": ["synthetic"]})     # resident batch of the device-timed region
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up (eager) + capture of the whole step in the model's own CUDA graph (opt.cuda_graph, the
    # documented fast path of the public API: the model warms up on a side stream, captures once, then replays)
    for _ in range(max(3, args.warmup)):
        m.optimize_parameters()
    torch.cuda.synchronize()
    launches_per_step = None
    if m.use_graph:
        try:
            n0 = lib.sgk_launch_count()
            m.optimize_parameters()                     # captures (launches are counted at capture) and replays once
            launches_per_step = lib.sgk_launch_count() - n0
        except Exception as e:  # capture not possible: stay eager
            sys.stderr.write("[bench] CUDA graph capture failed (%s); timing the eager step\n" % (e,))
            m.use_graph, m._graph = False, None
            torch.cuda.synchronize()
    used_graph = m._graph is not None

    for _ in range(args.warmup):
        m.optimize_parameters()
    # ---------------- timed region 1: device-resident inputs
    clocks = ClockSampler(local, enabled=(rank == 0))
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.sgk_launch_count()
    ev0.record()
    for _ in range(args.steps):
        m.optimize_parameters()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    eager_launches = lib.sgk_launch_count() - n0
    clk = clocks.stop()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    gpu_launches = launches_per_step * args.steps if used_graph else eager_launches

    # ---------------- timed region 2: end to end through the public API, H2D + D2H inside
    for i in range(max(10, args.warmup)):     # untimed: the same three calls as the timed loop (copy stream, loss read-back path)
        m.set_input({"A": host_batches[i % 2], "A_paths": ["synthetic"]})
        m.optimize_parameters()
        m.get_current_errors()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = args.steps
    e0.record()
    sink = 0.0
    errs = {}
    for i in range(e2e_steps):
        m.set_input({"A": host_batches[i % 2], "A_paths": ["synthetic"]})      # pinned host -> device
        m.optimize_parameters()
        errs = m.get_current_errors()                                           # device -> host (the losses)
        sink += errs["G_GAN"] if "G_GAN" in errs else next(iter(errs.values()))
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te)
    h2d = int(getattr(m, "h2d_bytes", B * 3 * 512 * 512 * 4))   # counted by set_input from the tensors it copies
    d2h = 4 * len(errs)

    # ---------------- sustained check: the same resident-input loop for >= 2 s (clocks settle to their loaded value)
    sustained = None
    if args.sustain > 0:
        n_s = max(args.steps, int(args.sustain / max(ms / args.steps * 1e-3, 1e-6)))
        clocks2 = ClockSampler(local, enabled=(rank == 0))
        barrier()
        clocks2.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_s):
            m.optimize_parameters()
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        c2 = clocks2.stop()
        sustained = {"steps": n_s, "seconds": float(ts) * 1e-3, "ms_per_step": float(ts) / n_s,
                     "value": B * world * n_s / (float(ts) * 1e-3), "sm_mhz": c2.get("sm_mhz"), "reasons": c2.get("reasons")}

    # ---------------- replicas must be identical after all those steps (data-parallel correctness on hardware)
    flat = torch.cat([p.detach().reshape(-1).double() for n in model_nets(m) for p in n.parameters()])
    checksum = torch.stack([flat.sum(), (flat * flat).sum()])
    replicas_identical = None
    if world > 1:
        allc = [torch.empty_like(checksum) for _ in range(world)]
        dist.all_gather(allc, checksum)
        replicas_identical = all(torch.equal(allc[0], c) for c in allc[1:])
        if not replicas_identical:
            raise RuntimeError("data-parallel replicas diverged: parameter checksums %s" % [c.tolist() for c in allc])
    param_checksum = [float(checksum[0]), float(checksum[1])]

    # ---------------- roofline leg: one instrumented eager step records every kernel call; each distinct call is then
    # replayed from its own CUDA graph (no host gaps).  Every rank runs the step (it contains the gradient all-reduces).
    roof = cpu = eager = None
    timer = S.ops.KernelTimer() if rank == 0 and not args.no_roofline else None
    if not args.no_roofline:
        S.ops.set_kernel_timer(timer)
        m._optimize_parameters_eager()
        S.ops.set_kernel_timer(None)
    barrier()
    if rank == 0:
        if timer is not None:
            try:
                roof = roofline(args, S, timer, ms / args.steps, B, W)
            except Exception as e:
                roof = {"error": "%s: %s" % (type(e).__name__, e)}
        del timer
        if world == 1 and not args.no_cpu_baseline and args.config == "fcgan":
            ips, s_per_step, threads, kind = cpu_reference_throughput(B, 12, 1, args.pool_size)
            cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": "%s step on the host CPU (torch CPU), the same batch-%d workload, 1 warm-up + 12 timed steps, "
                             "%.2f s/step" % ("unmodified reference FCGANModel" if kind == "reference" else "oracle port of the reference",
                                              B, s_per_step)}
        if world == 1 and not args.no_eager_cuda and args.config == "fcgan":
            try:
                eager = eager_cuda_throughput(B, max(args.steps, 10), 3, args.pool_size, local)
            except Exception as e:   # diagnostic leg: never takes the line down
                eager = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        ips = B * world * args.steps / (ms * 1e-3)
        line = {"metric": W["metric"], "value": ips, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "steps_per_sec": args.steps / (ms * 1e-3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
                "config": line_config(args, world), "cuda_graph": used_graph,
                "clocks": clk, "gpu_launches": int(gpu_launches),
                "e2e": {"value": B * world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps,
                        "api": "%s.set_input + optimize_parameters + get_current_errors (opt.cuda_graph=%s)" % (type(m).__name__, used_graph)},
                "sustained": sustained, "replicas_identical": replicas_identical, "param_checksum": param_checksum,
                "roofline": roof, "cpu_baseline": cpu, "eager_cuda": eager, "loss_G_checksum": sink / max(e2e_steps, 1)}
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # a captured graph keeps NCCL work alive; tear down in a fixed order and skip the interpreter's own
        # (occasionally hanging) NCCL finalisers
        barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


def roofline(args, S, timer, step_ms, B, W):
    """Roofline object from the recorded calls of one step (rank 0).  Also writes gpurun_out/layer_table_<config>.json."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    bf16_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    # per-call replays FIRST: the 1 s matmul burst of the peak probe leaves the chip hot / power-limited for a while, and the
    # calls measured right after it came out up to 1.6x slow (seen on the N=16 forward convs, which are replayed first)
    meas = timer.measure(reps=10, cold=False)
    if args.precision == "tf32":
        probe = probe_tf32_peak()
        peak = probe["sustained"]
        peak_source = ("measured live: torch.matmul tf32 8192^3 back to back for 1 s (burst %.0f); MEASURED_PEAKS.json has no tf32 "
                       "entry, bf16 sustained there is %.0f" % (probe["burst"], bf16_peak))
    else:
        peak, peak_source = bf16_peak, ("MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400")
    rows, bykern, fam = [], {}, {}
    for tag, e in meas.items():
        kern = S.ops.KernelTimer.main_kernel(e["kernels"])
        us = max(e["warm_us"], 1e-3)
        rows.append({"call": tag, "kernel": kern, "calls_per_step": e["calls"], "us": us,
                     "tflops": e["flops"] / (us * 1e-6) / 1e12 if e["flops"] else None,
                     "gbs": e["bytes"] / (us * 1e-6) / 1e9 if e["bytes"] else None,
                     "alg_mbytes": e["bytes"] / 1e6, "gflop": e["flops"] / 1e9})
        k = bykern.setdefault(kern, {"calls": 0, "us": 0.0, "flops": 0.0, "bytes": 0.0, "max_call_mb": 0.0})
        k["calls"] += e["calls"]; k["us"] += us * e["calls"]
        k["flops"] += e["flops"] * e["calls"]; k["bytes"] += e["bytes"] * e["calls"]
        k["max_call_mb"] = max(k["max_call_mb"], e["bytes"] / 1e6)
        if e["flops"]:
            f = fam.setdefault(tag.split(" ")[0], {"us": 0.0, "flops": 0.0, "calls": 0})
            f["us"] += us * e["calls"]; f["flops"] += e["flops"] * e["calls"]; f["calls"] += e["calls"]
    rows.sort(key=lambda r: -r["us"] * r["calls_per_step"])
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump({"step_ms": step_ms, "precision": args.precision, "batch": B, "peak_tflops": peak, "hbm_gbs": hbm, "rows": rows},
                  open(os.path.join(ROOT, "gpurun_out", "layer_table_%s.json" % args.config), "w"), indent=1)
    except OSError:
        pass
    conv = {k: v for k, v in bykern.items() if v["flops"] > 0}
    top_tag, top = max(conv.items(), key=lambda kv: kv[1]["us"])
    ach = top["flops"] / (top["us"] * 1e-6) / 1e12
    heavy = max((r for r in rows if r["kernel"] == top_tag), key=lambda r: r["us"])
    traffic = traffic_src = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        ent = tj["launches"].get(heavy["call"])
        if ent:
            traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
            traffic_src = "profiles/r2_ncu_traffic.json, launch '%s'" % heavy["call"]
    except Exception:
        pass
    bw = {}
    for kname, v in bykern.items():
        if v["flops"] == 0 and v["bytes"] > 0:
            gbs = v["bytes"] / (v["us"] * 1e-6) / 1e9
            bw[kname] = {"gbs_vs_algorithmic_bytes": gbs, "frac_of_hbm": gbs / hbm, "us_per_step": v["us"],
                         "calls_per_step": v["calls"], "alg_mbytes_per_step": v["bytes"] / 1e6, "largest_call_mbytes": v["max_call_mb"]}
    for r in rows:   # the image-layer kernels are HBM-bound although they carry FLOPs
        if r["kernel"] in ("conv_window_persist_kernel", "edge_wgrad_tma_kernel", "gather_thin_transposed_tile") and r["gbs"]:
            e = bw.setdefault(r["kernel"], {"gbs_vs_algorithmic_bytes": 0.0, "us_per_step": 0.0, "calls_per_step": 0,
                                            "alg_mbytes_per_step": 0.0, "largest_call_mbytes": 0.0})
            e["us_per_step"] += r["us"] * r["calls_per_step"]; e["calls_per_step"] += r["calls_per_step"]
            e["alg_mbytes_per_step"] += r["alg_mbytes"] * r["calls_per_step"]
            e["largest_call_mbytes"] = max(e["largest_call_mbytes"], r["alg_mbytes"])
            e["gbs_vs_algorithmic_bytes"] = e["alg_mbytes_per_step"] * 1e6 / (e["us_per_step"] * 1e-6) / 1e9
            e["frac_of_hbm"] = e["gbs_vs_algorithmic_bytes"] / hbm
    sum_us = sum(v["us"] for v in bykern.values()) * 1e-3
    fmt = lambda v: {"tflops": v["flops"] / (v["us"] * 1e-6) / 1e12, "frac_of_peak": v["flops"] / (v["us"] * 1e-6) / 1e12 / peak,
                     "ms_per_step": v["us"] * 1e-3, "calls_per_step": v["calls"]}
    return {"bound": "tensor", "kernel": top_tag, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_source,
            "frac_of_bf16_peak": ach / bf16_peak,
            "timing": ("every distinct kernel call of one step replayed 10x back to back from its own CUDA graph (no host gaps), CUDA "
                       "events on the launching stream, best of 3; a call whose tensors are smaller than the 126 MB L2 finds them "
                       "L2-resident, as it does inside the real step right after its producer (`largest_call_mbytes` says which "
                       "bandwidth figures can exceed the HBM peak for that reason); the sum over all calls is "
                       "`instrumented_ms_per_step.kernels` and must stay below `graph_step`"),
            "share_of_step": top["us"] * 1e-3 / step_ms,
            "heaviest_launch": {"call": heavy["call"], "us": heavy["us"], "tflops": heavy["tflops"], "frac": (heavy["tflops"] or 0.0) / peak},
            "by_kernel": {k: fmt(v) for k, v in sorted(conv.items(), key=lambda kv: -kv[1]["us"])},
            "conv_families": {k: fmt(v) for k, v in fam.items()},
            "instrumented_ms_per_step": {"kernels": sum_us, "graph_step": step_ms},
            "step_tensor_frac": (B / (step_ms * 1e-3)) * W["flop_per_sample_step"] / 1e12 / peak,
            "bandwidth": {"hbm_peak_gbs": hbm, "kernels": dict(sorted(bw.items(), key=lambda kv: -kv[1]["us_per_step"]))},
            "top_calls": [{k: r[k] for k in ("call", "kernel", "calls_per_step", "us", "tflops", "gbs")} for r in rows[:14]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager_cuda"])
    ap.add_argument("--config", default="fcgan", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default 8 for fcgan, 1 for cgan / twostage)")
    ap.add_argument("--pool-size", type=int, default=50, help="history buffer (reference default 50)")
    ap.add_argument("--precision", default=os.environ.get("SGK_PRECISION", "tf32"))
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the extra sustained-clock loop (0 = skip)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-cuda", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-kernel replay leg (ncu launch-list runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.batch is None:
        args.batch = 8 if args.config == "fcgan" else 1
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager_cuda":
        run_eager_cuda(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
