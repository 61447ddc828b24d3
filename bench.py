"""
bench.py -- BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2] [--nt T]
    python bench.py --impl reference ...          # the CPU formulation of the reference

metric : source x baseline x freq x time evaluations per second, forward+backward
         (evals = sum_t Nbl * Nf * Ns_t, sources counted after the FOV cut -- BASELINE.md)
step   : one forward + backward pass of the RIME over the rank's time group
         (loss = sum |V|^2; gradients to sky, beam and -- for c3 -- antenna positions)
value  : whole-job evals/s with parameters resident in HBM (CUDA events, max over ranks)
e2e    : the same, through the public API with parameters copied from pinned host memory every
         step and loss + gradients read back to the host inside the timed region
One JSON line is printed by rank 0.  See DESIGN.md section 6 for the roofline arithmetic.
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
sys.path.insert(0, ROOT)

FLOP_FWD, FLOP_BWD_SKY, FLOP_BWD_BL = 10, 10, 12


# --------------------------------------------------------------------------- helpers
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p.get("hbm_gbs"), sm_max_mhz=p.get("sm_max_mhz"), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(np.max(smax)) if smax else None,
                    power_w_max=float(np.max(power)) if power else None,
                    samples=len(sm), reasons=sorted(reasons))


class KernelTimer:
    """CUDA events around every libb200rime launch (on the launching = current stream)."""

    def __init__(self, ops):
        self.ops, self.records, self.orig = ops, [], ops._call

    def __enter__(self):
        def timed(name, sfx, *args):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self.orig(name, sfx, *args)
            e1.record()
            self.records.append((name, e0, e1))
        self.ops._call = timed
        return self

    def __exit__(self, *exc):
        self.ops._call = self.orig

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.records:
            d = out.setdefault(name, dict(ms=0.0, launches=0))
            d["ms"] += e0.elapsed_time(e1)
            d["launches"] += 1
        return out


def build_workload(name, nt, device, rank, world):
    import workloads
    if name == "c3":
        rime = workloads.pixel_interp(128, 1024, nt * world, device, torch.float32,
                                      antpos_param=True)
        desc = ("C3: HERA-350 (61075 cross baselines) x HEALPix nside-128 PixelSky x rect 1deg "
                "interpolated PixelBeam x 1024 freqs, fwd+bwd to sky, beam, antenna positions")
        grads = "sky,beam,antpos"
    elif name == "c2":
        rime = workloads.point_airy(10000, 256, nt * world, device, torch.float32)
        desc = ("C2: HERA-37 (666 cross baselines) x 10k point sources x Airy beam x 256 freqs, "
                "fwd+bwd to sky")
        grads = "sky"
    elif name == "c1":
        rime = workloads.point_airy(1000, 64, nt * world, device, torch.float32, bls='uniq')
        desc = ("C1: HERA-37 (63 unique baselines) x 1k point sources x Airy beam x 64 freqs, "
                "fwd+bwd to sky")
        grads = "sky"
    else:
        raise ValueError(name)
    if world > 1:                      # weak scaling: rank r owns times [r*nt, (r+1)*nt)
        rime.setup_sim_times(rime.all_sim_times[rank * nt:(rank + 1) * nt])
    params = [p for p in rime.parameters() if p.requires_grad]
    return rime, params, desc, grads


def flops_per_eval(grads):
    f = FLOP_FWD + FLOP_BWD_SKY
    if "antpos" in grads:
        f += FLOP_BWD_BL
    return f


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference(workload, steps, warmup, threads=None):
    """The reference's CPU formulation (oracle/rime_oracle.py: materialised fringe tensor, complex
    exp, multiply, sum; torch autograd backward) on a bounded sample of the workload."""
    from oracle import rime_oracle as orc
    import bayeslim_b200 as ba
    import workloads
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    dt = torch.float32
    if workload == "c3":
        nbl, nf = 32, 128
        rime = workloads.pixel_interp(128, nf, 1, 'cpu', dt, n_bl=nbl, antpos_param=True)
        sample = "C3 slice: %d baselines x %d freqs x 1 time x all sources above horizon" % (nbl, nf)
    elif workload == "c1":
        nbl, nf = 63, 64
        rime = workloads.point_airy(1000, nf, 10, 'cpu', dt, bls='uniq')
        sample = "C1 at full size: 63 unique baselines x 64 freqs x 10 times x 1k sources"
    else:
        nbl, nf = 63, 256
        rime = workloads.point_airy(10000, nf, 1, 'cpu', dt, bls='uniq')
        sample = "C2 slice: 63 unique baselines x %d freqs x 1 of 60 times x 10k sources" % nf
    ra, dec = rime.sky.angs[0], rime.sky.angs[1]
    zenaz = []
    for tm in rime.sim_times:
        za = rime.telescope.eq2top(tm, ra, dec)
        zenaz.append((za[0].to(dt), za[1].to(dt)))
    freqs = rime.array.freqs.to(dt)
    sp = rime.sky.params.detach().clone().requires_grad_(True)
    bp = rime.beam.params.detach().clone().requires_grad_(workload == "c3")
    antvecs = rime.array.antvecs.detach().to(dt).clone().requires_grad_(workload == "c3")
    bls = rime.sim_bls

    def step():
        for p in (sp, bp, antvecs):
            p.grad = None
        blvecs = orc.get_blvecs(antvecs, rime.array.ants, bls)
        if workload == "c3":
            sky = sp * float(rime.sky.px_area)
            bmap = orc.pixel_response_forward(bp, powerbeam=True)
            tg, pg = rime.beam.R.theta_grid, rime.beam.R.phi_grid

            def beam_fn(z, a):
                inds, wgts = orc.rect_interp_weights(tg, pg, z, a, 'linear')
                return orc.interp_map(bmap, inds, wgts.to(dt))
        else:
            sky = orc.point_sky_response(sp, freqs, 'powerlaw', f0=150e6)
            beam_fn = lambda z, a: orc.airy_response(bp, z, a, freqs, powerbeam=True)
        V = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=180.0)
        loss = (V.real ** 2 + V.imag ** 2).sum()
        loss.backward()
        return float(loss)

    ns = sum(int((z < 90).sum()) for z, _ in zenaz)
    evals = ns * len(bls) * nf
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt_s = (time.perf_counter() - t0) / steps
    return dict(value=evals / dt_s, ms_per_step=dt_s * 1e3, cores=threads, sample=sample,
                evals_per_step=evals)


# --------------------------------------------------------------------------- main
def emit(line):
    """Print the one JSON line on the process's ORIGINAL stdout (fd saved in main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # libraries (NCCL's version banner, torch warnings) may write to fd 1; keep stdout clean for
    # the single JSON line the driver parses by routing everything else to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c1"])
    ap.add_argument("--nt", type=int, default=None, help="times per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    nt = args.nt or {"c3": 2, "c2": 60, "c1": 10}[args.workload]
    unit = "source*baseline*freq*time evals/s (fwd+bwd)"

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        r = cpu_reference(args.workload, steps, min(args.warmup, 1))
        line = dict(metric="rime_evals_per_sec_fwd_bwd", value=r["value"], unit=unit,
                    n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 1),
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload=args.workload.upper() + " (CPU sample)", sample=r["sample"]),
                    cpu_baseline=dict(value=r["value"], unit=unit, cores=r["cores"], kind="port",
                                      sample=r["sample"]),
                    e2e=dict(value=r["value"], unit=unit, h2d_bytes_per_step=0,
                             d2h_bytes_per_step=0),
                    gpu_launches=0)
        emit(line)
        return

    import torch.distributed as dist
    import bayeslim_b200 as ba
    from bayeslim_b200 import ops, parallel, _lib
    import workloads

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))

    rime, params, desc, grads = build_workload(args.workload, nt, device, rank, world)

    def step(e2e_buffers=None):
        if e2e_buffers is not None:
            for p, h in zip(params, e2e_buffers["h_in"]):
                p.data.copy_(h, non_blocking=True)
        for p in params:
            p.grad = None
        V = rime().data
        loss = (V.real ** 2 + V.imag ** 2).sum()
        loss.backward()
        if world > 1:
            parallel.allreduce_gradients(params)
        if e2e_buffers is not None:
            for p, h in zip(params, e2e_buffers["h_out"]):
                h.copy_(p.grad, non_blocking=True)
            e2e_buffers["loss"] = float(loss.detach())  # D2H + sync
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(nsteps):
            step(**kw)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / nsteps

    # warm-up (also builds the per-time geometry tables, interpolation CSR, unit tables)
    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    evals_rank = workloads.count_evals(rime)
    ev = torch.tensor([evals_rank], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    evals_total = float(ev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    with KernelTimer(ops) as kt:
        ms_step = timed(args.steps)
        ksum = kt.summary()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end: parameters from pinned host memory in, loss + gradients out, every step
    bufs = dict(h_in=[p.detach().cpu().pin_memory() for p in params],
                h_out=[torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in params])
    step(e2e_buffers=bufs)
    ms_e2e = timed(args.steps, e2e_buffers=bufs)
    h2d = sum(h.numel() * h.element_size() for h in bufs["h_in"])
    d2h = sum(h.numel() * h.element_size() for h in bufs["h_out"]) + 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fringe_sum_fwd) and of the HBM-bound builder
    peaks = load_peaks()
    info = _lib.device_info(local)
    fp32_meas, _ = _lib.microbench("fp32", 4096)
    fp32x2_meas, _ = _lib.microbench("fp32x2", 4096)
    fp64_meas, _ = _lib.microbench("fp64", 1024)
    mufu_meas, _ = _lib.microbench("mufu", 2048)
    fp32_theory = 2 * 128 * info["sm_count"] * (peaks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
    k = {n: d for n, d in ksum.items()}

    def rate(name, flop_per_eval):
        if name not in k or k[name]["ms"] <= 0:
            return None
        return evals_rank * args.steps * flop_per_eval / (k[name]["ms"] * 1e-3) / 1e12

    # DRAM traffic per launch from the committed ncu --set full capture of this workload
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload:
            traffic = {kn.split("_kernel")[0].replace("ant_fringe", "antfringe"):
                       v["dram_bytes_per_launch"] for kn, v in tj["kernels"].items()}
    # FP32-bound kernels of the step and their ALGORITHMIC flop per evaluation (SURVEY 8(d):
    # rotation recurrence 6 + multiply-accumulate 4 forward, 10 backward to the sky, 12 backward
    # to the baseline / antenna vectors).  The antenna-factorised kernels do the same job with
    # fewer executed flops (one complex multiply-accumulate = 8 flop per computed antenna pair);
    # `achieved` uses the algorithmic count so that kernels doing the same work are comparable,
    # `executed_tflops` is what the FP32 pipes actually ran.
    fp32_kernels = (("fringe_sum_fwd", FLOP_FWD), ("fringe_sum_bwd_sky", FLOP_BWD_SKY),
                    ("fringe_sum_bwd_bl", FLOP_BWD_BL), ("antfringe_fwd", FLOP_FWD),
                    ("antfringe_bwd", FLOP_BWD_SKY + FLOP_BWD_BL))
    executed = {}
    tilings = [t for t in getattr(rime, "_ant_tilings", {}).values() if t is not None]
    if tilings:
        til = tilings[0]
        nsrc_pad = sum(rec.geom.S for rec in rime._geom_cache.values())
        nfp = -(-len(rime.array.freqs) // _lib.KC["f32"]) * _lib.KC["f32"]
        executed["antfringe_fwd"] = 8.0 * til.pair_slots * nsrc_pad * nfp
        executed["antfringe_bwd"] = 8.0 * til.bwd_rows * til.nm_pad * nsrc_pad * nfp

    def entry(name, fl):
        r = rate(name, fl)
        if not r:
            return None
        d = dict(bound="fp32", kernel=name + "_f32", achieved=r, peak=fp32_meas / 1e3,
                 unit="TFLOP/s", frac=r / (fp32_meas / 1e3),
                 peak_source="b200rime_microbench FFMA chains measured in this run "
                             "(MEASURED_PEAKS.json has no FP32 figure)",
                 peak_theoretical=fp32_theory, frac_of_theoretical=r / fp32_theory,
                 flop_per_eval=fl,
                 traffic=(traffic[name] * nt if name in traffic else None),
                 traffic_unit="DRAM bytes per launch: ncu dram__bytes_read+write of a 1-time launch "
                              "(profiles/r01_traffic.json) x times per launch",
                 ms_per_launch=k[name]["ms"] / max(k[name]["launches"], 1))
        if name in executed:
            ex = executed[name] * args.steps / (k[name]["ms"] * 1e-3) / 1e12
            d.update(executed_tflops=ex, executed_frac=ex / (fp32_meas / 1e3))
        return d

    entries = {n: entry(n, fl) for n, fl in fp32_kernels}
    entries = {n: e for n, e in entries.items() if e}
    dominant = max(entries, key=lambda n: k[n]["ms"]) if entries else None
    roofline = entries.get(dominant)
    others = {n: e for n, e in entries.items() if n != dominant}
    hbm = None
    bname = "build_airy"
    if args.workload == "c3":
        bname = "build_interp_t" if "build_interp_t" in k else "build_interp"
    if bname in k and k[bname]["ms"] > 0:
        rec = list(rime._geom_cache.values())[0]
        nf = len(rime.array.freqs)
        nsrc = sum(rec.geom.ns)
        npb = rime.beam.params.shape[-1] if args.workload == "c3" else 0
        # per launch (all times of the step): read the beam map once, read sky at the cut,
        # write A, read 4 idx + 4 wgt
        bytes_step = 4 * nf * (npb + 2 * nsrc) + 32 * nsrc
        gbs = bytes_step * args.steps / (k[bname]["ms"] * 1e-3) / 1e9
        hbm = dict(bound="hbm", kernel=bname + "_f32", achieved=gbs, peak=peaks["hbm_gbs"],
                   unit="GB/s", frac=gbs / peaks["hbm_gbs"], peak_source=peaks["source"],
                   traffic=traffic.get(bname),
                   traffic_unit="DRAM bytes of a 1-time launch (ncu, profiles/r01_traffic.json)")

    line = dict(
        metric="rime_evals_per_sec_fwd_bwd", value=evals_total / (ms_step * 1e-3), unit=unit,
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=desc, times_per_gpu=nt, grads=grads,
                    evals_per_step=evals_total, flop_per_eval=flops_per_eval(grads),
                    l2="inputs larger than L2 (perceived-sky slab %.0f MB, cotangent %.0f MB)"
                       % (4.0 * len(rime.array.freqs) * workloads.count_evals(rime) /
                          max(len(rime.sim_bls) * len(rime.array.freqs), 1) / 1e6,
                          8.0 * len(rime.sim_bls) * nt * len(rime.array.freqs) / 1e6),
                    parallelism="time-sharded x%d, 1 allreduce of gradients" % world),
        clocks=clocks,
        e2e=dict(value=evals_total / (ms_e2e * 1e-3), unit=unit, ms_per_step=ms_e2e,
                 h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
        gpu_launches=launches,
        roofline=roofline, roofline_other_kernels=others, roofline_hbm=hbm,
        fp32_tflops_algorithmic=evals_total * flops_per_eval(grads) / (ms_step * 1e-3) / 1e12 / world,
        peaks=dict(fp32_tflops_measured=fp32_meas / 1e3, fp32x2_tflops_measured=fp32x2_meas / 1e3,
                   fp64_tflops_measured=fp64_meas / 1e3,
                   mufu_gops_measured=mufu_meas, fp32_tflops_theoretical=fp32_theory,
                   sm_count=info["sm_count"]),
        kernel_ms_per_step={n: d["ms"] / args.steps for n, d in k.items()},
        kernel_launches_per_step={n: d["launches"] / args.steps for n, d in k.items()},
    )
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(args.workload, 1, 1)
        line["cpu_baseline"] = dict(value=r["value"], unit=unit, cores=r["cores"], kind="port",
                                    sample=r["sample"])
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
