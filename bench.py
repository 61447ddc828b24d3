"""
bench.py -- BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1|c4|c5] [--nt T]
                    [--pass fwdbwd|fwd]
    python bench.py --impl reference ...       # the UNMODIFIED reference on the host cores
    python bench.py --impl reference-gpu ...   # the reference's torch formulation on the B200

metric : source x baseline x freq x time evaluations per second, forward+backward (or forward
         only with --pass fwd); evals = sum_t Nbl * Nf * Ns_t, sources counted after the FOV cut
step   : one forward + backward pass of the RIME over the rank's work units
         (loss = sum |V|^2; gradients to sky, beam and -- for c3 -- antenna positions)
value  : whole-job evals/s with parameters resident in HBM (CUDA events, max over ranks)
e2e    : the same, through the public API with parameters copied from pinned host memory every
         step and loss + gradients read back to the host inside the timed region
One JSON line is printed by rank 0.  See DESIGN.md section 6 for the roofline arithmetic.

Workloads (BASELINE.json configs):  c3 (default) HERA-350 x nside-128 x 1024 ch, time-sharded,
weak scaling;  c4 4-pol Jones x nside-256 x 1024 ch, time-sharded, weak scaling;  c5 nside-256
x 1024 ch, the full-night time axis in single-time minibatches x 2 block-aligned baseline
groups = the reference's minibatch grid, a FIXED job of 8 times sharded over the ranks (strong
scaling);  c2 / c1 HERA-37 point-source cases.
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

FLOP_FWD, FLOP_BWD_SKY, FLOP_BWD_BL = 10, 10, 12      # SURVEY section 8(d), FP32 formulation
FLOP_MMA = 24     # executed tensor flop per (antenna pair slot, source, channel): 4 real products
#                   of a complex multiply x 3 float16 split terms x 2 (multiply-add)
UNIT = {"fwdbwd": "source*baseline*freq*time evals/s (fwd+bwd)",
        "fwd": "source*baseline*freq*time evals/s (fwd)"}
METRIC = {"fwdbwd": "rime_evals_per_sec_fwd_bwd", "fwd": "rime_evals_per_sec_fwd"}


# --------------------------------------------------------------------------- helpers
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p.get("hbm_gbs"), sm_max_mhz=p.get("sm_max_mhz"),
                    tensor_tflops=p.get("bf16_tflops_sustained") or p.get("bf16_tflops"),
                    tensor_tflops_burst=p.get("bf16_tflops"), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, tensor_tflops=1400.0, tensor_tflops_burst=1590.0,
                source="fallback")


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


DEFAULT_NT = {"c3": 2, "c2": 60, "c1": 10, "c4": 1, "c5": 8}


def build_workload(name, nt, device, rank, world):
    """(rime, params, description, grads, batch indices of this rank, scaling)."""
    import workloads
    from bayeslim_b200 import parallel
    scaling = "weak"
    if name == "c3":
        rime = workloads.pixel_interp(128, 1024, nt * world, device, torch.float32,
                                      antpos_param=True)
        desc = ("C3: HERA-350 (61075 cross baselines) x HEALPix nside-128 PixelSky x rect 1deg "
                "interpolated PixelBeam x 1024 freqs, fwd+bwd to sky, beam, antenna positions")
        grads = "sky,beam,antpos"
    elif name == "c4":
        rime = workloads.pixel_interp_pol(256, 1024, nt * world, device, torch.float32, dgrid=1.0)
        desc = ("C4: HERA-350 (61075 cross baselines) x 4-pol Jones PixelBeam (rect 1deg) x "
                "HEALPix nside-256 PixelSky with Stokes I, Q, U x 1024 freqs, fwd+bwd to sky, beam")
        grads = "sky,beam"
    elif name == "c5":
        rime = workloads.pixel_interp(256, 1024, nt, device, torch.float32, bl_groups=True,
                                      time_groups=True)
        desc = ("C5: HERA-350 x HEALPix nside-256 PixelSky x rect 1deg PixelBeam x 1024 freqs; "
                "minibatch grid of %d single times (of the 2000-time night) x %d block-aligned "
                "baseline groups, a fixed job sharded over the ranks" % (nt, rime.Nbl_groups))
        grads = "sky,beam"
        scaling = "strong"
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
    if name == "c5":
        batches = parallel.shard_units(rime.Nbatch, rank, world)
    else:
        if world > 1:                  # weak scaling: rank r owns times [r*nt, (r+1)*nt)
            rime.setup_sim_times(rime.all_sim_times[rank * nt:(rank + 1) * nt])
        batches = [0]
    params = [p for p in rime.parameters() if p.requires_grad]
    return rime, params, desc, grads, batches, scaling


def flops_per_eval(grads, mode):
    if mode == "fwd":
        return FLOP_FWD
    f = FLOP_FWD + FLOP_BWD_SKY
    if "antpos" in grads:
        f += FLOP_BWD_BL
    return f


# --------------------------------------------------------------------------- reference arms
def _reference_inputs(workload, nbl, nf, nt, dtype):
    """Synthetic inputs of a bounded slice of `workload`, built WITHOUT the product package
    (oracle helpers only): antenna layout, baselines, sky, beam map, per-time (zen, az)."""
    from oracle import rime_oracle as orc
    rng = np.random.default_rng(0)
    gen = torch.Generator(device='cpu').manual_seed(0)
    freqs = torch.linspace(100e6, 200e6, nf, dtype=torch.float64)
    inp = dict(freqs=freqs)
    if workload in ("c3", "c4", "c5"):
        nside = 128 if workload == "c3" else 256
        ants, vecs = orc.hera350()
        # the slice's baselines come from a subset of antennas spread over the array (core and
        # outriggers): the reference's ArrayModel groups all pairs of the antennas it is given
        # into redundant sets at construction, which takes minutes for all 350
        nsub = 2
        while nsub * (nsub - 1) // 2 < nbl:
            nsub += 1
        pick = np.unique(np.linspace(0, len(ants) - 1, nsub).round().astype(int))
        ants, vecs = [ants[i] for i in pick], vecs[pick]
        inp["bls"] = orc.cross_baselines(ants)[:nbl]
        theta, phi = orc.healpix_pix2ang(nside)
        dec = np.pi / 2 - theta
        keep = dec < np.radians(59.27852)
        inp["ra"], inp["dec"] = np.degrees(phi[keep]), np.degrees(dec[keep])
        npix = int(keep.sum())
        spec = (freqs / 150e6) ** -2.5
        base = torch.randn(npix, generator=gen).abs()
        inp["sky_params"] = (spec[:, None] * base[None, :]).to(dtype)[None, None]
        inp["px_area"] = orc.healpix_pixarea(nside)
        tg = torch.arange(0, 90.0 + 1e-6, 1.0, dtype=torch.float64)
        pg = torch.arange(0, 360.0 - 1e-6, 1.0, dtype=torch.float64)
        b_phi, b_theta = torch.meshgrid(pg, tg, indexing='xy')
        d2r = np.pi / 180
        airy = orc.airy_disk(b_theta.ravel() * d2r, b_phi.ravel() * d2r, 14.0, freqs, square=True)
        inp.update(theta_grid=tg, phi_grid=pg, b_theta=b_theta.ravel(), b_phi=b_phi.ravel(),
                   beam_params=airy.to(dtype)[None, None, None])
        inp["kind"] = "interp"
    else:
        ants, vecs = orc.make_hex(4, D=14.6)
        uniq = orc.unique_baselines(ants, vecs)
        inp["bls"] = uniq[:nbl]
        nsrc = 1000 if workload == "c1" else 10000
        inp["ra"] = rng.uniform(0, 360, nsrc)
        inp["dec"] = np.degrees(np.arcsin(rng.uniform(-1, np.sin(np.radians(29)), nsrc)))
        p = np.zeros((1, 1, 2, nsrc))
        p[0, 0, 0] = np.exp(rng.normal(size=nsrc))
        p[0, 0, 1] = rng.normal(-0.8, 0.2, nsrc)
        inp["sky_params"] = torch.as_tensor(p, dtype=dtype)
        inp["beam_params"] = torch.ones(1, 1, 1, 1, 1, dtype=dtype) * 14.0
        inp["kind"] = "airy"
    inp["ants"], inp["vecs"] = list(ants), np.asarray(vecs)
    inp["times"] = np.linspace(2458148.15, 2458148.25, max(nt, 2))[:nt]
    inp["zenaz"] = [orc.eq2top_synth(t, inp["ra"], inp["dec"], lat=-30.72148) for t in inp["times"]]
    return inp


def _load_reference():
    """The unmodified reference installed under baseline/_ref (third-party deps it never calls on
    this path are stubbed by tests/golden/_refshim.py).  None if it is not there."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "bayeslim")):
        return None
    os.environ["BAYESLIM_REFERENCE"] = ref_root
    from tests.golden import _refshim
    _refshim.REFERENCE_ROOT = ref_root
    return _refshim.load()


def _reference_rime(ba, inp, dtype, device):
    """rime_model.RIME of the unmodified reference on the slice `inp`."""
    torch.set_default_dtype(dtype)
    freqs = inp["freqs"].to(dtype)
    # float32 sessions need float32 antenna vectors and telescope angles
    antpos = dict(zip(inp["ants"], torch.as_tensor(inp["vecs"], dtype=dtype)))
    array = ba.telescope_model.ArrayModel(antpos, freqs=freqs)
    angs = torch.as_tensor(np.stack([inp["ra"], inp["dec"]]))
    if inp["kind"] == "interp":
        sky = ba.sky_model.PixelSky(inp["sky_params"].clone(), angs, inp["px_area"],
                                    R=ba.sky_model.PixelSkyResponse(freqs), parameter=True)
        R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta=inp["b_theta"],
                                        phi=inp["b_phi"], theta_grid=inp["theta_grid"],
                                        phi_grid=inp["phi_grid"], freq_mode='channel',
                                        powerbeam=True, realbeam=True, log=False)
        beam = ba.beam_model.PixelBeam(inp["beam_params"].clone(), freqs, R=R, pol='e',
                                       powerbeam=True, fov=180, parameter=True)
    else:
        R = ba.sky_model.PointSkyResponse(freqs, freq_mode='powerlaw', f0=150e6)
        sky = ba.sky_model.PointSky(inp["sky_params"].clone(), angs, R=R, parameter=True)
        beam = ba.beam_model.PixelBeam(inp["beam_params"].clone(), freqs,
                                       R=ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                       powerbeam=True, fov=180, parameter=False)
    tel = ba.telescope_model.TelescopeModel((21.42827, -30.72148, 1051.7), dtype=dtype)
    rime = ba.rime_model.RIME(sky, tel, beam, array, inp["bls"], inp["times"], freqs,
                              device=None if device == 'cpu' else device)
    for t, (zen, az) in zip(rime.sim_times, inp["zenaz"]):
        za = torch.stack([torch.as_tensor(zen), torch.as_tensor(az)]).to(dtype)
        rime.telescope.conv_cache[(sky.name, len(inp["ra"]), t)] = za
    if device != 'cpu':
        for obj in (sky, beam, array, tel, rime):
            obj.push(device)
    return rime, [p for p in (sky.params, beam.params) if p.requires_grad]


def reference_arm(workload, steps, warmup, mode="fwdbwd", device='cpu', size="full", threads=None):
    """Time the reference's own implementation of the path (rime_model.RIME.forward + autograd
    backward of the unmodified package under baseline/_ref; falls back to the restated port
    oracle/rime_oracle.py when it is absent) on a bounded slice of `workload`.  float32 is the
    headline (the GPU arm's dtype), float64 is timed beside it."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    slices = {"c3": (64, 128, 2), "c4": (64, 128, 2), "c5": (64, 128, 2), "c2": (63, 256, 2),
              "c1": (63, 64, 10)}
    nbl, nf, nt = slices[workload]
    if size == "small":                 # the cpu_baseline leg of the default GPU run (~10 s)
        nbl, nf, nt = {"c1": (63, 64, 10), "c2": (63, 256, 1)}.get(workload, (32, 128, 1))
    ba = _load_reference()
    kind = "reference" if ba is not None else "port"
    out = {}
    for dtype in (torch.float32, torch.float64):
        inp = _reference_inputs(workload, nbl, nf, nt, dtype)
        nsrc = sum(int((np.asarray(z) < 90).sum()) for z, _ in inp["zenaz"])
        evals = nsrc * len(inp["bls"]) * nf
        if ba is not None:
            rime, params = _reference_rime(ba, inp, dtype, device)

            def step():
                for p in params:
                    p.grad = None
                if mode == "fwd":
                    with torch.no_grad():
                        V = rime().data
                    return float((V.real ** 2 + V.imag ** 2).sum())
                V = rime().data
                loss = (V.real ** 2 + V.imag ** 2).sum()
                loss.backward()
                return float(loss)
        else:
            step = _port_step(inp, dtype, mode)
        nsteps = steps if dtype == torch.float32 else 1
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            step()
        if device != 'cpu':
            torch.cuda.synchronize()
        dt_s = (time.perf_counter() - t0) / nsteps
        out[str(dtype)[6:]] = dict(value=evals / dt_s, ms_per_step=dt_s * 1e3, steps=nsteps,
                                   evals_per_step=evals)
    torch.set_default_dtype(torch.float32)
    sample = ("%s slice: %d baselines x %d freqs x %d time(s) x all sources above the horizon "
              "(the reference materialises the (Nbl, Nf, Nsrc) fringe tensor: the full problem "
              "does not fit)" % (workload.upper(), nbl, nf, nt))
    return dict(kind=kind, cores=threads, sample=sample, float32=out["float32"],
                float64=out["float64"], value=out["float32"]["value"],
                ms_per_step=out["float32"]["ms_per_step"], steps=out["float32"]["steps"])


def _port_step(inp, dtype, mode):
    """Same slice through the restated port (oracle/rime_oracle.py)."""
    from oracle import rime_oracle as orc
    freqs = inp["freqs"].to(dtype)
    sp = inp["sky_params"].clone().requires_grad_(True)
    bp = inp["beam_params"].clone().requires_grad_(inp["kind"] == "interp")
    antvecs = torch.as_tensor(inp["vecs"], dtype=dtype)
    zenaz = [(torch.as_tensor(z).to(dtype), torch.as_tensor(a).to(dtype)) for z, a in inp["zenaz"]]

    def step():
        for p in (sp, bp):
            p.grad = None
        blvecs = orc.get_blvecs(antvecs, inp["ants"], inp["bls"])
        if inp["kind"] == "interp":
            sky = sp * float(inp["px_area"])
            bmap = orc.pixel_response_forward(bp, powerbeam=True)

            def beam_fn(z, a):
                inds, wgts = orc.rect_interp_weights(inp["theta_grid"], inp["phi_grid"], z, a, 'linear')
                return orc.interp_map(bmap, inds, wgts.to(dtype))
        else:
            sky = orc.point_sky_response(sp, freqs, 'powerlaw', f0=150e6)
            beam_fn = lambda z, a: orc.airy_response(bp, z, a, freqs, powerbeam=True)
        with torch.set_grad_enabled(mode != "fwd"):
            V = orc.rime_forward(sky, zenaz, beam_fn, inp["bls"], blvecs, freqs, fov=180.0)
            loss = (V.real ** 2 + V.imag ** 2).sum()
        if mode != "fwd":
            loss.backward()
        return float(loss)
    return step


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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c1", "c4", "c5"])
    ap.add_argument("--nt", type=int, default=None,
                    help="times per step per GPU (c5: times of the fixed job)")
    ap.add_argument("--pass", dest="mode", default="fwdbwd", choices=["fwdbwd", "fwd"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true",
                    help="capture the step (forward + loss + backward) into a CUDA graph and "
                         "time its replay (rime_model.GraphedStep; single-minibatch workloads)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    nt = args.nt or DEFAULT_NT[args.workload]
    unit, metric = UNIT[args.mode], METRIC[args.mode]

    if args.impl in ("reference", "reference-gpu"):
        if rank != 0:
            return
        gpu = args.impl == "reference-gpu"
        steps = 5 if not gpu else max(1, min(args.steps, 5))   # CPU arm: always 5 timed steps of ~10 s
        warm = max(1, min(args.warmup, 1))
        try:
            r = reference_arm(args.workload, steps, warm, mode=args.mode,
                              device=("cuda:0" if gpu else "cpu"))
        except Exception as e:      # noqa: BLE001 -- the informational GPU arm may not run
            if not gpu:
                raise
            emit(dict(impl=args.impl, unavailable="%s: %s" % (type(e).__name__, str(e)[:200])))
            return
        line = dict(metric=metric, value=r["value"], unit=unit, n_gpus=args.gpus, steps=r["steps"],
                    warmup=warm, ms_per_step=r["ms_per_step"], higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    impl=args.impl,
                    config=dict(workload=args.workload.upper() + " (bounded sample)",
                                sample=r["sample"], device=("B200 (torch eager)" if gpu else "host cores")),
                    cpu_baseline=dict(value=r["value"], unit=unit, cores=r["cores"], kind=r["kind"],
                                      sample=r["sample"], float32=r["float32"],
                                      float64=r["float64"]),
                    e2e=dict(value=r["value"], unit=unit, h2d_bytes_per_step=0,
                             d2h_bytes_per_step=0),
                    gpu_launches=0)
        emit(line)
        return

    import torch.distributed as dist
    import bayeslim_b200 as ba  # noqa: F401
    from bayeslim_b200 import ops, parallel, _lib
    import workloads

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))

    rime, params, desc, grads, batches, scaling = build_workload(args.workload, nt, device, rank,
                                                                 world)
    fwd_only = args.mode == "fwd"

    graphed = []

    def step(e2e_buffers=None):
        if e2e_buffers is not None:
            for p, h in zip(params, e2e_buffers["h_in"]):
                p.data.copy_(h, non_blocking=True)
        if graphed:
            total = graphed[0]()                       # one graph launch: forward + loss + backward
            if world > 1:
                parallel.allreduce_gradients(params)
            if e2e_buffers is not None:
                for p, h in zip(params, e2e_buffers["h_out"]):
                    h.copy_(p.grad, non_blocking=True)
                e2e_buffers["loss"] = float(total)
            return total
        for p in params:
            p.grad = None
        total = None
        for b in batches:
            if rime.Nbatch > 1:
                rime.batch_idx = b
            if fwd_only:
                with torch.no_grad():
                    V = rime().data
                    loss = (V.real ** 2 + V.imag ** 2).sum()
            else:
                V = rime().data
                loss = (V.real ** 2 + V.imag ** 2).sum()
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
            del V, loss
        if world > 1 and not fwd_only:
            parallel.allreduce_gradients(params)
        if e2e_buffers is not None:
            if not fwd_only:
                for p, h in zip(params, e2e_buffers["h_out"]):
                    h.copy_(p.grad, non_blocking=True)
            e2e_buffers["loss"] = float(total) if total is not None else 0.0   # D2H + sync
        return total

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
    evals_rank = 0
    for b in batches:
        if rime.Nbatch > 1:
            rime.batch_idx = b
        evals_rank += workloads.count_evals(rime)
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
    ms_eager = None
    if args.graph:
        # per-kernel durations come from the eager pass above; the reported step is the replay
        assert len(batches) == 1 and not fwd_only, "--graph: one minibatch, forward + backward"
        from bayeslim_b200.rime_model import GraphedStep
        if rime.Nbatch > 1:
            rime.batch_idx = batches[0]
        graphed.append(GraphedStep(rime, lambda vd: (vd.data.real ** 2 + vd.data.imag ** 2).sum(),
                                   params))
        for _ in range(3):
            step()
        ms_eager, ms_step = ms_step, timed(args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end: parameters from pinned host memory in, loss (+ gradients) out, every step
    bufs = dict(h_in=[p.detach().cpu().pin_memory() for p in params],
                h_out=[torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in params])
    step(e2e_buffers=bufs)
    ms_e2e = timed(args.steps, e2e_buffers=bufs)
    h2d = sum(h.numel() * h.element_size() for h in bufs["h_in"])
    d2h = (0 if fwd_only else sum(h.numel() * h.element_size() for h in bufs["h_out"])) + 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- post-timing parity check: a subset of this rank's visibilities against the fp64 port
    parity = None
    try:
        from oracle import workload_check
        if rime.Nbatch > 1:
            rime.batch_idx = batches[0]
        with torch.no_grad():
            V = rime().data
        nbl = len(rime.sim_bls)
        blen = rime.sim_blvecs.detach().norm(dim=1).cpu().numpy()
        order = np.argsort(blen)
        sel = sorted(set(order[:2].tolist() + order[-2:].tolist()
                         + np.random.default_rng(3).choice(nbl, min(4, nbl), replace=False).tolist()))
        nf = V.shape[-1]
        f_idx = list(range(0, nf, max(1, nf // 8)))[:8]
        zenaz = workloads.zenaz_of(rime)[:1]
        Vo = workload_check.oracle_vis_subset(rime, zenaz, sel, f_idx)
        sub = V[:, :, sel][:, :, :, :1][..., f_idx].cpu().to(torch.complex128)
        parity = dict(relmax_vs_fp64_port=float((sub - Vo).abs().max() / V.abs().max().cpu()),
                      baselines=len(sel), channels=len(f_idx), times=1, tolerance=1e-5,
                      norm="max |V| of the rank's visibility tensor")
        del V
    except Exception as e:      # noqa: BLE001 -- the check must not lose the measurement
        parity = dict(error="%s: %s" % (type(e).__name__, str(e)[:200]))

    # ---- rooflines
    peaks = load_peaks()
    info = _lib.device_info(local)
    fp32_meas, _ = _lib.microbench("fp32", 4096)
    fp32x2_meas, _ = _lib.microbench("fp32x2", 4096)
    fp64_meas, _ = _lib.microbench("fp64", 1024)
    mufu_meas, _ = _lib.microbench("mufu", 2048)
    fp32_theory = 2 * 128 * info["sm_count"] * (peaks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
    k = dict(ksum)

    # evaluations and padded sources the timed launches of each kernel family covered
    nf = len(rime.array.freqs)
    nfp = -(-nf // _lib.KC["f32"]) * _lib.KC["f32"]
    nplane = 1
    if args.workload == "c4":
        nplane = 4
    geo = {}
    for b in batches:
        if rime.Nbatch > 1:
            rime.batch_idx = b
        times = tuple(float(t) for t in rime.sim_times)
        recs = [r for key, r in rime._geom_cache.items() if key[2] == times]
        geo[b] = dict(ns=sum(sum(r.geom.ns) for r in recs), S=sum(r.geom.S for r in recs),
                      tc=rime._tc_tilings.get((rime._bl_key, torch.device(device))),
                      til=rime._ant_tilings.get((rime._bl_key, torch.device(device))),
                      nbl=len(rime.sim_bls))
    executed = dict(tcfringe_fwd=0.0, tcfringe_bwd=0.0, antfringe_fwd=0.0, antfringe_bwd=0.0)
    need_r = "antpos" in grads
    for g in geo.values():
        if g["tc"] is not None:
            executed["tcfringe_fwd"] += FLOP_MMA * g["tc"].pair_slots * g["S"] * nf * nplane
            stages = g["tc"].bwd_stages_full if need_r else g["tc"].bwd_stages_lower
            executed["tcfringe_bwd"] += FLOP_MMA * 128.0 * 16 * stages * g["S"] * nf * nplane
        if g["til"] is not None:
            executed["antfringe_fwd"] += 8.0 * g["til"].pair_slots * g["S"] * nfp * nplane
            executed["antfringe_bwd"] += 8.0 * g["til"].bwd_rows * g["til"].nm_pad * g["S"] * nfp * nplane

    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload:
            traffic = {kn: v for kn, v in tj["kernels"].items()}

    def entry(name, alg_flop):
        if name not in k or k[name]["ms"] <= 0:
            return None
        sec = k[name]["ms"] * 1e-3
        alg = evals_rank * args.steps * alg_flop / sec / 1e12
        tensor = name.startswith("tc")
        peak = peaks["tensor_tflops"] if tensor else fp32_theory
        ex = executed.get(name, 0.0) * args.steps / sec / 1e12 if executed.get(name) else alg
        d = dict(bound="tensor" if tensor else "fp32", kernel=name + "_f32",
                 achieved=ex, peak=peak, unit="TFLOP/s", frac=ex / peak,
                 achieved_is="executed flops: " + (
                     "24 per (antenna-pair slot, source, channel) = 4 real products x 3 float16 "
                     "split MMAs x 2, tile padding included" if tensor else
                     "8 per computed antenna pair (factorised) or the SURVEY count (baseline-owned)"),
                 peak_source=("MEASURED_PEAKS.json bf16_tflops_sustained (%s); float16 runs at the "
                              "same tensor rate" % peaks["source"]) if tensor else
                             "theoretical 2 x 128 lanes x SMs x max clock (MEASURED_PEAKS.json has no FP32 figure)",
                 algorithmic_tflops=alg, flop_per_eval_algorithmic=alg_flop,
                 frac_algorithmic_of_fp32_peak=alg / fp32_theory,
                 ms_per_launch=k[name]["ms"] / max(k[name]["launches"], 1),
                 launches_per_step=k[name]["launches"] / args.steps,
                 traffic=(traffic.get(name, {}).get("dram_bytes_per_launch")),
                 traffic_unit="DRAM bytes per launch (ncu dram__bytes_read+write, profiles/r02_traffic.json)")
        return d

    table = (("tcfringe_fwd", FLOP_FWD), ("tcfringe_bwd", FLOP_BWD_SKY + (FLOP_BWD_BL if need_r else 0)),
             ("antfringe_fwd", FLOP_FWD), ("antfringe_bwd", FLOP_BWD_SKY + (FLOP_BWD_BL if need_r else 0)),
             ("fringe_sum_fwd", FLOP_FWD), ("fringe_sum_bwd_sky", FLOP_BWD_SKY),
             ("fringe_sum_bwd_bl", FLOP_BWD_BL))
    entries = {n: entry(n, fl) for n, fl in table}
    entries = {n: e for n, e in entries.items() if e}
    dominant = max(entries, key=lambda n: k[n]["ms"]) if entries else None
    roofline = entries.get(dominant)
    others = {n: e for n, e in entries.items() if n != dominant}
    hbm = None
    bname = "build_airy"
    if args.workload in ("c3", "c5"):
        bname = "build_interp_t" if "build_interp_t" in k else "build_interp"
    if bname in k and k[bname]["ms"] > 0 and args.workload != "c4":
        nsrc = sum(g["ns"] for g in geo.values())
        npb = rime.beam.params.shape[-1] if args.workload in ("c3", "c5") else 0
        # per launch (all times of the step): read the beam map once, read sky at the cut,
        # write A, read 4 idx + 4 wgt
        bytes_step = 4 * nf * (npb * len(geo) + 2 * nsrc) + 32 * nsrc
        gbs = bytes_step * args.steps / (k[bname]["ms"] * 1e-3) / 1e9
        hbm = dict(bound="hbm", kernel=bname + "_f32", achieved=gbs, peak=peaks["hbm_gbs"],
                   unit="GB/s", frac=gbs / peaks["hbm_gbs"], peak_source=peaks["source"])

    kernel_ms = {n: d["ms"] / args.steps for n, d in k.items()}
    line = dict(
        metric=metric, value=evals_total / (ms_step * 1e-3), unit=unit,
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
        higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=desc, times_per_gpu=(nt if args.workload != "c5" else None),
                    work_units_rank0=len(batches), grads=("none (forward)" if fwd_only else grads),
                    evals_per_step=evals_total,
                    flop_per_eval=flops_per_eval(grads, args.mode),
                    l2="inputs larger than L2 (perceived-sky slab and cotangent / operand "
                       "matrices are 0.4 - 1.6 GB per time)",
                    parallelism=("minibatch grid (time x baseline group) sharded x%d" % world
                                 if args.workload == "c5" else "time-sharded x%d" % world)
                    + ("" if fwd_only else ", all-reduce of the gradients")),
        clocks=clocks,
        e2e=dict(value=evals_total / (ms_e2e * 1e-3), unit=unit, ms_per_step=ms_e2e,
                 h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
        gpu_launches=launches,
        roofline=roofline, roofline_other_kernels=others, roofline_hbm=hbm,
        parity_check=parity,
        fp32_tflops_algorithmic=evals_total * flops_per_eval(grads, args.mode) / (ms_step * 1e-3) / 1e12 / world,
        peaks=dict(fp32_tflops_measured=fp32_meas / 1e3, fp32x2_tflops_measured=fp32x2_meas / 1e3,
                   fp64_tflops_measured=fp64_meas / 1e3,
                   mufu_gops_measured=mufu_meas, fp32_tflops_theoretical=fp32_theory,
                   tensor_tflops_measured_sustained=peaks["tensor_tflops"],
                   tensor_tflops_measured_burst=peaks["tensor_tflops_burst"],
                   sm_count=info["sm_count"]),
        kernel_ms_per_step=kernel_ms,
        kernel_share_of_step={n: v / ms_step for n, v in kernel_ms.items()},
        kernel_launches_per_step={n: d["launches"] / args.steps for n, d in k.items()},
    )
    if args.graph:
        line["cuda_graph"] = dict(
            note="the timed step is ONE replay of a CUDA graph holding the forward, the loss and "
                 "the backward (rime_model.GraphedStep); gpu_launches counts the library kernels "
                 "inside the replayed graphs; per-kernel durations are from the eager pass",
            ms_per_step_eager=ms_eager, graph_replays_per_step=1)
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_arm(args.workload, 1, 1, mode=args.mode, size="small")
            line["cpu_baseline"] = dict(value=r["value"], unit=unit, cores=r["cores"], kind=r["kind"],
                                        sample=r["sample"], float64_value=r["float64"]["value"])
        except Exception as e:      # noqa: BLE001
            line["cpu_baseline"] = dict(error="%s: %s" % (type(e).__name__, str(e)[:200]))
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
